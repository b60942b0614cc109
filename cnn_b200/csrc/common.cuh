// common.cuh -- shared declarations of libcnn_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>

#include "cnn_b200.h"

struct cnn_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int conv_algo = CNN_CONV_AUTO;
    int tc_precision = CNN_TC_TF32X3;
    int sm_count = 148;
    long long launches = 0;
    int thin_slot = -1;  // __constant__ filter bank of the thin first-layer kernels (conv_thin.cu), -1 = none
    // scratch for split reductions (BN statistics, conv weight-gradient partials)
    float* scratch = nullptr;
    size_t scratch_bytes = 0;
    // side arena of the per-operator entry points (packed operands of conv_s2.cu, transposed Linear
    // weights): separate from `scratch`, which the kernels called underneath use themselves
    void* arena = nullptr;
    size_t arena_bytes = 0;
    // data parallelism (dist.cu): NCCL communicator of this rank, one process per GPU
    void* nccl_comm = nullptr;
    int dist_rank = 0, dist_world = 1;
    bool sync_bn = false;   // BatchNorm statistics over the global batch (all-reduced), see bn.cu
    // per-launch profile (cnn_prof_*): CUDA events around every CNN_LAUNCH while enabled
    struct cnn_prof* prof = nullptr;
    int prof_tag = -1;      // set by the engine: layer * 4 + pass (0 forward, 1 backward, 2 update / loss), -1 = none
};

struct cnn_prof {
    static constexpr int kMax = 512;
    cudaEvent_t ev[kMax + 1];
    const char* name[kMax];
    int tag[kMax];
    int n = 0;
    bool on = false;
};
void cnn_prof_mark(cnn_ctx* ctx, const char* name);   // event before a launch

// one process-wide lock for the library's few lazily initialised process-wide tables (kernel attribute flags,
// the conv_tc plan cache): contexts may be driven from different host threads, one per GPU
std::recursive_mutex& cnn_global_mutex();
void cnn_set_error(const char* fmt, ...);
int cnn_cuda_fail(cudaError_t e, const char* what, const char* file, int line);
float* cnn_scratch(cnn_ctx* ctx, size_t bytes);  // grows on demand; nullptr on failure
void* cnn_arena(cnn_ctx* ctx, size_t bytes);     // same for the side arena (256-byte aligned)

// 3-D fp32 TMA tensor map (dims / box innermost first, strides of dims 1 and 2 in bytes, multiples of 16)
int cnn_tmap_encode(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box);   // 4-byte elements, rank 2..5
int cnn_tmap_encode_3d(CUtensorMap* map, const void* base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                       const uint32_t box[3]);

#define CNN_CUDA(call)                                                             \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) return cnn_cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)

#define CNN_REQUIRE(cond, ...)       \
    do {                             \
        if (!(cond)) {               \
            cnn_set_error(__VA_ARGS__); \
            return CNN_ERR_ARG;      \
        }                            \
    } while (0)

// every kernel launch goes through this so the context can count launches and so that a
// launch-configuration error surfaces at the call site
#define CNN_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                                  \
        if ((ctx)->prof && (ctx)->prof->on) cnn_prof_mark((ctx), #kernel);                \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                  \
        ++(ctx)->launches;                                                                \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess) return cnn_cuda_fail(e_, #kernel, __FILE__, __LINE__);     \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Packed FP32 FMA (Blackwell FFMA2): two IEEE fused multiply-adds per instruction, bit-identical to
// two fmaf calls.  A 3-register FFMA issues every other cycle per scheduler, so fp32 CUDA-core kernels
// only reach the nominal FP32 rate through this form; operands may be a broadcast scalar register or a
// uniform-register pair loaded from the constant bank (ptxas picks those encodings by itself).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&r);
}

// block-wide sum; result valid in thread 0 (and broadcast to all when kBroadcast)
template <bool kBroadcast = false>
__device__ __forceinline__ float block_sum(float v, float* smem32) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem32[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float r = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
    if (wid == 0) r = warp_sum(r);
    if (kBroadcast) {
        if (threadIdx.x == 0) smem32[0] = r;
        __syncthreads();
        r = smem32[0];
        __syncthreads();
    } else {
        __syncthreads();
    }
    return r;
}

// conv algorithm entry points (conv_simt.cu / conv_tc.cu)
int conv_fwd_simt(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, int B,
                  int Cin, int H, int W, int Cout, int k, int s);
int conv_wgrad_simt(cnn_ctx*, const float* x, const float* delta, float* dw, float* db, int B,
                    int Cin, int H, int W, int Cout, int k, int s, float scale);
int conv_dgrad_simt(cnn_ctx*, const float* w, const float* delta, float* dx, int B, int Cin,
                    int H, int W, int Cout, int k, int s);
bool conv_tc_supported(int Cin, int Cout, int k, int s);
int conv_fwd_tc(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, int B,
                int Cin, int H, int W, int Cout, int k, int s);
int conv_wgrad_tc(cnn_ctx*, const float* x, const float* delta, float* dw, float* db, int B,
                  int Cin, int H, int W, int Cout, int k, int s, float scale);
int conv_dgrad_tc(cnn_ctx*, const float* w, const float* delta, float* dx, int B, int Cin, int H,
                  int W, int Cout, int k, int s);

// thin first layer (Cin=3, Cout=16, 3x3, stride 2): fp32 CUDA-core kernels with constant-bank filters
int conv_thin_acquire_slot(int device);
void conv_thin_release_slot(int device, int slot);
bool conv_thin_supported(const cnn_ctx*, int Cin, int H, int W, int Cout, int k, int s);
int conv_fwd_thin(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, int B, int H, int W);
int conv_dgrad_thin(cnn_ctx*, const float* w, const float* delta, float* dx, int B, int H, int W);
bool conv_thin_pool_supported(const cnn_ctx*, int Cin, int H, int W, int Cout, int k, int s, int pk, int pstep);
bool conv_thin_pool_preferred();
int conv_fwd_thin_relu_pool(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, float* y_relu,
                            float* y_pool, int32_t* mask, int B, int H, int W);
int conv_wgrad_thin(cnn_ctx*, const float* x, const float* delta, float* dw, float* db, int B, int H, int W,
                    float scale);
// lazy training head (conv 3->16 s2 -> ReLU -> MaxPool 2x2/2, conv_thin.cu): forward keeps only the pooled
// activations (packed for the following s2 conv and / or fp32) and one code byte per pool window and channel;
// the weight gradient reads the pooled delta through those codes.  w_save: [448] copy of filters + biases.
bool conv_head_lazy_supported(const cnn_ctx*, int Cin, int H, int W, int Cout, int k, int s, int pk, int pstep);
size_t conv_head_m8_bytes(int B, int H, int W);
int conv_head_fwd(cnn_ctx*, const float* x, const float* w, const float* bias, float* w_save, void* next_px,
                  float* pool, void* m8, int B, int H, int W);
int conv_head_wgrad(cnn_ctx*, const float* x, const float* dpool_nhwc, const void* m8, float* dw, float* db, int B,
                    int H, int W, float scale);

// 3x3 stride-2 layers with Cin, Cout multiples of 16: packed parity-plane operands + shifted-window
// tcgen05 GEMMs (conv_s2.cu).  y_relu / relu_y are optional fused ReLU outputs / masks.
bool conv_s2_supported(const cnn_ctx*, int Cin, int H, int W, int Cout, int k, int s);
int conv_fwd_s2(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, float* y_relu, int B, int Cin,
                int H, int W, int Cout);
int conv_dgrad_s2(cnn_ctx*, const float* w, const float* delta, float* dx, const float* relu_y, int B, int Cin, int H,
                  int W, int Cout);
int conv_wgrad_s2(cnn_ctx*, const float* x, const float* delta, float* dw, float* db, int B, int Cin, int H, int W,
                  int Cout, float scale);
// packed-operand interface used by the engine (net.cu): P(x) survives from forward to weight gradient,
// delta is packed once for both gradients
size_t conv_s2_px_bytes(int B, int Cin, int H, int W);
void conv_s2_px_geom(int B, int Cin, int H, int W, int* HP, int* PP, long long* RUNX);   // plane geometry of P(x)
size_t conv_s2_pd_bytes(int B, int Cout, int H, int W);
size_t conv_s2_dbp_bytes(int B, int Cout, int H, int W);
int conv_s2_pack_x(cnn_ctx*, const float* x, void* px, int B, int Cin, int H, int W);
int conv_s2_pack_d(cnn_ctx*, const float* delta, void* pd, float* dbp, int B, int Cout, int H, int W);
// wpk_ready: filter blocks packed earlier by conv_s2_pack_weights (null: packed into scratch by the call)
struct ConvS2PackJob { const float* w; void* out; int Cin, Cout, dgrad; };
size_t conv_s2_wpk_bytes(int Cin, int Cout, int dgrad);
int conv_s2_pack_weights(cnn_ctx*, const ConvS2PackJob* jobs, int n);   // one launch, n <= 8
// next_px: packed input buffer of a following s2 layer fed by this layer's ReLU output (written by the epilogue;
// its padding positions must have been zeroed once), or null
int conv_s2_fwd_packed(cnn_ctx*, const void* px, const float* w, const void* wpk_ready, const float* bias, float* y,
                       float* y_relu, int B, int Cin, int H, int W, int Cout, void* next_px);
// dx_nhwc: dx written channel-last [B][H][W][Cin] (relu_y must be null)
int conv_s2_dgrad_packed(cnn_ctx*, const void* pd, const float* w, const void* wpk_ready, float* dx, const float* relu_y,
                         int B, int Cin, int H, int W, int Cout, bool dx_nhwc = false);
int conv_s2_wgrad_packed(cnn_ctx*, const void* px, const void* pd, const float* dbp, float* dw, float* db, int B,
                         int Cin, int H, int W, int Cout, float scale);

// 3x3 stride-1 layers with Cin, Cout multiples of 32 (conv_s1.cu): packed one-plane operands, shifted-window
// tcgen05 GEMMs with chunked accumulation.  y_relu / relu_y are optional fused ReLU outputs / masks.
bool conv_s1_supported(const cnn_ctx*, int Cin, int H, int W, int Cout, int k, int s);
int conv_fwd_s1(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, float* y_relu, int B, int Cin,
                int H, int W, int Cout);
int conv_dgrad_s1(cnn_ctx*, const float* w, const float* delta, float* dx, const float* relu_y, int B, int Cin, int H,
                  int W, int Cout);
int conv_wgrad_s1(cnn_ctx*, const float* x, const float* delta, float* dw, float* db, int B, int Cin, int H, int W,
                  int Cout, float scale);
// thin first layer (3 -> 16 / 32 / 64 channels, 3x3, stride 1) with the following ReLU and, optionally, the packed input
// of a stride-1 layer that consumes the ReLU output -- one kernel, one pass over the outputs
bool conv_s1_first_supported(const cnn_ctx*, int Cin, int H, int W, int Cout, int k, int s);
int conv_s1_first_fwd(cnn_ctx*, const float* x, const float* w, const float* bias, float* y, float* y_relu, void* next_px, int B,
                      int H, int W, int Cout);
int conv_s1_first_wgrad(cnn_ctx*, const float* x, const float* delta, float* dw, float* db, int B, int H, int W, int Cout,
                        float scale);
// packed-operand interface used by the engine: src [B][C][SH][SW] sits in the top-left corner of the (H, W) pitch
// geometry of the layer INPUT; relu_y folds the ReLU backward of the layer above into the packing; dbp = bias-gradient partials
size_t conv_s1_pk_bytes(int B, int C, int H, int W);
size_t conv_s1_dbp_bytes(int B, int C, int H, int W);
int conv_s1_pack(cnn_ctx*, const float* src, const float* relu_y, void* dst, float* dbp, int B, int C, int H, int W, int SH,
                 int SW);
int conv_s1_fwd_packed(cnn_ctx*, const void* px, const float* w, const float* bias, float* y, float* y_relu, int B, int Cin,
                       int H, int W, int Cout);
int conv_s1_dgrad_packed(cnn_ctx*, const void* pd, const float* w, float* dx, const float* relu_y, int B, int Cin, int H, int W,
                         int Cout);
int conv_s1_wgrad_packed(cnn_ctx*, const void* px, const void* pd, const float* dbp, float* dw, float* db, int B, int Cin,
                         int H, int W, int Cout, float scale);

// SGD with the learning rate in device memory (elementwise.cu): graph replays are independent of its value
int cnn_sgd_step_dev_lr(cnn_ctx*, float* params, const float* grads, size_t n, const float* lr_dev);
int cnn_set_scalar(cnn_ctx*, float* dst, float v);
// one-shot gradient exchange + SGD over NVLink peer memory (dist.cu); setup is collective over the ranks
int cnn_peer_exchange_setup(cnn_ctx*, float* grads, float* params, size_t P, void** state_out);
int cnn_peer_exchange_bulk(cnn_ctx*, void* state, const float* lr_dev, int do_sgd, size_t lo);
int cnn_peer_exchange_step(cnn_ctx*, void* state, const float* lr_dev, int do_sgd, size_t hi);
void cnn_peer_exchange_destroy(void* state);

// LinearLayer::backward with the in-place ReLU backward of the layer below folded into dx (relu_y may be null)
int linear_backward_relu(cnn_ctx*, const float* x, const float* w, const float* delta, float* dw, float* db, float* dx,
                         const float* relu_y, int B, int in, int out, float scale);
