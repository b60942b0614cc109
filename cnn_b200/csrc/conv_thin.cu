// conv_thin.cu -- the "thin" first layer of the reference net (Conv2D 3 -> 16, 3x3, stride 2 on
// 224x224 images, alexnet.cpp:9), forward and input gradient, as fp32 CUDA-core kernels.
//
// Why not tensor cores here: with 3 input channels the GEMM is 27 deep and 16 wide.  The split-fp32
// tcgen05 path (conv_tc.cu) then spends its time writing and re-reading hi/lo operand tiles in shared
// memory (measured 137 us forward / 294 us input gradient at B=256, ncu + clock64 traces in
// profiles/), while the arithmetic itself is only 1.36 GFMA = 38 us of the FP32 pipe.  These kernels
// keep the FP32 pipe busy instead:
//   * the 432 filter taps + 16 biases sit in __constant__ memory and are FFMA constant operands
//     (no load instruction, no register per weight); the inner loops are fully unrolled
//   * the source rows of a tile (whole image rows) are streamed into shared memory by the TMA bulk
//     engine (cp.async.bulk + mbarrier, double-buffered, one tile ahead); a thread reads 27 (forward)
//     or 64 (input gradient) shared-memory words per 432 FFMAs
//   * one thread = one output pixel x 16 channels (forward) or one 2x2 input patch x 3 channels
//     (input gradient: each delta is read once, not 2.25x); every global store is coalesced.
// Results: fp32 FMA chains in the reference's loop order (conv2d.cpp:69-92, 161-201).
//
// __constant__ banks are per process and device, so a context owns one of kSlots banks for its life
// time (acquired in cnn_ctx_create); contexts beyond that simply keep using the generic kernels.
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kSlots = 4;
constexpr int kCin = 3, kCout = 16, kK = 3, kS = 2;
constexpr int kNW = kCout * kCin * kK * kK;
constexpr int kThinThreads = 256;

struct ThinConst {
    float w[kNW];      // [co][ci][ky][kx], the reference's filter order (input gradient: 9 taps per LDCU run)
    float wt[kNW];     // [ci][ky][kx][co]: the 16 channels of a tap are contiguous (forward: 4 x LDCU.128 per tap)
    float b[kCout];
};
__constant__ ThinConst c_thin[kSlots];

struct ThinArgs {
    const float* src;     // forward: x ; input gradient: delta
    float* dst;           // forward: y ; input gradient: dx
    // fused forward (conv -> ReLU -> 2x2/2 MaxPool, alexnet.cpp:12-16): the two following layers' outputs
    float* dst_relu;      // [B][16][OH][OW]
    float* dst_pool;      // [B][16][POH][POW]
    int32_t* mask;        // pool arg-max indices (pool2d.cpp:81), may be null (no_grad)
    int POH, POW;
    int OWp;              // thread mapping pitch: OW rounded up to even (fused) or OW
    int nbuf;             // forward: staged-row ring depth (2..4)
    int B, H, W, OH, OW;  // image and output geometry
    int GH, GW;           // tile row space: forward = (OH, OW), input gradient = patches (ceil(H/2), ceil(W/2))
    int TR, SCI;          // row-space rows per tile, tiles per image
    int seg;              // bytes per staged channel segment (multiple of 16)
    unsigned tiles;
    long long src_bytes16;
};

struct TileWalk {   // tile -> (image, row group) without divisions inside the loop
    int b, gi, step_b, step_gi, SCI;
    __device__ TileWalk(unsigned first, unsigned stride, int sci) : SCI(sci) {
        b = (int)(first / (unsigned)sci); gi = (int)(first % (unsigned)sci);
        step_b = (int)(stride / (unsigned)sci); step_gi = (int)(stride % (unsigned)sci);
    }
    __device__ void next() {
        gi += step_gi;
        if (gi >= SCI) { gi -= SCI; ++b; }
        b += step_b;
    }
};

// Warp 0 streams the source rows [r0, r1) of NCH channels of image b into one raw buffer.  NCHW rows
// are only 4-byte aligned: every copy starts at the preceding 16-byte boundary, the consumer adds
// (e0 & 3) back.  The transaction count is posted after the copies.
template <int NCH>
__device__ __forceinline__ void stream_rows(const ThinArgs& p, int lane, uint8_t* raw, uint64_t* bar, int b,
                                            int SH, int SW, int r0, int r1) {
    uint32_t mine = 0;
    const long long nfl = (long long)(r1 - r0) * SW;
    if (lane < NCH && nfl > 0) {
        const long long e0 = ((long long)(b * NCH + lane) * SH + r0) * SW;
        const long long ea = e0 & ~3ll;
        long long bytes = (((e0 - ea) + nfl) * 4 + 15) & ~15ll;
        if (ea * 4 + bytes > p.src_bytes16) bytes = p.src_bytes16 - ea * 4;
        tma_bulk_g2s(raw + (size_t)lane * p.seg, p.src + ea, (uint32_t)bytes, bar);
        mine = (uint32_t)bytes;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) mbar_expect_tx(bar, mine);
}

// Block layout shared by both kernels: warps 0-6 compute (224 threads), warp 7 streams rows.
// Buffers cycle through full[] (TMA transaction barrier) and empty[] (one arrive per compute warp),
// so compute warps never wait for each other, only for data.
constexpr int kComputeWarps = 7;
constexpr int kComputeThreads = kComputeWarps * 32;

// ------------------------------------------------------------------------------------ forward
// tile = TR (even) output rows of one image; thread = two vertically adjacent output pixels
// (rows 2*oyp, 2*oyp+1; they share one input row), all 16 channels: 864 FFMAs per 45 LDS.
template <int SLOT, bool FUSE>
__global__ void __launch_bounds__(kThinThreads) thin_fwd_kernel(const ThinArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);    // [nbuf <= 4]
    uint64_t* empty = full + 4;
    uint8_t* raw0 = smem + 128;
    const unsigned nbuf = (unsigned)p.nbuf;   // row ring depth: the streamer runs nbuf - 1 tiles ahead
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw_bytes = (uint32_t)(kCin * p.seg);
    const int segf = p.seg >> 2;
    if (tid == 0) {
        for (unsigned i = 0; i < nbuf; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kComputeWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned my_tiles = (p.tiles > blockIdx.x) ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileWalk tw(blockIdx.x, gridDim.x, p.SCI);

    if (warp == kComputeWarps) {
        // ---------------------------------------------------------------- row streamer
        unsigned sb = 0, sph = 0;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int oy0 = tw.gi * p.TR;
            const int nrows = min(p.TR, p.OH - oy0);
            if (ti >= nbuf) mbar_wait(&empty[sb], sph ^ 1);
            stream_rows<kCin>(p, lane, raw0 + (size_t)sb * raw_bytes, &full[sb], tw.b, p.H, p.W, oy0 * kS,
                              oy0 * kS + (nrows - 1) * kS + kK);
            tw.next();
            if (++sb == nbuf) { sb = 0; sph ^= 1; }
        }
        return;
    }
    // thread -> (row pair, column); the fused variant pads the row pitch to an even number of threads so
    // that horizontally adjacent output pixels sit in adjacent lanes of one warp (2x2 pooling by shuffle)
    const int oyp_raw = tid / p.OWp, ox_raw = tid - oyp_raw * p.OWp;
    const bool in_tile = oyp_raw < (p.TR >> 1) && ox_raw < p.OW;
    const int oyp = in_tile ? oyp_raw : 0;
    const int ox = in_tile ? ox_raw : 0;
    const int pix = (2 * oyp * kS) * p.W + ox * kS;
    const long long plane = (long long)p.H * p.W;
    const size_t oplane = (size_t)p.OH * p.OW;
    const ThinConst& c = c_thin[SLOT];
    unsigned cb = 0, cph = 0;
    for (unsigned ti = 0; ti < my_tiles; ++ti) {
        const int oy0 = tw.gi * p.TR;
        const int nrows = min(p.TR, p.OH - oy0);
        const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)cb * raw_bytes);
        const long long e00 = ((long long)tw.b * kCin * p.H + oy0 * kS) * p.W;
        const bool v0 = in_tile && 2 * oyp < nrows, v1 = in_tile && 2 * oyp + 1 < nrows;
        mbar_wait(&full[cb], cph);
        float2 a0[kCout / 2], a1[kCout / 2];   // channel pairs: one FFMA2 per two multiply-adds
#pragma unroll
        for (int co = 0; co < kCout / 2; ++co) a0[co] = a1[co] = make_float2(0.f, 0.f);
        if (v0) {
#pragma unroll
            for (int ci = 0; ci < kCin; ++ci) {
                const float* r = raw + ci * segf + (int)((e00 + ci * plane) & 3) + pix;
                float v[5][kK];   // 5 input rows x 3 columns
#pragma unroll
                for (int ry = 0; ry < 5; ++ry)
#pragma unroll
                    for (int kx = 0; kx < kK; ++kx) v[ry][kx] = r[ry * p.W + kx];   // rows 3,4 of an odd last pair: staged garbage, never stored
#pragma unroll
                for (int ky = 0; ky < kK; ++ky)
#pragma unroll
                    for (int kx = 0; kx < kK; ++kx) {
                        const float2* w2 = reinterpret_cast<const float2*>(&c.wt[((ci * kK + ky) * kK + kx) * kCout]);
                        const float2 x0 = make_float2(v[ky][kx], v[ky][kx]), x1 = make_float2(v[ky + kS][kx], v[ky + kS][kx]);
#pragma unroll
                        for (int co = 0; co < kCout / 2; ++co) {
                            a0[co] = ffma2(x0, w2[co], a0[co]);
                            a1[co] = ffma2(x1, w2[co], a1[co]);
                        }
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[cb]);   // staged rows are in registers: hand the buffer back
        {
            const unsigned nb = cb + 1 == nbuf ? 0 : cb + 1;
            if (nb == 0) cph ^= 1;
            cb = nb;
        }
        if constexpr (!FUSE) {
            if (v0) {
                float* o = p.dst + (size_t)tw.b * kCout * oplane + (size_t)(oy0 + 2 * oyp) * p.OW + ox;
#pragma unroll
                for (int co = 0; co < kCout; ++co) {
                    const float r0 = (co & 1) ? a0[co >> 1].y : a0[co >> 1].x, r1 = (co & 1) ? a1[co >> 1].y : a1[co >> 1].x;
                    o[(size_t)co * oplane] = r0 + c.b[co];
                    if (v1) o[(size_t)co * oplane + p.OW] = r1 + c.b[co];
                }
            }
        } else {
            // conv output, ReLU output (relu.cpp:25) and the 2x2/2 max-pool of the ReLU output with its
            // arg-max index (pool2d.cpp:53-87: scan order (0,0),(0,1),(1,0),(1,1), strict '<').  The thread
            // holds rows (oy, oy+1) of column ox; lane ^ 1 holds the other column of the window.  The even
            // lane pools channels 0-7, the odd lane channels 8-15: 16 shuffles per thread.
            // running pointers (one add per channel) instead of per-store 64-bit index arithmetic
            const size_t obase = (size_t)tw.b * kCout * oplane + (size_t)(oy0 + 2 * oyp) * p.OW + ox;
            float* oc = p.dst + obase;
            float* orl = p.dst_relu + obase;
            const int OW = p.OW;
            float q0[kCout], q1[kCout];
#pragma unroll
            for (int co = 0; co < kCout; ++co) {
                const float r0 = ((co & 1) ? a0[co >> 1].y : a0[co >> 1].x) + c.b[co];
                const float r1 = ((co & 1) ? a1[co >> 1].y : a1[co >> 1].x) + c.b[co];
                q0[co] = r0 >= 0.f ? r0 : 0.f;
                q1[co] = r1 >= 0.f ? r1 : 0.f;
                if (v0) {
                    oc[0] = r0;
                    orl[0] = q0[co];
                }
                if (v1) {
                    oc[OW] = r1;
                    orl[OW] = q1[co];
                }
                oc += oplane;
                orl += oplane;
            }
            const bool odd = ox_raw & 1;
            const int pc = ox >> 1, pr = (oy0 + 2 * oyp) >> 1;
            const bool pool_ok = v1 && pc < p.POW && pr < p.POH;     // both rows and both columns of the window exist
            const size_t pplane = (size_t)p.POH * p.POW;
            const size_t pbase = ((size_t)tw.b * kCout + (odd ? 8 : 0)) * pplane + (size_t)pr * p.POW + pc;
            float* op = p.dst_pool + pbase;
            int32_t* om = p.mask ? p.mask + pbase : nullptr;
            int mbase = (odd ? 8 : 0) * (int)oplane + (oy0 + 2 * oyp) * OW + 2 * pc;   // channel plane + window origin
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                // give the partner what it pools, keep what this lane pools
                const float give0 = odd ? q0[c8] : q0[c8 + 8], give1 = odd ? q1[c8] : q1[c8 + 8];
                const float keep0 = odd ? q0[c8 + 8] : q0[c8], keep1 = odd ? q1[c8 + 8] : q1[c8];
                const float got0 = __shfl_xor_sync(0xffffffffu, give0, 1), got1 = __shfl_xor_sync(0xffffffffu, give1, 1);
                const float v00 = odd ? got0 : keep0, v01 = odd ? keep0 : got0;
                const float v10 = odd ? got1 : keep1, v11 = odd ? keep1 : got1;
                float mv = v00;
                int mi = 0;
                if (mv < v01) { mv = v01; mi = 1; }
                if (mv < v10) { mv = v10; mi = OW; }
                if (mv < v11) { mv = v11; mi = OW + 1; }
                if (pool_ok) {
                    *op = mv;
                    if (om) *om = mbase + mi;
                }
                op += pplane;
                if (om) om += pplane;
                mbase += (int)oplane;
            }
        }
        tw.next();
    }
}

// ----------------------------------------------------------------------------- input gradient
// tile = TR (even) rows of 2x2 input patches of one image; thread = two vertically adjacent patches
// (py, py+1), 3 channels x 4 cells each: 864 FFMAs per 96 LDS.
// dx[ci][2py+pr][2px+pc] = sum_co sum_{ky%2==pr, kx%2==pc} w[co][ci][ky][kx] * delta[co][py-ky/2][px-kx/2]
template <int SLOT>
__global__ void __launch_bounds__(kThinThreads) thin_dgrad_kernel(const ThinArgs p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 2;
    uint8_t* raw0 = smem + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw_bytes = (uint32_t)(kCout * p.seg);
    const int segf = p.seg >> 2;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kComputeWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned my_tiles = (p.tiles > blockIdx.x) ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileWalk tw(blockIdx.x, gridDim.x, p.SCI);

    if (warp == kComputeWarps) {
        // ---------------------------------------------------------------- row streamer
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int py0 = tw.gi * p.TR;
            const int nrows = min(p.TR, p.GH - py0);
            if (ti >= 2) mbar_wait(&empty[ti & 1], ((ti >> 1) - 1) & 1);
            stream_rows<kCout>(p, lane, raw0 + (size_t)(ti & 1) * raw_bytes, &full[ti & 1], tw.b, p.OH, p.OW,
                               max(py0 - 1, 0), min(py0 + nrows, p.OH));
            tw.next();
        }
        return;
    }
    const bool in_tile = tid < (p.TR >> 1) * p.GW;
    const int pyp = in_tile ? tid / p.GW : 0;
    const int px = in_tile ? tid - pyp * p.GW : 0;
    const bool x0 = px < p.OW, x1 = px >= 1 && px - 1 < p.OW;     // delta columns px, px-1 exist
    const long long plane = (long long)p.OH * p.OW;
    const size_t iplane = (size_t)p.H * p.W;
    const bool vec2 = (p.W & 1) == 0;
    const ThinConst& c = c_thin[SLOT];
    for (unsigned ti = 0; ti < my_tiles; ++ti) {
        const int py0 = tw.gi * p.TR;
        const int nrows = min(p.TR, p.GH - py0);
        const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)(ti & 1) * raw_bytes);
        const int r0 = max(py0 - 1, 0);                       // first staged delta row
        const long long e00 = ((long long)tw.b * kCout * p.OH + r0) * p.OW;
        const int py = py0 + 2 * pyp;                         // patches py (a) and py + 1 (b)
        const bool va = in_tile && 2 * pyp < nrows, vb = in_tile && 2 * pyp + 1 < nrows;
        // delta rows py-1, py, py+1 (row index 0, 1, 2) x columns px, px-1
        const bool ym = py >= 1 && py - 1 < p.OH, y0 = py < p.OH, yp = py + 1 < p.OH;
        mbar_wait(&full[ti & 1], (ti >> 1) & 1);
        float2 acc[kCin][2][2];   // .x = patch a, .y = patch b: both take the same filter tap (one FFMA2)
#pragma unroll
        for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[ci][q >> 1][q & 1] = make_float2(0.f, 0.f);
        if (va) {
            // staged word of delta[co][py + j][px - dx] = base + co*segf + shift(co) + j*OW - dx
            const int base = (py - r0) * p.OW + px;
#pragma unroll
            for (int co = 0; co < kCout; ++co) {
                const float* r = raw + co * segf + (int)((e00 + co * plane) & 3) + base;
                float d[3][2];
                d[0][0] = (ym && x0) ? r[-p.OW] : 0.f;
                d[0][1] = (ym && x1) ? r[-p.OW - 1] : 0.f;
                d[1][0] = (y0 && x0) ? r[0] : 0.f;
                d[1][1] = (y0 && x1) ? r[-1] : 0.f;
                d[2][0] = (yp && x0) ? r[p.OW] : 0.f;
                d[2][1] = (yp && x1) ? r[p.OW - 1] : 0.f;
#pragma unroll
                for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
                    for (int ky = 0; ky < kK; ++ky)
#pragma unroll
                        for (int kx = 0; kx < kK; ++kx) {
                            const float wv = c.w[((co * kCin + ci) * kK + ky) * kK + kx];
                            acc[ci][ky & 1][kx & 1] = ffma2(make_float2(d[1 - (ky >> 1)][kx >> 1], d[2 - (ky >> 1)][kx >> 1]),
                                                            make_float2(wv, wv), acc[ci][ky & 1][kx & 1]);
                        }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[ti & 1]);
        if (va) {
            float* o = p.dst + (size_t)tw.b * kCin * iplane + (size_t)(2 * py) * p.W + 2 * px;
#pragma unroll
            for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {          // input rows 2py .. 2py+3 (patch a: 0,1 ; patch b: 2,3)
                    if ((pr < 2 || vb) && 2 * py + pr < p.H) {
                        float* q = o + (size_t)ci * iplane + (size_t)pr * p.W;
                        const float e0 = pr < 2 ? acc[ci][pr & 1][0].x : acc[ci][pr & 1][0].y;
                        const float e1 = pr < 2 ? acc[ci][pr & 1][1].x : acc[ci][pr & 1][1].y;
                        if (vec2) {   // W even: 2*px + 1 < W and the pair is 8-byte aligned
                            *reinterpret_cast<float2*>(q) = make_float2(e0, e1);
                        } else {
                            q[0] = e0;
                            if (2 * px + 1 < p.W) q[1] = e1;
                        }
                    }
                }
        }
        tw.next();
    }
}

// ---------------------------------------------------------------------------- weight gradient
// dw[co][ci][ky][kx] = scale * sum_{b,oy,ox} x[b][ci][2oy+ky][2ox+kx] * delta[b][co][oy][ox],
// db[co] = scale * sum delta (conv2d.cpp:108-159).  A 27 x 16 outer product per pixel is far too
// thin for a 128-row MMA tile (the tensor-core kernel spends its time building operand tiles), so
// it runs on the FP32 pipe: compute warp (ci, co-half) keeps a 9 tap x 8 channel accumulator tile in
// registers per lane, lanes walk the pixels of the staged rows (17 LDS per 72 FFMA), and the lane /
// CTA partial sums are reduced in a fixed order afterwards (deterministic).
constexpr int kWgWarps = 6;                 // (ci, co half)
constexpr int kWgThreads = (kWgWarps + 1) * 32;
constexpr int kWgRows = kCin * kK * kK + 1; // 27 taps + bias row

struct ThinWgrad {
    const float* x;
    const float* delta;
    float* partial;       // [grid][28][16]
    int B, H, W, OH, OW;
    int TR, SCI;          // output rows per tile, tiles per image
    int xseg, dseg;       // bytes per staged channel segment
    unsigned tiles;
    long long x_bytes16, d_bytes16;
};

__device__ __forceinline__ uint32_t stream_seg(const float* src, long long src_bytes16, long long e0, long long nfl,
                                               uint8_t* dst, uint64_t* bar) {
    const long long ea = e0 & ~3ll;
    long long bytes = (((e0 - ea) + nfl) * 4 + 15) & ~15ll;
    if (ea * 4 + bytes > src_bytes16) bytes = src_bytes16 - ea * 4;
    tma_bulk_g2s(dst, src + ea, (uint32_t)bytes, bar);
    return (uint32_t)bytes;
}

__global__ void __launch_bounds__(kWgThreads, 2) thin_wgrad_kernel(const ThinWgrad p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 2;
    uint8_t* raw0 = smem + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw_bytes = (uint32_t)(kCin * p.xseg + kCout * p.dseg);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kWgWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned my_tiles = (p.tiles > blockIdx.x) ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileWalk tw(blockIdx.x, gridDim.x, p.SCI);
    const long long dplane = (long long)p.OH * p.OW;

    if (warp == kWgWarps) {
        // ---------------------------------------------------------------- row streamer: lanes 0-2 x, 3-18 delta
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int oy0 = tw.gi * p.TR;
            const int nrows = min(p.TR, p.OH - oy0);
            uint8_t* raw = raw0 + (size_t)(ti & 1) * raw_bytes;
            if (ti >= 2) mbar_wait(&empty[ti & 1], ((ti >> 1) - 1) & 1);
            uint32_t mine = 0;
            if (lane < kCin) {
                mine = stream_seg(p.x, p.x_bytes16, ((long long)(tw.b * kCin + lane) * p.H + oy0 * kS) * p.W,
                                  (long long)((nrows - 1) * kS + kK) * p.W, raw + (size_t)lane * p.xseg, &full[ti & 1]);
            } else if (lane < kCin + kCout) {
                const int co = lane - kCin;
                mine = stream_seg(p.delta, p.d_bytes16, ((long long)(tw.b * kCout + co) * p.OH + oy0) * p.OW,
                                  (long long)nrows * p.OW, raw + (size_t)kCin * p.xseg + (size_t)co * p.dseg,
                                  &full[ti & 1]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            if (lane == 0) mbar_expect_tx(&full[ti & 1], mine);
            tw.next();
        }
        return;
    }
    const int ci = warp % kCin, co0 = (warp / kCin) * 8;
    float2 acc[kK * kK][4];   // channel pairs: one FFMA2 per two multiply-adds
    float bsum[8];
#pragma unroll
    for (int t = 0; t < kK * kK; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
    const int xsegf = p.xseg >> 2, dsegf = p.dseg >> 2;
    const int W = p.W, wrap = kS * p.W - kS * p.OW;
    for (unsigned ti = 0; ti < my_tiles; ++ti) {
        const int oy0 = tw.gi * p.TR;
        const int nrows = min(p.TR, p.OH - oy0);
        const int npx = nrows * p.OW;
        const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)(ti & 1) * raw_bytes);
        const long long ex = ((long long)(tw.b * kCin + ci) * p.H + oy0 * kS) * p.W;
        const float* rx = raw + ci * xsegf + (int)(ex & 3);
        const long long ed = ((long long)(tw.b * kCout + co0) * p.OH + oy0) * p.OW;
        // running shared-memory pointers: 8 delta rows (+32 pixels per iteration) and the x window
        const float* dp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            dp[j] = raw + kCin * xsegf + (co0 + j) * dsegf + (int)((ed + j * dplane) & 3) + lane;
        mbar_wait(&full[ti & 1], (ti >> 1) & 1);
        int oyl = 0, ox = lane;
        while (ox >= p.OW) { ox -= p.OW; ++oyl; }
        const float* r0 = rx + (oyl * kS) * W + ox * kS;
        // x window of a lane: columns 2*ox .. 2*ox+2 of three rows.  Lanes are 2 floats apart, so scalar
        // loads are 2-way bank conflicts; when the window start is 8-byte aligned (even row pitch and even
        // staging shift: the case for 224-wide images) the first two columns come as one conflict-free
        // 64-bit load -- 12 instead of 18 shared-memory wavefronts per pixel step
        auto pixel_loop = [&](auto vec_tag, auto bias_tag) {
            constexpr bool VEC = decltype(vec_tag)::value;
            constexpr bool BIAS = decltype(bias_tag)::value;   // only the ci == 0 warps sum delta for db
            for (int px = lane; px < npx; px += 32) {
                const float* r1 = r0 + W;
                const float* r2 = r1 + W;
                float xv[kK * kK], dv[8];
                if constexpr (VEC) {
                    const float2 a0 = *reinterpret_cast<const float2*>(r0), a1 = *reinterpret_cast<const float2*>(r1),
                                 a2 = *reinterpret_cast<const float2*>(r2);
                    xv[0] = a0.x; xv[1] = a0.y; xv[2] = r0[2];
                    xv[3] = a1.x; xv[4] = a1.y; xv[5] = r1[2];
                    xv[6] = a2.x; xv[7] = a2.y; xv[8] = r2[2];
                } else {
#pragma unroll
                    for (int kx = 0; kx < kK; ++kx) {
                        xv[kx] = r0[kx];
                        xv[kK + kx] = r1[kx];
                        xv[2 * kK + kx] = r2[kx];
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { dv[j] = *dp[j]; dp[j] += 32; }
#pragma unroll
                for (int t = 0; t < kK * kK; ++t)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        acc[t][j] = ffma2(make_float2(xv[t], xv[t]), make_float2(dv[2 * j], dv[2 * j + 1]), acc[t][j]);
                if constexpr (BIAS) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) bsum[j] += dv[j];
                }
                ox += 32;
                r0 += 32 * kS;
                while (ox >= p.OW) { ox -= p.OW; r0 += wrap; }   // next output row: skip the rest of two input rows
            }
        };
        // warp-uniform dispatch once per tile: no predicated-off instructions inside the loop
        const bool vec = ((ex | W) & 1) == 0;
        if (ci == 0) {
            if (vec) pixel_loop(std::true_type{}, std::true_type{});
            else pixel_loop(std::false_type{}, std::true_type{});
        } else {
            if (vec) pixel_loop(std::true_type{}, std::false_type{});
            else pixel_loop(std::false_type{}, std::false_type{});
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[ti & 1]);
        tw.next();
    }
    // lane partials -> one value per (tap, channel) of this CTA (fixed butterfly order)
    float* out = p.partial + (size_t)blockIdx.x * (kWgRows * kCout);
#pragma unroll
    for (int t = 0; t < kK * kK; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = warp_sum((j & 1) ? acc[t][j >> 1].y : acc[t][j >> 1].x);
            if (lane == 0) out[(ci * kK * kK + t) * kCout + co0 + j] = v;
        }
    if (ci == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = warp_sum(bsum[j]);
            if (lane == 0) out[(kWgRows - 1) * kCout + co0 + j] = v;
        }
    }
}

// dw / db = scale * sum over CTA partials: block = one (tap row), 16 channels x 16 split lanes
__global__ void __launch_bounds__(256) thin_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                                float* __restrict__ db, int splits, float scale) {
    const int kidx = blockIdx.x, co = threadIdx.x & 15, sl = threadIdx.x >> 4;
    float s = 0.f;
    for (int sp = sl; sp < splits; sp += 16) s += partial[((size_t)sp * kWgRows + kidx) * kCout + co];
    __shared__ float red[16][17];
    red[sl][co] = s;
    __syncthreads();
    if (sl == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) t += red[i][co];
        t *= scale;
        if (kidx == kWgRows - 1) db[co] = t;
        else dw[co * (kWgRows - 1) + kidx] = t;
    }
}

// =============================================================================== lazy head (training)
// The head of the reference net is Conv2D(3->16, 3x3, s2) -> ReLU -> MaxPool 2x2/2 (alexnet.cpp:12-16).
// A train step only needs three things from it: the pooled activations (the next conv's input), which
// window cell won each pool window, and whether the winner was positive (the ReLU gate of the backward
// pass, relu.cpp:39 keyed on the ReLU output = pooled value at the arg-max).  The two kernels below keep
// exactly that -- 254 -> 87 MB written per B=256 forward, 0.55 GB of dense delta traffic gone from the
// backward pass -- and the engine materialises conv / ReLU / pool outputs, the int32 mask and the image
// gradient on demand from the saved input and filters (net.cu, Layer::get_output semantics).
//
//   head_fwd_kernel   thread = one pool window = 2x2 conv outputs, two passes of 8 channels; the conv
//                     arithmetic is thin_fwd_kernel's (same FMA chain, bit-identical values); writes the
//                     packed bf16 pieces P(pool) of conv_s2.cu and one code byte per (window, channel):
//                     bits 0-1 = arg-max cell in the reference's scan order (pool2d.cpp:67-75), bit 2 = value > 0
//   head_wgrad_kernel the dense delta of the conv output is 75 % exact zeros (one survivor per window):
//                     lane = (channel, window) reads the 27 inputs under its survivor (data-dependent
//                     shared-memory window, conflict-free by construction) -- a quarter of the dense FMAs
struct HeadFwd {
    const float* x;       // [B][3][H][W], W % 4 == 0, 16-byte aligned
    uint4* px;            // P(pool output) of the following s2 conv (three bf16 pieces), or null
    float* pool;          // fp32 pool output [B][16][POH][POW], or null
    uint2* m8;            // [B][POH][POW][16] code bytes
    int B, H, W, POH, POW;
    int TRP, SCI;         // pool rows per tile, tiles per image
    int nbuf, seg;        // staged-row ring depth; bytes per staged channel segment
    unsigned tiles;
    int nx_HP, nx_PP;     // plane geometry of the following conv's input
    long long nx_RUNX;
};

// One pool window x 8 channels (half HALF of the 16).  One input row (5 values) at a time: value (ry, cx)
// feeds window cell (i, j) through tap (ky, kx) = (ry - 2i, cx - 2j); walking ci, ry, cx upwards keeps every
// accumulator's FMA chain in the reference's (ci, ky, kx) order (conv2d.cpp:78-86) -- the bits of thin_fwd_kernel.
template <int SLOT, int HALF>
__device__ __forceinline__ void head_window(const float* __restrict__ raw, int segf, int W, bool valid, float (&pv)[8],
                                            uint32_t& code_lo, uint32_t& code_hi) {
    const ThinConst& c = c_thin[SLOT];
    float2 acc[4][4];   // [cell i*2+j][channel pair]
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int cp = 0; cp < 4; ++cp) acc[q][cp] = make_float2(0.f, 0.f);
    if (valid) {
#pragma unroll
        for (int ci = 0; ci < kCin; ++ci) {
#pragma unroll
            for (int ry = 0; ry < 5; ++ry) {
                const float* r = raw + ci * segf + ry * W;
                const float4 a = *reinterpret_cast<const float4*>(r);
                const float v[5] = {a.x, a.y, a.z, a.w, r[4]};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int ky = ry - 2 * i;
                    if (ky < 0 || ky >= kK) continue;
#pragma unroll
                    for (int kx = 0; kx < kK; ++kx) {
                        const float2* w2 = reinterpret_cast<const float2*>(&c.wt[((ci * kK + ky) * kK + kx) * kCout + 8 * HALF]);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float xv = v[2 * j + kx];
#pragma unroll
                            for (int cp = 0; cp < 4; ++cp)
                                acc[i * 2 + j][cp] = ffma2(make_float2(xv, xv), w2[cp], acc[i * 2 + j][cp]);
                        }
                    }
                }
            }
        }
    }
    // bias, ReLU (relu.cpp:25), 2x2 max with the reference's scan order and strict '<' (pool2d.cpp:67-75)
    code_lo = code_hi = 0;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
        const float bias = c.b[8 * HALF + c8];
        float qv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float r = ((c8 & 1) ? acc[q][c8 >> 1].y : acc[q][c8 >> 1].x) + bias;
            qv[q] = r >= 0.f ? r : 0.f;
        }
        float mv = qv[0];
        uint32_t mi = 0;
        if (mv < qv[1]) { mv = qv[1]; mi = 1; }
        if (mv < qv[2]) { mv = qv[2]; mi = 2; }
        if (mv < qv[3]) { mv = qv[3]; mi = 3; }
        pv[c8] = mv;
        const uint32_t code = mi | (mv > 0.f ? 4u : 0u);
        if (c8 < 4) code_lo |= code << (8 * c8);
        else code_hi |= code << (8 * (c8 - 4));
    }
}

// Block: warps 0-6 = channels 0-7, warps 7-13 = channels 8-15 of the same tile (the same thread -> window
// map in both groups), warp 14 streams rows.  CW > 0: image width known at compile time.
constexpr int kHeadGroupWarps = 7;
constexpr int kHeadThreads = (2 * kHeadGroupWarps + 1) * 32;

template <int SLOT, int CW>
__global__ void __launch_bounds__(kHeadThreads, 2) head_fwd_kernel(const HeadFwd p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);    // [nbuf <= 4]
    uint64_t* empty = full + 4;
    uint8_t* raw0 = smem + 128;
    const unsigned nbuf = (unsigned)p.nbuf;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t raw_bytes = (uint32_t)(kCin * p.seg);
    const int segf = p.seg >> 2;
    if (tid == 0) {
        for (unsigned i = 0; i < nbuf; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 2 * kHeadGroupWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned my_tiles = (p.tiles > blockIdx.x) ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileWalk tw(blockIdx.x, gridDim.x, p.SCI);
    const int W = CW ? CW : p.W;

    if (warp == 2 * kHeadGroupWarps) {
        // ---------------------------------------------------------------- row streamer: input rows 4*py0 .. 4*(py0+nrows)
        unsigned sb = 0, sph = 0;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int py0 = tw.gi * p.TRP;
            const int nrows = min(p.TRP, p.POH - py0);
            const uint32_t bytes = (uint32_t)(4 * nrows + 1) * (uint32_t)W * 4u;   // W % 4 == 0: multiple of 16
            if (ti >= nbuf) mbar_wait(&empty[sb], sph ^ 1);
            if (lane < kCin)
                tma_bulk_g2s(raw0 + (size_t)sb * raw_bytes + (size_t)lane * p.seg,
                             p.x + ((size_t)(tw.b * kCin + lane) * p.H + 4 * py0) * W, bytes, &full[sb]);
            if (lane == 0) mbar_expect_tx(&full[sb], kCin * bytes);
            tw.next();
            if (++sb == nbuf) { sb = 0; sph ^= 1; }
            // the rows of the tile nbuf ahead go to L2 now: its copy can only start when a buffer frees up (the
            // slowest compute warp decides), and then it should not also wait for DRAM
            if (ti + nbuf < my_tiles && lane < kCin) {
                TileWalk nx = tw;
                for (unsigned k = 1; k < nbuf; ++k) nx.next();
                const int qy0 = nx.gi * p.TRP;
                const int qrows = min(p.TRP, p.POH - qy0);
                tma_prefetch_l2(p.x + ((size_t)(nx.b * kCin + lane) * p.H + 4 * qy0) * W, (uint32_t)(4 * qrows + 1) * (uint32_t)W * 4u);
            }
        }
        return;
    }
    const int half = warp >= kHeadGroupWarps ? 1 : 0;
    const int gtid = tid - half * (kHeadGroupWarps * 32);
    const int prow_raw = gtid / p.POW, ppx_raw = gtid - prow_raw * p.POW;
    const bool in_tile = prow_raw < p.TRP;
    const int prow = in_tile ? prow_raw : 0, ppx = in_tile ? ppx_raw : 0;
    const int pix = 4 * prow * W + 4 * ppx;
    const size_t pplane = (size_t)p.POH * p.POW;
    unsigned cb = 0, cph = 0;
    for (unsigned ti = 0; ti < my_tiles; ++ti) {
        const int py0 = tw.gi * p.TRP;
        const int nrows = min(p.TRP, p.POH - py0);
        const bool valid = in_tile && prow < nrows;
        const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)cb * raw_bytes) + pix;
        const int py = py0 + prow;
        float pv[8];
        uint32_t code_lo, code_hi;
        mbar_wait(&full[cb], cph);
        if (half == 0) head_window<SLOT, 0>(raw, segf, W, valid, pv, code_lo, code_hi);
        else head_window<SLOT, 1>(raw, segf, W, valid, pv, code_lo, code_hi);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[cb]);   // staged rows consumed: hand the buffer back before the stores
        {
            const unsigned nb = cb + 1 == nbuf ? 0 : cb + 1;
            if (nb == 0) cph ^= 1;
            cb = nb;
        }
        if (valid) {
            const size_t widx = ((size_t)tw.b * p.POH + py) * p.POW + ppx;
            p.m8[widx * 2 + half] = make_uint2(code_lo, code_hi);
            if (p.px) {
                uint4 hi, mid, lo;
                split8x3(pv, hi, mid, lo);
                const size_t gpos = (size_t)tw.b * p.nx_PP + (size_t)(py >> 1) * p.nx_HP + (ppx >> 1);
                const int qn = (py & 1) * 2 + (ppx & 1);
                p.px[((size_t)(0 * 2 + half) * 4 + qn) * p.nx_RUNX + gpos] = hi;
                p.px[((size_t)(1 * 2 + half) * 4 + qn) * p.nx_RUNX + gpos] = mid;
                p.px[((size_t)(2 * 2 + half) * 4 + qn) * p.nx_RUNX + gpos] = lo;
            }
            if (p.pool) {
                float* o = p.pool + ((size_t)tw.b * kCout + 8 * half) * pplane + (size_t)py * p.POW + ppx;
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) o[(size_t)c8 * pplane] = pv[c8];
            }
        }
        tw.next();
    }
}

// Sparse weight gradient of the head: dw[co][ci][ky][kx] = scale * sum over pool windows of
// delta_pool * [pooled value > 0] * x[ci][2*(2py+dy)+ky][2*(2px+dx)+kx], (dy, dx) = the window's arg-max
// cell -- the composition of pool2d.cpp:92-109, relu.cpp:30-44 and conv2d.cpp:108-159 without the dense
// intermediate.  Lane = (channel co = lane & 15, window slot = lane >> 4); the 16 lanes of a window read
// at most four distinct 8-byte-aligned addresses per load (broadcast).  The input rows of a tile arrive
// by ONE tensor-map TMA copy per channel whose box is 4 floats wider than the image: the zero-filled
// overhang makes the shared-memory row pitch W + 4, so the two candidate rows and the two windows of a warp
// fall into distinct banks.  delta arrives channel-last ([B][POH][POW][16], written that way by the
// following conv's input-gradient epilogue): one contiguous copy per tile, conflict-free lane reads.
constexpr int kHwWarps = 8;
constexpr int kHwThreads = (kHwWarps + 1) * 32;

struct HeadWgrad {
    const float* dpool;   // [B][POH][POW][16]
    const uint8_t* m8;    // [B][POH][POW][16]
    float* partial;       // [grid][28][16]
    int B, H, W, POH, POW;
    int TRP, SCI;
    int xseg;             // bytes per staged input channel (multiple of 128)
    unsigned tiles;
};

// CW > 0: image width known at compile time (row pitch and channel segment become immediates)
template <int CW, int CTRP>
__global__ void __launch_bounds__(kHwThreads) head_wgrad_kernel(const __grid_constant__ CUtensorMap xmap, const HeadWgrad p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 2;
    float* red = reinterpret_cast<float*>(smem + 128);                       // [kHwWarps][28][16] (after the loop)
    uint8_t* raw0 = smem + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int TRP = CTRP ? CTRP : p.TRP;
    const int xp = (CW ? CW : p.W) + 4;
    const int xseg = CW ? ((4 * CTRP + 1) * (CW + 4) * 4 + 127) / 128 * 128 : p.xseg;
    const int dbytes = TRP * p.POW * 64, mbytes = TRP * p.POW * 16;
    const uint32_t raw_bytes = (uint32_t)((kCin * xseg + dbytes + mbytes + 127) / 128 * 128);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], kHwWarps);
        }
        mbar_fence_init();
        tma_prefetch_desc(&xmap);
    }
    __syncthreads();
    const unsigned my_tiles = (p.tiles > blockIdx.x) ? (p.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    TileWalk tw(blockIdx.x, gridDim.x, p.SCI);

    if (warp == kHwWarps) {
        // ---------------------------------------------------------------- streamer: 3 tensor copies (x), delta, codes
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int py0 = tw.gi * TRP;
            const int nrows = min(TRP, p.POH - py0);
            uint8_t* raw = raw0 + (size_t)(ti & 1) * raw_bytes;
            if (ti >= 2) mbar_wait(&empty[ti & 1], ((ti >> 1) - 1) & 1);
            // the x box always has 4*TRP+1 rows (a short last tile reads rows it does not use; rows past the image
            // are zero-filled), so the transaction size is a constant
            const uint32_t xbytes = (uint32_t)((4 * TRP + 1) * xp * 4);
            const uint32_t db = (uint32_t)nrows * (uint32_t)p.POW * 64u, mb = (uint32_t)nrows * (uint32_t)p.POW * 16u;
            if (lane == 0) mbar_expect_tx(&full[ti & 1], kCin * xbytes + db + mb);
            __syncwarp();
            const size_t w0 = ((size_t)tw.b * p.POH + py0) * p.POW;
            if (lane < kCin)
                tma_tensor3d_g2s(raw + (size_t)lane * xseg, &xmap, 0, 4 * py0, tw.b * kCin + lane, &full[ti & 1]);
            else if (lane == kCin)
                tma_bulk_g2s(raw + (size_t)kCin * xseg, p.dpool + w0 * 16, db, &full[ti & 1]);
            else if (lane == kCin + 1)
                tma_bulk_g2s(raw + (size_t)kCin * xseg + dbytes, p.m8 + w0 * 16, mb, &full[ti & 1]);
            tw.next();
        }
    } else {
        const int co = lane & 15, slot = lane >> 4;
        float2 acc01[kCin][kK];   // taps kx = 0, 1
        float acc2[kCin][kK];     // tap kx = 2
        float bsum = 0.f;
#pragma unroll
        for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
            for (int ky = 0; ky < kK; ++ky) { acc01[ci][ky] = make_float2(0.f, 0.f); acc2[ci][ky] = 0.f; }
        const int xsegf = xseg >> 2;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int py0 = tw.gi * TRP;
            const int nrows = min(TRP, p.POH - py0);
            const int npix = nrows * p.POW;
            const float* xs = reinterpret_cast<const float*>(raw0 + (size_t)(ti & 1) * raw_bytes);
            const float* ds = xs + kCin * xsegf + co;
            const uint8_t* ms = reinterpret_cast<const uint8_t*>(xs + kCin * xsegf) + dbytes + co;
            mbar_wait(&full[ti & 1], (ti >> 1) & 1);
            int pi = 2 * warp + slot, prow = 0, ppx = pi;
            while (ppx >= p.POW) { ppx -= p.POW; ++prow; }
            // code / delta of the next window are requested one iteration ahead: the arg-max code heads the
            // dependent chain code -> window address -> 18 loads -> FMAs
            uint32_t code_n = ms[(pi < npix ? pi : 0) * 16];
            float d_n = ds[(pi < npix ? pi : 0) * 16];
            for (; pi - slot < npix; pi += 2 * kHwWarps) {
                const bool ok = pi < npix;
                const uint32_t code = code_n;
                const float d = d_n;
                {
                    const int pn = pi + 2 * kHwWarps < npix ? pi + 2 * kHwWarps : 0;
                    code_n = ms[pn * 16];
                    d_n = ds[pn * 16];
                }
                const float dv = (ok && (code & 4u)) ? d : 0.f;
                const float* r = xs + (4 * (ok ? prow : 0) + (int)(code & 2u)) * xp + 4 * (ok ? ppx : 0) + 2 * (int)(code & 1u);
#pragma unroll
                for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
                    for (int ky = 0; ky < kK; ++ky) {
                        const float2 a = *reinterpret_cast<const float2*>(r + ci * xsegf + ky * xp);
                        const float b2 = r[ci * xsegf + ky * xp + 2];
                        acc01[ci][ky] = ffma2(make_float2(dv, dv), a, acc01[ci][ky]);
                        acc2[ci][ky] = fmaf(dv, b2, acc2[ci][ky]);
                    }
                bsum += dv;
                ppx += 2 * kHwWarps;
                while (ppx >= p.POW) { ppx -= p.POW; ++prow; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ti & 1]);
            tw.next();
        }
        // the two window slots of a warp -> lanes 0-15; parked until every warp is done with the staged rows
#pragma unroll
        for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
            for (int ky = 0; ky < kK; ++ky) {
                acc01[ci][ky].x += __shfl_xor_sync(0xffffffffu, acc01[ci][ky].x, 16);
                acc01[ci][ky].y += __shfl_xor_sync(0xffffffffu, acc01[ci][ky].y, 16);
                acc2[ci][ky] += __shfl_xor_sync(0xffffffffu, acc2[ci][ky], 16);
            }
        bsum += __shfl_xor_sync(0xffffffffu, bsum, 16);
        asm volatile("bar.sync 1, %0;" ::"n"(kHwWarps * 32) : "memory");   // compute warps only
        if (slot == 0) {
            float* o = red + (size_t)warp * (kWgRows * kCout) + co;
#pragma unroll
            for (int ci = 0; ci < kCin; ++ci)
#pragma unroll
                for (int ky = 0; ky < kK; ++ky) {
                    o[((ci * kK + ky) * kK + 0) * kCout] = acc01[ci][ky].x;
                    o[((ci * kK + ky) * kK + 1) * kCout] = acc01[ci][ky].y;
                    o[((ci * kK + ky) * kK + 2) * kCout] = acc2[ci][ky];
                }
            o[(kWgRows - 1) * kCout] = bsum;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kHwWarps * 32) : "memory");
        float* out = p.partial + (size_t)blockIdx.x * (kWgRows * kCout);
        for (int i = tid; i < kWgRows * kCout; i += kHwWarps * 32) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kHwWarps; ++w) s += red[(size_t)w * (kWgRows * kCout) + i];
            out[i] = s;
        }
    }
}

std::mutex& slot_mutex() {
    static std::mutex m;
    return m;
}
bool g_slot_used[16][kSlots];

template <class K>
int thin_smem_attr(K kernel, size_t smem) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cnn_cuda_fail(e, "cudaFuncSetAttribute(thin)", __FILE__, __LINE__);
    }
    return CNN_OK;
}

// staged rows may need more than the default 48 KB of dynamic shared memory (once per device)
int thin_attrs(int device) {
    std::lock_guard<std::recursive_mutex> lk(cnn_global_mutex());
    static bool done[16];
    if (device < 0 || device >= 16 || done[device]) return CNN_OK;
    const size_t cap = 128 + 4 * 40 * 1024;
    if (int rc = thin_smem_attr(thin_fwd_kernel<0, false>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<1, false>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<2, false>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<3, false>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<0, true>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<1, true>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<2, true>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_fwd_kernel<3, true>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_dgrad_kernel<0>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_dgrad_kernel<1>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_dgrad_kernel<2>, cap)) return rc;
    if (int rc = thin_smem_attr(thin_dgrad_kernel<3>, cap)) return rc;
    done[device] = true;
    return CNN_OK;
}

// persistent grid = CTAs that are resident at once (shared memory and registers decide)
template <class K>
int resident_ctas(K kernel, size_t smem, int threads = kThinThreads) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n;
}

// filters -> this context's constant bank (written through the symbol's global address; the
// constant caches are coherent at kernel boundaries and everything here is ordered on ctx->stream)
// `save` (optional): a copy of filters + biases as this launch saw them, [432 + 16] floats -- the lazy head
// re-creates its outputs on demand after the SGD step has already changed the parameters
__global__ void thin_upload_kernel(const float* __restrict__ w, const float* __restrict__ bias, ThinConst* c,
                                   float* __restrict__ save) {
    const int i = threadIdx.x;
    if (i < kNW) {
        const float v = w[i];
        const int co = i / (kCin * kK * kK), tap = i % (kCin * kK * kK);
        c->w[i] = v;
        c->wt[tap * kCout + co] = v;
        if (save) save[i] = v;
    }
    if (bias && i < kCout) {
        c->b[i] = bias[i];
        if (save) save[kNW + i] = bias[i];
    }
}

int upload_filters(cnn_ctx* ctx, const float* w, const float* bias, float* save = nullptr) {
    ThinConst* sym = nullptr;
    // the symbol address is per device: resolve it against THIS context's device even if the calling thread last
    // touched another one (one process may hold contexts on several GPUs)
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx->device) CNN_CUDA(cudaSetDevice(ctx->device));
    CNN_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&sym), c_thin));
    CNN_LAUNCH(ctx, thin_upload_kernel, 1, 448, 0, w, bias, sym + ctx->thin_slot, save);
    return CNN_OK;
}

// tile height: as many row PAIRS as fit the 224 compute threads and ~40 KB of staged rows per buffer
bool plan_tiles(ThinArgs& p, int nch, int SW, int rows_per, int rows_extra, int pitch = 0) {
    if (pitch == 0) pitch = p.GW;
    if (pitch > kComputeThreads) return false;
    int TRp = std::min((p.GH + 1) / 2, kComputeThreads / pitch);
    for (; TRp >= 1; --TRp) {
        const size_t seg = (((size_t)(2 * TRp * rows_per + rows_extra) * SW * 4 + 12) + 15) / 16 * 16;
        if (seg * nch <= 40 * 1024) { p.seg = (int)seg; break; }
    }
    if (TRp < 1) return false;
    p.TR = 2 * TRp;
    p.SCI = (p.GH + p.TR - 1) / p.TR;
    return true;
}

#define THIN_DISPATCH(KERNEL, ...)                                                             \
    switch (ctx->thin_slot) {                                                                  \
        case 0: CNN_LAUNCH(ctx, KERNEL<0>, __VA_ARGS__); break;                                \
        case 1: CNN_LAUNCH(ctx, KERNEL<1>, __VA_ARGS__); break;                                \
        case 2: CNN_LAUNCH(ctx, KERNEL<2>, __VA_ARGS__); break;                                \
        default: CNN_LAUNCH(ctx, KERNEL<3>, __VA_ARGS__); break;                               \
    }
#define THIN_DISPATCH2(KERNEL, FLAG, ...)                                                      \
    switch (ctx->thin_slot) {                                                                  \
        case 0: CNN_LAUNCH(ctx, (KERNEL<0, FLAG>), __VA_ARGS__); break;                        \
        case 1: CNN_LAUNCH(ctx, (KERNEL<1, FLAG>), __VA_ARGS__); break;                        \
        case 2: CNN_LAUNCH(ctx, (KERNEL<2, FLAG>), __VA_ARGS__); break;                        \
        default: CNN_LAUNCH(ctx, (KERNEL<3, FLAG>), __VA_ARGS__); break;                       \
    }

}  // namespace

int conv_thin_acquire_slot(int device) {
    std::lock_guard<std::mutex> lk(slot_mutex());
    if (device < 0 || device >= 16) return -1;
    for (int i = 0; i < kSlots; ++i)
        if (!g_slot_used[device][i]) { g_slot_used[device][i] = true; return i; }
    return -1;
}

void conv_thin_release_slot(int device, int slot) {
    std::lock_guard<std::mutex> lk(slot_mutex());
    if (device >= 0 && device < 16 && slot >= 0 && slot < kSlots) g_slot_used[device][slot] = false;
}

bool conv_thin_supported(const cnn_ctx* ctx, int Cin, int H, int W, int Cout, int k, int s) {
    if (ctx->thin_slot < 0 || Cin != kCin || Cout != kCout || k != kK || s != kS) return false;
    if ((W + 1) / 2 > kComputeThreads || H < k || W < k) return false;
    return ((size_t)W * 4 * 5 + 16) * kCin <= 40 * 1024;   // at least one output row pair per tile fits
}

namespace {
int fwd_thin_launch(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, float* y_relu, float* y_pool,
                    int32_t* mask, int B, int H, int W) {
    const bool fuse = y_relu != nullptr;
    ThinArgs p{};
    p.src = x; p.dst = y; p.B = B; p.H = H; p.W = W;
    p.OH = (H - kK) / kS + 1; p.OW = (W - kK) / kS + 1;
    p.GH = p.OH; p.GW = p.OW;
    p.dst_relu = y_relu; p.dst_pool = y_pool; p.mask = mask;
    p.POH = (p.OH - 2) / 2 + 1; p.POW = (p.OW - 2) / 2 + 1;
    p.OWp = fuse ? (p.OW + 1) / 2 * 2 : p.OW;
    CNN_REQUIRE(plan_tiles(p, kCin, W, kS, kK - kS, p.OWp), "conv_thin: image too wide");
    CNN_REQUIRE(((uintptr_t)x & 15) == 0, "conv_thin: x must be 16-byte aligned");
    p.tiles = (unsigned)B * (unsigned)p.SCI;
    p.src_bytes16 = ((long long)B * kCin * H * W * 4 + 15) & ~15ll;
    if (int rc = upload_filters(ctx, w, bias)) return rc;
    // ring depth: the fused head writes 3.3x the bytes it reads, so its row reads queue behind a flood of
    // stores -- it streams two tiles ahead; the plain kernel keeps the double buffer (4 CTAs per SM)
    p.nbuf = fuse ? 3 : 2;
    if (const char* e = getenv("CNN_DBG_THIN_NBUF")) p.nbuf = std::max(2, std::min(4, atoi(e)));
    const size_t smem = 128 + (size_t)p.nbuf * kCin * p.seg;
    if (int rc = thin_attrs(ctx->device)) return rc;
    unsigned grid = (unsigned)(ctx->sm_count * (fuse ? resident_ctas(thin_fwd_kernel<0, true>, smem)
                                                      : resident_ctas(thin_fwd_kernel<0, false>, smem)));
    if (grid > p.tiles) grid = p.tiles;
    if (fuse) { THIN_DISPATCH2(thin_fwd_kernel, true, grid, kThinThreads, smem, p); }
    else { THIN_DISPATCH2(thin_fwd_kernel, false, grid, kThinThreads, smem, p); }
    return CNN_OK;
}
}  // namespace

int conv_fwd_thin(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int H, int W) {
    return fwd_thin_launch(ctx, x, w, bias, y, nullptr, nullptr, nullptr, B, H, W);
}

// conv (3 -> 16, 3x3, stride 2) -> ReLU -> MaxPool 2x2/2 in one pass: all three layers' outputs and the
// pool mask are written, the conv output is not read back (alexnet.cpp:12-16)
bool conv_thin_pool_supported(const cnn_ctx* ctx, int Cin, int H, int W, int Cout, int k, int s, int pk, int pstep) {
    if (!conv_thin_supported(ctx, Cin, H, W, Cout, k, s) || pk != 2 || pstep != 2) return false;
    const int OH = (H - kK) / kS + 1, OW = (W - kK) / kS + 1;
    return OH >= 2 && OW >= 2 && (OW + 1) / 2 * 2 <= kComputeThreads;
}

// Measured on B200 (profiles/r01_launch_list.md): the fused head takes 250 us at B=256 against 83 + 107 us
// for conv + (ReLU+pool) -- its compute warps sit in the full-barrier wait (ncu: 46 % long-scoreboard
// on the mbarrier try_wait), i.e. the one-tile-ahead row streamer no longer covers the longer epilogue.
// The engine therefore keeps the two-kernel head unless CNN_THIN_POOL_FUSE=1; the entry point stays
// (bit-identical results, tests/test_gpu_ops.py) for the next round's deeper row ring.
bool conv_thin_pool_preferred() { return getenv("CNN_THIN_POOL_FUSE") != nullptr; }

int conv_fwd_thin_relu_pool(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, float* y_relu,
                            float* y_pool, int32_t* mask, int B, int H, int W) {
    return fwd_thin_launch(ctx, x, w, bias, y, y_relu, y_pool, mask, B, H, W);
}

int conv_dgrad_thin(cnn_ctx* ctx, const float* w, const float* delta, float* dx, int B, int H, int W) {
    ThinArgs p{};
    p.src = delta; p.dst = dx; p.B = B; p.H = H; p.W = W;
    p.OH = (H - kK) / kS + 1; p.OW = (W - kK) / kS + 1;
    p.GH = (H + 1) / 2; p.GW = (W + 1) / 2;
    CNN_REQUIRE(plan_tiles(p, kCout, p.OW, 1, 1), "conv_thin: image too wide");
    CNN_REQUIRE(((uintptr_t)delta & 15) == 0, "conv_thin: delta must be 16-byte aligned");
    p.tiles = (unsigned)B * (unsigned)p.SCI;
    p.src_bytes16 = ((long long)B * kCout * p.OH * p.OW * 4 + 15) & ~15ll;
    if (int rc = upload_filters(ctx, w, nullptr)) return rc;
    const size_t smem = 128 + 2 * (size_t)kCout * p.seg;
    if (int rc = thin_attrs(ctx->device)) return rc;
    unsigned grid = (unsigned)(ctx->sm_count * resident_ctas(thin_dgrad_kernel<0>, smem));
    if (grid > p.tiles) grid = p.tiles;
    THIN_DISPATCH(thin_dgrad_kernel, grid, kThinThreads, smem, p);
    return CNN_OK;
}

int conv_wgrad_thin(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B, int H, int W,
                    float scale) {
    ThinWgrad p{};
    p.x = x; p.delta = delta; p.B = B; p.H = H; p.W = W;
    p.OH = (H - kK) / kS + 1; p.OW = (W - kK) / kS + 1;
    // rows per tile: ~48 KB of staged rows per buffer (two buffers, two CTAs per SM)
    int TR = std::min(p.OH, 8);
    for (; TR >= 1; --TR) {
        p.xseg = (int)((((size_t)((TR - 1) * kS + kK) * W * 4 + 12) + 15) / 16 * 16);
        p.dseg = (int)((((size_t)TR * p.OW * 4 + 12) + 15) / 16 * 16);
        if ((size_t)kCin * p.xseg + (size_t)kCout * p.dseg <= 52 * 1024) break;
    }
    CNN_REQUIRE(TR >= 1, "conv_thin: image too wide");
    CNN_REQUIRE((((uintptr_t)x | (uintptr_t)delta) & 15) == 0, "conv_thin: operands must be 16-byte aligned");
    p.TR = TR;
    p.SCI = (p.OH + TR - 1) / TR;
    p.tiles = (unsigned)B * (unsigned)p.SCI;
    p.x_bytes16 = ((long long)B * kCin * H * W * 4 + 15) & ~15ll;
    p.d_bytes16 = ((long long)B * kCout * p.OH * p.OW * 4 + 15) & ~15ll;
    const size_t smem = 128 + 2 * ((size_t)kCin * p.xseg + (size_t)kCout * p.dseg);
    static bool attr_done[16];
    std::unique_lock<std::recursive_mutex> alk(cnn_global_mutex());
    if (ctx->device >= 0 && ctx->device < 16 && !attr_done[ctx->device]) {
        CNN_CUDA(cudaFuncSetAttribute(thin_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + 2 * 52 * 1024));
        attr_done[ctx->device] = true;
    }
    alk.unlock();
    int res = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, thin_wgrad_kernel, kWgThreads, smem) != cudaSuccess || res < 1)
        res = 1;
    unsigned grid = (unsigned)(ctx->sm_count * res);
    if (grid > p.tiles) grid = p.tiles;
    float* partial = cnn_scratch(ctx, sizeof(float) * (size_t)grid * kWgRows * kCout + 64);
    CNN_REQUIRE(partial, "scratch allocation failed");
    p.partial = partial;
    CNN_LAUNCH(ctx, thin_wgrad_kernel, grid, kWgThreads, smem, p);
    CNN_LAUNCH(ctx, thin_wgrad_reduce_kernel, kWgRows, 256, 0, partial, dw, db, (int)grid, scale);
    return CNN_OK;
}

// ------------------------------------------------------------------------------- lazy head (host side)
bool conv_head_lazy_supported(const cnn_ctx* ctx, int Cin, int H, int W, int Cout, int k, int s, int pk, int pstep) {
    if (!conv_thin_supported(ctx, Cin, H, W, Cout, k, s) || pk != 2 || pstep != 2 || (W & 3)) return false;
    const int OH = (H - kK) / kS + 1, OW = (W - kK) / kS + 1;
    if (OH < 2 || OW < 2) return false;
    const int POW = (OW - 2) / 2 + 1;
    return POW <= kComputeThreads && W + 4 <= 256 && (size_t)5 * W * 4 * kCin * 2 + 128 <= 100 * 1024;
}

size_t conv_head_m8_bytes(int B, int H, int W) {
    const int OH = (H - kK) / kS + 1, OW = (W - kK) / kS + 1;
    return (size_t)B * ((OH - 2) / 2 + 1) * ((OW - 2) / 2 + 1) * 16;
}

int conv_head_fwd(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* w_save, void* next_px,
                  float* pool, void* m8, int B, int H, int W) {
    HeadFwd p{};
    p.x = x; p.px = static_cast<uint4*>(next_px); p.pool = pool; p.m8 = static_cast<uint2*>(m8);
    p.B = B; p.H = H; p.W = W;
    const int OH = (H - kK) / kS + 1, OW = (W - kK) / kS + 1;
    p.POH = (OH - 2) / 2 + 1; p.POW = (OW - 2) / 2 + 1;
    CNN_REQUIRE((W & 3) == 0 && ((uintptr_t)x & 15) == 0, "conv_head_fwd: rows must be 16-byte aligned");
    CNN_REQUIRE(p.POW <= kComputeThreads, "conv_head_fwd: image too wide");
    // pool rows per tile: fill the 224 compute threads, two buffers of staged rows, two CTAs per SM
    p.nbuf = 2;
    int TRP = std::min(p.POH, kComputeThreads / p.POW);
    if (const char* e = getenv("CNN_HEAD_TRP")) TRP = std::max(1, std::min(TRP, atoi(e)));
    if (const char* e = getenv("CNN_HEAD_NBUF")) p.nbuf = std::max(2, std::min(4, atoi(e)));
    for (; TRP >= 1; --TRP)
        if (128 + (size_t)p.nbuf * kCin * (4 * TRP + 1) * W * 4 <= 112 * 1024) break;
    CNN_REQUIRE(TRP >= 1, "conv_head_fwd: image too wide");
    p.TRP = TRP;
    p.seg = (4 * TRP + 1) * W * 4;
    p.SCI = (p.POH + TRP - 1) / TRP;
    p.tiles = (unsigned)B * (unsigned)p.SCI;
    if (next_px) conv_s2_px_geom(B, kCout, p.POH, p.POW, &p.nx_HP, &p.nx_PP, &p.nx_RUNX);
    if (int rc = upload_filters(ctx, w, bias, w_save)) return rc;
    const size_t smem = 128 + (size_t)p.nbuf * kCin * p.seg;
    {
        static std::mutex m;
        static bool done[16];
        std::lock_guard<std::mutex> lk(m);
        if (ctx->device >= 0 && ctx->device < 16 && !done[ctx->device]) {
            const int cap = 112 * 1024;
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<0, 224>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<1, 224>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<2, 224>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<3, 224>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            CNN_CUDA(cudaFuncSetAttribute(head_fwd_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
            done[ctx->device] = true;
        }
    }
    unsigned grid = (unsigned)(ctx->sm_count * (W == 224 ? resident_ctas(head_fwd_kernel<0, 224>, smem, kHeadThreads)
                                                         : resident_ctas(head_fwd_kernel<0, 0>, smem, kHeadThreads)));
    if (grid > p.tiles) grid = p.tiles;
    if (W == 224) { THIN_DISPATCH2(head_fwd_kernel, 224, grid, kHeadThreads, smem, p); }
    else { THIN_DISPATCH2(head_fwd_kernel, 0, grid, kHeadThreads, smem, p); }
    return CNN_OK;
}

int conv_head_wgrad(cnn_ctx* ctx, const float* x, const float* dpool_nhwc, const void* m8, float* dw, float* db, int B,
                    int H, int W, float scale) {
    HeadWgrad p{};
    p.dpool = dpool_nhwc; p.m8 = static_cast<const uint8_t*>(m8);
    p.B = B; p.H = H; p.W = W;
    const int OH = (H - kK) / kS + 1, OW = (W - kK) / kS + 1;
    p.POH = (OH - 2) / 2 + 1; p.POW = (OW - 2) / 2 + 1;
    CNN_REQUIRE((W & 3) == 0 && W + 4 <= 256 && (((uintptr_t)x | (uintptr_t)dpool_nhwc | (uintptr_t)m8) & 15) == 0,
                "conv_head_wgrad: unsupported width or unaligned operands");
    const bool fixed = W == 224;   // the reference's image size: compile-time pitch, two pool rows per tile
    int TRP = fixed ? 2 : std::min(p.POH, 2);
    size_t raw = 0;
    for (; TRP >= 1; --TRP) {
        p.xseg = ((4 * TRP + 1) * (W + 4) * 4 + 127) / 128 * 128;
        raw = ((size_t)kCin * p.xseg + (size_t)TRP * p.POW * 80 + 127) / 128 * 128;
        if (128 + 2 * raw <= 112 * 1024) break;
    }
    CNN_REQUIRE(TRP >= 1, "conv_head_wgrad: image too wide");
    p.TRP = TRP;
    p.SCI = (p.POH + TRP - 1) / TRP;
    p.tiles = (unsigned)B * (unsigned)p.SCI;
    // x as a 3-D tensor (W, H, B*3); the box is 4 columns wider than the image (zero-filled overhang = row padding)
    CUtensorMap xmap;
    const uint64_t dims[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)B * kCin};
    const uint64_t strides[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
    const uint32_t box[3] = {(uint32_t)W + 4, (uint32_t)(4 * TRP + 1), 1};
    if (int rc = cnn_tmap_encode_3d(&xmap, x, dims, strides, box)) return rc;
    const size_t smem = std::max(128 + 2 * raw, (size_t)128 + sizeof(float) * kHwWarps * kWgRows * kCout);
    {
        static std::mutex m;
        static bool done[16];
        std::lock_guard<std::mutex> lk(m);
        if (ctx->device >= 0 && ctx->device < 16 && !done[ctx->device]) {
            CNN_CUDA(cudaFuncSetAttribute(head_wgrad_kernel<224, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            CNN_CUDA(cudaFuncSetAttribute(head_wgrad_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            done[ctx->device] = true;
        }
    }
    int res = 0;
    cudaError_t oe = fixed ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, head_wgrad_kernel<224, 2>, kHwThreads, smem)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, head_wgrad_kernel<0, 0>, kHwThreads, smem);
    if (oe != cudaSuccess || res < 1) res = 1;
    unsigned grid = (unsigned)(ctx->sm_count * res);
    if (grid > p.tiles) grid = p.tiles;
    float* partial = cnn_scratch(ctx, sizeof(float) * (size_t)grid * kWgRows * kCout + 64);
    CNN_REQUIRE(partial, "scratch allocation failed");
    p.partial = partial;
    if (fixed) { CNN_LAUNCH(ctx, (head_wgrad_kernel<224, 2>), grid, kHwThreads, smem, xmap, p); }
    else { CNN_LAUNCH(ctx, (head_wgrad_kernel<0, 0>), grid, kHwThreads, smem, xmap, p); }
    CNN_LAUNCH(ctx, thin_wgrad_reduce_kernel, kWgRows, 256, 0, partial, dw, db, (int)grid, scale);
    return CNN_OK;
}
