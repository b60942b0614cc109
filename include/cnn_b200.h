/*
 * cnn_b200.h -- C ABI of libcnn_b200.so, the B200 (sm_100a) backend for the
 * hermosayhl/CNN train-step hot path.
 *
 * The reference has no FFI layer; its de-facto operator API is the C++ class set of
 * cpu/include/architectures.h:34-191, data_format.h:10-53 and func.h:8-18 (built as
 * the shared library `cnn_layers` by cpu/xmake.lua:62-77).  Each entry point below is
 * what the body of one of those methods calls in the replacement backend
 * (the .cpp files under cnn_b200/host); the reference member it replaces is cited per function.
 *
 * Conventions
 *  - plain C: opaque handles, raw DEVICE pointers unless a name says `host`, int dims.
 *  - every function returns 0 on success, a negative cnn_status otherwise;
 *    cnn_last_error() gives the message (thread-local).  Nothing throws.
 *  - a batch is ONE contiguous fp32 slab [B][C][H][W]; image b of the slab is the
 *    reference's b-th Tensor3D (CHW, index c*H*W + h*W + w, data_format.h:11-26).
 *  - all work is enqueued on the context's stream; only *_host / cnn_sync / cnn_d2h
 *    block the calling thread.  A context is driven by one host thread (the
 *    reference is single-threaded).
 *  - there is no CPU fallback: without a CUDA device every call fails with
 *    CNN_ERR_CUDA.
 */
#ifndef CNN_B200_H
#define CNN_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CNN_API __attribute__((visibility("default")))
#else
#define CNN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    CNN_OK = 0,
    CNN_ERR_ARG = -1,     /* bad shape / null pointer (the reference asserts, conv2d.cpp:14-15) */
    CNN_ERR_CUDA = -2,    /* CUDA runtime / driver error, or no device */
    CNN_ERR_STATE = -3,   /* call order (e.g. backward before forward) */
    CNN_ERR_UNSUPPORTED = -4,
    CNN_ERR_NCCL = -5     /* NCCL error, or libnccl.so.2 not loadable */
} cnn_status;

typedef struct cnn_ctx cnn_ctx;
typedef struct cnn_net cnn_net;

/* layer codes used by cnn_net_create specs: {type, a, b, c, d} */
enum { CNN_CONV = 0, CNN_BN = 1, CNN_RELU = 2, CNN_POOL = 3, CNN_LINEAR = 4,
       CNN_PAD = 5, CNN_AVGPOOL = 6 /* extensions: items 7-8 of the reference's TODO list, cnn.cpp:15-24 */ };

/* conv algorithm selection (cnn_ctx_set_conv_algo) */
enum { CNN_CONV_AUTO = 0, CNN_CONV_SIMT = 1, CNN_CONV_TCGEN05 = 2 };

/* operand split of the generic tensor-core path, conv_tc.cu (cnn_ctx_set_tc_precision).  fp32 operands
 * are split x = hi + lo and every K-step issues hi*hi + hi*lo + lo*hi:
 *   TF32X3 (default)  hi = tf32(x), lo = x - hi exact: products good to ~2^-22.
 *   BF16X3  hi = bf16(x), lo = bf16(x - hi): 2x MMA rate, products good to ~2^-16.
 *   MIXED   forward / input gradient TF32X3, weight gradient BF16X3.
 *   BF16X1  single pass: hi = bf16(x) only, fp32 accumulate (BASELINE config 5, "bf16 accumulate fp32"):
 *           ~2^-9 per product, graded at a bf16 tolerance; 3x fewer MMAs than BF16X3.
 * A gradient whose per-image contributions cancel across the batch amplifies the product error by the
 * cancellation factor (measured: 6e-4 on the weight gradients of a B=256 step with 2^-16 products), so
 * the two-piece bf16 modes are opt-in speed modes, not parity modes.  The packed stride-2 kernels
 * (conv_s2.cu) always split into three bf16 pieces (2^-24). */
enum { CNN_TC_TF32X3 = 0, CNN_TC_BF16X3 = 1, CNN_TC_MIXED = 2, CNN_TC_BF16X1 = 3 };

/* ---- context, errors, memory ------------------------------------------------ */

CNN_API const char* cnn_last_error(void);
CNN_API const char* cnn_version(void);

/* stream: a cudaStream_t the caller owns (e.g. torch's current stream), or NULL to let
 * the context create its own non-blocking stream. */
CNN_API int cnn_ctx_create(int device, void* stream, cnn_ctx** out);
CNN_API int cnn_ctx_destroy(cnn_ctx* ctx);
CNN_API int cnn_ctx_set_stream(cnn_ctx* ctx, void* stream);
CNN_API void* cnn_ctx_stream(cnn_ctx* ctx);
/* one process per GPU: bind the calling thread to the CPUs of the GPU's NUMA node (pages it touches first -- pinned
 * staging buffers -- then live next to the GPU); a no-op where the topology cannot be read */
CNN_API int cnn_ctx_bind_numa(cnn_ctx* ctx);
CNN_API int cnn_ctx_set_conv_algo(cnn_ctx* ctx, int algo);
CNN_API int cnn_ctx_set_tc_precision(cnn_ctx* ctx, int mode);
CNN_API int cnn_sync(cnn_ctx* ctx);
/* number of kernels this library launched on the context so far (bench `gpu_launches`) */
CNN_API long long cnn_launch_count(cnn_ctx* ctx);
/* Per-launch device times of everything launched on the context between _begin and _end (CUDA events on
 * the launching stream around every kernel; eager launches only -- a graph replay has no events inside):
 * names = '\n'-separated kernel identifiers, us[i] = microseconds from launch i to launch i+1. */
CNN_API int cnn_prof_begin(cnn_ctx* ctx);
CNN_API int cnn_prof_end(cnn_ctx* ctx, char* names, size_t names_cap, float* us, int max_entries, int* n_out);

/* Replaces `new data_type[C*H*W]` / the Tensor3D destructor (data_format.h:17-26,
 * data_format.cpp:152-158): device slabs, pinned host staging, copies. */
CNN_API int cnn_malloc(cnn_ctx* ctx, size_t bytes, void** dptr);
CNN_API int cnn_free(cnn_ctx* ctx, void* dptr);
CNN_API int cnn_host_alloc(cnn_ctx* ctx, size_t bytes, void** hptr); /* pinned */
CNN_API int cnn_host_free(cnn_ctx* ctx, void* hptr);
CNN_API int cnn_memset(cnn_ctx* ctx, void* dptr, int byte, size_t bytes);            /* Tensor3D::set_zero */
CNN_API int cnn_h2d(cnn_ctx* ctx, void* dst, const void* host_src, size_t bytes);    /* async on the stream */
CNN_API int cnn_d2h(cnn_ctx* ctx, void* host_dst, const void* src, size_t bytes);    /* blocks until done */
CNN_API int cnn_d2d(cnn_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---- layer operators --------------------------------------------------------- */

/* Conv2D::forward, conv2d.cpp:34-94.  w [Cout][Cin][k][k], no padding, k odd >= 3
 * (k == 1 also accepted), OH = (H-k)/stride + 1. */
CNN_API int cnn_conv2d_forward(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y,
                       int B, int Cin, int H, int W, int Cout, int k, int stride);
/* Conv2D::forward -> ReLU::forward -> MaxPool2D::forward (the head of the reference model,
 * alexnet.cpp:12-16) in one pass: all three layers' output buffers and the pool arg-max mask
 * (pool2d.cpp:81; may be NULL under WithoutGrad) are written, bit-identical to the three calls.
 * Served for the reference's first layer (3 -> 16, k 3, stride 2) with a 2x2 / step-2 pool;
 * CNN_ERR_UNSUPPORTED otherwise (callers fall back to the separate entry points). */
CNN_API int cnn_conv2d_relu_maxpool_forward(cnn_ctx* ctx, const float* x, const float* w, const float* bias,
                                    float* y_conv, float* y_relu, float* y_pool, int32_t* mask, int B, int Cin,
                                    int H, int W, int Cout, int k, int stride, int pool_k, int pool_step);

/* Conv2D::backward, weight + bias gradient, conv2d.cpp:108-159.  Overwrites dw/db with
 * scale * sum over the batch; the reference's scale is 1/B (under data parallelism
 * 1/B_global). */
CNN_API int cnn_conv2d_backward_weights(cnn_ctx* ctx, const float* x, const float* delta, float* dw,
                                float* db, int B, int Cin, int H, int W, int Cout, int k,
                                int stride, float scale);
/* Conv2D::backward, input gradient, conv2d.cpp:161-201 (gather form of the scatter at
 * :192; cells no window covers stay 0). */
CNN_API int cnn_conv2d_backward_data(cnn_ctx* ctx, const float* w, const float* delta, float* dx, int B,
                             int Cin, int H, int W, int Cout, int k, int stride);

/* MaxPool2D::forward, pool2d.cpp:7-89.  mask (int32 [B][C][OH][OW], may be NULL) holds the
 * reference's flat CHW index of the first maximum (pool2d.cpp:71,79-82). */
CNN_API int cnn_maxpool_forward(cnn_ctx* ctx, const float* x, float* y, int32_t* mask, int B, int C,
                        int H, int W, int k, int step);
/* MaxPool2D::backward, pool2d.cpp:92-109: zero + assign (last writer wins if step < k). */
CNN_API int cnn_maxpool_backward(cnn_ctx* ctx, const float* delta, const int32_t* mask, float* dx, int B,
                         int C, int H, int W, int k, int step);

/* Fused ReLU::forward + MaxPool2D::forward for a pool that directly follows a ReLU with
 * non-overlapping windows (step >= k): one pass over x writes BOTH layers' outputs (and the mask),
 * bit-identical to the two separate calls.  Used by the engine; the layer classes stay separate. */
CNN_API int cnn_relu_maxpool_forward(cnn_ctx* ctx, const float* x, float* y_relu, float* y_pool, int32_t* mask,
                             int B, int C, int H, int W, int k, int step);
/* Fused MaxPool2D::backward + ReLU::backward (in place on the pool's delta_output, relu.cpp:39):
 * dx[mask[i]] = (pool_out[i] <= 0) ? 0 : delta[i], 0 elsewhere.  pool_out[i] is the ReLU output at the
 * arg-max cell, so the ReLU output itself is not read. */
CNN_API int cnn_maxpool_relu_backward(cnn_ctx* ctx, const float* delta, const int32_t* mask, const float* pool_out,
                              float* dx, int B, int C, int H, int W, int k, int step);

/* ReLU::forward relu.cpp:9-28 (x >= 0 ? x : 0) and ReLU::backward relu.cpp:30-44
 * (in place on delta, keyed on the saved OUTPUT: y <= 0 ? 0 : delta). */
CNN_API int cnn_relu_forward(cnn_ctx* ctx, const float* x, float* y, size_t n);
CNN_API int cnn_relu_backward(cnn_ctx* ctx, float* delta, const float* y, size_t n);

/* LinearLayer::forward linear.cpp:22-45; w is [in][out] row-major. */
CNN_API int cnn_linear_forward(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y,
                       int B, int in, int out);
/* LinearLayer::backward linear.cpp:47-93; dw/db overwritten with scale*sum; dx may be NULL. */
CNN_API int cnn_linear_backward(cnn_ctx* ctx, const float* x, const float* w, const float* delta,
                        float* dw, float* db, float* dx, int B, int in, int out, float scale);

/* BatchNorm2D::forward batchnorm2d.cpp:24-97.  Train: two-pass biased batch statistics,
 * moving = (1-momentum)*moving + momentum*batch; saves xhat, batch_mean, batch_var. */
CNN_API int cnn_bn_forward_train(cnn_ctx* ctx, const float* x, const float* gamma, const float* beta,
                         float* moving_mean, float* moving_var, float* batch_mean,
                         float* batch_var, float* xhat, float* y, int B, int C, int H, int W,
                         float eps, float momentum);
CNN_API int cnn_bn_forward_eval(cnn_ctx* ctx, const float* x, const float* gamma, const float* beta,
                        const float* moving_mean, const float* moving_var, float* xhat, float* y,
                        int B, int C, int H, int W, float eps);
/* BatchNorm2D::backward batchnorm2d.cpp:100-158: in place on delta; dgamma/dbeta are plain
 * sums (NOT divided by B, :123-124). */
CNN_API int cnn_bn_backward(cnn_ctx* ctx, float* delta, const float* x, const float* xhat,
                    const float* gamma, const float* batch_mean, const float* batch_var,
                    float* dgamma, float* dbeta, int B, int C, int H, int W, float eps);

/* Tensor3D::read_from_opencv_mat (data_format.cpp:13-23) for a batch: src [B][H][W][C] uint8
 * (OpenCV's interleaved layout) -> dst [B][C][H][W] fp32, dst = src * 1.f / 255. */
CNN_API int cnn_u8hwc_to_chw(cnn_ctx* ctx, const uint8_t* src, float* dst, int B, int C, int H, int W);

/* softmax (func.cpp:16-37) + one_hot (:40-53) + cross_entroy_backward (:56-73) +
 * Tensor3D::argmax (data_format.cpp:37-48) in one launch.  labels may be NULL (inference:
 * probs + pred only).  loss_sum receives sum_b log(p[b][label]) (with the reference's
 * 0*log(0) = NaN behaviour); the caller's loss is -loss_sum / B_global.  delta = p - onehot. */
CNN_API int cnn_softmax_xent(cnn_ctx* ctx, const float* logits, const int32_t* labels, float* probs,
                     float* delta, float* loss_sum, int32_t* pred, int B, int classes);

/* cross_entroy_backward exactly as func.cpp:56-73 takes it: probabilities and one-hot label rows
 * (both [B][classes]); delta = p - y, loss_sum = sum_b sum_i log(p) * y (0*log(0) = NaN kept). */
CNN_API int cnn_xent_backward(cnn_ctx* ctx, const float* probs, const float* onehot, float* delta,
                      float* loss_sum, int B, int classes);

/* <Layer>::update_gradients: p -= lr * g (conv2d.cpp:205-217, linear.cpp:95-102,
 * batchnorm2d.cpp:161-166), one launch over a flat slab. */
CNN_API int cnn_sgd_step(cnn_ctx* ctx, float* params, const float* grads, size_t n, float lr);
/* AvgPool2D / global average pool -- item 7 of the reference's TODO list (cnn.cpp:15-24); window geometry of
 * MaxPool2D (pool2d.cpp:14), k = H = W is the global pool.  backward: dx = sum over covering windows of delta / k^2. */
CNN_API int cnn_avgpool_forward(cnn_ctx* ctx, const float* x, float* y, int B, int C, int H, int W, int k, int step);
CNN_API int cnn_avgpool_backward(cnn_ctx* ctx, const float* delta, float* dx, int B, int C, int H, int W, int k, int step);
/* Zero padding as a layer (item 8 of the same list): y[B][C][H+2p][W+2p] with x in the middle; the backward crops.
 * A padded convolution is this layer in front of the unpadded one. */
CNN_API int cnn_pad2d_forward(cnn_ctx* ctx, const float* x, float* y, int B, int C, int H, int W, int pad);
CNN_API int cnn_pad2d_backward(cnn_ctx* ctx, const float* delta, float* dx, int B, int C, int H, int W, int pad);
/* Optimizer extensions -- item 2 of the reference's TODO list (cnn.cpp:15-24: "momentum, Adam"), over the same
 * flat slabs: v = momentum*v + g, p -= lr*v; and Adam with bias correction at step t >= 1.  State slabs are
 * caller-owned device buffers of n floats, zero before the first step. */
CNN_API int cnn_sgd_momentum_step(cnn_ctx* ctx, float* params, const float* grads, float* velocity, size_t n,
                          float lr, float momentum);
CNN_API int cnn_adam_step(cnn_ctx* ctx, float* params, const float* grads, float* m, float* v, size_t n, float lr,
                  float beta1, float beta2, float eps, int t);

/* ---- whole-network engine ------------------------------------------------------
 * What AlexNet::{forward,backward,update_gradients,save_weights,load_weights}
 * (alexnet.cpp:35-90) plus the step body cnn.cpp:81-92 do, with all buffers resident:
 * one flat parameter slab and one flat gradient slab in checkpoint order
 * (alexnet.cpp:69-77), activations in [B][C][H][W] slabs, the step captured in a CUDA
 * graph.  specs: n_layers x 5 ints {type,a,b,c,d}:
 *   CONV cin,cout,k,stride | BN channels | RELU | POOL k,step | LINEAR in,out
 *   extensions: PAD border | AVGPOOL k,step                                             */
CNN_API int cnn_net_create(cnn_ctx* ctx, const int* specs, int n_layers, int B, int C, int H, int W,
                   cnn_net** out);
CNN_API int cnn_net_destroy(cnn_net* net);
CNN_API long long cnn_net_param_count(const cnn_net* net);
CNN_API int cnn_net_num_classes(const cnn_net* net);
CNN_API float* cnn_net_params(cnn_net* net);   /* device, param_count floats */
CNN_API float* cnn_net_grads(cnn_net* net);    /* device, param_count floats (+ tail: see below) */
/* grad slab tail: [param_count] = sum_b log p[label] of the last step (rides the same
 * all-reduce); cnn_net_grad_slab_count = param_count + 1. */
CNN_API long long cnn_net_grad_slab_count(const cnn_net* net);
CNN_API int cnn_net_set_params_host(cnn_net* net, const float* host_src);  /* checkpoint bytes */
CNN_API int cnn_net_get_params_host(cnn_net* net, float* host_dst);
CNN_API int cnn_net_get_grads_host(cnn_net* net, float* host_dst);
CNN_API int cnn_net_use_graph(cnn_net* net, int enable);
/* forward over a device batch; no_grad != 0 == WithoutGrad (BN eval branch, no masks). */
CNN_API int cnn_net_forward(cnn_net* net, const float* x, int no_grad);
CNN_API const float* cnn_net_logits(cnn_net* net);                       /* device [B][classes] */
CNN_API const float* cnn_net_probs(cnn_net* net);                        /* device [B][classes] */
/* Layer::get_output (architectures.h:45) of layer idx copied to host as [B][C][H][W]. */
CNN_API int cnn_net_layer_output_host(cnn_net* net, int idx, float* host_dst, long long* count);
/* softmax-xent + backward of every layer: fills the gradient slab with
 * grad_scale * sum over the LOCAL batch (grad_scale = 1/B_global). */
CNN_API int cnn_net_backward(cnn_net* net, const int32_t* labels, float grad_scale);
CNN_API const float* cnn_net_input_grad(cnn_net* net);                   /* device dL/d image */
/* Lazy head (default on): cnn_net_train_step* keep only what the step consumes from a
 * Conv2D(3->16,k3,s2) -> ReLU -> MaxPool(2,2) head (alexnet.cpp:12-16): pooled activations and one
 * arg-max/sign code per pool window.  The head's Layer::get_output tensors, the pool mask
 * (pool2d.cpp:79-82) and the image gradient AlexNet::backward returns (alexnet.cpp:55) are re-created
 * bit-identically on demand by cnn_net_layer_output_host / cnn_net_input_grad / cnn_net_pool_mask /
 * cnn_net_materialize from the step's input batch (which must still be intact) and the pre-update
 * filters.  enable = 0: every step writes all of them, like the reference. */
CNN_API int cnn_net_set_lazy(cnn_net* net, int enable);
CNN_API int cnn_net_materialize(cnn_net* net);
CNN_API const int32_t* cnn_net_pool_mask(cnn_net* net, int idx);           /* device int32 mask of pool layer idx */
CNN_API int cnn_net_update(cnn_net* net, float lr);
/* forward + backward (+ update when do_update & 1) as one graph launch.  do_update & 2: the gradient
 * slab (gradients + loss tail) is summed over the ranks of cnn_dist_init between backward and update,
 * inside the same graph -- the data-parallel step of SURVEY §8e; grad_scale is then 1 / B_global. */
CNN_API int cnn_net_train_step(cnn_net* net, const float* x, const int32_t* labels, float lr,
                       float grad_scale, int do_update);
/* The reference-facing call: HOST images [B][C][H][W] and labels in, loss (=-sum/B) and
 * probabilities out; H2D and D2H inside.  host buffers should be pinned (cnn_host_alloc). */
CNN_API int cnn_net_train_step_host(cnn_net* net, const float* host_x, const int32_t* host_labels,
                            float lr, float* host_loss, float* host_probs);
/* The same call, pipelined: _submit enqueues the H2D of the batch on a copy stream plus the step
 * and returns; _wait blocks for the OLDEST submitted step and returns its loss / probabilities.
 * At most two steps in flight; the host buffers of a submission must stay valid (and should be
 * pinned) until its _wait returns.  A loop `submit(i+1); wait(i)` overlaps the PCIe transfer of
 * the next batch with the current step.  Results are identical to cnn_net_train_step_host. */
CNN_API int cnn_net_train_step_host_submit(cnn_net* net, const float* host_x, const int32_t* host_labels,
                                   float lr);
CNN_API int cnn_net_train_step_host_wait(cnn_net* net, float* host_loss, float* host_probs);
/* _submit for images as the reference's loader holds them before Tensor3D::read_from_opencv_mat
 * (data_format.cpp:13-23, pipeline.cpp:143-164): B interleaved [H][W][C] uint8 images; the planar
 * `v * 1.f / 255` conversion runs on the device (cnn_u8hwc_to_chw), bit-identical. */
CNN_API int cnn_net_train_step_host_submit_u8(cnn_net* net, const uint8_t* host_hwc, const int32_t* host_labels,
                                      float lr);
CNN_API int cnn_net_predict_host(cnn_net* net, const float* host_x, float* host_probs,
                         int32_t* host_pred);

/* ---- data parallelism: one process per GPU, ONE all-reduce of the gradient slab per step -------
 * The batch mean of the weight gradients (conv2d.cpp:148,157, linear.cpp:62,70) is the path's only
 * exchange.  NCCL (libnccl.so.2) is bound at run time.  Rank 0 calls cnn_dist_unique_id and ships the
 * 128 bytes to the other ranks by any means; every rank then calls cnn_dist_init (collective). */
CNN_API int cnn_dist_unique_id(void* out128);
CNN_API int cnn_dist_init(cnn_ctx* ctx, int rank, int world, const void* id128);
CNN_API int cnn_dist_world(const cnn_ctx* ctx);
/* SyncBN: BatchNorm2D batch statistics (batchnorm2d.cpp:46-61) and the backward sums (:118-147) are
 * taken over the GLOBAL batch -- 2 all-reduces of C floats forward, 1 of 3C floats backward -- so N
 * ranks at B_local reproduce the single-process reference at B_global.  Off (default): per-rank
 * statistics = the reference at B_local.  Collective: all ranks set it and call BN alike. */
CNN_API int cnn_dist_set_sync_bn(cnn_ctx* ctx, int enable);
CNN_API int cnn_dist_allreduce_sum(cnn_ctx* ctx, float* buf, size_t n);   /* in stream order, in place */
/* One-shot gradient exchange fused with the SGD step over NVLink peer memory (collective over the ranks of
 * cnn_dist_init, one node): every rank maps all peers' gradient slabs (cudaIpc) and a data-parallel step
 * (do_update & 2) then runs ONE kernel that reads the slabs of all ranks, adds them in rank order (identical on
 * every rank: replicas stay bit-identical) and applies p -= lr*g, instead of ncclAllReduce + the SGD kernel.
 * Returns CNN_ERR_UNSUPPORTED -- on every rank alike -- where peer mapping is not available; the NCCL path stays. */
CNN_API int cnn_net_enable_peer_exchange(cnn_net* net);
CNN_API int cnn_dist_finalize(cnn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* CNN_B200_H */
