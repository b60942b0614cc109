// net.cu -- whole-network engine: the resident-buffer equivalent of
// AlexNet::{forward,backward,update_gradients,save_weights,load_weights} (alexnet.cpp:35-90)
// and of the train-step body cnn.cpp:81-92.
//
// HBM layout (all fp32, one cudaMalloc each, sized for 180 GB parts):
//   params  [P]      checkpoint order: conv W[Cout][Cin][k][k], bias | linear W[in][out], bias |
//                    BN gamma, beta, moving_mean, moving_var (alexnet.cpp:69-77)
//   grads   [P + 1]  same order (BN moving-stat slots stay 0); tail slot = sum_b log p[label],
//                    so one all-reduce of this slab carries gradients and the loss
//   per layer: out [B][OC][OH][OW]; conv/pool/linear also dx [B][C][H][W] (the reference's
//   delta_output); pool: int32 mask; BN: xhat + batch mean/var
// Layers keep separate outputs exactly like the reference, so Layer::get_output() of any
// layer (gradCAM reads the pre-ReLU conv output, alexnet.cpp:105) is one D2H copy.
// The step is stream-ordered with no host synchronisation and is replayed from a CUDA graph.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

struct LayerRt {
    int type = 0, a = 0, b = 0, c = 0, d = 0;
    int C = 0, H = 0, W = 0, OC = 0, OH = 0, OW = 0;
    size_t w_off = 0, b_off = 0, mm_off = 0, mv_off = 0, w_cnt = 0, b_cnt = 0;
    float* out = nullptr;
    float* dx = nullptr;
    int32_t* mask = nullptr;
    float *xhat = nullptr, *bmean = nullptr, *bvar = nullptr;
    // 3x3 stride-2 layers served by conv_s2.cu: packed input (kept for the weight gradient), packed
    // delta (shared by both gradients) and the bias-gradient partial sums
    bool s2 = false;
    bool s2_px_ready = false;   // the producing layer's epilogue has already written s2_px this pass
    void *s2_px = nullptr, *s2_pd = nullptr, *s2_wf = nullptr, *s2_wd = nullptr;
    float* s2_dbp = nullptr;
    // 3x3 stride-1 layers served by conv_s1.cu: packed input, kept from the forward pass for the weight gradient
    bool s1 = false;
    void* s1_px = nullptr;
    bool s1_px_ready = false;   // this forward pass's packed input was already written by the producing layer
    const float* in = nullptr;
    size_t in_count(int B) const { return (size_t)B * C * H * W; }
    size_t out_count(int B) const { return (size_t)B * OC * OH * OW; }
};

struct GraphKey {
    const float* x = nullptr;
    const int32_t* labels = nullptr;
    float scale = 0.f;             // (the learning rate lives in device memory: not part of the key)
    int do_update = 0;
    const float* scratch = nullptr;
    const void* arena = nullptr;   // both arenas may be re-allocated by later per-operator calls
    // everything else the captured kernels were chosen by (a setter called between steps must not replay
    // the old configuration)
    int conv_algo = 0, tc_precision = 0, sync_bn = 0, world = 1, fuse = 1, lazy = 1, peer = 0;
    bool same_config(const GraphKey& o) const {
        return conv_algo == o.conv_algo && tc_precision == o.tc_precision && sync_bn == o.sync_bn && world == o.world &&
               fuse == o.fuse && lazy == o.lazy && peer == o.peer;
    }
    bool operator==(const GraphKey& o) const {
        return x == o.x && labels == o.labels && scale == o.scale && do_update == o.do_update &&
               scratch == o.scratch && arena == o.arena && same_config(o);
    }
};

}  // namespace

struct cnn_net {
    cnn_ctx* ctx = nullptr;
    int B = 0, C = 0, H = 0, W = 0, classes = 0;
    std::vector<LayerRt> layers;
    size_t P = 0;
    float *params = nullptr, *grads = nullptr, *probs = nullptr, *delta0 = nullptr;
    int32_t* pred = nullptr;
    float* x_in = nullptr;       // device staging for host-fed steps
    int32_t* labels_in = nullptr;
    float* pin_loss = nullptr;   // pinned host word for the loss read-back
    const float* input_grad = nullptr;
    bool forwarded = false, forwarded_train = false;
    bool use_graph = true, warmed = false;
    bool fuse = true;            // ReLU+MaxPool peepholes (results identical to the separate layers)
    // Lazy head: train steps keep only what the step itself consumes from the conv -> ReLU ->
    // MaxPool head (pooled activations in packed form + one code byte per pool window); the head's layer
    // outputs, the pool mask and the image gradient are re-created on demand (net_materialize).
    bool lazy = true;            // cnn_net_set_lazy
    bool lazy_step = false;      // the step in flight is a lazy train step (set by net_step_eager)
    bool head_ok = false;        // layers 0..3 are thin conv, ReLU, 2x2/2 pool, s2 conv
    uint8_t* head_m8 = nullptr;  // [B][POH][POW][16] codes
    float* head_wsave = nullptr; // conv1 filters + biases as the lazy forward saw them
    float* head_dtmp = nullptr;  // transposition scratch of the materialising path (allocated on first use)
    const float* head_x = nullptr;
    bool head_fwd_stale = false; // conv/ReLU/pool outputs + mask of the last forward not materialised
    bool head_bwd_stale = false; // pool / conv1 delta_output (image gradient) of the last backward not materialised
    bool head_lazy_fwd = false;  // the last forward ran the lazy head
    // pipelined host-fed steps (cnn_net_train_step_host_submit / _wait): two staging slots, the H2D
    // of slot i+1 runs on a copy stream while the step of slot i computes
    struct HostSlot {
        float* x = nullptr;          // device [B][C][H][W]
        uint8_t* x_u8 = nullptr;     // device [B][H][W][C] (u8 submissions)
        int32_t* labels = nullptr;   // device [B]
        float* pin = nullptr;        // pinned host: [0] = loss sum, [16...] = probabilities
        cudaEvent_t copied = nullptr, stepped = nullptr;
        bool busy = false;
    };
    HostSlot slots[2];
    cudaStream_t copy_stream = nullptr;
    // weight gradients run on a side stream next to the input-gradient chain (fork / join inside the
    // step, also under graph capture): wgrad(L) overlaps dgrad(L), pack(L-1), ...
    cudaStream_t wg_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool has_bn = false;
    void* s1_pd = nullptr;       // packed delta of the s1 layer in flight (shared: one layer at a time), + bias partials
    float* s1_dbp = nullptr;
    // generic lazy first layer: the image gradient (alexnet.cpp:55) of a first conv layer is computed on demand
    float* first_wsave = nullptr;
    const float* first_delta = nullptr;
    bool first_dgrad_stale = false;
    float* lr_dev = nullptr;         // learning rate of the step in flight (read by the SGD / exchange kernels)
    float lr_host = -1.f;
    void* peer = nullptr;            // one-shot peer-memory gradient exchange + SGD (dist.cu), replaces NCCL all-reduce + sgd_kernel
    bool allreduce_in_bwd = false;   // set by the step when the slab all-reduce is part of it (do_update & 2)
    bool peer_in_bwd = false;        // same for the peer-memory exchange; peer_head = slab elements left to its closing phase
    int peer_sgd = 0;
    size_t peer_head = 0;
    bool allreduce_done = false;     // the backward pass already issued it (overlapped with the first layer)
    unsigned long long submitted = 0, retired = 0;
    struct CachedGraph {
        GraphKey key; cudaGraphExec_t exec = nullptr; long long kernels = 0; bool lazy_head = false;
        bool first_stale = false; const float* first_delta = nullptr;
    };
    GraphKey warm_key;           // configuration of the last eager (warm) step: plans / arenas exist for it
    std::vector<CachedGraph> graphs;  // a few (input buffer, lr, ...) variants, e.g. double-buffered inputs
    std::vector<void*> allocs;
};

namespace {

// [B][HW][C] -> [B][C][HW]
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C,
                                                            int HW) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * C * HW) return;
    const int pos = (int)(i % HW);
    const size_t t = i / HW;
    const int c = (int)(t % C);
    const size_t b = t / C;
    dst[i] = src[(b * HW + pos) * C + c];
}

template <class T>
int dalloc(cnn_net* n, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
    if (e != cudaSuccess) return cnn_cuda_fail(e, "cudaMalloc(net buffer)", __FILE__, __LINE__);
    n->allocs.push_back(q);
    *p = (T*)q;
    return CNN_OK;
}

bool use_s2(const cnn_net* n, const LayerRt& l) {
    return l.s2 && n->ctx->conv_algo == CNN_CONV_AUTO;
}

bool use_s1(const cnn_net* n, const LayerRt& l) {
    return l.s1 && n->ctx->conv_algo == CNN_CONV_AUTO && n->fuse;
}

int net_forward(cnn_net* n, const float* x, bool no_grad, bool lazy = false) {
    cnn_ctx* ctx = n->ctx;
    const int B = n->B;
    const float* cur = x;
    const bool lazy_head = lazy && !no_grad && n->lazy && n->fuse && n->head_ok && n->layers.size() > 3 && use_s2(n, n->layers[3]);
    {   // filter blocks of every packed-path layer, forward and input gradient, in one launch per 8 jobs
        ConvS2PackJob jobs[8];
        int nj = 0;
        for (auto& l : n->layers) {
            if (l.type != CNN_CONV || !use_s2(n, l)) continue;
            for (int dg = 0; dg < (no_grad ? 1 : 2); ++dg) {
                jobs[nj++] = ConvS2PackJob{n->params + l.w_off, dg ? l.s2_wd : l.s2_wf, l.C, l.b, dg};
                if (nj == 8) {
                    if (int rc = conv_s2_pack_weights(ctx, jobs, nj)) return rc;
                    nj = 0;
                }
            }
        }
        if (int rc = conv_s2_pack_weights(ctx, jobs, nj)) return rc;
    }
    n->head_lazy_fwd = false;
    n->head_fwd_stale = n->head_bwd_stale = false;
    for (size_t li = 0; li < n->layers.size(); ++li) {
        LayerRt& l = n->layers[li];
        l.in = cur;
        int rc = CNN_OK;
        ctx->prof_tag = (int)li * 4;
        // lazy training head: one kernel, pooled activations straight into the next conv's packed input
        if (li == 0 && lazy_head) {
            LayerRt& r = n->layers[1];
            LayerRt& p = n->layers[2];
            r.in = l.out;
            p.in = r.out;
            rc = conv_head_fwd(ctx, cur, n->params + l.w_off, n->params + l.b_off, n->head_wsave, n->layers[3].s2_px,
                               nullptr, n->head_m8, B, l.H, l.W);
            if (rc) return rc;
            n->layers[3].s2_px_ready = true;
            n->head_lazy_fwd = n->head_fwd_stale = true;
            n->head_x = cur;
            cur = p.out;
            li += 2;
            continue;
        }
        // ReLU directly followed by a non-overlapping MaxPool: one pass writes both layers' outputs
        if (n->fuse && l.type == CNN_RELU && li + 1 < n->layers.size() && n->layers[li + 1].type == CNN_POOL &&
            n->layers[li + 1].b >= n->layers[li + 1].a) {
            LayerRt& p = n->layers[li + 1];
            p.in = l.out;
            rc = cnn_relu_maxpool_forward(ctx, cur, l.out, p.out, no_grad ? nullptr : p.mask, B, l.C, l.H, l.W, p.a, p.b);
            if (rc) return rc;
            cur = p.out;
            ++li;
            continue;
        }
        // head of the reference model: thin conv -> ReLU -> 2x2/2 MaxPool in one kernel (all outputs written)
        if (n->fuse && l.type == CNN_CONV && li + 2 < n->layers.size() && n->layers[li + 1].type == CNN_RELU &&
            n->layers[li + 2].type == CNN_POOL && ctx->conv_algo == CNN_CONV_AUTO && conv_thin_pool_preferred() &&
            conv_thin_pool_supported(ctx, l.C, l.H, l.W, l.b, l.c, l.d, n->layers[li + 2].a, n->layers[li + 2].b)) {
            LayerRt& r = n->layers[li + 1];
            LayerRt& p = n->layers[li + 2];
            r.in = l.out;
            p.in = r.out;
            rc = conv_fwd_thin_relu_pool(ctx, cur, n->params + l.w_off, n->params + l.b_off, l.out, r.out, p.out,
                                         no_grad ? nullptr : p.mask, B, l.H, l.W);
            if (rc) return rc;
            cur = p.out;
            li += 2;
            continue;
        }
        // thin stride-1 first layer + ReLU (VGG-style nets): one kernel writes both layers' outputs and, when a packed
        // stride-1 layer follows, that layer's packed input
        if (li == 0 && n->fuse && l.type == CNN_CONV && ctx->conv_algo == CNN_CONV_AUTO && n->layers.size() > 1 &&
            n->layers[1].type == CNN_RELU && conv_s1_first_supported(ctx, l.C, l.H, l.W, l.b, l.c, l.d)) {
            LayerRt& r = n->layers[1];
            LayerRt* nxt = (n->layers.size() > 2 && n->layers[2].type == CNN_CONV && use_s1(n, n->layers[2])) ? &n->layers[2] : nullptr;
            r.in = l.out;
            rc = conv_s1_first_fwd(ctx, cur, n->params + l.w_off, n->params + l.b_off, l.out, r.out, nxt ? nxt->s1_px : nullptr, B,
                                   l.H, l.W, l.b);
            if (rc) return rc;
            if (nxt) nxt->s1_px_ready = true;
            cur = r.out;
            ++li;
            continue;
        }
        if (l.type == CNN_CONV && use_s2(n, l)) {
            // packed shifted-window path; a directly following ReLU is written by the same epilogue
            // (both layers' outputs materialise, relu.cpp:25 applied to the stored value)
            const bool relu_next = n->fuse && li + 1 < n->layers.size() && n->layers[li + 1].type == CNN_RELU;
            // conv -> ReLU -> s2 conv: this epilogue also writes the next conv's packed input (no pack kernel)
            LayerRt* nxt = (relu_next && li + 2 < n->layers.size() && n->layers[li + 2].type == CNN_CONV &&
                            use_s2(n, n->layers[li + 2]) && !getenv("CNN_DBG_NOPACKFUSE"))
                               ? &n->layers[li + 2] : nullptr;
            if (!l.s2_px_ready && (rc = conv_s2_pack_x(ctx, cur, l.s2_px, B, l.C, l.H, l.W))) return rc;
            l.s2_px_ready = false;
            rc = conv_s2_fwd_packed(ctx, l.s2_px, n->params + l.w_off, l.s2_wf, n->params + l.b_off, l.out,
                                    relu_next ? n->layers[li + 1].out : nullptr, B, l.C, l.H, l.W, l.b,
                                    nxt ? nxt->s2_px : nullptr);
            if (nxt) nxt->s2_px_ready = true;
            if (rc) return rc;
            cur = l.out;
            if (relu_next) {
                n->layers[li + 1].in = l.out;
                cur = n->layers[li + 1].out;
                ++li;
            }
            continue;
        }
        if (l.type == CNN_CONV && use_s1(n, l)) {
            // packed one-plane path: P(x) is packed once and kept for the weight gradient; a directly following ReLU
            // is written by the same epilogue (both layers' outputs materialise, relu.cpp:25 on the stored value)
            const bool relu_next = li + 1 < n->layers.size() && n->layers[li + 1].type == CNN_RELU;
            if (!l.s1_px_ready && (rc = conv_s1_pack(ctx, cur, nullptr, l.s1_px, nullptr, B, l.C, l.H, l.W, l.H, l.W))) return rc;
            l.s1_px_ready = false;
            rc = conv_s1_fwd_packed(ctx, l.s1_px, n->params + l.w_off, n->params + l.b_off, l.out,
                                    relu_next ? n->layers[li + 1].out : nullptr, B, l.C, l.H, l.W, l.b);
            if (rc) return rc;
            cur = l.out;
            if (relu_next) {
                n->layers[li + 1].in = l.out;
                cur = n->layers[li + 1].out;
                ++li;
            }
            continue;
        }
        switch (l.type) {
            case CNN_CONV:
                rc = cnn_conv2d_forward(ctx, cur, n->params + l.w_off, n->params + l.b_off, l.out, B, l.C,
                                        l.H, l.W, l.b, l.c, l.d);
                break;
            case CNN_BN:
                if (!no_grad)
                    rc = cnn_bn_forward_train(ctx, cur, n->params + l.w_off, n->params + l.b_off,
                                              n->params + l.mm_off, n->params + l.mv_off, l.bmean, l.bvar,
                                              l.xhat, l.out, B, l.C, l.H, l.W, 1e-5f, 0.1f);
                else
                    rc = cnn_bn_forward_eval(ctx, cur, n->params + l.w_off, n->params + l.b_off,
                                             n->params + l.mm_off, n->params + l.mv_off, l.xhat, l.out, B,
                                             l.C, l.H, l.W, 1e-5f);
                break;
            case CNN_RELU:
                rc = cnn_relu_forward(ctx, cur, l.out, l.in_count(B));
                break;
            case CNN_POOL:
                rc = cnn_maxpool_forward(ctx, cur, l.out, no_grad ? nullptr : l.mask, B, l.C, l.H, l.W, l.a,
                                         l.b);
                break;
            case CNN_LINEAR:
                rc = cnn_linear_forward(ctx, cur, n->params + l.w_off, n->params + l.b_off, l.out, B, l.a,
                                        l.b);
                break;
            case CNN_PAD:
                rc = cnn_pad2d_forward(ctx, cur, l.out, B, l.C, l.H, l.W, l.a);
                break;
            case CNN_AVGPOOL:
                rc = cnn_avgpool_forward(ctx, cur, l.out, B, l.C, l.H, l.W, l.a, l.b);
                break;
        }
        if (rc) return rc;
        cur = l.out;
    }
    n->forwarded = true;
    n->forwarded_train = !no_grad;
    return CNN_OK;
}

int net_backward(cnn_net* n, const int32_t* labels, float scale) {
    cnn_ctx* ctx = n->ctx;
    const int B = n->B;
    int rc = cnn_softmax_xent(ctx, n->layers.back().out, labels, n->probs, n->delta0, n->grads + n->P,
                              n->pred, B, n->classes);
    if (rc) return rc;
    float* delta = n->delta0;
    // Fork: the weight gradient of a layer only feeds the gradient slab, so it runs on wg_stream while
    // the main stream continues with the input gradient and the layers below.  All forked kernels
    // share the scratch arena among themselves (one stream: ordered); scratch users on the main
    // stream (generic tcgen05 convs, BN) join first.
    cudaStream_t main_stream = ctx->stream;
    // measured: no gain on AlexNet-lite (the kernels of one layer fill the SMs / TMEM one after the other,
    // 0.7818 vs 0.7834 ms per step), so the fork is opt-in
    const bool can_fork = n->fuse && !n->has_bn && n->wg_stream && getenv("CNN_FORK_WGRAD");
    bool pending = false;
    auto fork_begin = [&]() -> int {
        CNN_CUDA(cudaEventRecord(n->ev_fork, main_stream));
        CNN_CUDA(cudaStreamWaitEvent(n->wg_stream, n->ev_fork, 0));
        ctx->stream = n->wg_stream;
        return CNN_OK;
    };
    auto fork_end = [&]() { ctx->stream = main_stream; pending = true; };
    auto join = [&]() -> int {
        if (!pending) return CNN_OK;
        CNN_CUDA(cudaEventRecord(n->ev_join, n->wg_stream));
        CNN_CUDA(cudaStreamWaitEvent(main_stream, n->ev_join, 0));
        pending = false;
        return CNN_OK;
    };
    // data parallel: every gradient except the first parameter layer's is final once the backward pass
    // reaches that layer -- their all-reduce (>99 % of the slab, loss tail included) starts there on the
    // side stream and overlaps the first layer's weight / input gradient; the small head follows.
    int first_param = -1;
    for (int i = 0; i < (int)n->layers.size() && first_param < 0; ++i)
        if (n->layers[i].w_cnt) first_param = i;
    n->allreduce_done = false;
    n->peer_head = 0;
    bool ar_pending = false;
    size_t ar_head = 0;
    // the exchange of slab elements [head, P + 1] starts on the side stream: the library's ncclAllReduce, or the early
    // phase of the peer-memory exchange (which also applies SGD to those parameters -- nothing below reads them)
    auto start_bulk = [&](size_t head) -> int {
        if (!(n->allreduce_in_bwd || n->peer_in_bwd) || cnn_dist_world(ctx) < 2 || !n->wg_stream || getenv("CNN_DBG_NOAROVERLAP"))
            return CNN_OK;
        // measured at N=2 (256 / 128 images per GPU): the early phase of the peer exchange costs more than it hides
        // (0.429 / 0.309 ms per step against 0.419 / 0.298 with one exchange after the backward pass: a second flag
        // round, and its spinning CTAs share the SMs with the first layer's weight gradient) -- opt-in
        if (n->peer_in_bwd && !getenv("CNN_PEER_OVERLAP")) return CNN_OK;
        if (n->peer_in_bwd) head &= ~(size_t)3;
        if (head == 0 || head >= n->P + 1) return CNN_OK;
        CNN_CUDA(cudaEventRecord(n->ev_fork, main_stream));
        CNN_CUDA(cudaStreamWaitEvent(n->wg_stream, n->ev_fork, 0));
        ctx->stream = n->wg_stream;
        int r = n->peer_in_bwd ? cnn_peer_exchange_bulk(ctx, n->peer, n->lr_dev, n->peer_sgd, head)
                               : cnn_dist_allreduce_sum(ctx, n->grads + head, n->P + 1 - head);
        ctx->stream = main_stream;
        if (r) return r;
        ar_pending = true;
        ar_head = head;
        return CNN_OK;
    };
    for (int i = (int)n->layers.size() - 1; i >= 0; --i) {
        LayerRt& l = n->layers[i];
        ctx->prof_tag = i * 4 + 1;
        if (i == first_param && !pending) {
            const size_t head = l.w_off + l.w_cnt + l.b_cnt + (l.type == CNN_BN ? 2 * l.b_cnt : 0);   // slab elements of this layer
            if (l.w_off == 0 && (rc = start_bulk(head))) return rc;
        }
        if (n->head_lazy_fwd && i == 2) {
            // lazy head: weight / bias gradient of conv1 straight from the pooled delta (pool, ReLU and conv
            // backward composed, no dense intermediate); the image gradient is re-created on demand
            if (!pending) {
                LayerRt& c1 = n->layers[0];
                if (c1.w_off == 0 && (rc = start_bulk(c1.w_off + c1.w_cnt + c1.b_cnt))) return rc;
            }
            LayerRt& c1 = n->layers[0];
            if ((rc = join())) return rc;
            rc = conv_head_wgrad(ctx, c1.in, delta, n->head_m8, n->grads + c1.w_off, n->grads + c1.b_off, B, c1.H, c1.W, scale);
            if (rc) { ctx->stream = main_stream; return rc; }
            n->head_bwd_stale = true;
            delta = nullptr;
            break;
        }
        switch (l.type) {
            case CNN_CONV:
                if (use_s2(n, l)) {
                    // delta packed once for both gradients; P(x) is the forward pass's; the in-place ReLU
                    // backward of the layer below (relu.cpp:39) folds into the input-gradient epilogue
                    const bool relu_below = n->fuse && i > 0 && n->layers[i - 1].type == CNN_RELU;
                    if ((rc = conv_s2_pack_d(ctx, delta, l.s2_pd, l.s2_dbp, B, l.b, l.H, l.W))) return rc;
                    if (can_fork && (rc = fork_begin())) return rc;
                    rc = conv_s2_wgrad_packed(ctx, l.s2_px, l.s2_pd, l.s2_dbp, n->grads + l.w_off, n->grads + l.b_off, B,
                                              l.C, l.H, l.W, l.b, scale);
                    if (can_fork) fork_end();
                    if (rc) return rc;
                    // above a lazy head the pooled delta goes out channel-last (what head_wgrad_kernel streams)
                    rc = conv_s2_dgrad_packed(ctx, l.s2_pd, n->params + l.w_off, l.s2_wd, l.dx,
                                              relu_below ? n->layers[i - 1].out : nullptr, B, l.C, l.H, l.W, l.b,
                                              n->head_lazy_fwd && i == 3);
                    delta = l.dx;
                    if (relu_below) --i;
                    break;
                }
                if (use_s1(n, l)) {
                    // delta packed once for both gradients (bias-gradient partials come with it); the in-place ReLU
                    // backward of the layer below (relu.cpp:39) folds into the input-gradient epilogue
                    const bool relu_below = i > 0 && n->layers[i - 1].type == CNN_RELU;
                    if ((rc = join())) return rc;
                    if ((rc = conv_s1_pack(ctx, delta, nullptr, n->s1_pd, n->s1_dbp, B, l.b, l.H, l.W, l.OH, l.OW))) return rc;
                    rc = conv_s1_wgrad_packed(ctx, l.s1_px, n->s1_pd, n->s1_dbp, n->grads + l.w_off, n->grads + l.b_off, B, l.C,
                                              l.H, l.W, l.b, scale);
                    if (rc) return rc;
                    rc = conv_s1_dgrad_packed(ctx, n->s1_pd, n->params + l.w_off, l.dx, relu_below ? n->layers[i - 1].out : nullptr,
                                              B, l.C, l.H, l.W, l.b);
                    delta = l.dx;
                    if (relu_below) --i;
                    break;
                }
                {
                    // thin first layer: its input gradient needs no scratch, so the weight gradient forks too
                    const bool thin = ctx->conv_algo == CNN_CONV_AUTO && conv_thin_supported(ctx, l.C, l.H, l.W, l.b, l.c, l.d);
                    if (can_fork && thin) {
                        if ((rc = fork_begin())) return rc;
                    } else if ((rc = join())) {
                        return rc;
                    }
                    rc = cnn_conv2d_backward_weights(ctx, l.in, delta, n->grads + l.w_off, n->grads + l.b_off, B, l.C, l.H,
                                                     l.W, l.b, l.c, l.d, scale);
                    if (can_fork && thin) fork_end();
                    if (rc) return rc;
                }
                if (i == 0 && n->lazy_step && n->first_wsave) {
                    // lazy step: nothing in the step consumes the image gradient AlexNet::backward returns (alexnet.cpp:55,
                    // discarded by cnn.cpp:88); cnn_net_input_grad re-creates it from the saved pre-update filters
                    CNN_CUDA(cudaMemcpyAsync(n->first_wsave, n->params + l.w_off, l.w_cnt * sizeof(float), cudaMemcpyDeviceToDevice,
                                             ctx->stream));
                    n->first_delta = delta;
                    n->first_dgrad_stale = true;
                    delta = nullptr;
                    break;
                }
                rc = cnn_conv2d_backward_data(ctx, n->params + l.w_off, delta, l.dx, B, l.C, l.H, l.W, l.b,
                                              l.c, l.d);
                delta = l.dx;
                break;
            case CNN_BN:
                if ((rc = join())) return rc;
                rc = cnn_bn_backward(ctx, delta, l.in, l.xhat, n->params + l.w_off, l.bmean, l.bvar,
                                     n->grads + l.w_off, n->grads + l.b_off, B, l.C, l.H, l.W, 1e-5f);
                break;
            case CNN_RELU:
                rc = cnn_relu_backward(ctx, delta, l.out, l.in_count(B));
                break;
            case CNN_POOL:
                // pool right after a ReLU: the ReLU backward (in place on this layer's delta_output,
                // relu.cpp:39) folds into the scatter, keyed on the pooled value = ReLU output at arg-max
                if (n->fuse && i > 0 && n->layers[i - 1].type == CNN_RELU && l.b >= l.a) {
                    rc = cnn_maxpool_relu_backward(ctx, delta, l.mask, l.out, l.dx, B, l.C, l.H, l.W, l.a, l.b);
                    delta = l.dx;
                    --i;  // the ReLU layer is done
                    break;
                }
                rc = cnn_maxpool_backward(ctx, delta, l.mask, l.dx, B, l.C, l.H, l.W, l.a, l.b);
                delta = l.dx;
                break;
            case CNN_LINEAR: {
                const bool relu_below = n->fuse && i > 0 && n->layers[i - 1].type == CNN_RELU;
                rc = linear_backward_relu(ctx, l.in, n->params + l.w_off, delta, n->grads + l.w_off, n->grads + l.b_off,
                                          l.dx, relu_below ? n->layers[i - 1].out : nullptr, B, l.a, l.b, scale);
                delta = l.dx;
                if (relu_below) --i;
                break;
            }
            case CNN_PAD:
                rc = cnn_pad2d_backward(ctx, delta, l.dx, B, l.C, l.H, l.W, l.a);
                delta = l.dx;
                break;
            case CNN_AVGPOOL:
                rc = cnn_avgpool_backward(ctx, delta, l.dx, B, l.C, l.H, l.W, l.a, l.b);
                delta = l.dx;
                break;
        }
        if (rc) { ctx->stream = main_stream; return rc; }
    }
    if ((rc = join())) return rc;
    if (ar_pending) {   // join the tail all-reduce, then the first layer's few gradients
        CNN_CUDA(cudaEventRecord(n->ev_join, n->wg_stream));
        CNN_CUDA(cudaStreamWaitEvent(main_stream, n->ev_join, 0));
        if (n->peer_in_bwd) {
            n->peer_head = ar_head;   // the closing phase (net_step_eager) takes [0, head)
        } else {
            if ((rc = cnn_dist_allreduce_sum(ctx, n->grads, ar_head))) return rc;
            n->allreduce_done = true;
        }
    }
    n->input_grad = delta;
    return CNN_OK;
}

// Re-create what the lazy head skipped, bit-identical to the eager layers: conv / ReLU / pool outputs and the
// int32 mask from the saved input pointer and the saved pre-update filters; after a backward pass also the
// pool's and conv1's delta_output (the image gradient the reference returns, alexnet.cpp:55).  The input
// batch of the last step must still be intact.
int net_materialize(cnn_net* n, bool want_bwd) {
    cnn_ctx* ctx = n->ctx;
    const int B = n->B;
    LayerRt &c1 = n->layers[0], &r = n->layers[1], &p = n->layers[2];
    int rc;
    if (n->head_fwd_stale) {
        if ((rc = conv_fwd_thin(ctx, n->head_x, n->head_wsave, n->head_wsave + c1.w_cnt, c1.out, B, c1.H, c1.W))) return rc;
        if ((rc = cnn_relu_maxpool_forward(ctx, c1.out, r.out, p.out, p.mask, B, r.C, r.H, r.W, p.a, p.b))) return rc;
        n->head_fwd_stale = false;
    }
    if (want_bwd && n->first_dgrad_stale) {
        LayerRt& l = n->layers[0];
        if ((rc = cnn_conv2d_backward_data(ctx, n->first_wsave, n->first_delta, l.dx, B, l.C, l.H, l.W, l.b, l.c, l.d))) return rc;
        n->input_grad = l.dx;
        n->first_dgrad_stale = false;
    }
    if (want_bwd && n->head_bwd_stale) {
        // the lazy backward left conv2's delta_output channel-last: bring it back to the reference's CHW order
        LayerRt& c2 = n->layers[3];
        const size_t cnt = c2.in_count(B);
        if (!n->head_dtmp && (rc = dalloc(n, &n->head_dtmp, cnt))) return rc;
        CNN_LAUNCH(ctx, nhwc_to_nchw_kernel, cdiv((long long)cnt, 256), 256, 0, c2.dx, n->head_dtmp, B, c2.C, c2.H * c2.W);
        CNN_CUDA(cudaMemcpyAsync(c2.dx, n->head_dtmp, cnt * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = cnn_maxpool_relu_backward(ctx, c2.dx, p.mask, p.out, p.dx, B, p.C, p.H, p.W, p.a, p.b))) return rc;
        if ((rc = conv_dgrad_thin(ctx, n->head_wsave, p.dx, c1.dx, B, c1.H, c1.W))) return rc;
        n->input_grad = c1.dx;
        n->head_bwd_stale = false;
    }
    return CNN_OK;
}

// the learning rate goes to device memory BEFORE a step is captured / replayed (never inside the graph)
int net_set_lr(cnn_net* n, float lr) {
    if (lr == n->lr_host) return CNN_OK;
    n->lr_host = lr;
    return cnn_set_scalar(n->ctx, n->lr_dev, lr);
}

int net_step_eager(cnn_net* n, const float* x, const int32_t* labels, float scale, int do_update) {
    struct TagReset { cnn_ctx* c; ~TagReset() { c->prof_tag = -1; } } reset{n->ctx};
    n->ctx->prof_tag = -1;
    int rc = net_forward(n, x, false, true);
    if (rc) return rc;
    n->ctx->prof_tag = (int)n->layers.size() * 4 + 2;
    n->lazy_step = n->lazy && n->fuse;
    n->first_dgrad_stale = false;
    const bool peer = (do_update & 2) && n->peer && cnn_dist_world(n->ctx) > 1;
    n->allreduce_in_bwd = (do_update & 2) != 0 && !peer;
    n->peer_in_bwd = peer;
    n->peer_sgd = do_update & 1;
    rc = net_backward(n, labels, scale);
    n->allreduce_in_bwd = false;
    n->peer_in_bwd = false;
    n->lazy_step = false;
    if (rc) return rc;
    n->ctx->prof_tag = (int)n->layers.size() * 4 + 2;
    if (peer) return cnn_peer_exchange_step(n->ctx, n->peer, n->lr_dev, do_update & 1, n->peer_head);   // sum over ranks + SGD in one pass
    if ((do_update & 2) && !n->allreduce_done && (rc = cnn_dist_allreduce_sum(n->ctx, n->grads, n->P + 1))) return rc;
    if (do_update & 1) rc = cnn_sgd_step_dev_lr(n->ctx, n->params, n->grads, n->P, n->lr_dev);
    return rc;
}

}  // namespace

namespace {
// the reference asserts 0 <= label < classes in one_hot (func.cpp:40-53); a device-side label cannot be checked
// without a synchronisation, a host-side one can
int check_labels(const cnn_net* n, const int32_t* host_labels) {
    for (int b = 0; b < n->B; ++b)
        if (host_labels[b] < 0 || host_labels[b] >= n->classes) {
            cnn_set_error("label %d of image %d is outside [0, %d)", (int)host_labels[b], b, n->classes);
            return CNN_ERR_ARG;
        }
    return CNN_OK;
}
}  // namespace

extern "C" {

int cnn_net_create(cnn_ctx* ctx, const int* specs, int n_layers, int B, int C, int H, int W, cnn_net** out) {
    CNN_REQUIRE(ctx && specs && out && n_layers > 0, "cnn_net_create: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cnn_net_create: bad input shape");
    CNN_CUDA(cudaSetDevice(ctx->device));
    cnn_net* n = new cnn_net;
    n->ctx = ctx;
    n->B = B; n->C = C; n->H = H; n->W = W;
    size_t off = 0;
    int c = C, h = H, w = W;
    for (int i = 0; i < n_layers; ++i) {
        LayerRt l;
        const int* s = specs + 5 * i;
        l.type = s[0]; l.a = s[1]; l.b = s[2]; l.c = s[3]; l.d = s[4];
        l.C = c; l.H = h; l.W = w;
        bool ok = true;
        switch (l.type) {
            case CNN_CONV:
                ok = l.a == c && l.b > 0 && l.c > 0 && (l.c & 1) && l.d > 0 && h >= l.c && w >= l.c;
                l.OC = l.b; l.OH = (h - l.c) / l.d + 1; l.OW = (w - l.c) / l.d + 1;
                l.w_cnt = (size_t)l.b * l.a * l.c * l.c; l.b_cnt = l.b;
                break;
            case CNN_BN:
                ok = l.a == c;
                l.OC = c; l.OH = h; l.OW = w;
                l.w_cnt = c; l.b_cnt = c;
                break;
            case CNN_RELU:
                l.OC = c; l.OH = h; l.OW = w;
                break;
            case CNN_POOL:
                ok = l.a > 0 && l.b > 0 && h >= l.a && w >= l.a;
                l.OC = c; l.OH = (h - l.a) / l.b + 1; l.OW = (w - l.a) / l.b + 1;
                break;
            case CNN_LINEAR:
                ok = l.a == c * h * w && l.b > 0;
                l.OC = l.b; l.OH = 1; l.OW = 1;
                l.w_cnt = (size_t)l.a * l.b; l.b_cnt = l.b;
                break;
            case CNN_PAD:
                ok = l.a >= 0;
                l.OC = c; l.OH = h + 2 * l.a; l.OW = w + 2 * l.a;
                break;
            case CNN_AVGPOOL:
                ok = l.a > 0 && l.b > 0 && h >= l.a && w >= l.a;
                l.OC = c; l.OH = (h - l.a) / l.b + 1; l.OW = (w - l.a) / l.b + 1;
                break;
            default: ok = false;
        }
        if (!ok) {
            cnn_set_error("cnn_net_create: layer %d (type %d) does not fit input %dx%dx%d", i, l.type, c, h, w);
            delete n;
            return CNN_ERR_ARG;
        }
        if (l.w_cnt) {
            l.w_off = off; off += l.w_cnt;
            l.b_off = off; off += l.b_cnt;
            if (l.type == CNN_BN) {
                l.mm_off = off; off += l.b_cnt;
                l.mv_off = off; off += l.b_cnt;
            }
        }
        c = l.OC; h = l.OH; w = l.OW;
        n->layers.push_back(l);
    }
    n->P = off;
    n->classes = c * h * w;
    int rc = CNN_OK;
    auto fail = [&](int code) { cnn_net_destroy(n); return code; };
    if ((rc = dalloc(n, &n->params, n->P))) return fail(rc);
    if ((rc = dalloc(n, &n->grads, n->P + 1))) return fail(rc);
    if ((rc = dalloc(n, &n->probs, (size_t)B * n->classes))) return fail(rc);
    if ((rc = dalloc(n, &n->delta0, (size_t)B * n->classes))) return fail(rc);
    if ((rc = dalloc(n, &n->pred, (size_t)B))) return fail(rc);
    size_t s1_pd_max = 0, s1_dbp_max = 0;
    for (auto& l : n->layers) {
        if ((rc = dalloc(n, &l.out, l.out_count(B)))) return fail(rc);
        if (l.type == CNN_CONV || l.type == CNN_POOL || l.type == CNN_LINEAR || l.type == CNN_PAD || l.type == CNN_AVGPOOL)
            if ((rc = dalloc(n, &l.dx, l.in_count(B)))) return fail(rc);
        if (l.type == CNN_POOL)
            if ((rc = dalloc(n, &l.mask, l.out_count(B)))) return fail(rc);
        if (l.type == CNN_CONV && conv_s2_supported(ctx, l.C, l.H, l.W, l.b, l.c, l.d)) {
            uint8_t *px = nullptr, *pd = nullptr;
            if ((rc = dalloc(n, &px, conv_s2_px_bytes(B, l.C, l.H, l.W)))) return fail(rc);
            if ((rc = dalloc(n, &pd, conv_s2_pd_bytes(B, l.b, l.H, l.W)))) return fail(rc);
            if ((rc = dalloc(n, &l.s2_dbp, conv_s2_dbp_bytes(B, l.b, l.H, l.W) / sizeof(float)))) return fail(rc);
            uint8_t *wf = nullptr, *wd = nullptr;
            if ((rc = dalloc(n, &wf, conv_s2_wpk_bytes(l.C, l.b, 0)))) return fail(rc);
            if ((rc = dalloc(n, &wd, conv_s2_wpk_bytes(l.C, l.b, 1)))) return fail(rc);
            // the spare plane positions and the tail slack must read as zero when a producer epilogue
            // (which only writes real pixels) fills this buffer instead of the pack kernel
            if (cudaMemsetAsync(px, 0, conv_s2_px_bytes(B, l.C, l.H, l.W), ctx->stream) != cudaSuccess) return fail(CNN_ERR_CUDA);
            l.s2_px = px; l.s2_pd = pd; l.s2_wf = wf; l.s2_wd = wd; l.s2 = true;
        }
        if (l.type == CNN_CONV && !l.s2 && conv_s1_supported(ctx, l.C, l.H, l.W, l.b, l.c, l.d)) {
            uint8_t* px = nullptr;
            if ((rc = dalloc(n, &px, conv_s1_pk_bytes(B, l.C, l.H, l.W)))) return fail(rc);
            // guard band and tail must read as zero when a producing layer (which only writes real pixels) fills this buffer
            if (cudaMemsetAsync(px, 0, conv_s1_pk_bytes(B, l.C, l.H, l.W), ctx->stream) != cudaSuccess) return fail(CNN_ERR_CUDA);
            l.s1_px = px; l.s1 = true;
            s1_pd_max = std::max(s1_pd_max, conv_s1_pk_bytes(B, l.b, l.H, l.W));
            s1_dbp_max = std::max(s1_dbp_max, conv_s1_dbp_bytes(B, l.b, l.H, l.W));
        }
        if (l.type == CNN_BN) {
            if ((rc = dalloc(n, &l.xhat, l.in_count(B)))) return fail(rc);
            if ((rc = dalloc(n, &l.bmean, (size_t)l.C))) return fail(rc);
            if ((rc = dalloc(n, &l.bvar, (size_t)l.C))) return fail(rc);
        }
    }
    for (auto& l : n->layers) n->has_bn = n->has_bn || l.type == CNN_BN;
    if (s1_pd_max) {
        uint8_t* pd = nullptr;
        if ((rc = dalloc(n, &pd, s1_pd_max))) return fail(rc);
        if ((rc = dalloc(n, &n->s1_dbp, s1_dbp_max / sizeof(float)))) return fail(rc);
        n->s1_pd = pd;
    }
    if (n->layers[0].type == CNN_CONV && (rc = dalloc(n, &n->first_wsave, n->layers[0].w_cnt))) return fail(rc);
    if ((rc = dalloc(n, &n->lr_dev, 1))) return fail(rc);
    if (n->layers.size() >= 4) {
        const LayerRt &c1 = n->layers[0], &r = n->layers[1], &p = n->layers[2], &c2 = n->layers[3];
        n->head_ok = c1.type == CNN_CONV && r.type == CNN_RELU && p.type == CNN_POOL && c2.type == CNN_CONV && c2.s2 &&
                     conv_head_lazy_supported(ctx, c1.C, c1.H, c1.W, c1.b, c1.c, c1.d, p.a, p.b);
        if (n->head_ok) {
            if ((rc = dalloc(n, &n->head_m8, conv_head_m8_bytes(B, c1.H, c1.W)))) return fail(rc);
            if ((rc = dalloc(n, &n->head_wsave, c1.w_cnt + c1.b_cnt))) return fail(rc);
        }
    }
    if (getenv("CNN_LAZY_HEAD") && atoi(getenv("CNN_LAZY_HEAD")) == 0) n->lazy = false;
    if (cudaStreamCreateWithFlags(&n->wg_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&n->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&n->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cnn_set_error("cnn_net_create: stream / event creation failed");
        return fail(CNN_ERR_CUDA);
    }
    if (cudaMemsetAsync(n->params, 0, n->P * sizeof(float), ctx->stream) != cudaSuccess ||
        cudaMemsetAsync(n->grads, 0, (n->P + 1) * sizeof(float), ctx->stream) != cudaSuccess ||
        cudaHostAlloc((void**)&n->pin_loss, 64, cudaHostAllocDefault) != cudaSuccess) {
        cnn_set_error("cnn_net_create: buffer initialisation failed");
        return fail(CNN_ERR_CUDA);
    }
    *out = n;
    return CNN_OK;
}

int cnn_net_destroy(cnn_net* n) {
    if (!n) return CNN_OK;
    cudaStreamSynchronize(n->ctx->stream);
    for (auto& g : n->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (n->peer) cnn_peer_exchange_destroy(n->peer);
    if (n->copy_stream) {
        cudaStreamSynchronize(n->copy_stream);
        cudaStreamDestroy(n->copy_stream);
    }
    if (n->wg_stream) {
        cudaStreamSynchronize(n->wg_stream);
        cudaStreamDestroy(n->wg_stream);
    }
    if (n->ev_fork) cudaEventDestroy(n->ev_fork);
    if (n->ev_join) cudaEventDestroy(n->ev_join);
    for (auto& sl : n->slots) {
        if (sl.copied) cudaEventDestroy(sl.copied);
        if (sl.stepped) cudaEventDestroy(sl.stepped);
        if (sl.pin) cudaFreeHost(sl.pin);
    }
    for (void* p : n->allocs) cudaFree(p);
    if (n->pin_loss) cudaFreeHost(n->pin_loss);
    delete n;
    return CNN_OK;
}

long long cnn_net_param_count(const cnn_net* n) { return n ? (long long)n->P : 0; }
long long cnn_net_grad_slab_count(const cnn_net* n) { return n ? (long long)n->P + 1 : 0; }
int cnn_net_num_classes(const cnn_net* n) { return n ? n->classes : 0; }
float* cnn_net_params(cnn_net* n) { return n ? n->params : nullptr; }
float* cnn_net_grads(cnn_net* n) { return n ? n->grads : nullptr; }
const float* cnn_net_logits(cnn_net* n) { return n ? n->layers.back().out : nullptr; }
const float* cnn_net_probs(cnn_net* n) { return n ? n->probs : nullptr; }
const float* cnn_net_input_grad(cnn_net* n) {
    if (!n) return nullptr;
    if ((n->head_bwd_stale || n->first_dgrad_stale) && net_materialize(n, true)) return nullptr;
    return n->input_grad;
}

int cnn_net_enable_peer_exchange(cnn_net* n) {
    CNN_REQUIRE(n, "net is NULL");
    if (n->peer) return CNN_OK;
    CNN_CUDA(cudaStreamSynchronize(n->ctx->stream));
    return cnn_peer_exchange_setup(n->ctx, n->grads, n->params, n->P, &n->peer);
}

int cnn_net_set_lazy(cnn_net* n, int enable) {
    CNN_REQUIRE(n, "net is NULL");
    n->lazy = enable != 0;
    return CNN_OK;
}

int cnn_net_materialize(cnn_net* n) {
    CNN_REQUIRE(n, "net is NULL");
    return net_materialize(n, true);
}

const int32_t* cnn_net_pool_mask(cnn_net* n, int idx) {
    if (!n || idx < 0 || idx >= (int)n->layers.size() || n->layers[idx].type != CNN_POOL) return nullptr;
    if (idx < 3 && n->head_fwd_stale && net_materialize(n, false)) return nullptr;
    return n->layers[idx].mask;
}

int cnn_net_set_params_host(cnn_net* n, const float* host_src) {
    CNN_REQUIRE(n && host_src, "cnn_net_set_params_host: NULL argument");
    CNN_CUDA(cudaMemcpyAsync(n->params, host_src, n->P * sizeof(float), cudaMemcpyHostToDevice, n->ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(n->ctx->stream));
    return CNN_OK;
}

int cnn_net_get_params_host(cnn_net* n, float* host_dst) {
    CNN_REQUIRE(n && host_dst, "cnn_net_get_params_host: NULL argument");
    return cnn_d2h(n->ctx, host_dst, n->params, n->P * sizeof(float));
}

int cnn_net_get_grads_host(cnn_net* n, float* host_dst) {
    CNN_REQUIRE(n && host_dst, "cnn_net_get_grads_host: NULL argument");
    return cnn_d2h(n->ctx, host_dst, n->grads, n->P * sizeof(float));
}

int cnn_net_use_graph(cnn_net* n, int enable) {
    CNN_REQUIRE(n, "net is NULL");
    n->use_graph = enable != 0;
    return CNN_OK;
}

int cnn_net_forward(cnn_net* n, const float* x, int no_grad) {
    CNN_REQUIRE(n && x, "cnn_net_forward: NULL argument");
    return net_forward(n, x, no_grad != 0);
}

int cnn_net_layer_output_host(cnn_net* n, int idx, float* host_dst, long long* count) {
    CNN_REQUIRE(n && idx >= 0 && idx < (int)n->layers.size(), "cnn_net_layer_output_host: bad layer index");
    const LayerRt& l = n->layers[idx];
    if (count) *count = (long long)l.out_count(n->B);
    if (!host_dst) return CNN_OK;
    if (idx < 3 && n->head_fwd_stale)
        if (int rc = net_materialize(n, false)) return rc;
    return cnn_d2h(n->ctx, host_dst, l.out, l.out_count(n->B) * sizeof(float));
}

int cnn_net_backward(cnn_net* n, const int32_t* labels, float grad_scale) {
    CNN_REQUIRE(n && labels, "cnn_net_backward: NULL argument");
    if (!n->forwarded_train) {
        cnn_set_error("cnn_net_backward: no preceding forward with gradients enabled");
        return CNN_ERR_STATE;
    }
    return net_backward(n, labels, grad_scale);
}

int cnn_net_update(cnn_net* n, float lr) {
    CNN_REQUIRE(n, "net is NULL");
    return cnn_sgd_step(n->ctx, n->params, n->grads, n->P, lr);
}

int cnn_net_train_step(cnn_net* n, const float* x, const int32_t* labels, float lr, float grad_scale,
                       int do_update) {
    CNN_REQUIRE(n && x && labels, "cnn_net_train_step: NULL argument");
    cnn_ctx* ctx = n->ctx;
    if (int rc = net_set_lr(n, lr)) return rc;
    if (!n->use_graph) return net_step_eager(n, x, labels, grad_scale, do_update);
    GraphKey k{x, labels, grad_scale, do_update, nullptr, nullptr, ctx->conv_algo, ctx->tc_precision,
               (int)ctx->sync_bn, cnn_dist_world(ctx), (int)n->fuse, (int)n->lazy, n->peer ? 1 : 0};
    if (!n->warmed || !n->warm_key.same_config(k)) {
        // first step (and the first one after a configuration change) runs eagerly: sizes the scratch arena,
        // builds kernel plans (cudaMalloc / synchronous uploads are illegal inside a capture), surfaces launch errors
        n->warmed = true;
        n->warm_key = k;
        return net_step_eager(n, x, labels, grad_scale, do_update);
    }
    k.scratch = ctx->scratch;
    k.arena = ctx->arena;
    cnn_net::CachedGraph* hit = nullptr;
    for (auto& g : n->graphs)
        if (g.key == k) hit = &g;
    if (!hit) {
        if (n->graphs.size() >= 8) {  // evict the oldest variant
            cudaGraphExecDestroy(n->graphs.front().exec);
            n->graphs.erase(n->graphs.begin());
        }
        const long long before = ctx->launches;
        CNN_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = net_step_eager(n, x, labels, grad_scale, do_update);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        cnn_net::CachedGraph g;
        g.key = k;
        g.kernels = ctx->launches - before;
        g.lazy_head = n->head_lazy_fwd;
        g.first_stale = n->first_dgrad_stale;
        g.first_delta = n->first_delta;
        ctx->launches = before;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return cnn_cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return cnn_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
        n->graphs.push_back(g);
        hit = &n->graphs.back();
    }
    CNN_CUDA(cudaGraphLaunch(hit->exec, ctx->stream));
    ctx->launches += hit->kernels;
    n->forwarded = n->forwarded_train = true;
    n->head_lazy_fwd = n->head_fwd_stale = n->head_bwd_stale = hit->lazy_head;
    n->head_x = x;
    n->first_dgrad_stale = hit->first_stale;
    n->first_delta = hit->first_delta;
    if (hit->lazy_head || hit->first_stale) n->input_grad = nullptr;
    return CNN_OK;
}

int cnn_net_train_step_host(cnn_net* n, const float* host_x, const int32_t* host_labels, float lr,
                            float* host_loss, float* host_probs) {
    CNN_REQUIRE(n && host_x && host_labels, "cnn_net_train_step_host: NULL argument");
    cnn_ctx* ctx = n->ctx;
    int rc;
    if ((rc = check_labels(n, host_labels))) return rc;
    if (!n->x_in) {
        if ((rc = dalloc(n, &n->x_in, (size_t)n->B * n->C * n->H * n->W))) return rc;
        if ((rc = dalloc(n, &n->labels_in, (size_t)n->B))) return rc;
    }
    const size_t xb = sizeof(float) * (size_t)n->B * n->C * n->H * n->W;
    CNN_CUDA(cudaMemcpyAsync(n->x_in, host_x, xb, cudaMemcpyHostToDevice, ctx->stream));
    CNN_CUDA(cudaMemcpyAsync(n->labels_in, host_labels, sizeof(int32_t) * n->B, cudaMemcpyHostToDevice,
                             ctx->stream));
    // data parallel (cnn_dist_init done): the slab all-reduce rides inside the step, gradients are batch means
    // over the GLOBAL batch and the loss tail is the global sum
    const int world = cnn_dist_world(ctx);
    if ((rc = cnn_net_train_step(n, n->x_in, n->labels_in, lr, 1.f / ((float)n->B * world), world > 1 ? 3 : 1))) return rc;
    CNN_CUDA(cudaMemcpyAsync(n->pin_loss, n->grads + n->P, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (host_probs)
        CNN_CUDA(cudaMemcpyAsync(host_probs, n->probs, sizeof(float) * (size_t)n->B * n->classes,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    // func.cpp:71: loss_value * (-1.0) / batch_size, evaluated in double
    if (host_loss) *host_loss = (float)((double)*n->pin_loss * (-1.0) / ((double)n->B * world));
    return CNN_OK;
}

// ---- pipelined host-fed steps -------------------------------------------------------------------
// submit(i+1) may be called before wait(i): the batch of step i+1 crosses PCIe on the copy stream
// while step i computes, so a training loop is bound by max(H2D, step) instead of their sum.
// u8 submissions take the image as the reference's loader holds it before
// Tensor3D::read_from_opencv_mat (data_format.cpp:13-23): interleaved HWC bytes; the planar
// float conversion (x * 1.f / 255, bit-identical) runs on the device, a quarter of the PCIe bytes.
namespace {
int host_pipe_init(cnn_net* n, bool want_u8) {
    int rc;
    if (!n->copy_stream) CNN_CUDA(cudaStreamCreateWithFlags(&n->copy_stream, cudaStreamNonBlocking));
    const size_t cnt = (size_t)n->B * n->C * n->H * n->W;
    for (auto& sl : n->slots) {
        if (!sl.x) {
            if ((rc = dalloc(n, &sl.x, cnt))) return rc;
            if ((rc = dalloc(n, &sl.labels, (size_t)n->B))) return rc;
            CNN_CUDA(cudaHostAlloc((void**)&sl.pin, sizeof(float) * (16 + (size_t)n->B * n->classes), cudaHostAllocDefault));
            CNN_CUDA(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
            CNN_CUDA(cudaEventCreateWithFlags(&sl.stepped, cudaEventDisableTiming));
        }
        if (want_u8 && !sl.x_u8)
            if ((rc = dalloc(n, &sl.x_u8, cnt))) return rc;
    }
    return CNN_OK;
}

int host_submit(cnn_net* n, const void* host_x, bool u8, const int32_t* host_labels, float lr) {
    cnn_ctx* ctx = n->ctx;
    if (n->submitted - n->retired >= 2) {
        cnn_set_error("cnn_net_train_step_host_submit: two steps already in flight, call _wait first");
        return CNN_ERR_STATE;
    }
    if (int rc = check_labels(n, host_labels)) return rc;
    if (int rc = host_pipe_init(n, u8)) return rc;
    cnn_net::HostSlot& sl = n->slots[n->submitted & 1];
    const size_t cnt = (size_t)n->B * n->C * n->H * n->W;
    // the slot's previous step (two submissions ago) has been waited for, so its buffers are free;
    // the copy stream still orders behind that step's event for callers that skip results
    if (n->submitted >= 2) CNN_CUDA(cudaStreamWaitEvent(n->copy_stream, sl.stepped, 0));
    if (u8) CNN_CUDA(cudaMemcpyAsync(sl.x_u8, host_x, cnt, cudaMemcpyHostToDevice, n->copy_stream));
    else CNN_CUDA(cudaMemcpyAsync(sl.x, host_x, cnt * sizeof(float), cudaMemcpyHostToDevice, n->copy_stream));
    CNN_CUDA(cudaMemcpyAsync(sl.labels, host_labels, sizeof(int32_t) * n->B, cudaMemcpyHostToDevice, n->copy_stream));
    CNN_CUDA(cudaEventRecord(sl.copied, n->copy_stream));
    CNN_CUDA(cudaStreamWaitEvent(ctx->stream, sl.copied, 0));
    int rc;
    if (u8 && (rc = cnn_u8hwc_to_chw(ctx, sl.x_u8, sl.x, n->B, n->C, n->H, n->W))) return rc;
    const int world = cnn_dist_world(ctx);   // > 1: global-batch gradients, all-reduce inside the step graph
    if ((rc = cnn_net_train_step(n, sl.x, sl.labels, lr, 1.f / ((float)n->B * world), world > 1 ? 3 : 1))) return rc;
    CNN_CUDA(cudaMemcpyAsync(sl.pin, n->grads + n->P, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaMemcpyAsync(sl.pin + 16, n->probs, sizeof(float) * (size_t)n->B * n->classes,
                             cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaEventRecord(sl.stepped, ctx->stream));
    sl.busy = true;
    ++n->submitted;
    return CNN_OK;
}
}  // namespace

int cnn_net_train_step_host_submit(cnn_net* n, const float* host_x, const int32_t* host_labels, float lr) {
    CNN_REQUIRE(n && host_x && host_labels, "cnn_net_train_step_host_submit: NULL argument");
    return host_submit(n, host_x, false, host_labels, lr);
}

int cnn_net_train_step_host_submit_u8(cnn_net* n, const uint8_t* host_hwc, const int32_t* host_labels, float lr) {
    CNN_REQUIRE(n && host_hwc && host_labels, "cnn_net_train_step_host_submit_u8: NULL argument");
    return host_submit(n, host_hwc, true, host_labels, lr);
}

int cnn_net_train_step_host_wait(cnn_net* n, float* host_loss, float* host_probs) {
    CNN_REQUIRE(n, "cnn_net_train_step_host_wait: NULL argument");
    if (n->retired == n->submitted) {
        cnn_set_error("cnn_net_train_step_host_wait: nothing in flight");
        return CNN_ERR_STATE;
    }
    cnn_net::HostSlot& sl = n->slots[n->retired & 1];
    CNN_CUDA(cudaEventSynchronize(sl.stepped));
    if (host_loss) *host_loss = (float)((double)sl.pin[0] * (-1.0) / ((double)n->B * cnn_dist_world(n->ctx)));   // func.cpp:71
    if (host_probs) memcpy(host_probs, sl.pin + 16, sizeof(float) * (size_t)n->B * n->classes);
    sl.busy = false;
    ++n->retired;
    return CNN_OK;
}

int cnn_net_predict_host(cnn_net* n, const float* host_x, float* host_probs, int32_t* host_pred) {
    CNN_REQUIRE(n && host_x, "cnn_net_predict_host: NULL argument");
    cnn_ctx* ctx = n->ctx;
    int rc;
    if (!n->x_in) {
        if ((rc = dalloc(n, &n->x_in, (size_t)n->B * n->C * n->H * n->W))) return rc;
        if ((rc = dalloc(n, &n->labels_in, (size_t)n->B))) return rc;
    }
    const size_t xb = sizeof(float) * (size_t)n->B * n->C * n->H * n->W;
    CNN_CUDA(cudaMemcpyAsync(n->x_in, host_x, xb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = net_forward(n, n->x_in, true))) return rc;
    if ((rc = cnn_softmax_xent(ctx, n->layers.back().out, nullptr, n->probs, nullptr, nullptr, n->pred, n->B,
                               n->classes)))
        return rc;
    if (host_probs)
        CNN_CUDA(cudaMemcpyAsync(host_probs, n->probs, sizeof(float) * (size_t)n->B * n->classes,
                                 cudaMemcpyDeviceToHost, ctx->stream));
    if (host_pred)
        CNN_CUDA(cudaMemcpyAsync(host_pred, n->pred, sizeof(int32_t) * n->B, cudaMemcpyDeviceToHost, ctx->stream));
    CNN_CUDA(cudaStreamSynchronize(ctx->stream));
    return CNN_OK;
}

}  // extern "C"
