// bn.cu -- BatchNorm2D forward (train / eval) and backward (batchnorm2d.cpp:24-158).
//
// Per channel the reference reduces over N = B*H*W elements three times (mean, biased
// variance, backward sums).  Here every reduction is a deterministic two-level sum:
// grid (C, S) blocks each reduce one contiguous slice of the channel with warp shuffles,
// a finalize kernel adds the S partials in order.  The two-pass mean -> variance order of
// the reference is kept (it is also the numerically safer one).  Elementwise passes are
// plane-contiguous float4 streams.
#include "common.cuh"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ size_t chan_addr(size_t e, int c, int C, int HW) {
    const size_t b = e / HW, i = e % HW;
    return (b * C + c) * (size_t)HW + i;
}

// mode 0: sum x ; mode 1: sum (x - mean[c])^2
template <int MODE>
__global__ void bn_partial(const float* __restrict__ x, const float* __restrict__ mean,
                           float* __restrict__ partial, int C, int HW, size_t N, int S) {
    __shared__ float red[32];
    const int c = blockIdx.x, s = blockIdx.y;
    const size_t per = (N + S - 1) / S;
    const size_t beg = per * s, end = min(N, beg + per);
    const float mu = MODE ? mean[c] : 0.f;
    float acc = 0.f;
    for (size_t e = beg + threadIdx.x; e < end; e += kT) {
        const float v = x[chan_addr(e, c, C, HW)];
        if (MODE) { const float d = v - mu; acc = fmaf(d, d, acc); } else acc += v;
    }
    const float r = block_sum(acc, red);
    if (threadIdx.x == 0) partial[c * S + s] = r;
}

// MODE 0: mean[c] = sum/N.  MODE 1: var[c] = sum/N and the moving statistics update
// (batchnorm2d.cpp:78-79).
template <int MODE>
__global__ void bn_finalize(const float* __restrict__ partial, float* __restrict__ out,
                            const float* __restrict__ mean, float* __restrict__ moving_mean,
                            float* __restrict__ moving_var, int C, int S, float invN, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    for (int i = 0; i < S; ++i) s += partial[c * S + i];
    const float v = s * invN;
    out[c] = v;
    if (MODE) {
        moving_mean[c] = (1.f - momentum) * moving_mean[c] + momentum * mean[c];
        moving_var[c] = (1.f - momentum) * moving_var[c] + momentum * v;
    }
}

// xhat = (x - mean) * inv ; y = gamma * xhat + beta   (batchnorm2d.cpp:67-75, :84-93)
__global__ void bn_normalize(const float* __restrict__ x, const float* __restrict__ gamma,
                             const float* __restrict__ beta, const float* __restrict__ mean,
                             const float* __restrict__ var, float* __restrict__ xhat,
                             float* __restrict__ y, int C, int HW, float eps, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int c = (int)((idx / HW) % C);
        const float inv = 1.f / sqrtf(var[c] + eps);
        const float n = (x[idx] - mean[c]) * inv;
        xhat[idx] = n;
        y[idx] = gamma[c] * n + beta[c];
    }
}

// backward partials: S1 = sum delta, S2 = sum delta*xhat, S3 = sum (x - mean)
__global__ void bn_bwd_partial(const float* __restrict__ delta, const float* __restrict__ x,
                               const float* __restrict__ xhat, const float* __restrict__ mean,
                               float* __restrict__ partial, int C, int HW, size_t N, int S) {
    __shared__ float red[32];
    const int c = blockIdx.x, s = blockIdx.y;
    const size_t per = (N + S - 1) / S;
    const size_t beg = per * s, end = min(N, beg + per);
    const float mu = mean[c];
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (size_t e = beg + threadIdx.x; e < end; e += kT) {
        const size_t a = chan_addr(e, c, C, HW);
        const float d = delta[a];
        a1 += d;
        a2 = fmaf(d, xhat[a], a2);
        a3 += x[a] - mu;
    }
    a1 = block_sum(a1, red);
    a2 = block_sum(a2, red);
    a3 = block_sum(a3, red);
    if (threadIdx.x == 0) {
        partial[(0 * C + c) * S + s] = a1;
        partial[(1 * C + c) * S + s] = a2;
        partial[(2 * C + c) * S + s] = a3;
    }
}

// Per channel: dgamma, dbeta and the three coefficients of the elementwise pass.  With
// g = delta*gamma and xhat = (x-mean)*inv (batchnorm2d.cpp:129-155):
//   dvar  = sum g (x-mean) (-1/2) inv^3 = -1/2 inv^2 gamma S2
//   dmean = sum(-g inv) + (dvar/N)(-2) sum(x-mean) = -inv gamma S1 - 2 (dvar/N) S3
//   dx    = g inv + (dvar/N) 2 (x-mean) + dmean/N
__global__ void bn_bwd_finalize(const float* __restrict__ partial, const float* __restrict__ gamma,
                                const float* __restrict__ mean, const float* __restrict__ var,
                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                float* __restrict__ coef, int C, int S, float invN, float eps, int write_dgrad) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int i = 0; i < S; ++i) {
        s1 += partial[(0 * C + c) * S + i];
        s2 += partial[(1 * C + c) * S + i];
        s3 += partial[(2 * C + c) * S + i];
    }
    if (write_dgrad) {
        dgamma[c] = s2;
        dbeta[c] = s1;
    }
    const float inv = 1.f / sqrtf(var[c] + eps);
    const float g = gamma[c];
    const float dvar = -0.5f * inv * inv * g * s2;
    const float dvn = dvar * invN;
    const float dmean = -inv * g * s1 - 2.f * dvn * s3;
    coef[c * 4 + 0] = g * inv;
    coef[c * 4 + 1] = 2.f * dvn;
    coef[c * 4 + 2] = dmean * invN;
    coef[c * 4 + 3] = mean[c];
}

__global__ void bn_bwd_apply(float* __restrict__ delta, const float* __restrict__ x,
                             const float* __restrict__ coef, int C, int HW, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4* cf = reinterpret_cast<const float4*>(coef);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int c = (int)((idx / HW) % C);
        const float4 k = cf[c];
        delta[idx] = delta[idx] * k.x + k.y * (x[idx] - k.w) + k.z;
    }
}

inline int pick_splits(const cnn_ctx* ctx, int C, size_t N) {
    int S = cdiv(ctx->sm_count * 4, C);
    const int by_size = (int)((N + 4095) / 4096);
    if (S > by_size) S = by_size;
    if (S > 64) S = 64;
    if (S < 1) S = 1;
    return S;
}

inline int ew_grid(const cnn_ctx* ctx, size_t total) {
    long long g = (long long)((total + kT - 1) / kT);
    const long long cap = (long long)ctx->sm_count * 8;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" {

int cnn_bn_forward_train(cnn_ctx* ctx, const float* x, const float* gamma, const float* beta,
                         float* moving_mean, float* moving_var, float* batch_mean, float* batch_var,
                         float* xhat, float* y, int B, int C, int H, int W, float eps, float momentum) {
    CNN_REQUIRE(ctx && x && gamma && beta && moving_mean && moving_var && batch_mean && batch_var &&
                    xhat && y, "cnn_bn_forward_train: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cnn_bn_forward_train: bad shape");
    const int HW = H * W;
    const size_t N = (size_t)B * HW, total = N * C;
    const int S = pick_splits(ctx, C, N);
    float* partial = cnn_scratch(ctx, sizeof(float) * ((size_t)C * S + (size_t)C));
    CNN_REQUIRE(partial, "scratch allocation failed");
    const float invN = 1.f / (float)N;
    dim3 grid(C, S);
    const int world = ctx->sync_bn ? cnn_dist_world(ctx) : 1;
    if (world > 1) {
        // SyncBN: the statistics are those of the GLOBAL batch, as the single-process
        // reference at B_global computes them -- per-channel sums meet in two C-float all-reduces
        // (sum x, then sum (x - mean)^2 with the global mean: the reference's two-pass order)
        float* sums = partial + (size_t)C * S;
        const float invNg = 1.f / ((float)N * (float)world);
        CNN_LAUNCH(ctx, bn_partial<0>, grid, kT, 0, x, nullptr, partial, C, HW, N, S);
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, partial, sums, nullptr, nullptr, nullptr, C, S, 1.f, momentum);
        if (int rc = cnn_dist_allreduce_sum(ctx, sums, (size_t)C)) return rc;
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, sums, batch_mean, nullptr, nullptr, nullptr, C, 1, invNg, momentum);
        CNN_LAUNCH(ctx, bn_partial<1>, grid, kT, 0, x, batch_mean, partial, C, HW, N, S);
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, partial, sums, nullptr, nullptr, nullptr, C, S, 1.f, momentum);
        if (int rc = cnn_dist_allreduce_sum(ctx, sums, (size_t)C)) return rc;
        CNN_LAUNCH(ctx, bn_finalize<1>, cdiv(C, 128), 128, 0, sums, batch_var, batch_mean, moving_mean, moving_var, C, 1,
                   invNg, momentum);
    } else {
        CNN_LAUNCH(ctx, bn_partial<0>, grid, kT, 0, x, nullptr, partial, C, HW, N, S);
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, partial, batch_mean, nullptr, nullptr, nullptr,
                   C, S, invN, momentum);
        CNN_LAUNCH(ctx, bn_partial<1>, grid, kT, 0, x, batch_mean, partial, C, HW, N, S);
        CNN_LAUNCH(ctx, bn_finalize<1>, cdiv(C, 128), 128, 0, partial, batch_var, batch_mean, moving_mean,
                   moving_var, C, S, invN, momentum);
    }
    CNN_LAUNCH(ctx, bn_normalize, ew_grid(ctx, total), kT, 0, x, gamma, beta, batch_mean, batch_var, xhat,
               y, C, HW, eps, total);
    return CNN_OK;
}

int cnn_bn_forward_eval(cnn_ctx* ctx, const float* x, const float* gamma, const float* beta,
                        const float* moving_mean, const float* moving_var, float* xhat, float* y, int B,
                        int C, int H, int W, float eps) {
    CNN_REQUIRE(ctx && x && gamma && beta && moving_mean && moving_var && xhat && y,
                "cnn_bn_forward_eval: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cnn_bn_forward_eval: bad shape");
    const size_t total = (size_t)B * C * H * W;
    CNN_LAUNCH(ctx, bn_normalize, ew_grid(ctx, total), kT, 0, x, gamma, beta, moving_mean, moving_var,
               xhat, y, C, H * W, eps, total);
    return CNN_OK;
}

int cnn_bn_backward(cnn_ctx* ctx, float* delta, const float* x, const float* xhat, const float* gamma,
                    const float* batch_mean, const float* batch_var, float* dgamma, float* dbeta, int B,
                    int C, int H, int W, float eps) {
    CNN_REQUIRE(ctx && delta && x && xhat && gamma && batch_mean && batch_var && dgamma && dbeta,
                "cnn_bn_backward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cnn_bn_backward: bad shape");
    const int HW = H * W;
    const size_t N = (size_t)B * HW, total = N * C;
    const int S = pick_splits(ctx, C, N);
    float* partial = cnn_scratch(ctx, sizeof(float) * ((size_t)3 * C * S + 4 * (size_t)C + 4 + 3 * (size_t)C));
    CNN_REQUIRE(partial, "scratch allocation failed");
    // keep coef 16-byte aligned for the float4 loads
    float* coef = partial + (((size_t)3 * C * S + 3) / 4) * 4;
    float* sums = coef + 4 * (size_t)C;
    dim3 grid(C, S);
    CNN_LAUNCH(ctx, bn_bwd_partial, grid, kT, 0, delta, x, xhat, batch_mean, partial, C, HW, N, S);
    const int world = ctx->sync_bn ? cnn_dist_world(ctx) : 1;
    if (world > 1) {
        // SyncBN: dgamma / dbeta stay LOCAL sums (the gradient-slab all-reduce adds them up, like every
        // other parameter gradient); the coefficients of dx need the GLOBAL sums S1, S2, S3 and N_global
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(3 * C, 128), 128, 0, partial, sums, nullptr, nullptr, nullptr, 3 * C, S, 1.f, 0.f);
        CNN_LAUNCH(ctx, bn_bwd_finalize, cdiv(C, 128), 128, 0, sums, gamma, batch_mean, batch_var, dgamma, dbeta, coef, C, 1,
                   1.f / (float)N, eps, 1);
        if (int rc = cnn_dist_allreduce_sum(ctx, sums, (size_t)3 * C)) return rc;
        CNN_LAUNCH(ctx, bn_bwd_finalize, cdiv(C, 128), 128, 0, sums, gamma, batch_mean, batch_var, dgamma, dbeta, coef, C, 1,
                   1.f / ((float)N * (float)world), eps, 0);
    } else {
        CNN_LAUNCH(ctx, bn_bwd_finalize, cdiv(C, 128), 128, 0, partial, gamma, batch_mean, batch_var, dgamma,
                   dbeta, coef, C, S, 1.f / (float)N, eps, 1);
    }
    CNN_LAUNCH(ctx, bn_bwd_apply, ew_grid(ctx, total), kT, 0, delta, x, coef, C, HW, total);
    return CNN_OK;
}

}  // extern "C"
