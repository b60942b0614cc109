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
constexpr int kChunk = 8 * kT;   // elements of one plane a block handles at a time (8 coalesced loads in flight per thread)

// A channel's N = B*HW elements are B plane runs of HW contiguous floats.  Work is cut into units = (image b, chunk k of
// the plane): index arithmetic (one division) per unit, none per element; loads are plane-contiguous and coalesced.
struct Units {
    int nchunk, total;   // chunks per plane, B * nchunk
    __host__ __device__ Units(int B, int HW) : nchunk((HW + kChunk - 1) / kChunk), total(B * ((HW + kChunk - 1) / kChunk)) {}
};

// mode 0: sum x ; mode 1: sum (x - mean[c])^2.  Block (c, s) reduces units [u0, u1) of channel c.
template <int MODE>
__global__ void __launch_bounds__(kT) bn_partial(const float* __restrict__ x, const float* __restrict__ mean,
                                                  float* __restrict__ partial, int B, int C, int HW, int S) {
    __shared__ float red[32];
    const int c = blockIdx.x, s = blockIdx.y;
    const Units U(B, HW);
    const int per = (U.total + S - 1) / S, u0 = per * s, u1 = min(U.total, u0 + per);
    const float mu = MODE ? mean[c] : 0.f;
    float acc = 0.f;
    for (int u = u0; u < u1; ++u) {
        const int b = u / U.nchunk, k = u - b * U.nchunk;
        const float* p = x + ((size_t)b * C + c) * HW;
        const int i0 = k * kChunk + threadIdx.x, lim = min(HW, (k + 1) * kChunk);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = i0 + j * kT < lim ? p[i0 + j * kT] : (MODE ? mu : 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE) { const float d = v[j] - mu; acc = fmaf(d, d, acc); } else acc += v[j];
        }
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

// xhat = (x - mean) * inv ; y = gamma * xhat + beta   (batchnorm2d.cpp:67-75, :84-93).  Block = one unit of one plane.
__global__ void __launch_bounds__(kT) bn_normalize(const float* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                                    const float* __restrict__ var, float* __restrict__ xhat,
                                                    float* __restrict__ y, int C, int HW, float eps, int nchunk, int planes) {
    for (int blk = blockIdx.x; blk < planes * nchunk; blk += gridDim.x) {
        const int pl = blk / nchunk, k = blk - pl * nchunk, c = pl % C;
        const float inv = 1.f / sqrtf(var[c] + eps), mu = mean[c], g = gamma[c], bt = beta[c];
        const size_t base = (size_t)pl * HW;
        const int i0 = k * kChunk + threadIdx.x, lim = min(HW, (k + 1) * kChunk);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = i0 + j * kT < lim ? x[base + i0 + j * kT] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j * kT < lim) {
                const float n = (v[j] - mu) * inv;
                xhat[base + i0 + j * kT] = n;
                y[base + i0 + j * kT] = g * n + bt;
            }
    }
}

// backward partials: S1 = sum delta, S2 = sum delta*xhat, S3 = sum (x - mean)
__global__ void __launch_bounds__(kT) bn_bwd_partial(const float* __restrict__ delta, const float* __restrict__ x,
                                                      const float* __restrict__ xhat, const float* __restrict__ mean,
                                                      float* __restrict__ partial, int B, int C, int HW, int S) {
    __shared__ float red[32];
    const int c = blockIdx.x, s = blockIdx.y;
    const Units U(B, HW);
    const int per = (U.total + S - 1) / S, u0 = per * s, u1 = min(U.total, u0 + per);
    const float mu = mean[c];
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int u = u0; u < u1; ++u) {
        const int b = u / U.nchunk, k = u - b * U.nchunk;
        const size_t base = ((size_t)b * C + c) * HW;
        const int i0 = k * kChunk + threadIdx.x, lim = min(HW, (k + 1) * kChunk);
        float d[8], h[8], v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = i0 + j * kT < lim;
            d[j] = ok ? delta[base + i0 + j * kT] : 0.f;
            h[j] = ok ? xhat[base + i0 + j * kT] : 0.f;
            v[j] = ok ? x[base + i0 + j * kT] : mu;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a1 += d[j];
            a2 = fmaf(d[j], h[j], a2);
            a3 += v[j] - mu;
        }
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

__global__ void __launch_bounds__(kT) bn_bwd_apply(float* __restrict__ delta, const float* __restrict__ x,
                                                    const float* __restrict__ coef, int C, int HW, int nchunk, int planes) {
    const float4* cf = reinterpret_cast<const float4*>(coef);
    for (int blk = blockIdx.x; blk < planes * nchunk; blk += gridDim.x) {
        const int pl = blk / nchunk, k = blk - pl * nchunk;
        const float4 kf = cf[pl % C];
        const size_t base = (size_t)pl * HW;
        const int i0 = k * kChunk + threadIdx.x, lim = min(HW, (k + 1) * kChunk);
        float d[8], v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool ok = i0 + j * kT < lim;
            d[j] = ok ? delta[base + i0 + j * kT] : 0.f;
            v[j] = ok ? x[base + i0 + j * kT] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j * kT < lim) delta[base + i0 + j * kT] = d[j] * kf.x + kf.y * (v[j] - kf.w) + kf.z;
    }
}

inline int pick_splits(const cnn_ctx* ctx, int C, int B, int HW) {
    int S = cdiv(ctx->sm_count * 8, C);
    const int units = Units(B, HW).total;
    if (S > units) S = units;
    if (S > 128) S = 128;
    if (S < 1) S = 1;
    return S;
}

inline int ew_grid(const cnn_ctx* ctx, int planes, int HW) {
    const long long g = (long long)planes * Units(1, HW).nchunk;
    const long long cap = (long long)ctx->sm_count * 16;
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
    const int S = pick_splits(ctx, C, B, HW);
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
        CNN_LAUNCH(ctx, bn_partial<0>, grid, kT, 0, x, nullptr, partial, B, C, HW, S);
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, partial, sums, nullptr, nullptr, nullptr, C, S, 1.f, momentum);
        if (int rc = cnn_dist_allreduce_sum(ctx, sums, (size_t)C)) return rc;
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, sums, batch_mean, nullptr, nullptr, nullptr, C, 1, invNg, momentum);
        CNN_LAUNCH(ctx, bn_partial<1>, grid, kT, 0, x, batch_mean, partial, B, C, HW, S);
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, partial, sums, nullptr, nullptr, nullptr, C, S, 1.f, momentum);
        if (int rc = cnn_dist_allreduce_sum(ctx, sums, (size_t)C)) return rc;
        CNN_LAUNCH(ctx, bn_finalize<1>, cdiv(C, 128), 128, 0, sums, batch_var, batch_mean, moving_mean, moving_var, C, 1,
                   invNg, momentum);
    } else {
        CNN_LAUNCH(ctx, bn_partial<0>, grid, kT, 0, x, nullptr, partial, B, C, HW, S);
        CNN_LAUNCH(ctx, bn_finalize<0>, cdiv(C, 128), 128, 0, partial, batch_mean, nullptr, nullptr, nullptr,
                   C, S, invN, momentum);
        CNN_LAUNCH(ctx, bn_partial<1>, grid, kT, 0, x, batch_mean, partial, B, C, HW, S);
        CNN_LAUNCH(ctx, bn_finalize<1>, cdiv(C, 128), 128, 0, partial, batch_var, batch_mean, moving_mean,
                   moving_var, C, S, invN, momentum);
    }
    CNN_LAUNCH(ctx, bn_normalize, ew_grid(ctx, B * C, HW), kT, 0, x, gamma, beta, batch_mean, batch_var, xhat,
               y, C, HW, eps, Units(1, HW).nchunk, B * C);
    (void)total;
    return CNN_OK;
}

int cnn_bn_forward_eval(cnn_ctx* ctx, const float* x, const float* gamma, const float* beta,
                        const float* moving_mean, const float* moving_var, float* xhat, float* y, int B,
                        int C, int H, int W, float eps) {
    CNN_REQUIRE(ctx && x && gamma && beta && moving_mean && moving_var && xhat && y,
                "cnn_bn_forward_eval: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cnn_bn_forward_eval: bad shape");
    CNN_LAUNCH(ctx, bn_normalize, ew_grid(ctx, B * C, H * W), kT, 0, x, gamma, beta, moving_mean, moving_var,
               xhat, y, C, H * W, eps, Units(1, H * W).nchunk, B * C);
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
    const int S = pick_splits(ctx, C, B, HW);
    float* partial = cnn_scratch(ctx, sizeof(float) * ((size_t)3 * C * S + 4 * (size_t)C + 4 + 3 * (size_t)C));
    CNN_REQUIRE(partial, "scratch allocation failed");
    // keep coef 16-byte aligned for the float4 loads
    float* coef = partial + (((size_t)3 * C * S + 3) / 4) * 4;
    float* sums = coef + 4 * (size_t)C;
    dim3 grid(C, S);
    CNN_LAUNCH(ctx, bn_bwd_partial, grid, kT, 0, delta, x, xhat, batch_mean, partial, B, C, HW, S);
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
    CNN_LAUNCH(ctx, bn_bwd_apply, ew_grid(ctx, B * C, HW), kT, 0, delta, x, coef, C, HW, Units(1, HW).nchunk, B * C);
    (void)total;
    return CNN_OK;
}

}  // extern "C"
