// conv_simt.cu -- fp32 SIMT direct convolution (forward, weight gradient, input gradient).
//
// This is the general-shape path (any odd k, any stride, any channel count) and the
// numerical cross-check for the tcgen05 implicit-GEMM kernels in conv_tc.cu.  Register
// blocking over output channels, filter slices staged in shared memory and read back as
// broadcast float4, pixel-contiguous (coalesced) activation accesses in NCHW.
#include "common.cuh"

namespace {

constexpr int kPix = 128;  // pixels (threads) per block

template <int CO_BLK>
__device__ __forceinline__ void fma_row(float (&acc)[CO_BLK], float xv, const float* wrow) {
    const float4* wv = reinterpret_cast<const float4*>(wrow);
#pragma unroll
    for (int c4 = 0; c4 < CO_BLK / 4; ++c4) {
        const float4 q = wv[c4];
        acc[c4 * 4 + 0] = fmaf(xv, q.x, acc[c4 * 4 + 0]);
        acc[c4 * 4 + 1] = fmaf(xv, q.y, acc[c4 * 4 + 1]);
        acc[c4 * 4 + 2] = fmaf(xv, q.z, acc[c4 * 4 + 2]);
        acc[c4 * 4 + 3] = fmaf(xv, q.w, acc[c4 * 4 + 3]);
    }
}

// ------------------------------------------------------------------ forward
// y[b][o][oy][ox] = bias[o] + sum_i sum_{ky,kx} x[b][i][oy*s+ky][ox*s+kx] * w[o][i][ky][kx]
// (conv2d.cpp:69-92; accumulation i-major, bias last, as the reference does)
template <int CO_BLK, int KS>
__global__ void __launch_bounds__(kPix)
conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                float* __restrict__ y, int Cin, int H, int W, int Cout, int OH, int OW, int k_rt,
                int s) {
    constexpr int CI_CHUNK = 8;
    const int k = KS ? KS : k_rt;
    const int kk = k * k;
    extern __shared__ __align__(16) float ws[];  // [CI_CHUNK*kk][CO_BLK]
    const int b = blockIdx.z, co0 = blockIdx.y * CO_BLK;
    const int p = blockIdx.x * kPix + threadIdx.x;
    const bool valid = p < OH * OW;
    const int oy = valid ? p / OW : 0, ox = valid ? p % OW : 0;
    const size_t plane = (size_t)H * W;
    const float* base = x + (size_t)b * Cin * plane + (size_t)(oy * s) * W + ox * s;
    float acc[CO_BLK];
#pragma unroll
    for (int c = 0; c < CO_BLK; ++c) acc[c] = 0.f;

    for (int ci0 = 0; ci0 < Cin; ci0 += CI_CHUNK) {
        const int nci = min(CI_CHUNK, Cin - ci0);
        const int run = nci * kk;  // contiguous floats per filter in this slice
        __syncthreads();
        for (int t = threadIdx.x; t < CO_BLK * run; t += kPix) {
            const int co = t / run, r = t % run;
            ws[r * CO_BLK + co] =
                (co0 + co < Cout) ? w[((size_t)(co0 + co) * Cin + ci0) * kk + r] : 0.f;
        }
        __syncthreads();
        if (valid) {
            for (int ci = 0; ci < nci; ++ci) {
                const float* xp = base + (size_t)(ci0 + ci) * plane;
                const float* wp = ws + ci * kk * CO_BLK;
                if constexpr (KS > 0) {
#pragma unroll
                    for (int i = 0; i < KS; ++i)
#pragma unroll
                        for (int j = 0; j < KS; ++j)
                            fma_row<CO_BLK>(acc, xp[i * W + j], wp + (i * KS + j) * CO_BLK);
                } else {
                    for (int i = 0; i < k; ++i)
                        for (int j = 0; j < k; ++j)
                            fma_row<CO_BLK>(acc, xp[i * W + j], wp + (i * k + j) * CO_BLK);
                }
            }
        }
    }
    if (!valid) return;
    const size_t oplane = (size_t)OH * OW;
#pragma unroll
    for (int c = 0; c < CO_BLK; ++c)
        if (co0 + c < Cout) y[((size_t)b * Cout + co0 + c) * oplane + p] = acc[c] + bias[co0 + c];
}

// ------------------------------------------------------------- weight gradient
// dw[o][i][ky][kx] = scale * sum_b sum_{oy,ox} delta[b][o][oy][ox] * x[b][i][oy*s+ky][ox*s+kx]
// db[o] = scale * sum delta[b][o][:]     (conv2d.cpp:120-157)
// grid: x = slices of the flattened (b, pixel) range, y = output-channel blocks, z = Cin.
// Per-thread register tile [CO_BLK][9], block reduction, one atomicAdd per (o,i,tap) and
// slice (dw/db are zeroed by the caller).
template <int CO_BLK>
__global__ void __launch_bounds__(kPix)
conv_wgrad3_kernel(const float* __restrict__ x, const float* __restrict__ delta, float* __restrict__ dw,
                   float* __restrict__ db, int Cin, int H, int W, int Cout, int OH, int OW, int s,
                   size_t P, size_t per_slice, float scale) {
    __shared__ float red[kPix / 32][CO_BLK * 10];
    const int ci = blockIdx.z, co0 = blockIdx.y * CO_BLK;
    const size_t beg = per_slice * blockIdx.x, end = min(P, beg + per_slice);
    const int opl = OH * OW;
    const size_t plane = (size_t)H * W;
    float acc[CO_BLK][9];
    float dsum[CO_BLK];
#pragma unroll
    for (int c = 0; c < CO_BLK; ++c) {
        dsum[c] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[c][t] = 0.f;
    }
    for (size_t e = beg + threadIdx.x; e < end; e += kPix) {
        const int b = (int)(e / opl), p = (int)(e % opl);
        const int oy = p / OW, ox = p % OW;
        const float* xp = x + ((size_t)b * Cin + ci) * plane + (size_t)(oy * s) * W + ox * s;
        float xv[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) xv[i * 3 + j] = xp[i * W + j];
        const float* dp = delta + ((size_t)b * Cout + co0) * opl + p;
#pragma unroll
        for (int c = 0; c < CO_BLK; ++c) {
            const float d = (co0 + c < Cout) ? dp[(size_t)c * opl] : 0.f;
            dsum[c] += d;
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[c][t] = fmaf(d, xv[t], acc[c][t]);
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CO_BLK; ++c) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float v = warp_sum(acc[c][t]);
            if (lane == 0) red[wid][c * 10 + t] = v;
        }
        const float v = warp_sum(dsum[c]);
        if (lane == 0) red[wid][c * 10 + 9] = v;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CO_BLK * 10; t += kPix) {
        const int c = t / 10, tap = t % 10;
        if (co0 + c >= Cout) continue;
        float v = 0.f;
#pragma unroll
        for (int wgi = 0; wgi < kPix / 32; ++wgi) v += red[wgi][t];
        if (tap < 9) atomicAdd(&dw[((size_t)(co0 + c) * Cin + ci) * 9 + tap], v * scale);
        else if (ci == 0) atomicAdd(&db[co0 + c], v * scale);
    }
}

// generic k: grid z = Cin*k*k (one (i,tap) per block), register tile [CO_BLK]
template <int CO_BLK>
__global__ void __launch_bounds__(kPix)
conv_wgrad_generic_kernel(const float* __restrict__ x, const float* __restrict__ delta,
                          float* __restrict__ dw, float* __restrict__ db, int Cin, int H, int W,
                          int Cout, int OH, int OW, int k, int s, size_t P, size_t per_slice,
                          float scale) {
    __shared__ float red[kPix / 32][CO_BLK * 2];
    const int kk = k * k;
    const int ci = blockIdx.z / kk, tap = blockIdx.z % kk, ky = tap / k, kx = tap % k;
    const int co0 = blockIdx.y * CO_BLK;
    const size_t beg = per_slice * blockIdx.x, end = min(P, beg + per_slice);
    const int opl = OH * OW;
    const size_t plane = (size_t)H * W;
    float acc[CO_BLK], dsum[CO_BLK];
#pragma unroll
    for (int c = 0; c < CO_BLK; ++c) acc[c] = dsum[c] = 0.f;
    for (size_t e = beg + threadIdx.x; e < end; e += kPix) {
        const int b = (int)(e / opl), p = (int)(e % opl);
        const int oy = p / OW, ox = p % OW;
        const float xv = x[((size_t)b * Cin + ci) * plane + (size_t)(oy * s + ky) * W + ox * s + kx];
        const float* dp = delta + ((size_t)b * Cout + co0) * opl + p;
#pragma unroll
        for (int c = 0; c < CO_BLK; ++c) {
            const float d = (co0 + c < Cout) ? dp[(size_t)c * opl] : 0.f;
            dsum[c] += d;
            acc[c] = fmaf(d, xv, acc[c]);
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CO_BLK; ++c) {
        const float v = warp_sum(acc[c]), u = warp_sum(dsum[c]);
        if (lane == 0) { red[wid][c * 2] = v; red[wid][c * 2 + 1] = u; }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < CO_BLK * 2; t += kPix) {
        const int c = t / 2;
        if (co0 + c >= Cout) continue;
        float v = 0.f;
#pragma unroll
        for (int wgi = 0; wgi < kPix / 32; ++wgi) v += red[wgi][t];
        if ((t & 1) == 0) atomicAdd(&dw[((size_t)(co0 + c) * Cin + ci) * kk + tap], v * scale);
        else if (blockIdx.z == 0) atomicAdd(&db[co0 + c], v * scale);
    }
}

// --------------------------------------------------------------- input gradient
// Gather form of the scatter at conv2d.cpp:192.  One thread owns an S x S patch of the input
// plane: tap (ky,kx) of window (oy,ox) lands on patch cell (ky%S, kx%S) of patch
// (oy + ky/S, ox + kx/S), so each tap feeds exactly one cell -- no wasted MACs and no
// divergence for stride 2.  Cells no window covers keep 0 (last row/col for even H, k3 s2).
template <int S, int KS, int CI_BLK>
__global__ void __launch_bounds__(kPix)
conv_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ delta, float* __restrict__ dx,
                  int Cin, int H, int W, int Cout, int OH, int OW) {
    constexpr int KK = KS * KS, CO_CHUNK = 8;
    __shared__ __align__(16) float ws[CO_CHUNK * KK * CI_BLK];  // [co][tap][ci]
    const int b = blockIdx.z, ci0 = blockIdx.y * CI_BLK;
    const int PH = (H + S - 1) / S, PW = (W + S - 1) / S;
    const int q = blockIdx.x * kPix + threadIdx.x;
    const bool valid = q < PH * PW;
    const int py = valid ? q / PW : 0, px = valid ? q % PW : 0;
    const int opl = OH * OW;
    float acc[S * S][CI_BLK];
#pragma unroll
    for (int a = 0; a < S * S; ++a)
#pragma unroll
        for (int c = 0; c < CI_BLK; ++c) acc[a][c] = 0.f;

    for (int co0 = 0; co0 < Cout; co0 += CO_CHUNK) {
        const int nco = min(CO_CHUNK, Cout - co0);
        __syncthreads();
        for (int t = threadIdx.x; t < CO_CHUNK * KK * CI_BLK; t += kPix) {
            const int c = t % CI_BLK, tap = (t / CI_BLK) % KK, co = t / (CI_BLK * KK);
            ws[t] = (co < nco && ci0 + c < Cin)
                        ? w[((size_t)(co0 + co) * Cin + ci0 + c) * KK + tap] : 0.f;
        }
        __syncthreads();
        if (!valid) continue;
        for (int co = 0; co < nco; ++co) {
            const float* dp = delta + ((size_t)b * Cout + co0 + co) * opl;
#pragma unroll
            for (int ky = 0; ky < KS; ++ky) {
                const int oy = py - ky / S;
                if (oy < 0 || oy >= OH) continue;
#pragma unroll
                for (int kx = 0; kx < KS; ++kx) {
                    const int ox = px - kx / S;
                    if (ox < 0 || ox >= OW) continue;
                    const float d = dp[oy * OW + ox];
                    const float* wv = ws + (co * KK + ky * KS + kx) * CI_BLK;
#pragma unroll
                    for (int c = 0; c < CI_BLK; ++c)
                        acc[(ky % S) * S + (kx % S)][c] = fmaf(d, wv[c], acc[(ky % S) * S + (kx % S)][c]);
                }
            }
        }
    }
    if (!valid) return;
    const size_t plane = (size_t)H * W;
#pragma unroll
    for (int pr = 0; pr < S; ++pr)
#pragma unroll
        for (int pc = 0; pc < S; ++pc) {
            const int r = py * S + pr, c = px * S + pc;
            if (r >= H || c >= W) continue;
#pragma unroll
            for (int ci = 0; ci < CI_BLK; ++ci)
                if (ci0 + ci < Cin)
                    dx[((size_t)b * Cin + ci0 + ci) * plane + (size_t)r * W + c] = acc[pr * S + pc][ci];
        }
}

// any k / stride: one thread per input element
__global__ void conv_dgrad_naive_kernel(const float* __restrict__ w, const float* __restrict__ delta,
                                        float* __restrict__ dx, int Cin, int H, int W, int Cout, int OH,
                                        int OW, int k, int s, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int kk = k * k, opl = OH * OW;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int c = (int)(idx % W);
        size_t t = idx / W;
        const int r = (int)(t % H);
        t /= H;
        const int ci = (int)(t % Cin);
        const size_t b = t / Cin;
        float acc = 0.f;
        for (int ky = 0; ky < k; ++ky) {
            const int ry = r - ky;
            if (ry < 0 || ry % s) continue;
            const int oy = ry / s;
            if (oy >= OH) continue;
            for (int kx = 0; kx < k; ++kx) {
                const int rx = c - kx;
                if (rx < 0 || rx % s) continue;
                const int ox = rx / s;
                if (ox >= OW) continue;
                const float* dp = delta + (b * Cout) * opl + oy * OW + ox;
                const float* wp = w + (size_t)ci * kk + ky * k + kx;
                for (int co = 0; co < Cout; ++co)
                    acc = fmaf(dp[(size_t)co * opl], wp[(size_t)co * Cin * kk], acc);
            }
        }
        dx[idx] = acc;
    }
}

}  // namespace

int conv_fwd_simt(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B,
                  int Cin, int H, int W, int Cout, int k, int s) {
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1;
    constexpr int CO_BLK = 16;
    dim3 grid(cdiv(OH * OW, kPix), cdiv(Cout, CO_BLK), B);
    const size_t smem = sizeof(float) * 8 * k * k * CO_BLK;
    if (k == 3) {
        CNN_LAUNCH(ctx, (conv_fwd_kernel<CO_BLK, 3>), grid, kPix, smem, x, w, bias, y, Cin, H, W, Cout, OH,
                   OW, k, s);
    } else {
        CNN_REQUIRE(smem <= 48 * 1024, "conv forward: kernel size %d too large", k);
        CNN_LAUNCH(ctx, (conv_fwd_kernel<CO_BLK, 0>), grid, kPix, smem, x, w, bias, y, Cin, H, W, Cout, OH,
                   OW, k, s);
    }
    return CNN_OK;
}

int conv_wgrad_simt(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B,
                    int Cin, int H, int W, int Cout, int k, int s, float scale) {
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1;
    constexpr int CO_BLK = 8;
    const size_t P = (size_t)B * OH * OW;
    CNN_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * k * k, ctx->stream));
    CNN_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, ctx->stream));
    const int gy = cdiv(Cout, CO_BLK);
    const int gz = (k == 3) ? Cin : Cin * k * k;
    CNN_REQUIRE(gz <= 65535, "conv weight gradient: Cin*k*k too large for the SIMT path");
    long long slices = cdiv((long long)ctx->sm_count * 8, (long long)gy * gz);
    const long long max_slices = (long long)((P + 2047) / 2048);
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    const size_t per = (P + slices - 1) / slices;
    dim3 grid((unsigned)slices, gy, gz);
    if (k == 3) {
        CNN_LAUNCH(ctx, conv_wgrad3_kernel<CO_BLK>, grid, kPix, 0, x, delta, dw, db, Cin, H, W, Cout, OH, OW,
                   s, P, per, scale);
    } else {
        CNN_LAUNCH(ctx, conv_wgrad_generic_kernel<CO_BLK>, grid, kPix, 0, x, delta, dw, db, Cin, H, W, Cout,
                   OH, OW, k, s, P, per, scale);
    }
    return CNN_OK;
}

int conv_dgrad_simt(cnn_ctx* ctx, const float* w, const float* delta, float* dx, int B, int Cin, int H,
                    int W, int Cout, int k, int s) {
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1;
    if (k == 3 && (s == 1 || s == 2)) {
        const int PH = cdiv(H, s), PW = cdiv(W, s);
        if (s == 2 && Cin <= 4) {   // first layer (image gradient): no FMAs wasted on absent channels
            constexpr int CI_BLK = 4;
            dim3 grid(cdiv(PH * PW, kPix), cdiv(Cin, CI_BLK), B);
            CNN_LAUNCH(ctx, (conv_dgrad_kernel<2, 3, CI_BLK>), grid, kPix, 0, w, delta, dx, Cin, H, W, Cout,
                       OH, OW);
        } else if (s == 2) {
            constexpr int CI_BLK = 8;
            dim3 grid(cdiv(PH * PW, kPix), cdiv(Cin, CI_BLK), B);
            CNN_LAUNCH(ctx, (conv_dgrad_kernel<2, 3, CI_BLK>), grid, kPix, 0, w, delta, dx, Cin, H, W, Cout,
                       OH, OW);
        } else {
            constexpr int CI_BLK = 16;
            dim3 grid(cdiv(PH * PW, kPix), cdiv(Cin, CI_BLK), B);
            CNN_LAUNCH(ctx, (conv_dgrad_kernel<1, 3, CI_BLK>), grid, kPix, 0, w, delta, dx, Cin, H, W, Cout,
                       OH, OW);
        }
        return CNN_OK;
    }
    const size_t total = (size_t)B * Cin * H * W;
    long long g = (long long)((total + 255) / 256);
    if (g > (long long)ctx->sm_count * 16) g = (long long)ctx->sm_count * 16;
    CNN_LAUNCH(ctx, conv_dgrad_naive_kernel, (int)g, 256, 0, w, delta, dx, Cin, H, W, Cout, OH, OW, k, s,
               total);
    return CNN_OK;
}
