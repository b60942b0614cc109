// elementwise.cu -- HBM-bound operators: ReLU, MaxPool, softmax-cross-entropy, SGD.
// All are streaming kernels: 128-bit accesses where alignment allows, grid sized in
// multiples of the SM count with a grid-stride loop.
#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;

inline int stream_grid(const cnn_ctx* ctx, size_t work_items) {
    // 8 resident CTAs of 256 threads per SM saturate HBM for pure streaming kernels
    long long want = (long long)((work_items + kThreads - 1) / kThreads);
    long long cap = (long long)ctx->sm_count * 8;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// ---- ReLU ------------------------------------------------------------------
// relu.cpp:25: y = x >= 0 ? x : 0  (so -0.0 passes through, NaN -> 0)
__device__ __forceinline__ float relu1(float v) { return v >= 0.f ? v : 0.f; }

__global__ void relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n4,
                                size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (size_t j = i; j < n4; j += stride) {
        float4 v = __ldcs(x4 + j);
        v.x = relu1(v.x); v.y = relu1(v.y); v.z = relu1(v.z); v.w = relu1(v.w);
        y4[j] = v;
    }
    for (size_t j = n4 * 4 + i; j < n; j += stride) y[j] = relu1(x[j]);
}

// relu.cpp:39: delta = (y <= 0) ? 0 : delta, in place
__global__ void relu_bwd_kernel(float* __restrict__ d, const float* __restrict__ y, size_t n4,
                                size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4* d4 = reinterpret_cast<float4*>(d);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    for (size_t j = i; j < n4; j += stride) {
        float4 g = d4[j];
        const float4 o = __ldcs(y4 + j);
        g.x = o.x <= 0.f ? 0.f : g.x; g.y = o.y <= 0.f ? 0.f : g.y;
        g.z = o.z <= 0.f ? 0.f : g.z; g.w = o.w <= 0.f ? 0.f : g.w;
        d4[j] = g;
    }
    for (size_t j = n4 * 4 + i; j < n; j += stride) d[j] = y[j] <= 0.f ? 0.f : d[j];
}

// ---- SGD ---------------------------------------------------------------------
// w -= lr * g with the product rounded before the subtraction, exactly like the
// reference's two fp32 operations (conv2d.cpp:212): no FMA contraction.
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, float lr) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        p[j] = __fsub_rn(p[j], __fmul_rn(lr, g[j]));
}

// the same step with the learning rate read from device memory: a captured step graph then serves every learning rate
// (an LR schedule would otherwise re-capture the whole step for every new value)
__global__ void sgd_dev_lr_kernel(float* __restrict__ p, const float* __restrict__ g, size_t n, const float* __restrict__ lr_ptr) {
    const float lr = *lr_ptr;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
        p[j] = __fsub_rn(p[j], __fmul_rn(lr, g[j]));
}
__global__ void set_scalar_kernel(float* dst, float v) { *dst = v; }

// ---- optimizer extensions (the reference's TODO item 2, cnn.cpp:15-24: "momentum, Adam") --------------------
// Same slab layout as sgd_kernel; state slabs are caller-owned (zero-initialised).  Classic heavy-ball momentum
// v = mu*v + g; p -= lr*v, and Adam (Kingma & Ba) with bias correction; plain fp32, one pass, 20 / 28 B per parameter.
__global__ void sgd_momentum_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, size_t n,
                                    float lr, float mu) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const float vv = __fadd_rn(__fmul_rn(mu, v[j]), g[j]);
        v[j] = vv;
        p[j] = __fsub_rn(p[j], __fmul_rn(lr, vv));
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float b1, float b2, float eps, float c1, float c2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const float gj = g[j];
        const float mj = b1 * m[j] + (1.f - b1) * gj;
        const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
        m[j] = mj;
        v[j] = vj;
        p[j] -= lr * (mj * c1) / (sqrtf(vj * c2) + eps);   // c1 = 1/(1-b1^t), c2 = 1/(1-b2^t)
    }
}

// ---- AvgPool / global average pool (the reference's TODO item 7, cnn.cpp:15-24) -------------------------------
// Same window geometry as MaxPool2D (OH = (H-k)/step + 1, trailing rows / columns dropped); k = H = W is the global pool.
// Backward adds delta / k^2 to every cell of the window (overlapping windows accumulate: gather form, no atomics).
__global__ void avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int OH, int OW, int k,
                                   int step, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int ox = (int)(i % OW);
        const size_t t = i / OW;
        const int oy = (int)(t % OH);
        const float* p = x + ((t / OH) * H + (size_t)oy * step) * W + (size_t)ox * step;
        float s = 0.f;
        for (int a = 0; a < k; ++a)
            for (int b = 0; b < k; ++b) s += p[(size_t)a * W + b];
        y[i] = s / (float)(k * k);
    }
}
__global__ void avgpool_bwd_kernel(const float* __restrict__ delta, float* __restrict__ dx, int H, int W, int OH, int OW, int k,
                                   int step, size_t total) {
    const float inv = 1.f / (float)(k * k);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W);
        const size_t t = i / W;
        const int y = (int)(t % H);
        const float* d = delta + (t / H) * (size_t)OH * OW;
        float s = 0.f;
        // windows (oy, ox) with oy*step <= y < oy*step + k
        const int oy1 = min(OH - 1, y / step), ox1 = min(OW - 1, xx / step);
        for (int oy = oy1; oy >= 0 && oy * step + k > y; --oy)
            for (int ox = ox1; ox >= 0 && ox * step + k > xx; --ox) s += d[(size_t)oy * OW + ox];
        dx[i] = s * inv;
    }
}

// ---- zero padding (item 8 of the reference's TODO list, cnn.cpp:15-24) ------------------------------------
// A padded convolution = this layer in front of the unpadded one, so every fast convolution path applies.
// One thread per element of the larger (padded) tensor; rows are contiguous on both sides.
__global__ void pad2d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int pad, size_t total) {
    const int PH = H + 2 * pad, PW = W + 2 * pad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int px = (int)(i % PW);
        const size_t t = i / PW;
        const int py = (int)(t % PH);
        const int xx = px - pad, yy = py - pad;
        y[i] = ((unsigned)xx < (unsigned)W && (unsigned)yy < (unsigned)H) ? x[((t / PH) * H + yy) * W + xx] : 0.f;
    }
}
__global__ void pad2d_bwd_kernel(const float* __restrict__ delta, float* __restrict__ dx, int H, int W, int pad, size_t total) {
    const int PH = H + 2 * pad, PW = W + 2 * pad;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W);
        const size_t t = i / W;
        const int yy = (int)(t % H);
        dx[i] = delta[((t / H) * PH + yy + pad) * PW + xx + pad];
    }
}

// ---- MaxPool -------------------------------------------------------------------
// One thread per output element, ox fastest (coalesced row reads).  Scan order and the
// strict '<' follow pool2d.cpp:67-75: the first element seeds the maximum, a later
// element replaces it only if strictly greater, NaN never replaces.
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                   int32_t* __restrict__ mask, int C, int H, int W, int OH, int OW,
                                   int k, int step, int planes) {
    // blockIdx.y strides over the (b, c) planes, threads over the pixels of one output plane: a
    // single integer division per output instead of three (the index math was the bottleneck)
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= OH * OW) return;
    const int oy = p / OW, ox = p - oy * OW;
    const int r0 = oy * step, c0 = ox * step;
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        const int c = pl % C;
        const float* plane = x + (size_t)pl * H * W;
        float mv = plane[r0 * W + c0];
        int mi = 0;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                if (i == 0 && j == 0) continue;
                const float v = plane[(r0 + i) * W + c0 + j];
                if (mv < v) { mv = v; mi = i * W + j; }
            }
        const size_t o = (size_t)pl * OH * OW + p;
        y[o] = mv;
        if (mask) mask[o] = c * H * W + mi + r0 * W + c0;
    }
}

// Gather form of pool2d.cpp:96-107 (zero, then dx[mask[i]] = delta[i] for ascending i):
// each input cell looks at the windows that cover it, in DESCENDING output order, and takes
// the first whose mask points at it -- i.e. the last writer of the reference's loop.  Cells
// outside every window, or never a maximum, get 0.  Fully coalesced writes, no memset.
__global__ void maxpool_bwd_kernel(const float* __restrict__ delta, const int32_t* __restrict__ mask,
                                   float* __restrict__ dx, int C, int H, int W, int OH, int OW, int k,
                                   int step, size_t total) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int col = (int)(idx % W);
        size_t t = idx / W;
        const int row = (int)(t % H);
        t /= H;  // b*C + c
        const int c = (int)(t % C);
        const int self = c * H * W + row * W + col;
        int oy_hi = row / step, ox_hi = col / step;
        if (oy_hi > OH - 1) oy_hi = OH - 1;
        if (ox_hi > OW - 1) ox_hi = OW - 1;
        int oy_lo = (row - k + step) / step;  // ceil((row-k+1)/step) for row-k+1 >= 0
        int ox_lo = (col - k + step) / step;
        if (row - k + 1 <= 0) oy_lo = 0;
        if (col - k + 1 <= 0) ox_lo = 0;
        const size_t obase = t * (size_t)OH * OW;
        float g = 0.f;
        bool found = false;
        for (int oy = oy_hi; oy >= oy_lo && !found; --oy)
            for (int ox = ox_hi; ox >= ox_lo; --ox) {
                const size_t o = obase + (size_t)oy * OW + ox;
                if (mask[o] == self) { g = delta[o]; found = true; break; }
            }
        dx[idx] = g;
    }
}

// Non-overlapping windows (step >= k): the step x step blocks anchored at (oy*step, ox*step),
// oy < ceil(H/step), tile the input plane exactly, so one thread owns one block: it writes
// delta at the recorded arg-max cell and 0 everywhere else (cells in the gap between windows
// and in trailing rows/cols included).  mask and delta are read once, coalesced.
__global__ void maxpool_bwd_tiled_kernel(const float* __restrict__ delta, const int32_t* __restrict__ mask,
                                         float* __restrict__ dx, int H, int W, int OH, int OW, int step,
                                         int BH, int BW, int planes) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= BH * BW) return;
    const int by = q / BW, bx = q - by * BW;
    const int r0 = by * step, c0 = bx * step;
    const bool has = by < OH && bx < OW;
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        int target = -1;
        float g = 0.f;
        if (has) {
            const size_t o = (size_t)pl * OH * OW + (size_t)by * OW + bx;
            target = mask[o] % (H * W);  // position inside this plane
            g = delta[o];
        }
        float* plane = dx + (size_t)pl * H * W;
        for (int i = 0; i < step && r0 + i < H; ++i)
            for (int j = 0; j < step && c0 + j < W; ++j) {
                const int pos = (r0 + i) * W + c0 + j;
                plane[pos] = (pos == target) ? g : 0.f;
            }
    }
}

// ReLU + MaxPool in one pass (step >= k): thread per step x step block as in the tiled backward.
// Every cell of the block gets its ReLU output; the k x k window inside the block is reduced with the
// reference's scan order / strict '<' on the ReLU'd values.
__global__ void relu_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ yr,
                                        float* __restrict__ yp, int32_t* __restrict__ mask, int C, int H,
                                        int W, int OH, int OW, int k, int step, int BH, int BW, int planes) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= BH * BW) return;
    const int by = q / BW, bx = q - by * BW;
    const int r0 = by * step, c0 = bx * step;
    const bool has = by < OH && bx < OW;
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        const float* src = x + (size_t)pl * H * W;
        float* dst = yr + (size_t)pl * H * W;
        float mv = 0.f;
        int mi = 0;
        for (int i = 0; i < step && r0 + i < H; ++i)
            for (int j = 0; j < step && c0 + j < W; ++j) {
                const int pos = (r0 + i) * W + c0 + j;
                const float v = relu1(src[pos]);
                dst[pos] = v;
                if (i < k && j < k) {
                    if (i == 0 && j == 0) { mv = v; mi = 0; }
                    else if (mv < v) { mv = v; mi = i * W + j; }
                }
            }
        if (has) {
            const size_t o = (size_t)pl * OH * OW + (size_t)by * OW + bx;
            yp[o] = mv;
            if (mask) mask[o] = (pl % C) * H * W + mi + r0 * W + c0;
        }
    }
}

__global__ void maxpool_relu_bwd_kernel(const float* __restrict__ delta, const int32_t* __restrict__ mask,
                                        const float* __restrict__ pool_out, float* __restrict__ dx, int H,
                                        int W, int OH, int OW, int step, int BH, int BW, int planes) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= BH * BW) return;
    const int by = q / BW, bx = q - by * BW;
    const int r0 = by * step, c0 = bx * step;
    const bool has = by < OH && bx < OW;
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        int target = -1;
        float g = 0.f;
        if (has) {
            const size_t o = (size_t)pl * OH * OW + (size_t)by * OW + bx;
            target = mask[o] % (H * W);
            g = pool_out[o] <= 0.f ? 0.f : delta[o];   // relu.cpp:39 on the arg-max cell
        }
        float* plane = dx + (size_t)pl * H * W;
        for (int i = 0; i < step && r0 + i < H; ++i)
            for (int j = 0; j < step && c0 + j < W; ++j) {
                const int pos = (r0 + i) * W + c0 + j;
                plane[pos] = (pos == target) ? g : 0.f;
            }
    }
}

// ---- 2x2 / step-2 pooling on row bands (the reference's MaxPool2D default, architectures.h:97) ----
// One warp owns one band = two consecutive input rows of one plane = 2*W CONTIGUOUS floats.  With odd
// W (111 in AlexNet-lite) neither rows nor planes are 8-byte aligned, and a thread-per-window kernel
// touches every 32-byte sector with two half-used instructions.  Here every global access is a fully
// coalesced 128-byte warp access: the band goes through shared memory, the 2x2 windows are reduced
// from there (scan order and strict '<' of pool2d.cpp:67-75 on the ReLU'd values).
// NW = windows per lane (ceil(OW / 32)); a band has at most 256 floats (8 slots per lane).  All loops are
// unrolled over compile-time slots with one base pointer per array (immediate offsets), the band ->
// (plane, row pair) walk has no divisions: the kernels are bound by memory, not by address arithmetic.
template <int NW>
__global__ void __launch_bounds__(256) relu_maxpool2_fwd_band_kernel(const float* __restrict__ x, float* __restrict__ yr,
                                                                     float* __restrict__ yp, int32_t* __restrict__ mask,
                                                                     int C, int H, int W, int OH, int OW, int bpp,
                                                                     int bands) {
    extern __shared__ float band_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = band_smem + warp * 256;
    const int stride = gridDim.x * 8, spl = stride / bpp, sby = stride % bpp;
    const int HW = H * W, OHW = OH * OW;
    int band = blockIdx.x * 8 + warp;
    int pl = band / bpp, by = band % bpp;
    int cpl = pl % C;                       // channel of the plane, advanced with the walk
    const int scpl = spl % C;
    for (; band < bands; band += stride) {
        const int r0 = 2 * by, n = (r0 + 1 < H ? 2 : 1) * W;
        const size_t base = (size_t)pl * HW + (size_t)r0 * W + lane;
        const float* xs = x + base;
        float* ys = yr + base;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (lane + 32 * j < n) ? __ldg(xs + 32 * j) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float r = relu1(v[j]);
            if (lane + 32 * j < n) ys[32 * j] = r;
            buf[lane + 32 * j] = r;
        }
        __syncwarp();
        if (by < OH) {
            const int cbase = cpl * HW + r0 * W;
            const size_t o = (size_t)pl * OHW + (size_t)by * OW + lane;
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const int ox = lane + 32 * k;
                if (ox < OW) {
                    const int c0 = 2 * ox;
                    const float2 top = *reinterpret_cast<const float2*>(buf + c0);
                    const float v10 = buf[W + c0], v11 = buf[W + c0 + 1];
                    float mv = top.x;
                    int mi = 0;
                    if (mv < top.y) { mv = top.y; mi = 1; }
                    if (mv < v10) { mv = v10; mi = W; }
                    if (mv < v11) { mv = v11; mi = W + 1; }
                    yp[o + 32 * k] = mv;
                    if (mask) mask[o + 32 * k] = cbase + mi + c0;
                }
            }
        }
        __syncwarp();
        by += sby; pl += spl; cpl += scpl;
        if (by >= bpp) { by -= bpp; ++pl; ++cpl; }
        while (cpl >= C) cpl -= C;
    }
}

// backward of the same pooling (pool2d.cpp:96-107), optionally with the ReLU backward of the layer
// below folded in (pool_out = ReLU output at the arg-max cell, relu.cpp:39): the band is zeroed in
// shared memory, the <= OW gradients are scattered there, one coalesced store pass writes it out.
// The (mask, delta, pool) loads of the next band are in flight while the current one is stored.
template <int NW>
__global__ void __launch_bounds__(256) maxpool2_bwd_band_kernel(const float* __restrict__ delta,
                                                                const int32_t* __restrict__ mask,
                                                                const float* __restrict__ pool_out, float* __restrict__ dx,
                                                                int C, int H, int W, int OH, int OW, int bpp, int bands) {
    extern __shared__ float band_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = band_smem + warp * 256;
    const int HW = H * W, OHW = OH * OW;
    const int stride = gridDim.x * 8, spl = stride / bpp, sby = stride % bpp, scpl = spl % C;
    int t_n[NW];
    float g_n[NW];
    auto fetch = [&](bool live, int pl, int by, int cpl) {
        // mask = c*H*W + position in the plane (pool2d.cpp:81): position relative to this band
        const int off = cpl * HW + 2 * by * W;
        const size_t o = (size_t)pl * OHW + (size_t)by * OW + lane;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            t_n[k] = -1;
            g_n[k] = 0.f;
            if (live && by < OH && lane + 32 * k < OW) {
                t_n[k] = __ldg(mask + o + 32 * k) - off;
                float g = __ldg(delta + o + 32 * k);
                if (pool_out && __ldg(pool_out + o + 32 * k) <= 0.f) g = 0.f;
                g_n[k] = g;
            }
        }
    };
    int band = blockIdx.x * 8 + warp;
    int pl = band / bpp, by = band % bpp, cpl = pl % C;
    fetch(band < bands, pl, by, cpl);
    for (; band < bands; band += stride) {
        const int r0 = 2 * by, n = (r0 + 1 < H ? 2 : 1) * W;
        float* out = dx + (size_t)pl * HW + (size_t)r0 * W + lane;
        int t_c[NW];
        float g_c[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) { t_c[k] = t_n[k]; g_c[k] = g_n[k]; }
        int nby = by + sby, npl = pl + spl, ncpl = cpl + scpl;
        if (nby >= bpp) { nby -= bpp; ++npl; ++ncpl; }
        while (ncpl >= C) ncpl -= C;
        fetch(band + stride < bands, npl, nby, ncpl);
        *reinterpret_cast<float4*>(buf + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(buf + 128 + 4 * lane) = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NW; ++k)
            if ((unsigned)t_c[k] < (unsigned)n) buf[t_c[k]] = g_c[k];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lane + 32 * j < n) out[32 * j] = buf[lane + 32 * j];
        __syncwarp();
        by = nby; pl = npl; cpl = ncpl;
    }
}

// ---- softmax + cross entropy + argmax ---------------------------------------------
// One thread per row, loops in the reference's order (func.cpp:24-33, :62-69) so that with
// identical logits only expf/logf ulps can differ.  Row terms are then added in ascending b
// by one thread, reproducing the reference's sequential fp32 loss accumulation.
__device__ __forceinline__ float clamped_exp(float v) {  // func.cpp:7-11
    if (v >= 88.f) return FLT_MAX;
    if (v <= -50.f) return 0.f;
    return expf(v);
}

__global__ void softmax_xent_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                                    float* __restrict__ probs, float* __restrict__ delta,
                                    float* __restrict__ loss_sum, int32_t* __restrict__ pred, int B,
                                    int n) {
    extern __shared__ float row_term[];  // B floats
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const float* z = logits + (size_t)b * n;
        float* p = probs + (size_t)b * n;
        float mv = z[0];
        for (int i = 1; i < n; ++i)
            if (z[i] > mv) mv = z[i];  // Tensor3D::max via argmax, strict '>'
        float sum = 0.f;
        for (int i = 0; i < n; ++i) {
            const float e = clamped_exp(z[i] - mv);
            p[i] = e;
            sum += e;
        }
        float best = 0.f;
        int besti = 0;
        for (int i = 0; i < n; ++i) {
            float v = p[i] / sum;
            if (isnan(v)) v = 0.f;
            p[i] = v;
            if (i == 0 || v > best) { best = v; besti = i; }  // data_format.cpp:37-48, first max
        }
        if (pred) pred[b] = besti;
        if (labels) {
            const int lab = labels[b];
            float term = 0.f;
            for (int i = 0; i < n; ++i) {
                const float yv = (i == lab) ? 1.f : 0.f;
                delta[(size_t)b * n + i] = p[i] - yv;
                term += logf(p[i]) * yv;  // 0 * log(0) = NaN, as in the reference
            }
            row_term[b] = term;
        }
    }
    if (!labels) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        float loss = 0.f;
        for (int b = 0; b < B; ++b) loss += row_term[b];
        *loss_sum = loss;
    }
}

// func.cpp:56-73 on probabilities + one-hot rows; one thread per row, rows summed in order
__global__ void xent_backward_kernel(const float* __restrict__ probs, const float* __restrict__ onehot,
                                     float* __restrict__ delta, float* __restrict__ loss_sum, int B, int n) {
    extern __shared__ float row_term[];
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float term = 0.f;
        for (int i = 0; i < n; ++i) {
            const float p = probs[(size_t)b * n + i], yv = onehot[(size_t)b * n + i];
            delta[(size_t)b * n + i] = p - yv;
            term += logf(p) * yv;
        }
        row_term[b] = term;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float loss = 0.f;
        for (int b = 0; b < B; ++b) loss += row_term[b];
        *loss_sum = loss;
    }
}

}  // namespace

namespace {
// ---- Tensor3D::read_from_opencv_mat for a batch (data_format.cpp:13-23) ------------------------
// src [B][H][W][C] bytes -> dst [B][C][H][W] floats, dst = src * 1.f / 255 (IEEE division, so the
// floats are the reference's bit for bit).  C == 3 fast path: a thread converts 4 pixels = three
// 32-bit loads -> one float4 store per colour plane.
__global__ void u8hwc_to_chw3_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t quads,
                                     int plane4) {   // plane4 = H*W/4
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += stride) {
        const size_t b = q / (size_t)plane4;
        const int i4 = (int)(q - b * (size_t)plane4);
        const uint32_t* s = reinterpret_cast<const uint32_t*>(src) + q * 3;
        const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);   // bytes p0c0 p0c1 p0c2 p1c0 | ...
        uint8_t by[12];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            by[j] = (uint8_t)(w0 >> (8 * j));
            by[4 + j] = (uint8_t)(w1 >> (8 * j));
            by[8 + j] = (uint8_t)(w2 >> (8 * j));
        }
        float4* d = reinterpret_cast<float4*>(dst) + b * 3 * (size_t)plane4 + i4;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float4 v;
            v.x = __fdiv_rn((float)by[c] * 1.f, 255.f);
            v.y = __fdiv_rn((float)by[3 + c] * 1.f, 255.f);
            v.z = __fdiv_rn((float)by[6 + c] * 1.f, 255.f);
            v.w = __fdiv_rn((float)by[9 + c] * 1.f, 255.f);
            d[(size_t)c * plane4] = v;
        }
    }
}

__global__ void u8hwc_to_chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t total, int C,
                                    int HW) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int i = (int)(idx % HW);
        const size_t t = idx / HW;
        const int c = (int)(t % C);
        const size_t b = t / C;
        dst[idx] = __fdiv_rn((float)src[(b * HW + i) * C + c] * 1.f, 255.f);
    }
}
}  // namespace

extern "C" {

int cnn_xent_backward(cnn_ctx* ctx, const float* probs, const float* onehot, float* delta, float* loss_sum,
                      int B, int classes) {
    CNN_REQUIRE(ctx && probs && onehot && delta && loss_sum, "cnn_xent_backward: NULL argument");
    CNN_REQUIRE(B > 0 && classes > 0 && (size_t)B * sizeof(float) <= 48 * 1024, "cnn_xent_backward: bad shape");
    const int threads = B < 1024 ? ((B + 31) / 32) * 32 : 1024;
    CNN_LAUNCH(ctx, xent_backward_kernel, 1, threads, (size_t)B * sizeof(float), probs, onehot, delta, loss_sum,
               B, classes);
    return CNN_OK;
}

int cnn_u8hwc_to_chw(cnn_ctx* ctx, const uint8_t* src, float* dst, int B, int C, int H, int W) {
    CNN_REQUIRE(ctx && src && dst, "cnn_u8hwc_to_chw: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "cnn_u8hwc_to_chw: bad shape");
    const int HW = H * W;
    if (C == 3 && HW % 4 == 0 && (((uintptr_t)src & 3) == 0) && (((uintptr_t)dst & 15) == 0)) {
        const size_t quads = (size_t)B * (HW / 4);
        CNN_LAUNCH(ctx, u8hwc_to_chw3_kernel, stream_grid(ctx, quads), kThreads, 0, src, dst, quads, HW / 4);
    } else {
        const size_t total = (size_t)B * C * HW;
        CNN_LAUNCH(ctx, u8hwc_to_chw_kernel, stream_grid(ctx, total), kThreads, 0, src, dst, total, C, HW);
    }
    return CNN_OK;
}

int cnn_relu_forward(cnn_ctx* ctx, const float* x, float* y, size_t n) {
    CNN_REQUIRE(ctx && x && y, "cnn_relu_forward: NULL argument");
    if (n == 0) return CNN_OK;
    const bool aligned = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
    const size_t n4 = aligned ? n / 4 : 0;
    CNN_LAUNCH(ctx, relu_fwd_kernel, stream_grid(ctx, n / 4 + 1), kThreads, 0, x, y, n4, n);
    return CNN_OK;
}

int cnn_relu_backward(cnn_ctx* ctx, float* delta, const float* y, size_t n) {
    CNN_REQUIRE(ctx && delta && y, "cnn_relu_backward: NULL argument");
    if (n == 0) return CNN_OK;
    const bool aligned = (((uintptr_t)delta | (uintptr_t)y) & 15) == 0;
    const size_t n4 = aligned ? n / 4 : 0;
    CNN_LAUNCH(ctx, relu_bwd_kernel, stream_grid(ctx, n / 4 + 1), kThreads, 0, delta, y, n4, n);
    return CNN_OK;
}

int cnn_sgd_step(cnn_ctx* ctx, float* params, const float* grads, size_t n, float lr) {
    CNN_REQUIRE(ctx && params && grads, "cnn_sgd_step: NULL argument");
    if (n == 0) return CNN_OK;
    CNN_LAUNCH(ctx, sgd_kernel, stream_grid(ctx, n), kThreads, 0, params, grads, n, lr);
    return CNN_OK;
}

}  // extern "C"

int cnn_sgd_step_dev_lr(cnn_ctx* ctx, float* params, const float* grads, size_t n, const float* lr_dev) {
    if (n == 0) return CNN_OK;
    CNN_LAUNCH(ctx, sgd_dev_lr_kernel, stream_grid(ctx, n), kThreads, 0, params, grads, n, lr_dev);
    return CNN_OK;
}

int cnn_set_scalar(cnn_ctx* ctx, float* dst, float v) {
    CNN_LAUNCH(ctx, set_scalar_kernel, 1, 1, 0, dst, v);
    return CNN_OK;
}

extern "C" {

int cnn_avgpool_forward(cnn_ctx* ctx, const float* x, float* y, int B, int C, int H, int W, int k, int step) {
    CNN_REQUIRE(ctx && x && y, "cnn_avgpool_forward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && k > 0 && step > 0 && H >= k && W >= k, "cnn_avgpool_forward: bad shape");
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const size_t total = (size_t)B * C * OH * OW;
    CNN_LAUNCH(ctx, avgpool_fwd_kernel, stream_grid(ctx, total), kThreads, 0, x, y, H, W, OH, OW, k, step, total);
    return CNN_OK;
}

int cnn_avgpool_backward(cnn_ctx* ctx, const float* delta, float* dx, int B, int C, int H, int W, int k, int step) {
    CNN_REQUIRE(ctx && delta && dx, "cnn_avgpool_backward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && k > 0 && step > 0 && H >= k && W >= k, "cnn_avgpool_backward: bad shape");
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const size_t total = (size_t)B * C * H * W;
    CNN_LAUNCH(ctx, avgpool_bwd_kernel, stream_grid(ctx, total), kThreads, 0, delta, dx, H, W, OH, OW, k, step, total);
    return CNN_OK;
}

int cnn_pad2d_forward(cnn_ctx* ctx, const float* x, float* y, int B, int C, int H, int W, int pad) {
    CNN_REQUIRE(ctx && x && y, "cnn_pad2d_forward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && pad >= 0, "cnn_pad2d_forward: bad shape");
    const size_t total = (size_t)B * C * (H + 2 * pad) * (W + 2 * pad);
    CNN_LAUNCH(ctx, pad2d_fwd_kernel, stream_grid(ctx, total), kThreads, 0, x, y, H, W, pad, total);
    return CNN_OK;
}

int cnn_pad2d_backward(cnn_ctx* ctx, const float* delta, float* dx, int B, int C, int H, int W, int pad) {
    CNN_REQUIRE(ctx && delta && dx, "cnn_pad2d_backward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && pad >= 0, "cnn_pad2d_backward: bad shape");
    const size_t total = (size_t)B * C * H * W;
    CNN_LAUNCH(ctx, pad2d_bwd_kernel, stream_grid(ctx, total), kThreads, 0, delta, dx, H, W, pad, total);
    return CNN_OK;
}

int cnn_sgd_momentum_step(cnn_ctx* ctx, float* params, const float* grads, float* velocity, size_t n, float lr, float momentum) {
    CNN_REQUIRE(ctx && params && grads && velocity, "cnn_sgd_momentum_step: NULL argument");
    if (n == 0) return CNN_OK;
    CNN_LAUNCH(ctx, sgd_momentum_kernel, stream_grid(ctx, n), kThreads, 0, params, grads, velocity, n, lr, momentum);
    return CNN_OK;
}

int cnn_adam_step(cnn_ctx* ctx, float* params, const float* grads, float* m, float* v, size_t n, float lr, float beta1,
                  float beta2, float eps, int t) {
    CNN_REQUIRE(ctx && params && grads && m && v && t >= 1, "cnn_adam_step: bad argument");
    if (n == 0) return CNN_OK;
    const float c1 = (float)(1.0 / (1.0 - pow((double)beta1, t))), c2 = (float)(1.0 / (1.0 - pow((double)beta2, t)));
    CNN_LAUNCH(ctx, adam_kernel, stream_grid(ctx, n), kThreads, 0, params, grads, m, v, n, lr, beta1, beta2, eps, c1, c2);
    return CNN_OK;
}

int cnn_maxpool_forward(cnn_ctx* ctx, const float* x, float* y, int32_t* mask, int B, int C, int H,
                        int W, int k, int step) {
    CNN_REQUIRE(ctx && x && y, "cnn_maxpool_forward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && k > 0 && step > 0 && H >= k && W >= k, "cnn_maxpool_forward: bad shape");
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const int planes = B * C;
    dim3 grid(cdiv((long long)OH * OW, kThreads), planes < 65535 ? planes : 65535);
    CNN_LAUNCH(ctx, maxpool_fwd_kernel, grid, kThreads, 0, x, y, mask, C, H, W, OH, OW, k, step, planes);
    return CNN_OK;
}

int cnn_maxpool_backward(cnn_ctx* ctx, const float* delta, const int32_t* mask, float* dx, int B,
                         int C, int H, int W, int k, int step) {
    CNN_REQUIRE(ctx && delta && mask && dx, "cnn_maxpool_backward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && k > 0 && step > 0 && H >= k && W >= k, "cnn_maxpool_backward: bad shape");
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    if (k == 2 && step == 2 && W <= 128 && (long long)B * C * ((H + 1) / 2) < (1ll << 30)) {
        const int bpp = (H + 1) / 2;
        const long long bands = (long long)B * C * bpp;
        const int g = (int)std::min<long long>((bands + 7) / 8, (long long)ctx->sm_count * 8);
        const size_t sm = 8 * 256 * sizeof(float);
        const float* none = nullptr;
        switch ((OW + 31) / 32) {
            case 1: CNN_LAUNCH(ctx, maxpool2_bwd_band_kernel<1>, g, 256, sm, delta, mask, none, dx, C, H, W, OH, OW, bpp, (int)bands); break;
            case 2: CNN_LAUNCH(ctx, maxpool2_bwd_band_kernel<2>, g, 256, sm, delta, mask, none, dx, C, H, W, OH, OW, bpp, (int)bands); break;
            default: CNN_LAUNCH(ctx, maxpool2_bwd_band_kernel<4>, g, 256, sm, delta, mask, none, dx, C, H, W, OH, OW, bpp, (int)bands); break;
        }
        return CNN_OK;
    }
    if (step >= k) {
        const int BH = (H + step - 1) / step, BW = (W + step - 1) / step;
        const int planes = B * C;
        dim3 grid(cdiv((long long)BH * BW, kThreads), planes < 65535 ? planes : 65535);
        CNN_LAUNCH(ctx, maxpool_bwd_tiled_kernel, grid, kThreads, 0, delta, mask, dx, H, W, OH, OW, step, BH, BW,
                   planes);
        return CNN_OK;
    }
    const size_t total = (size_t)B * C * H * W;
    CNN_LAUNCH(ctx, maxpool_bwd_kernel, stream_grid(ctx, total), kThreads, 0, delta, mask, dx, C, H, W,
               OH, OW, k, step, total);
    return CNN_OK;
}

int cnn_relu_maxpool_forward(cnn_ctx* ctx, const float* x, float* y_relu, float* y_pool, int32_t* mask, int B,
                             int C, int H, int W, int k, int step) {
    CNN_REQUIRE(ctx && x && y_relu && y_pool, "cnn_relu_maxpool_forward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && k > 0 && step >= k && H >= k && W >= k, "cnn_relu_maxpool_forward: needs step >= k");
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const int BH = (H + step - 1) / step, BW = (W + step - 1) / step, planes = B * C;
    if (k == 2 && step == 2 && W <= 128 && (long long)planes * ((H + 1) / 2) < (1ll << 30)) {
        const int bpp = (H + 1) / 2;
        const long long bands = (long long)planes * bpp;
        const int g = (int)std::min<long long>((bands + 7) / 8, (long long)ctx->sm_count * 8);
        const size_t sm = 8 * 256 * sizeof(float);
        switch ((OW + 31) / 32) {
            case 1: CNN_LAUNCH(ctx, relu_maxpool2_fwd_band_kernel<1>, g, 256, sm, x, y_relu, y_pool, mask, C, H, W, OH, OW, bpp, (int)bands); break;
            case 2: CNN_LAUNCH(ctx, relu_maxpool2_fwd_band_kernel<2>, g, 256, sm, x, y_relu, y_pool, mask, C, H, W, OH, OW, bpp, (int)bands); break;
            default: CNN_LAUNCH(ctx, relu_maxpool2_fwd_band_kernel<4>, g, 256, sm, x, y_relu, y_pool, mask, C, H, W, OH, OW, bpp, (int)bands); break;
        }
        return CNN_OK;
    }
    dim3 grid(cdiv((long long)BH * BW, kThreads), planes < 65535 ? planes : 65535);
    CNN_LAUNCH(ctx, relu_maxpool_fwd_kernel, grid, kThreads, 0, x, y_relu, y_pool, mask, C, H, W, OH, OW, k, step,
               BH, BW, planes);
    return CNN_OK;
}

int cnn_maxpool_relu_backward(cnn_ctx* ctx, const float* delta, const int32_t* mask, const float* pool_out,
                              float* dx, int B, int C, int H, int W, int k, int step) {
    CNN_REQUIRE(ctx && delta && mask && pool_out && dx, "cnn_maxpool_relu_backward: NULL argument");
    CNN_REQUIRE(B > 0 && C > 0 && k > 0 && step >= k && H >= k && W >= k, "cnn_maxpool_relu_backward: needs step >= k");
    const int OH = (H - k) / step + 1, OW = (W - k) / step + 1;
    const int BH = (H + step - 1) / step, BW = (W + step - 1) / step, planes = B * C;
    if (k == 2 && step == 2 && W <= 128 && (long long)B * C * ((H + 1) / 2) < (1ll << 30)) {
        const int bpp = (H + 1) / 2;
        const long long bands = (long long)planes * bpp;
        const int g = (int)std::min<long long>((bands + 7) / 8, (long long)ctx->sm_count * 8);
        const size_t sm = 8 * 256 * sizeof(float);
        switch ((OW + 31) / 32) {
            case 1: CNN_LAUNCH(ctx, maxpool2_bwd_band_kernel<1>, g, 256, sm, delta, mask, pool_out, dx, C, H, W, OH, OW, bpp, (int)bands); break;
            case 2: CNN_LAUNCH(ctx, maxpool2_bwd_band_kernel<2>, g, 256, sm, delta, mask, pool_out, dx, C, H, W, OH, OW, bpp, (int)bands); break;
            default: CNN_LAUNCH(ctx, maxpool2_bwd_band_kernel<4>, g, 256, sm, delta, mask, pool_out, dx, C, H, W, OH, OW, bpp, (int)bands); break;
        }
        return CNN_OK;
    }
    dim3 grid(cdiv((long long)BH * BW, kThreads), planes < 65535 ? planes : 65535);
    CNN_LAUNCH(ctx, maxpool_relu_bwd_kernel, grid, kThreads, 0, delta, mask, pool_out, dx, H, W, OH, OW, step, BH,
               BW, planes);
    return CNN_OK;
}

int cnn_softmax_xent(cnn_ctx* ctx, const float* logits, const int32_t* labels, float* probs,
                     float* delta, float* loss_sum, int32_t* pred, int B, int classes) {
    CNN_REQUIRE(ctx && logits && probs, "cnn_softmax_xent: NULL argument");
    CNN_REQUIRE(B > 0 && classes > 0, "cnn_softmax_xent: bad shape");
    CNN_REQUIRE(!labels || (delta && loss_sum), "cnn_softmax_xent: labels need delta and loss_sum");
    CNN_REQUIRE((size_t)B * sizeof(float) <= 200 * 1024, "cnn_softmax_xent: batch too large");
    const size_t smem = (size_t)B * sizeof(float);
    if (smem > 48 * 1024)
        CNN_CUDA(cudaFuncSetAttribute(softmax_xent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    const int threads = B < 1024 ? ((B + 31) / 32) * 32 : 1024;
    CNN_LAUNCH(ctx, softmax_xent_kernel, 1, threads, smem, logits, labels, probs, delta, loss_sum, pred,
               B, classes);
    return CNN_OK;
}

}  // extern "C"
