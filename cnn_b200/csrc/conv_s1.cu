// conv_s1.cu -- 3x3 stride-1 Conv2D (the VGG-style and ResNet-shaped configs of BASELINE.json: Conv2D ctor
// with stride = 1, architectures.h:69) forward, input gradient and weight gradient as shifted-window implicit
// GEMMs on tcgen05, for layers with Cin % 16 == 0 and Cout % 32 == 0.  Same idea as conv_s2.cu -- activations
// are re-laid once into 16-byte chunks of 8 channels (bf16 pieces), and a filter tap is then only a start
// address -- with what the deep / wide stride-1 layers need on top:
//
//   * ONE plane: a tensor laid out at the input pitch, position m = b*H*W + y*W + x; tap (ky, kx) reads
//     position m + ky*W + kx.  Operand runs are staged per (16-channel block, filter row ky): the three kx
//     taps of a row are start-address offsets into the same 264 staged positions, so the halo is 8
//     positions instead of two image rows.
//   * Two 128-row M tiles (256 consecutive positions) share every filter stage: the filters of a wide
//     layer (up to 512 x 512 x 9 x 6 bytes) cannot stay in shared memory, and re-streaming them for every
//     128 rows would need more L2 bandwidth than an SM gets.
//   * Chunked accumulation.  tcgen05.mma adds into its fp32 TMEM accumulator with truncation, a bias of
//     ~2^-26 per accumulate step that the batch-mean gradients amplify by two orders of magnitude
//     (conv_s2.cu); a 512-channel layer would put 288 steps into one accumulator.  Here the MMA
//     warp closes an accumulator after ONE stage (3 large hi*hi products, issued after the stage's small
//     correction products so that those are added while the accumulator is still small) and 16 epilogue
//     warps add the closed chunk into fp32 registers with round-to-nearest while the next chunk fills the
//     other TMEM buffer.  Reduction length per truncating chain: 3, for any channel count.  Two MMA warps take
//     turns (one per TMEM buffer), so the issue overhead of one stage runs under the MMAs of the other.
//   * The thin first layer of such nets (3 -> 16/32/64 channels) has its own CUDA-core kernels at the end of the
//     packing section: first_s1_fwd_kernel (conv + ReLU + the next layer's packed input) and
//     first_s1_wgrad_kernel.
//
// Numerics: forward = three bf16 pieces per operand (all product terms down to 2^-24), gradients = two
// pieces (hi*hi + hi*lo + lo*hi), fp32 accumulation as above.  CNN_TC_BF16X1 (BASELINE config 5, "bf16,
// accumulate fp32") issues the hi*hi products only and stages only the hi pieces.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kMT = 128;                 // rows per MMA == TMEM lanes
constexpr int kItemRows = 2 * kMT;       // positions per work item
constexpr int kStageRows = kItemRows + 8;   // + kx shifts 0..2, rounded to 8
constexpr int kEpiWarps = 16;
constexpr int kMmaWarps = 2;                       // each owns one of the two TMEM chunk buffers (every other stage)
constexpr int kS1Threads = (1 + kMmaWarps + kEpiWarps) * 32;   // warp 0 TMA, warps 1-2 MMA, warps 3-18 epilogue
constexpr int kS1Header = 256 + 2048;    // barriers + bias (<= 512 channels)
constexpr uint32_t kRunBytes = kStageRows * 16;

struct S1Geom {
    int B, H, W;           // pitch geometry: position m = b*PP + y*W + x
    int PP, G;             // positions per image; guard / tail positions of a run (>= 2W + 8, multiple of 8)
    long long NPOS, NPOSR; // B*PP ; rounded up to kItemRows
    long long RUN;         // positions per (piece, channel group) run: G + NPOSR + G
};

S1Geom make_geom1(int B, int H, int W) {
    S1Geom g{};
    g.B = B; g.H = H; g.W = W;
    g.PP = H * W;
    g.G = (2 * W + 8 + 7) / 8 * 8;
    g.NPOS = (long long)B * g.PP;
    g.NPOSR = (g.NPOS + kItemRows - 1) / kItemRows * kItemRows;
    g.RUN = g.G + g.NPOSR + g.G;
    return g;
}

// ------------------------------------------------------------------------------------ packing
// src[B][C][SH][SW] fp32 (SH <= H, SW <= W: a delta tensor sits in the top-left corner of the pitch
// geometry, everything else is zero) -> P[piece][C/8][RUN].  Thread = (run position, channel group).
// PIECES = 3: forward activations; 2: gradients.  relu_y (optional, same shape as src): the in-place ReLU
// backward of the layer above folded into the packing (relu.cpp:39: delta = y <= 0 ? 0 : delta).
// db_partial (optional): per-block channel sums of the packed values (bias gradient, conv2d.cpp:153-157).
template <int PIECES>
__global__ void __launch_bounds__(256) s1_pack_kernel(const float* __restrict__ src, const float* __restrict__ relu_y,
                                                       uint4* __restrict__ dst, float* __restrict__ db_partial, const S1Geom g,
                                                       int C, int SH, int SW) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = blockIdx.y, ncg = C >> 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (r < g.RUN) {
        const long long m = r - g.G;
        if (m >= 0 && m < g.NPOS) {
            const int b = (int)(m / g.PP);
            const int rem = (int)(m - (long long)b * g.PP);
            const int y = rem / g.W, x = rem - y * g.W;
            if (y < SH && x < SW) {
                const size_t plane = (size_t)SH * SW;
                const size_t o = ((size_t)b * C + (size_t)cg * 8) * plane + (size_t)y * SW + x;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = __ldg(src + o + (size_t)j * plane);
                    if (relu_y && __ldg(relu_y + o + (size_t)j * plane) <= 0.f) t = 0.f;
                    v[j] = t;
                }
            }
        }
        if (PIECES == 3) {
            uint4 hi, mid, lo;
            split8x3(v, hi, mid, lo);
            dst[(size_t)(0 * ncg + cg) * g.RUN + r] = hi;
            dst[(size_t)(1 * ncg + cg) * g.RUN + r] = mid;
            dst[(size_t)(2 * ncg + cg) * g.RUN + r] = lo;
        } else {
            uint4 hi, lo;
            split8(v, hi, lo);
            dst[(size_t)(0 * ncg + cg) * g.RUN + r] = hi;
            dst[(size_t)(1 * ncg + cg) * g.RUN + r] = lo;
        }
    }
    if (db_partial) {   // fixed-order block reduction: deterministic
        __shared__ float red[8][8];
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float s = warp_sum(v[j]);
            if (lane == 0) red[wid][j] = s;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
            db_partial[(size_t)blockIdx.x * C + cg * 8 + threadIdx.x] = s;
        }
    }
}

// ----------------------------------------------------------------------- thin first layer, stride 1
// Conv2D(3 -> COUT <= 64, 3x3, stride 1) + the ReLU that follows it, for nets whose first layer feeds a packed
// stride-1 layer (VGG-style: conv2d.cpp:69-92 + relu.cpp:21-28).  K = 27 is no tensor-core shape (see conv_thin.cu);
// the pass is bound by what it writes: conv output, ReLU output and -- so that the next layer needs no pack pass
// over them -- that layer's packed input P(relu) (three bf16 pieces), 896 B per pixel at 64 channels.
// Thread = one output pixel x all channels: 27 inputs in registers, filters [tap][co] in shared memory read as
// float4 (one LDS.128 per four FMAs), fp32 FMA chain per output in the reference's loop order (ci, ky, kx), + bias.
struct FirstS1 {
    const float* x;
    const float* w;       // [COUT][3][3][3]
    const float* bias;
    float* y;             // [B][COUT][OH][OW]
    float* y_relu;        // same shape, may be null
    uint4* next_px;       // P(relu output) in the pitch geometry (OH, OW) of the next layer's input, may be null
    long long next_run;   // positions per (piece, channel group) run of that buffer
    int next_g;           // its guard positions
    int B, H, W, OH, OW;
    long long npix;
};

template <int COUT>
__global__ void __launch_bounds__(256, 2) first_s1_fwd_kernel(const FirstS1 p) {
    __shared__ __align__(16) float w_s[27 * COUT];
    __shared__ float b_s[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += 256) {
        const int tap = i / COUT, co = i - tap * COUT;
        w_s[i] = p.w[co * 27 + tap];
    }
    for (int i = threadIdx.x; i < COUT; i += 256) b_s[i] = p.bias ? p.bias[i] : 0.f;
    __syncthreads();
    const long long m = (long long)blockIdx.x * 256 + threadIdx.x;
    if (m >= p.npix) return;
    const int opl = p.OH * p.OW;
    const int b = (int)(m / opl);
    const int rem = (int)(m - (long long)b * opl);
    const int oy = rem / p.OW, ox = rem - oy * p.OW;
    float xin[27];
    {
        const float* xb = p.x + ((size_t)b * 3 * p.H + oy) * p.W + ox;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) xin[(ci * 3 + ky) * 3 + kx] = __ldg(xb + ((size_t)ci * p.H + ky) * p.W + kx);
    }
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 27; ++tap) {
        const float4* w4 = reinterpret_cast<const float4*>(w_s + tap * COUT);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
            const float4 wv = w4[c4];
            acc[4 * c4 + 0] = fmaf(xin[tap], wv.x, acc[4 * c4 + 0]);
            acc[4 * c4 + 1] = fmaf(xin[tap], wv.y, acc[4 * c4 + 1]);
            acc[4 * c4 + 2] = fmaf(xin[tap], wv.z, acc[4 * c4 + 2]);
            acc[4 * c4 + 3] = fmaf(xin[tap], wv.w, acc[4 * c4 + 3]);
        }
    }
    const size_t o = (size_t)b * COUT * opl + (size_t)oy * p.OW + ox;
#pragma unroll
    for (int cg = 0; cg < COUT / 8; ++cg) {
        float q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int co = cg * 8 + j;
            const float r = acc[co] + b_s[co];
            p.y[o + (size_t)co * opl] = r;
            q[j] = r >= 0.f ? r : 0.f;                 // relu.cpp:25
            if (p.y_relu) p.y_relu[o + (size_t)co * opl] = q[j];
        }
        if (p.next_px) {
            uint4 hi, mid, lo;
            split8x3(q, hi, mid, lo);
            const size_t r0 = (size_t)p.next_g + (size_t)m;      // same position numbering: PP = OH * OW
            p.next_px[(size_t)(0 * (COUT / 8) + cg) * p.next_run + r0] = hi;
            p.next_px[(size_t)(1 * (COUT / 8) + cg) * p.next_run + r0] = mid;
            p.next_px[(size_t)(2 * (COUT / 8) + cg) * p.next_run + r0] = lo;
        }
    }
}

// Weight / bias gradient of the same layer (conv2d.cpp:108-159): dw[co][ci][ky][kx] = scale * sum over pixels of
// delta[b][co][y][x] * x[b][ci][y+ky][x+kx].  1728 FMAs per pixel at 64 channels and nothing to reuse across pixels but
// the 27 inputs: a register-tile kernel without shared memory.  A group of COUT/4 lanes walks a contiguous range of
// pixels; lane = four output channels, 4 x 27 accumulators in registers; per pixel it reads the 27 inputs (the same
// addresses for all lanes of the group: one broadcast transaction each) and its four deltas (consecutive pixels of a
// channel plane over the iterations: every 32-byte sector is used eight times).  The groups of a warp are added by
// shuffles, the warps of a block one after the other in shared memory, the blocks by first_s1_wgrad_reduce_kernel in
// block order: deterministic.
template <int COUT>
__global__ void __launch_bounds__(256, 1) first_s1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ delta,
                                                                 float* __restrict__ partial, int B, int H, int W, int OH,
                                                                 int OW, long long npix, long long per_group) {
    constexpr int LANES = COUT / 4;          // lanes per pixel group
    constexpr int GROUPS = 32 / LANES;       // pixel groups per warp
    constexpr int NOUT = COUT * 27 + COUT;   // dw then db
    __shared__ float red[NOUT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cog = lane % LANES, grp = lane / LANES;
    const long long group = ((long long)blockIdx.x * 8 + warp) * GROUPS + grp;
    long long m = group * per_group;
    const long long m_end = m + per_group < npix ? m + per_group : npix;
    float acc[4][27], dbs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        dbs[i] = 0.f;
#pragma unroll
        for (int t = 0; t < 27; ++t) acc[i][t] = 0.f;
    }
    const int opl = OH * OW;
    if (m < m_end) {
        int b = (int)(m / opl);
        int rem = (int)(m - (long long)b * opl);
        int oy = rem / OW, ox = rem - oy * OW;
        for (; m < m_end; ++m) {
            const float* xb = x + ((size_t)b * 3 * H + oy) * W + ox;
            const float* dp = delta + ((size_t)b * COUT + 4 * cog) * opl + (size_t)oy * OW + ox;
            float xin[27], d[4];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) xin[(ci * 3 + ky) * 3 + kx] = __ldg(xb + ((size_t)ci * H + ky) * W + kx);
#pragma unroll
            for (int i = 0; i < 4; ++i) d[i] = __ldg(dp + (size_t)i * opl);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dbs[i] += d[i];
#pragma unroll
                for (int t = 0; t < 27; ++t) acc[i][t] = fmaf(d[i], xin[t], acc[i][t]);
            }
            if (++ox == OW) {
                ox = 0;
                if (++oy == OH) { oy = 0; ++b; }
            }
        }
    }
    // pixel groups of the warp (fixed shuffle tree), then the warps of the block in warp order
#pragma unroll
    for (int off = LANES; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dbs[i] += __shfl_xor_sync(0xffffffffu, dbs[i], off);
#pragma unroll
            for (int t = 0; t < 27; ++t) acc[i][t] += __shfl_xor_sync(0xffffffffu, acc[i][t], off);
        }
    }
    for (int w = 0; w < 8; ++w) {
        if (warp == w && grp == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = 4 * cog + i;
#pragma unroll
                for (int t = 0; t < 27; ++t) red[co * 27 + t] = (w == 0 ? 0.f : red[co * 27 + t]) + acc[i][t];
                red[COUT * 27 + co] = (w == 0 ? 0.f : red[COUT * 27 + co]) + dbs[i];
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < NOUT; i += 256) partial[(size_t)blockIdx.x * NOUT + i] = red[i];
}

__global__ void __launch_bounds__(256) first_s1_wgrad_reduce_kernel(const float* __restrict__ partial, int nblocks, int cout,
                                                                    float* __restrict__ dw, float* __restrict__ db, float scale) {
    const int nout = cout * 27 + cout;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nout) return;
    float s = 0.f;
    for (int blk = 0; blk < nblocks; ++blk) s += partial[(size_t)blk * nout + i];
    if (i < cout * 27) dw[i] = s * scale;
    else db[i - cout * 27] = s * scale;
}

// filters -> per-stage operand blocks [nt][kc][ky][piece][kx][cgl = 2][n < Ntile] of 16-byte chunks (8 k values):
//   forward: n = co, k = ci, value W[co][ci][ky][kx], three pieces
//   input gradient: n = ci, k = co, value W[co][ci][ky][kx], three pieces
__global__ void __launch_bounds__(256) s1_pack_w_kernel(const float* __restrict__ w, uint4* __restrict__ out, int Cin, int Cout,
                                                         int dgrad, int Ntile) {
    const int N = dgrad ? Cin : Cout, KC = (dgrad ? Cout : Cin) >> 4, ntn = N / Ntile;
    const int pieces = 3;
    const long long total = (long long)ntn * KC * 3 * 3 * 2 * Ntile;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total; id += (long long)gridDim.x * blockDim.x) {
        const int nl = (int)(id % Ntile);
        long long t = id / Ntile;
        const int cgl = (int)(t & 1);
        t >>= 1;
        const int kx = (int)(t % 3);
        t /= 3;
        const int ky = (int)(t % 3);
        t /= 3;
        const int kc = (int)(t % KC), nt = (int)(t / KC);
        const int n = nt * Ntile + nl;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 16 + cgl * 8 + j;
            v[j] = dgrad ? w[((size_t)k * Cin + n) * 9 + ky * 3 + kx] : w[((size_t)n * Cin + k) * 9 + ky * 3 + kx];
        }
        uint4 pc[3];
        split8x3(v, pc[0], pc[1], pc[2]);
        const size_t blk = (((size_t)nt * KC + kc) * 3 + ky) * pieces;
        for (int h = 0; h < pieces; ++h) out[(((blk + h) * 3 + kx) * 2 + cgl) * Ntile + nl] = pc[h];
    }
}

// ----------------------------------------------------------------------------- forward / dgrad
struct S1Gemm {
    const uint4* act;     // forward: P(x), 3 pieces ; dgrad: dP, 2 pieces
    const uint4* wpk;     // packed filters
    const float* bias;    // forward only
    const float* relu_y;  // dgrad: ReLU output of the layer below (mask y <= 0 -> 0, relu.cpp:39) or null
    float* dst;           // forward: y [B][N][VH][VW] ; dgrad: dx [B][N][H][W]
    float* dst_relu;      // forward: optional ReLU output (relu.cpp:25)
    S1Geom g;
    int K, N;             // reduction channels, output channels
    int Ntile, ntn, KC;
    int VH, VW;           // valid output rows / columns (forward: OH, OW ; dgrad: H, W)
    int nstage, items;
    int single;           // CNN_TC_BF16X1: hi * hi products only
    int pieces;           // bf16 pieces multiplied per operand: 3 (every product term down to 2^-24), 2 (2^-16) or 1 (single pass)
    int dbg;              // CNN_DBG_S1 timing experiments (WRONG results): 1 = no MMAs, 2 = no TMEM drain, 4 = no operand loads
    uint32_t a_bytes, b_bytes;   // staged bytes per stage (all three pieces are laid out, `pieces` of them are loaded)
};

// One work item = 256 consecutive positions x Ntile channels; one stage = (16-channel block kc, filter row ky).
template <bool DGRAD>
__global__ void __launch_bounds__(kS1Threads, 1) s1_gemm_kernel(const S1Gemm p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // [8]
    uint64_t* empty = full + 8;                               // [8]
    uint64_t* acc_full = full + 16;                           // [2]
    uint64_t* acc_empty = full + 18;                          // [2]
    uint32_t* tslot = reinterpret_cast<uint32_t*>(full + 20);
    float* sbias = reinterpret_cast<float*>(smem + 256);
    uint8_t* stages = smem + kS1Header;
    constexpr int PIECES = 3;
    const S1Geom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    const int Ntile = p.Ntile;

    if (warp == 0) {
        tmem_alloc(tslot, 512u);
        if (lane == 0) {
            for (int i = 0; i < p.nstage; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&acc_full[i], 1);
                mbar_init(&acc_empty[i], kEpiWarps);
            }
            mbar_fence_init();
        }
    }
    if (!DGRAD)
        for (int i = tid; i < p.N; i += kS1Threads) sbias[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA: one run per lane + the filter block
        const int ncgK = p.K >> 3;
        uint32_t s = 0, ph = 0;
        bool wrapped = false;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int mt = item / p.ntn, nt = item - mt * p.ntn;
            const long long m0 = (long long)mt * kItemRows;
            for (int kc = 0; kc < p.KC; ++kc)
                for (int ky = 0; ky < 3; ++ky) {
                    if (wrapped) mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = stages + (size_t)s * stage_bytes;
                    if (p.dbg & 4) {
                        if (lane == 0) mbar_arrive(&full[s]);
                        if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; wrapped = true; }
                        continue;
                    }
                    if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)p.pieces * 2 * kRunBytes + (uint32_t)p.pieces * (p.b_bytes / 3));
                    __syncwarp();
                    if (lane < p.pieces * 2) {
                        const int h = lane >> 1, cgl = lane & 1;
                        const long long start = g.G + m0 + (DGRAD ? -(long long)ky * g.W - 2 : (long long)ky * g.W);
                        tma_bulk_g2s(st + (size_t)lane * kRunBytes, p.act + ((size_t)h * ncgK + (size_t)kc * 2 + cgl) * g.RUN + start,
                                     kRunBytes, &full[s]);
                    } else if (lane == PIECES * 2) {   // filter block: pieces are contiguous, the first `pieces` are loaded
                        tma_bulk_g2s(st + p.a_bytes,
                                     reinterpret_cast<const uint8_t*>(p.wpk) + (((size_t)nt * p.KC + kc) * 3 + ky) * p.b_bytes,
                                     (uint32_t)p.pieces * (p.b_bytes / 3), &full[s]);
                    }
                    if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; wrapped = true; }
                }
        }
    } else if (warp <= kMmaWarps) {
        // ------------------------------------------------------------ MMA issue: one accumulator chunk per stage.
        // Two issuing warps, each owning one TMEM chunk buffer = every other stage.  What a warp does between the last
        // MMA of a stage and the first of its next one (two commits, two barrier waits, ~150 descriptor instructions:
        // about 1100 cycles per stage, measured with the loads and the TMEM drain switched off, CNN_DBG_S1=6) then runs
        // under the other warp's MMAs instead of leaving the tensor pipe idle -- the pipe queues only a few instructions
        // ahead of the issuing thread.  Forward pass of a VGG-style step: 11.06 -> 10.16 ms.
        const uint32_t idesc = idesc_bf16(kMT, Ntile);
        const uint32_t st0 = smem_u32(stages);
        const uint32_t b_lbo = (uint32_t)Ntile * 16;
        const uint32_t mine = (uint32_t)(warp - 1);
        uint32_t nchunk = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            for (int c = 0; c < p.KC * 3; ++c, ++nchunk) {
                if ((nchunk & 1) != mine) continue;
                const uint32_t buf = mine;
                const uint32_t s = nchunk % (uint32_t)p.nstage, ph = (nchunk / (uint32_t)p.nstage) & 1;
                if (nchunk >= 2) mbar_wait(&acc_empty[buf], ((nchunk >> 1) - 1) & 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = st0 + s * stage_bytes, sb = sa + p.a_bytes;
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        if (p.dbg & 1) break;
                        const uint32_t d = tmem + buf * (uint32_t)(2 * Ntile) + (uint32_t)(mi * Ntile);
                        uint64_t A[PIECES][3], Bd[PIECES][3];
#pragma unroll
                        for (int h = 0; h < PIECES; ++h)
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                const uint32_t aoff = (uint32_t)(mi * kMT + (DGRAD ? 2 - kx : kx)) * 16;
                                A[h][kx] = desc_nosw(sa + (uint32_t)h * 2 * kRunBytes + aoff, kRunBytes, 128);
                                Bd[h][kx] = desc_nosw(sb + (uint32_t)((h * 3 + kx) * 2) * b_lbo, b_lbo, 128);
                            }
                        bool acc = false;
                        if (!p.single) {   // the small correction products first: added while the accumulator is small
#pragma unroll
                            for (int kx = 0; kx < 3; ++kx) {
                                if (p.pieces == 3) {
                                    mma_bf16(d, A[2][kx], Bd[0][kx], idesc, acc); acc = true;
                                    mma_bf16(d, A[0][kx], Bd[2][kx], idesc, true);
                                    mma_bf16(d, A[1][kx], Bd[1][kx], idesc, true);
                                }
                                mma_bf16(d, A[1][kx], Bd[0][kx], idesc, acc); acc = true;
                                mma_bf16(d, A[0][kx], Bd[1][kx], idesc, true);
                            }
                        }
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) { mma_bf16(d, A[0][kx], Bd[0][kx], idesc, acc); acc = true; }
                    }
                    mma_commit(&empty[s]);
                    mma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: 16 warps, chunk sums in registers
        // warp -> TMEM lane group (warp & 3, fixed by the hardware) and slice (M tile, column half)
        const int e = warp - 1 - kMmaWarps, lg = warp & 3, slice = e >> 2;
        const int mi = slice >> 1, ncol = Ntile >> 1, c0 = (slice & 1) * ncol;
        const int row = lg * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(mi * Ntile + c0);
        uint32_t nchunk = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const int mt = item / p.ntn, nt = item - mt * p.ntn;
            // row = position m -> (b, y, x); only the valid VH x VW corner of the pitch geometry exists in dst
            const long long m = (long long)mt * kItemRows + mi * kMT + row;
            const size_t plane = (size_t)p.VH * p.VW;
            const int ch0 = nt * Ntile + c0;
            bool valid = m < g.NPOS;
            size_t o = 0;
            if (valid) {
                const int b = (int)(m / g.PP);
                const int rem = (int)(m - (long long)b * g.PP);
                const int y = rem / g.W, x = rem - y * g.W;
                valid = y < p.VH && x < p.VW;
                o = ((size_t)b * p.N + ch0) * plane + (size_t)y * p.VW + x;
            }
            // input gradient: the ReLU outputs that gate this row's columns (relu.cpp:39) are requested 8 at a time
            // while the first chunks accumulate -- at write-out time they are bits, not loads on the critical path
            // (with the loads there, every item stalled the chunk pipeline: 4.6 instead of 2.1 ms on a VGG-style layer)
            unsigned long long keep = ~0ull;
            const bool masked = DGRAD && p.relu_y != nullptr && valid;
            float racc[64];
            for (int c = 0; c < p.KC * 3; ++c, ++nchunk) {
                const uint32_t buf = nchunk & 1;
                float ry[8];
                const bool fetch = masked && c < 8 && c * 8 < ncol;
                if (fetch) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) ry[i] = __ldg(p.relu_y + o + (size_t)(c * 8 + i) * plane);
                }
                mbar_wait(&acc_full[buf], (nchunk >> 1) & 1);
                tc_fence_after();
                // 16 columns at a time (registers: 64 running sums + one group in flight)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j * 16 < ncol && !(p.dbg & 2)) {
                        uint32_t v[16];
                        tmem_ld16_async(tbase + buf * (uint32_t)(2 * Ntile) + j * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float t = __uint_as_float(v[i]);
                            racc[j * 16 + i] = c == 0 ? t : __fadd_rn(racc[j * 16 + i], t);
                        }
                    }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);   // chunk is in registers: the buffer can be refilled
                if (fetch) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (ry[i] <= 0.f) keep &= ~(1ull << (c * 8 + i));
                }
            }
            if (masked)   // fewer chunks than column groups (K < 48): the rest of the gates now
                for (int j = p.KC * 3 * 8; j < ncol; ++j)
                    if (__ldg(p.relu_y + o + (size_t)j * plane) <= 0.f) keep &= ~(1ull << j);
            if (valid) {
#pragma unroll
                for (int j = 0; j < 64; ++j)
                    if (j < ncol) {
                        if (!DGRAD) {
                            const float r = racc[j] + sbias[ch0 + j];
                            p.dst[o + (size_t)j * plane] = r;
                            if (p.dst_relu) p.dst_relu[o + (size_t)j * plane] = r >= 0.f ? r : 0.f;
                        } else {
                            p.dst[o + (size_t)j * plane] = ((keep >> j) & 1ull) ? racc[j] : 0.f;
                        }
                    }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512u);
}

// -------------------------------------------------------------------------------- weight gradient
// D[kx][co][ci] = sum_m dP[m][co] * P(x)[m + ky*W + kx][ci]  (conv2d.cpp:108-159; dP is zero wherever delta does
// not exist, so the sum may run over every position).  K = positions: both operands are MN-major views of the
// packed runs.  One work item = (128 output channels, Nci input channels, filter row ky): three accumulators
// (kx) in TMEM, the kx shift is a start-address offset on the x side (K rows are 16 bytes apart).  The position
// range of an item is split over several CTAs; a CTA walks its range in chunks of at most kAccPos positions --
// the reduction-length cap of a truncating TMEM accumulator chain -- and adds every closed chunk
// to ITS partial block in global memory (L2 resident, read-add-write by the same threads: deterministic);
// s1_wgrad_reduce_kernel sums the blocks of the CTAs of an item in a fixed order.
// Operands arrive by tensor-map TMA: one 4-D box [16 B][positions][16 channel groups][piece] per operand and
// piece; channel groups past the tensor read as zero (layers with fewer than 128 output channels).
constexpr int kWT = 64;                  // positions per stage
constexpr int kWTX = kWT + 8;            // x positions staged (+ kx shifts)
constexpr int kAccPos = 4096;            // positions per accumulator chain (1024 measured no closer to fp64: tools/fullstep_parity_vgg.py)
constexpr int kWgThreadsS1 = 6 * 32;     // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct S1Wgrad {
    float* partial;       // [item][split][3][128][Nci]
    S1Geom g;
    int Cin, Cout, Nci;
    int nct, nnt;         // output-channel tiles of 128, input-channel tiles of Nci
    int G, nky;           // Cout <= 64: G = 128 / Cout copies of delta, each shifted one image row further, fill the 128 MMA
                          // rows -- row group g of an item yields filter row kyA + g; nky = items along ky (3, 2 or 1)
    int nsplit;           // CTAs per item
    int stages_total;     // NPOSR / kWT
    int nstage;
    int single;           // CNN_TC_BF16X1: hi * hi products only
    int pieces;           // 3 or 2 pieces multiplied per operand
    uint32_t a_bytes, b_bytes;
};

__global__ void __launch_bounds__(kWgThreadsS1, 1) s1_wgrad_kernel(const __grid_constant__ CUtensorMap dmap,
                                                                    const __grid_constant__ CUtensorMap xmap, const S1Wgrad p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // [4]
    uint64_t* empty = full + 4;                               // [4]
    uint64_t* acc_full = full + 8;
    uint64_t* acc_empty = full + 9;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(full + 10);
    uint8_t* stages = smem + 128;
    const S1Geom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    const int Nci = p.Nci;
    // item = (ct, nt, ky block), split = position range
    const int item = blockIdx.x / p.nsplit, split = blockIdx.x - item * p.nsplit;
    const int ky = (item % p.nky) * p.G, nt = (item / p.nky) % p.nnt, ct = item / (p.nky * p.nnt);
    const int per = (p.stages_total + p.nsplit - 1) / p.nsplit;
    const int s_begin = split * per, s_end = min(p.stages_total, s_begin + per);
    const int n_it = max(0, s_end - s_begin);
    constexpr int kChunkStages = kAccPos / kWT;
    const int nchunks = (n_it + kChunkStages - 1) / kChunkStages;

    if (warp == 0) {
        tmem_alloc(tslot, 512u);
        if (lane == 0) {
            for (int i = 0; i < p.nstage; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            mbar_init(acc_full, 1);
            mbar_init(acc_empty, 4);
            mbar_fence_init();
            tma_prefetch_desc(&dmap);
            tma_prefetch_desc(&xmap);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA: one box per operand and piece
        if (lane == 0) {
            for (int it = 0; it < n_it; ++it) {
                const int s = it % p.nstage;
                if (it >= p.nstage) mbar_wait(&empty[s], ((it / p.nstage) - 1) & 1);
                uint8_t* st = stages + (size_t)s * stage_bytes;
                const long long k0 = (long long)(s_begin + it) * kWT;
                const int ncopy = p.G == 1 ? 1 : min(p.G, 3 - ky);          // row groups with a real filter row
                const uint32_t a_copy = p.a_bytes / 3 / (uint32_t)p.G;      // bytes of one row group of one piece
                mbar_expect_tx(&full[s], (uint32_t)p.pieces * ((uint32_t)ncopy * a_copy + p.b_bytes / 3));
                for (int h = 0; h < p.pieces; ++h) {
                    // row group c: delta shifted c image rows back pairs with x staged for filter row ky -> filter row ky + c
                    for (int c = 0; c < ncopy; ++c)
                        tma_tensor4d_g2s(st + (size_t)h * (p.a_bytes / 3) + (size_t)c * a_copy, &dmap, 0, (int)(g.G + k0 - (long long)c * g.W),
                                         ct * 16, h, &full[s]);
                    tma_tensor4d_g2s(st + p.a_bytes + (size_t)h * (p.b_bytes / 3), &xmap, 0, (int)(g.G + k0 + (long long)ky * g.W),
                                     nt * (Nci / 8), h, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issue
        const uint32_t idesc = idesc_bf16_mn(kMT, Nci);
        const uint32_t st0 = smem_u32(stages);
        const uint32_t a_sbo = kWT * 16, b_sbo = kWTX * 16;
        const uint32_t a_pc = p.a_bytes / 3, b_pc = p.b_bytes / 3;   // bytes between the pieces of an operand
        for (int it = 0; it < n_it; ++it) {
            const int s = it % p.nstage;
            const int cpos = it % kChunkStages;       // stage index inside the accumulator chunk
            if (cpos == 0 && it > 0) {                // previous chunk must have been drained
                mbar_wait(acc_empty, ((it / kChunkStages) - 1) & 1);
                tc_fence_after();
            }
            mbar_wait(&full[s], (it / p.nstage) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = st0 + (uint32_t)s * stage_bytes, sb = sa + p.a_bytes;
#pragma unroll
                for (int j = 0; j < kWT / 16; ++j) {
                    const uint32_t ao = sa + (uint32_t)j * 256;
                    const uint64_t a0 = desc_nosw(ao, 128, a_sbo), a1 = desc_nosw(ao + a_pc, 128, a_sbo),
                                   a2 = desc_nosw(ao + 2 * a_pc, 128, a_sbo);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint32_t bo = sb + (uint32_t)j * 256 + (uint32_t)kx * 16;
                        const uint64_t b0 = desc_nosw(bo, 128, b_sbo), b1 = desc_nosw(bo + b_pc, 128, b_sbo),
                                       b2 = desc_nosw(bo + 2 * b_pc, 128, b_sbo);
                        const uint32_t d = tmem + (uint32_t)(kx * Nci);
                        const bool acc = (cpos | j) != 0;
                        if (p.single) {
                            mma_bf16(d, a0, b0, idesc, acc);
                        } else if (p.pieces == 3) {   // every product term down to 2^-24
                            mma_bf16(d, a2, b0, idesc, acc);
                            mma_bf16(d, a0, b2, idesc, true);
                            mma_bf16(d, a1, b1, idesc, true);
                            mma_bf16(d, a1, b0, idesc, true);
                            mma_bf16(d, a0, b1, idesc, true);
                            mma_bf16(d, a0, b0, idesc, true);
                        } else {
                            mma_bf16(d, a1, b0, idesc, acc);
                            mma_bf16(d, a0, b1, idesc, true);
                            mma_bf16(d, a0, b0, idesc, true);
                        }
                    }
                }
                mma_commit(&empty[s]);
                if (cpos == kChunkStages - 1 || it == n_it - 1) mma_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------ epilogue: chunk -> this CTA's partial block
        const int lg = warp & 3, row = lg * 32 + lane;
        float* out = p.partial + ((size_t)blockIdx.x * 3 * kMT + row) * Nci;
        const uint32_t trow = tmem + ((uint32_t)(lg * 32) << 16);
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(acc_full, c & 1);
            tc_fence_after();
            for (int kx = 0; kx < 3; ++kx)
                for (int c0 = 0; c0 < Nci; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + kx * Nci + c0, v);
                    float4* q = reinterpret_cast<float4*>(out + (size_t)kx * kMT * Nci + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 t = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        if (c > 0) {
                            const float4 o = q[j];
                            t.x = __fadd_rn(t.x, o.x); t.y = __fadd_rn(t.y, o.y); t.z = __fadd_rn(t.z, o.z); t.w = __fadd_rn(t.w, o.w);
                        }
                        q[j] = t;
                    }
                }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
        if (nchunks == 0) {   // empty position range: the reduce kernel still reads this block
            for (int kx = 0; kx < 3; ++kx)
                for (int c0 = 0; c0 < Nci; ++c0) out[(size_t)kx * kMT * Nci + c0] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512u);
}

// dw[co][ci][ky][kx] = scale * sum over the splits of item (ct, nt, ky); db = scale * sum over pack-block partials.
// Block = 32 consecutive elements x 8 split lanes (coalesced 128-byte reads, eight sums in flight), fixed order.
__global__ void __launch_bounds__(256) s1_wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ db_partial,
                                                               int nblocks, float* __restrict__ dw, float* __restrict__ db, int Cin,
                                                               int Cout, int Nci, int nnt, int nsplit, int G, int nky, float scale) {
    __shared__ float red[8][33];
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const long long id = (long long)blockIdx.x * 32 + col;
    const long long nw = (long long)((Cout + kMT - 1) / kMT) * nky * kMT * Cin * 3;   // rows past Cout / filter rows past 2 are padding
    float s = 0.f;
    int co = -1, ci = 0, tap = 0;
    if (id < nw) {
        // id enumerates (ct, nt, ky block, kx, row, cil) with cil fastest; row = (row group g, channel) when G > 1
        const int cil = (int)(id % Nci);
        long long t = id / Nci;
        const int row = (int)(t % kMT);
        t /= kMT;
        const int kx = (int)(t % 3);
        t /= 3;
        const int kyb = (int)(t % nky);
        t /= nky;
        const int nt = (int)(t % nnt), ct = (int)(t / nnt);
        const int grp = G > 1 ? row / Cout : 0;
        const int ky = kyb * G + grp;
        co = G > 1 ? row - grp * Cout : ct * kMT + row;
        ci = nt * Nci + cil; tap = ky * 3 + kx;
        if (ky > 2 || grp >= G) co = -1;
        if (co >= 0 && co < Cout) {
            const int item = (ct * nnt + nt) * nky + kyb;
            const float* src = partial + (((size_t)item * nsplit * 3 + kx) * kMT + row) * Nci + cil;
            for (int sp = sl; sp < nsplit; sp += 8) s += src[(size_t)sp * 3 * kMT * Nci];
        }
    } else if (id - nw < Cout) {
        co = (int)(id - nw);
        for (int b = sl; b < nblocks; b += 8) s += db_partial[(size_t)b * Cout + co];
    }
    red[sl][col] = s;
    __syncthreads();
    if (sl == 0 && co >= 0 && co < Cout) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][col];
        if (id < nw) dw[((size_t)co * Cin + ci) * 9 + tap] = t * scale;
        else if (db) db[co] = t * scale;
    }
}

size_t align_up1(size_t v, size_t a) { return (v + a - 1) / a * a; }

int s1_attrs(int device) {
    std::lock_guard<std::recursive_mutex> lk(cnn_global_mutex());
    static bool done[16];
    if (device < 0 || device >= 16 || done[device]) return CNN_OK;
    CNN_CUDA(cudaFuncSetAttribute(s1_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CNN_CUDA(cudaFuncSetAttribute(s1_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CNN_CUDA(cudaFuncSetAttribute(s1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    done[device] = true;
    return CNN_OK;
}

// bf16 pieces multiplied per operand in the two gradient passes: two (hi*hi + hi*lo + lo*hi, 2^-16 per product).
// Measured on a full-size VGG-style step against an fp64 evaluation (tools/fullstep_parity_vgg.py): every gradient
// tensor comes out as close to fp64 with two pieces as with three (the error that remains is the forward pass's
// fp32-grade rounding seen through the cancellation of the batch-mean gradients), at half the MMAs.  The forward
// pass keeps three pieces.  CNN_S1_GRAD_PIECES=3 switches the gradients to three for such measurements.
int grad_pieces(const cnn_ctx* ctx) {
    (void)ctx;
    if (const char* e = getenv("CNN_S1_GRAD_PIECES")) return atoi(e) == 3 ? 3 : 2;
    return 2;
}

int pick_ntile(int N) {   // columns per work item: <= 128 and a multiple of 32 (two epilogue column halves of 16-column groups)
    for (int t = 128; t >= 32; t -= 32)
        if (N % t == 0) return t;
    return 0;
}

template <int PIECES>
int launch_pack(cnn_ctx* ctx, const S1Geom& g, const float* src, const float* relu_y, uint4* dst, float* dbp, int C, int SH, int SW) {
    dim3 grid((unsigned)cdiv(g.RUN, 256), (unsigned)(C / 8));
    CNN_LAUNCH(ctx, s1_pack_kernel<PIECES>, grid, 256, 0, src, relu_y, dst, dbp, g, C, SH, SW);
    return CNN_OK;
}

int launch_gemm1(cnn_ctx* ctx, const S1Geom& g, bool dgrad, const uint4* act, const uint4* wpk, const float* bias,
                 const float* relu_y, float* dst, float* dst_relu, int K, int N, int VH, int VW) {
    S1Gemm p{};
    p.act = act; p.wpk = wpk; p.bias = bias; p.relu_y = relu_y; p.dst = dst; p.dst_relu = dst_relu; p.g = g;
    p.K = K; p.N = N; p.VH = VH; p.VW = VW;
    p.Ntile = pick_ntile(N);
    CNN_REQUIRE(p.Ntile > 0 && K % 16 == 0 && N <= 512, "conv_s1: unsupported channel counts %d -> %d", K, N);
    p.ntn = N / p.Ntile;
    p.KC = K / 16;
    p.single = ctx->tc_precision == CNN_TC_BF16X1 ? 1 : 0;
    const int pieces = 3;
    p.pieces = p.single ? 1 : (dgrad ? grad_pieces(ctx) : 3);   // single pass: only the hi pieces are read
    if (const char* e = getenv("CNN_DBG_S1")) p.dbg = atoi(e);
    p.a_bytes = (uint32_t)pieces * 2 * kRunBytes;
    p.b_bytes = (uint32_t)pieces * 3 * 2 * p.Ntile * 16;
    const size_t stage = (size_t)p.a_bytes + p.b_bytes;
    p.nstage = (int)std::min<size_t>(8, (227 * 1024 - kS1Header) / stage);
    CNN_REQUIRE(p.nstage >= 2, "conv_s1: stage does not fit in shared memory");
    p.items = (int)(g.NPOSR / kItemRows) * p.ntn;
    const int grid = std::min(p.items, ctx->sm_count);
    const size_t smem = kS1Header + (size_t)p.nstage * stage;
    if (int rc = s1_attrs(ctx->device)) return rc;
    if (dgrad) { CNN_LAUNCH(ctx, s1_gemm_kernel<true>, grid, kS1Threads, smem, p); }
    else { CNN_LAUNCH(ctx, s1_gemm_kernel<false>, grid, kS1Threads, smem, p); }
    return CNN_OK;
}

size_t pk_bytes(const S1Geom& g, int C, int pieces) { return align_up1((size_t)pieces * (C / 8) * g.RUN * 16, 256); }
size_t wpk_bytes(int Cin, int Cout) { return align_up1((size_t)3 * 9 * Cin * Cout * 2, 256); }

}  // namespace

bool conv_s1_supported(const cnn_ctx* ctx, int Cin, int H, int W, int Cout, int k, int s) {
    (void)ctx;
    if (k != 3 || s != 1 || H < 3 || W < 3) return false;
    if (Cin % 32 || Cout % 32 || Cin > 512 || Cout > 512) return false;
    if ((long long)H * W > (1 << 24)) return false;
    return getenv("CNN_DBG_NOS1") == nullptr;
}

bool conv_s1_first_supported(const cnn_ctx* ctx, int Cin, int H, int W, int Cout, int k, int s) {
    (void)ctx;
    if (Cin != 3 || k != 3 || s != 1 || H < 3 || W < 3) return false;
    if (Cout != 16 && Cout != 32 && Cout != 64) return false;
    return getenv("CNN_DBG_NOS1FIRST") == nullptr;
}

// y = conv(x) + bias, y_relu = ReLU(y) (optional), next_px = P(y_relu) for a packed stride-1 layer that takes
// y_relu [B][Cout][OH][OW] as its input (optional; guard band and tail of that buffer must already be zero)
int conv_s1_first_fwd(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, float* y_relu, void* next_px,
                      int B, int H, int W, int Cout) {
    FirstS1 p{};
    p.x = x; p.w = w; p.bias = bias; p.y = y; p.y_relu = y_relu; p.next_px = static_cast<uint4*>(next_px);
    p.B = B; p.H = H; p.W = W; p.OH = H - 2; p.OW = W - 2;
    p.npix = (long long)B * p.OH * p.OW;
    if (next_px) {
        const S1Geom ng = make_geom1(B, p.OH, p.OW);
        p.next_run = ng.RUN;
        p.next_g = ng.G;
    }
    const unsigned grid = (unsigned)cdiv(p.npix, 256);
    switch (Cout) {
        case 16: CNN_LAUNCH(ctx, first_s1_fwd_kernel<16>, grid, 256, 0, p); break;
        case 32: CNN_LAUNCH(ctx, first_s1_fwd_kernel<32>, grid, 256, 0, p); break;
        case 64: CNN_LAUNCH(ctx, first_s1_fwd_kernel<64>, grid, 256, 0, p); break;
        default: CNN_REQUIRE(false, "conv_s1_first_fwd: unsupported channel count %d", Cout);
    }
    return CNN_OK;
}

// dw [Cout][3][3][3], db [Cout] of the thin stride-1 first layer; delta [B][Cout][H-2][W-2]
int conv_s1_first_wgrad(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B, int H, int W, int Cout,
                        float scale) {
    const int OH = H - 2, OW = W - 2;
    const long long npix = (long long)B * OH * OW;
    const int lanes = Cout / 4, groups_per_block = 8 * (32 / lanes);
    // one block per SM (4 x 27 accumulators per thread); every pixel group gets the same number of consecutive pixels
    int nblocks = ctx->sm_count;
    if ((long long)nblocks * groups_per_block > npix) nblocks = (int)cdiv(npix, groups_per_block);
    const long long per_group = cdiv(npix, (long long)nblocks * groups_per_block);
    const int nout = Cout * 27 + Cout;
    uint8_t* scratch = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, (size_t)nblocks * nout * sizeof(float) + 256));
    CNN_REQUIRE(scratch, "scratch allocation failed");
    float* partial = reinterpret_cast<float*>(align_up1((uintptr_t)scratch, 256));
    switch (Cout) {
        case 16: CNN_LAUNCH(ctx, first_s1_wgrad_kernel<16>, nblocks, 256, 0, x, delta, partial, B, H, W, OH, OW, npix, per_group); break;
        case 32: CNN_LAUNCH(ctx, first_s1_wgrad_kernel<32>, nblocks, 256, 0, x, delta, partial, B, H, W, OH, OW, npix, per_group); break;
        case 64: CNN_LAUNCH(ctx, first_s1_wgrad_kernel<64>, nblocks, 256, 0, x, delta, partial, B, H, W, OH, OW, npix, per_group); break;
        default: CNN_REQUIRE(false, "conv_s1_first_wgrad: unsupported channel count %d", Cout);
    }
    CNN_LAUNCH(ctx, first_s1_wgrad_reduce_kernel, cdiv(nout, 256), 256, 0, partial, nblocks, Cout, dw, db, scale);
    return CNN_OK;
}

// ---- packed-buffer interface (the engine keeps P(x) from the forward pass for the weight gradient and packs
// delta once for both gradients, with the ReLU backward of the layer above folded into the packing) -----------
size_t conv_s1_pk_bytes(int B, int C, int H, int W) { return pk_bytes(make_geom1(B, H, W), C, 3); }
size_t conv_s1_dbp_bytes(int B, int C, int H, int W) {
    return align_up1((size_t)cdiv(make_geom1(B, H, W).RUN, 256) * C * sizeof(float), 256);
}

int conv_s1_pack(cnn_ctx* ctx, const float* src, const float* relu_y, void* dst, float* dbp, int B, int C, int H, int W,
                 int SH, int SW) {
    return launch_pack<3>(ctx, make_geom1(B, H, W), src, relu_y, reinterpret_cast<uint4*>(dst), dbp, C, SH, SW);
}

namespace {
int pack_filters(cnn_ctx* ctx, const float* w, int Cin, int Cout, int dgrad, uint4** out) {
    uint8_t* scratch = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, wpk_bytes(Cin, Cout) + 256));
    CNN_REQUIRE(scratch, "scratch allocation failed");
    uint4* wpk = reinterpret_cast<uint4*>(align_up1((uintptr_t)scratch, 256));
    const int Ntile = pick_ntile(dgrad ? Cin : Cout);
    CNN_REQUIRE(Ntile > 0, "conv_s1: channel counts must be multiples of 32");
    CNN_LAUNCH(ctx, s1_pack_w_kernel, std::min(cdiv((long long)Cin * Cout * 9 / 8, 256), 1024), 256, 0, w, wpk, Cin, Cout, dgrad, Ntile);
    *out = wpk;
    return CNN_OK;
}
}  // namespace

int conv_s1_fwd_packed(cnn_ctx* ctx, const void* px, const float* w, const float* bias, float* y, float* y_relu, int B,
                       int Cin, int H, int W, int Cout) {
    uint4* wpk = nullptr;
    if (int rc = pack_filters(ctx, w, Cin, Cout, 0, &wpk)) return rc;
    return launch_gemm1(ctx, make_geom1(B, H, W), false, reinterpret_cast<const uint4*>(px), wpk, bias, nullptr, y, y_relu, Cin, Cout,
                        H - 2, W - 2);
}

int conv_s1_dgrad_packed(cnn_ctx* ctx, const void* pd, const float* w, float* dx, const float* relu_y, int B, int Cin, int H,
                         int W, int Cout) {
    uint4* wpk = nullptr;
    if (int rc = pack_filters(ctx, w, Cin, Cout, 1, &wpk)) return rc;
    return launch_gemm1(ctx, make_geom1(B, H, W), true, reinterpret_cast<const uint4*>(pd), wpk, nullptr, relu_y, dx, nullptr, Cout, Cin,
                        H, W);
}

int conv_fwd_s1(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, float* y_relu, int B, int Cin,
                int H, int W, int Cout) {
    void* px = cnn_arena(ctx, conv_s1_pk_bytes(B, Cin, H, W));
    CNN_REQUIRE(px, "conv_s1: arena allocation failed");
    if (int rc = conv_s1_pack(ctx, x, nullptr, px, nullptr, B, Cin, H, W, H, W)) return rc;
    return conv_s1_fwd_packed(ctx, px, w, bias, y, y_relu, B, Cin, H, W, Cout);
}

int conv_dgrad_s1(cnn_ctx* ctx, const float* w, const float* delta, float* dx, const float* relu_y, int B, int Cin, int H,
                  int W, int Cout) {
    void* pd = cnn_arena(ctx, conv_s1_pk_bytes(B, Cout, H, W));
    CNN_REQUIRE(pd, "conv_s1: arena allocation failed");
    if (int rc = conv_s1_pack(ctx, delta, nullptr, pd, nullptr, B, Cout, H, W, H - 2, W - 2)) return rc;
    return conv_s1_dgrad_packed(ctx, pd, w, dx, relu_y, B, Cin, H, W, Cout);
}

int conv_s1_wgrad_packed(cnn_ctx* ctx, const void* px, const void* pd, const float* dbp, float* dw, float* db, int B, int Cin,
                         int H, int W, int Cout, float scale) {
    const S1Geom g = make_geom1(B, H, W);
    CNN_REQUIRE(g.RUN < (1ll << 31), "conv_s1: tensor too large for the weight-gradient tensor maps");
    const unsigned pack_blocks = (unsigned)cdiv(g.RUN, 256);
    S1Wgrad p{};
    p.g = g; p.Cin = Cin; p.Cout = Cout;
    p.Nci = Cin % 128 == 0 ? 128 : (Cin % 64 == 0 ? 64 : 32);
    p.nct = (Cout + kMT - 1) / kMT;
    p.nnt = Cin / p.Nci;
    p.stages_total = (int)(g.NPOSR / kWT);
    p.G = Cout == 64 ? 2 : (Cout == 32 ? 4 : 1);
    p.nky = (3 + p.G - 1) / p.G;
    const int items = p.nct * p.nnt * p.nky;
    p.nsplit = std::max(1, std::min(ctx->sm_count / items, p.stages_total));
    p.a_bytes = 3u * 16 * kWT * 16;
    p.b_bytes = 3u * (uint32_t)(p.Nci / 8) * kWTX * 16;
    p.single = ctx->tc_precision == CNN_TC_BF16X1 ? 1 : 0;
    p.pieces = p.single ? 1 : grad_pieces(ctx);   // single pass: only the hi pieces are read
    const size_t stage = (size_t)p.a_bytes + p.b_bytes;
    p.nstage = (int)std::min<size_t>(4, (227 * 1024 - 128) / stage);
    CNN_REQUIRE(p.nstage >= 2, "conv_s1: weight-gradient stage does not fit in shared memory");
    const size_t partb = (size_t)items * p.nsplit * 3 * kMT * p.Nci * sizeof(float);
    uint8_t* scratch = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, partb + 256));
    CNN_REQUIRE(scratch, "scratch allocation failed");
    p.partial = reinterpret_cast<float*>(align_up1((uintptr_t)scratch, 256));
    // packed buffers as 4-D tensors (16-byte chunk, position, channel group, piece)
    CUtensorMap dmap, xmap;
    {
        const uint64_t dims[4] = {4, (uint64_t)g.RUN, (uint64_t)(Cout / 8), 3};
        const uint64_t str[3] = {16, (uint64_t)g.RUN * 16, (uint64_t)g.RUN * 16 * (Cout / 8)};
        const uint32_t box[4] = {4, kWT, (uint32_t)(p.G > 1 ? Cout / 8 : 16), 1};
        if (int rc = cnn_tmap_encode(&dmap, pd, 4, dims, str, box)) return rc;
    }
    {
        const uint64_t dims[4] = {4, (uint64_t)g.RUN, (uint64_t)(Cin / 8), 3};
        const uint64_t str[3] = {16, (uint64_t)g.RUN * 16, (uint64_t)g.RUN * 16 * (Cin / 8)};
        const uint32_t box[4] = {4, kWTX, (uint32_t)(p.Nci / 8), 1};
        if (int rc = cnn_tmap_encode(&xmap, px, 4, dims, str, box)) return rc;
    }
    if (int rc = s1_attrs(ctx->device)) return rc;
    const size_t smem = 128 + (size_t)p.nstage * stage;
    CNN_LAUNCH(ctx, s1_wgrad_kernel, items * p.nsplit, kWgThreadsS1, smem, dmap, xmap, p);
    const long long total = (long long)p.nct * p.nky * kMT * Cin * 3 + Cout;
    CNN_LAUNCH(ctx, s1_wgrad_reduce_kernel, cdiv(total, 32), 256, 0, p.partial, dbp, (int)pack_blocks, dw, db, Cin, Cout, p.Nci,
               p.nnt, p.nsplit, p.G, p.nky, scale);
    return CNN_OK;
}

int conv_wgrad_s1(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B, int Cin, int H, int W,
                  int Cout, float scale) {
    const size_t pxb = conv_s1_pk_bytes(B, Cin, H, W), pdb = conv_s1_pk_bytes(B, Cout, H, W);
    uint8_t* a = reinterpret_cast<uint8_t*>(cnn_arena(ctx, pxb + pdb + conv_s1_dbp_bytes(B, Cout, H, W)));
    CNN_REQUIRE(a, "conv_s1: arena allocation failed");
    float* dbp = reinterpret_cast<float*>(a + pxb + pdb);
    if (int rc = conv_s1_pack(ctx, x, nullptr, a, nullptr, B, Cin, H, W, H, W)) return rc;
    if (int rc = conv_s1_pack(ctx, delta, nullptr, a + pxb, dbp, B, Cout, H, W, H - 2, W - 2)) return rc;
    return conv_s1_wgrad_packed(ctx, a, a + pxb, dbp, dw, db, B, Cin, H, W, Cout, scale);
}
