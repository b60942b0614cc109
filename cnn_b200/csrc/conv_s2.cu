// conv_s2.cu -- 3x3 stride-2 Conv2D (the reference's default geometry, architectures.h:69) forward,
// input gradient and weight gradient as *shifted-window* implicit GEMMs on tcgen05, for layers with
// Cin % 16 == 0 and Cout % 16 == 0 (AlexNet-lite conv2..conv4, alexnet.cpp:17-26).
//
// Idea.  The gather kernels of conv_tc.cu rebuild an im2col tile per K block with SIMT producer
// warps (2.25x duplicated, latency bound).  Here an activation tensor is first re-laid ("packed")
// into a tensor-core-native format and the GEMM kernels then need NO producer warps at all:
//
//   P(x)  "parity-plane chunks" of x[B][C][H][W]:   [half][cg = C/8][plane q = (y&1, x&1)][gpos]
//         one 16-byte chunk = 8 consecutive channels (bf16) of one pixel; gpos = b*PP + (y>>1)*HP + (x>>1),
//         HP = (W+1)/2, PP = ((H+1)/2)*HP; half 0 = bf16(x), half 1 = bf16(x - hi)  (BF16x3 split)
//   dP    the same for delta[B][Cout][OH][OW] with ONE plane at the SAME pitch: gpos = b*PP + oy*HP + ox
//         (OH = (H+1)/2 - 1, OW = HP - 1: the spare row / column of every image is zero and doubles as
//         the halo of the input-gradient taps), preceded by a zero guard band.
//
// In this format a row of 128 consecutive gpos is a ready-made UMMA operand in the *no-swizzle*
// canonical layouts (8 rows x 16 bytes core matrices, rows 16 bytes apart):
//   forward   D[m][co]   = sum_tap sum_ci P(x)[q(tap)][m + shift(tap)][ci] * W[co][ci][tap]     (conv2d.cpp:69-92)
//             K-major A; a filter tap is just a different plane + start address: 9 x Cin/16 x 3 MMAs per tile
//   dgrad     Dcell[m][ci] = sum_tap(cell) sum_co dP[m - shift(tap)][co] * W[co][ci][tap]       (conv2d.cpp:161-201)
//             m = 2x2 input patch, 4 accumulators (patch cells) side by side in TMEM, no wasted taps
//   wgrad     Dtap[co][ci] = sum_m dP[m][co] * P(x)[q(tap)][m + shift(tap)][ci]                 (conv2d.cpp:108-159)
//             K = pixels: both operands MN-major views of the same buffers, 9 accumulators in TMEM
// All operand traffic is cp.async.bulk (TMA engine) from the packed buffers -- one elected thread
// issues copies and MMAs, the CTA's four warps only run the epilogue.  Results: BF16x3 split
// (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), normwise error ~5e-6 against the fp32 reference.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kTile = 128;   // GEMM rows per tile == TMEM lanes

struct S2Geom {
    int B, Cin, H, W, Cout, OH, OW;
    int HPr, HP, PP;       // plane rows / cols / positions per image
    int SH;                // shift span: round8(HP + 1)
    int NPT;               // positions staged per tile: kTile + SH
    long long NPOS;        // B * PP
    long long NPOS128;     // NPOS rounded up to a multiple of 128
    long long RUNX;        // positions per (half, cg, plane) run of P(x)   = NPOS128 + SH
    long long RUND;        // positions per (half, cg) run of dP            = SH (guard) + NPOS128
};

S2Geom make_geom(int B, int Cin, int H, int W, int Cout) {
    S2Geom g{};
    g.B = B; g.Cin = Cin; g.H = H; g.W = W; g.Cout = Cout;
    g.OH = (H - 3) / 2 + 1; g.OW = (W - 3) / 2 + 1;
    g.HPr = (H + 1) / 2; g.HP = (W + 1) / 2; g.PP = g.HPr * g.HP;
    g.SH = (g.HP + 1 + 7) / 8 * 8;
    g.NPT = kTile + g.SH;
    g.NPOS = (long long)B * g.PP;
    g.NPOS128 = (g.NPOS + 127) / 128 * 128;
    g.RUNX = g.NPOS128 + g.SH;
    g.RUND = g.SH + g.NPOS128;
    return g;
}

// ------------------------------------------------------------------------------------ packing
// x[B][C][H][W] fp32 -> P(x).  Thread = (gpos, cg): 4 planes x 8 channels; the slack behind the last
// image is written as zeros (the GEMMs read it for their garbage rows, and 0 * garbage must stay 0).
__global__ void __launch_bounds__(256) s2_pack_x_kernel(const float* __restrict__ x, uint4* __restrict__ px,
                                                         const S2Geom g) {
    const long long gpos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gpos >= g.RUNX) return;
    const int cg = blockIdx.y, ncg = g.Cin >> 3;
    const bool in = gpos < g.NPOS;
    int b = 0, py = 0, pxx = 0;
    if (in) {
        b = (int)(gpos / g.PP);
        const int r = (int)(gpos - (long long)b * g.PP);
        py = r / g.HP;
        pxx = r - py * g.HP;
    }
    const size_t plane = (size_t)g.H * g.W;
    const float* xb = x + ((size_t)b * g.Cin + (size_t)cg * 8) * plane;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int y = 2 * py + (q >> 1), xx = 2 * pxx + (q & 1);
        const bool ok = in && y < g.H && xx < g.W;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = ok ? __ldg(xb + (size_t)j * plane + (size_t)y * g.W + xx) : 0.f;
        uint4 hi, mid, lo;
        split8x3(v, hi, mid, lo);
        px[((size_t)(0 * ncg + cg) * 4 + q) * g.RUNX + gpos] = hi;
        px[((size_t)(1 * ncg + cg) * 4 + q) * g.RUNX + gpos] = mid;
        px[((size_t)(2 * ncg + cg) * 4 + q) * g.RUNX + gpos] = lo;
    }
}

// delta[B][Cout][OH][OW] fp32 -> dP (guard band, spare row / column and tail slack written as zeros),
// plus per-block partial sums of delta per channel for the bias gradient (conv2d.cpp:153-157).
__global__ void __launch_bounds__(256) s2_pack_d_kernel(const float* __restrict__ d, uint4* __restrict__ pd,
                                                         float* __restrict__ db_partial, const S2Geom g) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cg = blockIdx.y, ncg = g.Cout >> 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (r < g.RUND) {
        const long long gpos = r - g.SH;
        if (gpos >= 0 && gpos < g.NPOS) {
            const int b = (int)(gpos / g.PP);
            const int rr = (int)(gpos - (long long)b * g.PP);
            const int oy = rr / g.HP, ox = rr - oy * g.HP;
            if (oy < g.OH && ox < g.OW) {
                const size_t plane = (size_t)g.OH * g.OW;
                const float* p = d + ((size_t)b * g.Cout + (size_t)cg * 8) * plane + (size_t)oy * g.OW + ox;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(p + (size_t)j * plane);
            }
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        pd[(size_t)(0 * ncg + cg) * g.RUND + r] = hi;
        pd[(size_t)(1 * ncg + cg) * g.RUND + r] = lo;
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
            db_partial[(size_t)blockIdx.x * g.Cout + cg * 8 + threadIdx.x] = s;
        }
    }
}

// filters -> per-K-stage operand blocks [kc][half][tap][cgl = 2][n] of 16-byte chunks (8 k values):
//   forward: n = co, k = ci (W[co][ci][tap]), three pieces ; input gradient: n = ci, k = co, two pieces
__device__ __forceinline__ void s2_pack_w_body(const float* __restrict__ w, uint4* __restrict__ out, int Cin, int Cout,
                                               int dgrad, int first, int step) {
    const int N = dgrad ? Cin : Cout, Kc = (dgrad ? Cout : Cin) >> 4;
    const int total = Kc * 9 * 2 * N;
    for (int id = first; id < total; id += step) {
        const int n = id % N;
        int t = id / N;
        const int cgl = t & 1;
        t >>= 1;
        const int tap = t % 9, kc = t / 9;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kc * 16 + cgl * 8 + j;
            v[j] = dgrad ? w[((size_t)k * Cin + n) * 9 + tap] : w[((size_t)n * Cin + k) * 9 + tap];
        }
        if (dgrad) {
            uint4 hi, lo;
            split8(v, hi, lo);
            out[((size_t)((kc * 2 + 0) * 9 + tap) * 2 + cgl) * N + n] = hi;
            out[((size_t)((kc * 2 + 1) * 9 + tap) * 2 + cgl) * N + n] = lo;
        } else {
            uint4 hi, mid, lo;
            split8x3(v, hi, mid, lo);
            out[((size_t)((kc * 3 + 0) * 9 + tap) * 2 + cgl) * N + n] = hi;
            out[((size_t)((kc * 3 + 1) * 9 + tap) * 2 + cgl) * N + n] = mid;
            out[((size_t)((kc * 3 + 2) * 9 + tap) * 2 + cgl) * N + n] = lo;
        }
    }
}

__global__ void __launch_bounds__(256) s2_pack_w_kernel(const float* __restrict__ w, uint4* __restrict__ out, int Cin,
                                                         int Cout, int dgrad) {
    s2_pack_w_body(w, out, Cin, Cout, dgrad, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// all filter packs of a step in one launch (the engine packs every s2 layer's forward and
// input-gradient blocks up front: the parameters do not change inside a step); blockIdx.y = job
struct S2PackJobs {
    const float* w[8];
    uint4* out[8];
    int Cin[8], Cout[8], dgrad[8];
};
__global__ void __launch_bounds__(256) s2_pack_w_multi_kernel(const S2PackJobs j) {
    const int k = blockIdx.y;
    s2_pack_w_body(j.w[k], j.out[k], j.Cin[k], j.Cout[k], j.dgrad[k], blockIdx.x * blockDim.x + threadIdx.x,
                   gridDim.x * blockDim.x);
}

// ----------------------------------------------------------------------------- forward / dgrad
struct S2Gemm {
    const uint4* act;     // forward: P(x) ; dgrad: dP
    const uint4* wpk;     // packed filters
    const float* bias;    // forward only
    const float* relu_y;  // dgrad: ReLU output of the layer below (mask y <= 0 -> 0, relu.cpp:39) or null
    float* dst;           // forward: y ; dgrad: dx
    float* dst_relu;      // forward: optional ReLU output (relu.cpp:25)
    int dst_nhwc;         // dgrad: dx is written channel-last [B][H][W][Cin] (the lazy head's weight gradient reads it so)
    // forward: optional packed copy P(relu output) for a following s2 layer (its conv input), written by
    // the same epilogue so that layer needs no pack kernel: geometry of THAT layer's input planes
    uint4* next_px;
    int nx_HP, nx_PP, nx_ncg;
    long long nx_RUNX;
    S2Geom g;
    int KC;               // K stages of 16 channels
    int N;                // accumulator columns per cell (forward: Cout, dgrad: Cin)
    int nstage;           // activation (+ filter) stage ring depth, <= 8
    int wres;             // 1: all filter blocks stay resident in shared memory (loaded once per CTA)
    int accbufs;          // TMEM accumulator buffers (2: the epilogue overlaps the next tile's MMAs)
    int acc_cols;         // columns per accumulator buffer
    int tmem_cols;
    int tiles;
    int dbg;              // CNN_DBG_S2 experiment knob: 1 = no MMAs, 2 = no epilogue stores
    uint32_t a_bytes, b_bytes;
};

constexpr int kS2Threads = 128;          // weight-gradient kernel
constexpr int kS2GemmThreads = 320;      // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane group)
constexpr int kS2Header = 256 + 1024;    // barriers + bias

// Persistent, warp-specialised: tiles are strided over the grid; warp 0 streams the operand runs of
// (tile, K stage) items through a shared-memory ring with cp.async.bulk, warp 1 issues the MMAs of an
// item as soon as it has landed, warps 2-5 drain one of two TMEM accumulator buffers while the MMAs of
// the next tile fill the other.  DGRAD = false: forward ; true: input gradient (4 patch-cell accumulators).
template <bool DGRAD>
__global__ void __launch_bounds__(kS2GemmThreads) s2_gemm_kernel(const S2Gemm p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // [8]
    uint64_t* empty = full + 8;                               // [8]
    uint64_t* wfull = full + 16;
    uint64_t* acc_full = full + 17;                           // [2]
    uint64_t* acc_empty = full + 19;                          // [2]
    uint32_t* tslot = reinterpret_cast<uint32_t*>(full + 21);
    float* sbias = reinterpret_cast<float*>(smem + 256);
    uint8_t* wreg = smem + kS2Header;                         // resident filter blocks (wres)
    const uint32_t w_all = p.wres ? (uint32_t)p.KC * p.b_bytes : 0u;
    uint8_t* stages = wreg + w_all;
    const S2Geom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t stage_bytes = p.a_bytes + (p.wres ? 0u : p.b_bytes);

    if (warp == 0) {
        tmem_alloc(tslot, (uint32_t)p.tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < p.nstage; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            mbar_init(wfull, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&acc_full[i], 1);
                mbar_init(&acc_empty[i], 8);
            }
            mbar_fence_init();
        }
    }
    if (!DGRAD)
        for (int i = tid; i < p.N; i += kS2GemmThreads) sbias[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA: one run per lane
        constexpr int NRUN = DGRAD ? 4 : 24;      // (piece, [plane,] cgl) runs of NPT positions
        const int ncg = (DGRAD ? g.Cout : g.Cin) >> 3;
        const long long run = DGRAD ? g.RUND : g.RUNX;
        if (p.wres && lane == 0) {
            mbar_expect_tx(wfull, w_all);
            for (int kc = 0; kc < p.KC; ++kc)
                tma_bulk_g2s(wreg + (size_t)kc * p.b_bytes, reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)kc * p.b_bytes,
                             p.b_bytes, wfull);
        }
        uint32_t s = 0, ph = 0;
        bool wrapped = false;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const long long m0 = (long long)tile * kTile;
            for (int kc = 0; kc < p.KC; ++kc) {
                if (wrapped) mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = stages + (size_t)s * stage_bytes;
                if (lane == 0) mbar_expect_tx(&full[s], stage_bytes);
                __syncwarp();
                if (lane < NRUN) {
                    const int h = DGRAD ? (lane >> 1) : (lane >> 3), cgl = lane & 1;
                    const int q = DGRAD ? 0 : ((lane >> 1) & 3);
                    const size_t chan = (size_t)h * ncg + (size_t)kc * 2 + cgl;
                    const uint4* src = p.act + (DGRAD ? chan : chan * 4 + q) * run + m0;   // dP: run offset m0 == gpos m0 - SH
                    tma_bulk_g2s(st + (size_t)lane * g.NPT * 16, src, (uint32_t)g.NPT * 16, &full[s]);
                } else if (lane == NRUN && !p.wres) {
                    tma_bulk_g2s(st + p.a_bytes, reinterpret_cast<const uint8_t*>(p.wpk) + (size_t)kc * p.b_bytes, p.b_bytes,
                                 &full[s]);
                }
                if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; wrapped = true; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issue
        const uint32_t idesc = idesc_bf16(kTile, p.N);
        const uint32_t st0 = smem_u32(stages), w0 = smem_u32(wreg);
        const uint32_t a_lbo = (uint32_t)g.NPT * 16, b_lbo = (uint32_t)p.N * 16;
        const uint32_t a_half = (DGRAD ? 2u : 8u) * a_lbo;           // bytes between the pieces of an operand
        const uint32_t b_half = 9u * 2u * b_lbo;
        if (p.wres) mbar_wait(wfull, 0);
        uint32_t s = 0, ph = 0, a = 0, aph = 0;
        bool awrapped = false;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            if (awrapped) mbar_wait(&acc_empty[a], aph ^ 1);
            const uint32_t dbase = tmem + a * (uint32_t)p.acc_cols;
            for (int kc = 0; kc < p.KC; ++kc) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = st0 + s * stage_bytes;
                    const uint32_t sb = p.wres ? w0 + (uint32_t)kc * p.b_bytes : sa + p.a_bytes;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int ky = tap / 3, kx = tap % 3;
                        uint32_t a_off, cell;
                        if (DGRAD) {
                            cell = (uint32_t)((ky & 1) * 2 + (kx & 1));
                            a_off = (uint32_t)(g.SH - ((ky >> 1) * g.HP + (kx >> 1))) * 16;
                        } else {
                            cell = (uint32_t)ky;   // hi*hi products of filter row ky; accumulator 3 = correction terms
                            const int q = (ky & 1) * 2 + (kx & 1);
                            a_off = (uint32_t)q * 2 * a_lbo + (uint32_t)((ky >> 1) * g.HP + (kx >> 1)) * 16;
                        }
                        // first MMA into an accumulator (cell) of this tile overwrites it
                        const bool first = kc == 0 && (DGRAD ? (ky < 2 && kx < 2) : kx == 0);
                        const uint64_t a0 = desc_nosw(sa + a_off, a_lbo, 128), a1 = desc_nosw(sa + a_half + a_off, a_lbo, 128);
                        const uint32_t bo = sb + (uint32_t)tap * 2 * b_lbo;
                        const uint64_t b0 = desc_nosw(bo, b_lbo, 128), b1 = desc_nosw(bo + b_half, b_lbo, 128);
                        const uint32_t d = dbase + cell * (uint32_t)p.N;
                        if (p.dbg & 1) continue;
                        if (DGRAD) {   // hi*hi + hi*lo + lo*hi, small terms first
                            mma_bf16(d, a1, b0, idesc, !first);
                            mma_bf16(d, a0, b1, idesc, true);
                            mma_bf16(d, a0, b0, idesc, true);
                        } else {
                            // Three pieces: every product term down to 2^-24.  tcgen05 adds into its fp32
                            // accumulator with TRUNCATION, a systematic shrink of ~2^-25 per accumulate step;
                            // 54-216 steps into one accumulator biased the logits by ~5e-6, which the softmax
                            // and the cancellation of the batch-mean gradients amplified to 6e-4 on a B=256
                            // step (tools/fullstep_parity_b256.py).  The large hi*hi products therefore go to
                            // one accumulator per filter row (3 x KC steps each), all small correction terms to
                            // a fourth whose magnitude (and ulp) is 2^-8 of theirs; the epilogue adds the four
                            // in fp32 with round-to-nearest.
                            const uint64_t a2 = desc_nosw(sa + 2 * a_half + a_off, a_lbo, 128);
                            const uint64_t b2 = desc_nosw(bo + 2 * b_half, b_lbo, 128);
                            const uint32_t dc = dbase + 3u * (uint32_t)p.N;
                            mma_bf16(dc, a2, b0, idesc, !(kc == 0 && tap == 0));
                            mma_bf16(dc, a0, b2, idesc, true);
                            mma_bf16(dc, a1, b1, idesc, true);
                            mma_bf16(dc, a1, b0, idesc, true);
                            mma_bf16(dc, a0, b1, idesc, true);
                            mma_bf16(d, a0, b0, idesc, !first);
                        }
                    }
                    mma_commit(&empty[s]);
                    if (kc == p.KC - 1) mma_commit(&acc_full[a]);
                }
                __syncwarp();
                if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; }
            }
            if (++a == (uint32_t)p.accbufs) { a = 0; aph ^= 1; awrapped = true; }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2-9)
        // warp w reads TMEM lanes 32*(w & 3)...; the two warps of a lane group split the work of a row:
        // forward = halves of the channel chunks, input gradient = patch cells {0,1} / {2,3}
        const int wq = warp & 3, half = (warp - 2) >> 2;
        const int row = wq * 32 + lane;
        uint32_t a = 0, aph = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
            const long long m = (long long)tile * kTile + row;
            const bool in = m < g.NPOS;
            int b = 0, py = 0, pxx = 0;
            if (in) {
                b = (int)(m / g.PP);
                const int r = (int)(m - (long long)b * g.PP);
                py = r / g.HP;
                pxx = r - py * g.HP;
            }
            // input gradient: the ReLU outputs that gate this row (relu.cpp:39) are fetched and folded into mask bits
            // BEFORE waiting for the accumulator -- the loads overlap the tile's MMAs instead of sitting between the
            // TMEM reads and the stores (12 % of this kernel's samples were the compare behind those loads).
            // keep = 16 bits per work unit (cell, 16-channel chunk), unit 0 in the low bits.
            uint32_t keep[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) keep[i] = 0xFFFFFFFFu;
            if (DGRAD && p.relu_y) {
                const size_t iplane = (size_t)g.H * g.W;
                const int nch = p.N >> 4, units = 2 * nch;
                for (int u = 0; u < units; ++u) {
                    const int cell = 2 * half + u / nch, c0 = (u % nch) * 16;
                    const int y = 2 * py + (cell >> 1), xx = 2 * pxx + (cell & 1);
                    uint32_t m16 = 0xFFFFu;
                    if (in && y < g.H && xx < g.W) {
                        const float* ry = p.relu_y + (size_t)b * g.Cin * iplane + (size_t)c0 * iplane + (size_t)y * g.W + xx;
                        float t[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) t[j] = __ldg(ry + (size_t)j * iplane);
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (t[j] <= 0.f) m16 &= ~(1u << j);
                    }
                    // shift the 256-bit register array right by 16 and enter the new unit at the top
#pragma unroll
                    for (int i = 0; i < 7; ++i) keep[i] = __funnelshift_r(keep[i], keep[i + 1], 16);
                    keep[7] = (keep[7] >> 16) | (m16 << 16);
                }
                for (int u = units; u < 16; ++u) {   // align unit 0 with bit 0
#pragma unroll
                    for (int i = 0; i < 7; ++i) keep[i] = __funnelshift_r(keep[i], keep[i + 1], 16);
                    keep[7] >>= 16;
                }
            }
            mbar_wait(&acc_full[a], aph);
            tc_fence_after();
            const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16) + a * (uint32_t)p.acc_cols;
            if (!DGRAD) {
                const bool ok = in && py < g.OH && pxx < g.OW;
                const size_t oplane = (size_t)g.OH * g.OW;
                const size_t o0 = (size_t)b * g.Cout * oplane + (size_t)py * g.OW + pxx;
                const int nch = p.N >> 4, cbeg = half ? (nch + 1) / 2 : 0, cend = half ? nch : (nch + 1) / 2;
                for (int c = cbeg; c < cend; ++c) {
                    const int c0 = c * 16;
                    float v[16];
                    {   // ((row 0 + row 1) + row 2) + correction terms, round-to-nearest fp32 adds
                        float v1[16], v2[16], vc[16];
                        tmem_ld16(trow + c0, v);
                        tmem_ld16(trow + p.N + c0, v1);
                        tmem_ld16(trow + 2 * p.N + c0, v2);
                        tmem_ld16(trow + 3 * p.N + c0, vc);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __fadd_rn(__fadd_rn(__fadd_rn(v[j], v1[j]), v2[j]), vc[j]);
                    }
                    if (ok && !(p.dbg & 2)) {
                        float q[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float r = v[j] + sbias[c0 + j];
                            const size_t o = o0 + (size_t)(c0 + j) * oplane;
                            p.dst[o] = r;
                            q[j] = r >= 0.f ? r : 0.f;
                            if (p.dst_relu) p.dst_relu[o] = q[j];
                        }
                        if (p.next_px) {
                            // this pixel of the next layer's input: plane (oy&1, ox&1), position (oy>>1, ox>>1)
                            const size_t gpos = (size_t)b * p.nx_PP + (size_t)(py >> 1) * p.nx_HP + (pxx >> 1);
                            const int qn = (py & 1) * 2 + (pxx & 1);
#pragma unroll
                            for (int h8 = 0; h8 < 2; ++h8) {
                                uint4 hi, mid, lo;
                                split8x3(q + 8 * h8, hi, mid, lo);
                                const size_t cgn = (size_t)(c0 >> 3) + h8;
                                p.next_px[((0 * (size_t)p.nx_ncg + cgn) * 4 + qn) * p.nx_RUNX + gpos] = hi;
                                p.next_px[((1 * (size_t)p.nx_ncg + cgn) * 4 + qn) * p.nx_RUNX + gpos] = mid;
                                p.next_px[((2 * (size_t)p.nx_ncg + cgn) * 4 + qn) * p.nx_RUNX + gpos] = lo;
                            }
                        }
                    }
                }
            } else {
                // work units (cell, 16-channel chunk), gated by the mask bits gathered above
                const size_t iplane = (size_t)g.H * g.W;
                const int nch = p.N >> 4, units = 2 * nch;
                for (int u = 0; u < units; ++u) {
                    const int cell = 2 * half + u / nch, c0 = (u % nch) * 16;
                    const int y = 2 * py + (cell >> 1), xx = 2 * pxx + (cell & 1);
                    const bool ok_c = in && y < g.H && xx < g.W;
                    const size_t o_c = (size_t)b * g.Cin * iplane + (size_t)c0 * iplane + (size_t)y * g.W + xx;
                    float v[16];
                    tmem_ld16(trow + cell * p.N + c0, v);
                    const uint32_t m16 = keep[0];
                    if (ok_c && !(p.dbg & 2)) {
                        if (p.dst_nhwc) {   // 16 consecutive channels of one pixel: four 16-byte stores
                            float4* q = reinterpret_cast<float4*>(p.dst + (((size_t)b * g.H + y) * g.W + xx) * g.Cin + c0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) q[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) p.dst[o_c + (size_t)j * iplane] = ((m16 >> j) & 1u) ? v[j] : 0.f;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 7; ++i) keep[i] = __funnelshift_r(keep[i], keep[i + 1], 16);
                    keep[7] >>= 16;
                }
            }
            // accumulator fully read: hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
            if (++a == (uint32_t)p.accbufs) { a = 0; aph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// -------------------------------------------------------------------------------- weight gradient
// D[row = (g, co)][ci] (+)= sum_k dP[co][k - s_g - s_addr] * P(x)[plane q][ci][k]
// The x operand is read unshifted; the tap shift S = s_g + s_addr sits on the delta side, where it is
// either a start-address offset (K rows are 16 bytes apart in the MN-major layout) or -- to fill all
// 128 MMA rows when Cout < 128 -- one of G shifted copies of the delta run stacked as row groups:
//   Cout <= 32: G = 4 copies (shifts 0, 1, HP, HP+1)  -> one MMA per parity plane covers up to 4 taps
//   Cout <= 64: G = 2 copies (shifts 0, 1), s_addr in {0, HP}                       -> 6 MMAs per K step
//   else       : G = 1, s_addr in {0, 1, HP, HP+1}                                   -> 9 MMAs per K step
struct S2Wgrad {
    const uint4* px;      // P(x)
    const uint4* pd;      // dP
    float* partial;       // [cta][Cout][Cin][9]
    S2Geom g;
    int KT;               // pixels per K chunk
    int KTA;              // staged delta positions per run: KT + halo
    int halo;             // delta positions staged in front of a chunk (start-address shifts)
    int G, rows_per;      // row groups per MMA, rows per group (128 / G)
    int cin_per;          // input channels per CTA column (grid.y splits Cin)
    int nstage;
    int tmem_cols;
    int chunks;           // total K chunks
    int dbg;              // CNN_DBG_S2: 1 = no MMAs, 2 = no partial stores
    int nmma;             // MMAs (accumulators) per K step
    int gshift[4];        // s_g per row group
    signed char mq[9];    // plane of MMA u
    int maddr[9];         // s_addr of MMA u
    signed char mtap[9][4];   // filter tap (ky*3+kx) produced by row group g of MMA u, -1 = none
    uint32_t a_bytes, b_bytes;
};

__global__ void __launch_bounds__(kS2Threads) s2_wgrad_kernel(const S2Wgrad p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 4;
    uint64_t* accbar = full + 8;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(full + 9);
    uint8_t* stages = smem + 128;
    const S2Geom& g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t stage_bytes = p.a_bytes + p.b_bytes;
    // contiguous chunk range of this CTA
    const int per = (p.chunks + gridDim.x - 1) / gridDim.x;
    const int c_begin = blockIdx.x * per, c_end = min(p.chunks, c_begin + per);
    const int n_it = max(c_end - c_begin, 0);
    const int ci0 = blockIdx.y * p.cin_per;

    if (warp == 0) {
        tmem_alloc(tslot, (uint32_t)p.tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < p.nstage; ++i) {
                mbar_init(&full[i], 1);
                mbar_init(&empty[i], 1);
            }
            mbar_init(accbar, 1);
            mbar_fence_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    const int ncg_real = g.Cout >> 3, ncg_pad = p.rows_per >> 3, ncg_b = p.cin_per >> 3, ncg_x = g.Cin >> 3;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA
        const int runs_a = 2 * p.G * ncg_real, runs_b = 2 * 4 * ncg_b;
        const uint32_t tx = (uint32_t)runs_a * p.KTA * 16 + (uint32_t)runs_b * p.KT * 16;
        for (int it = 0; it < n_it; ++it) {
            const int s = it % p.nstage;
            if (it >= p.nstage) mbar_wait(&empty[s], ((it / p.nstage) - 1) & 1);
            uint8_t* st = stages + (size_t)s * stage_bytes;
            const long long k0 = (long long)(c_begin + it) * p.KT;
            if (lane == 0) mbar_expect_tx(&full[s], tx);
            __syncwarp();
            for (int r = lane; r < runs_a + runs_b; r += 32) {
                if (r < runs_a) {          // dP copies [piece][group][cg][KTA]
                    const int cg = r % ncg_real, gi = (r / ncg_real) % p.G, h = r / (ncg_real * p.G);
                    tma_bulk_g2s(st + (size_t)((h * p.G + gi) * ncg_pad + cg) * p.KTA * 16,
                                 p.pd + (size_t)(h * ncg_real + cg) * g.RUND + g.SH + k0 - p.gshift[gi] - p.halo,
                                 (uint32_t)p.KTA * 16, &full[s]);
                } else {                   // P(x) [piece][plane][cg][KT]
                    const int rb = r - runs_a;
                    const int cg = rb % ncg_b, q = (rb / ncg_b) & 3, h = rb / (4 * ncg_b);
                    tma_bulk_g2s(st + p.a_bytes + (size_t)rb * p.KT * 16,
                                 p.px + ((size_t)(h * ncg_x + (ci0 >> 3) + cg) * 4 + q) * g.RUNX + k0, (uint32_t)p.KT * 16,
                                 &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issue
        const uint32_t idesc = idesc_bf16_mn(kTile, p.cin_per);
        const uint32_t st0 = smem_u32(stages);
        const uint32_t a_sbo = (uint32_t)p.KTA * 16, b_sbo = (uint32_t)p.KT * 16;
        const uint32_t a_half = (uint32_t)(p.G * ncg_pad) * a_sbo, b_half = 4u * (uint32_t)ncg_b * b_sbo;
        for (int it = 0; it < n_it; ++it) {
            const int s = it % p.nstage;
            mbar_wait(&full[s], (it / p.nstage) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sa = st0 + (uint32_t)s * stage_bytes, sb = sa + p.a_bytes;
                for (int j = 0; j < p.KT / 16; ++j) {
                    for (int u = 0; u < p.nmma; ++u) {
                        if (p.dbg & 1) continue;
                        const uint32_t ao = sa + (uint32_t)(p.halo - p.maddr[u] + j * 16) * 16;
                        const uint64_t ahi = desc_nosw(ao, 128, a_sbo), alo = desc_nosw(ao + a_half, 128, a_sbo);
                        const uint32_t bo = sb + (uint32_t)p.mq[u] * ncg_b * b_sbo + (uint32_t)j * 256;
                        const uint64_t bhi = desc_nosw(bo, 128, b_sbo), blo = desc_nosw(bo + b_half, 128, b_sbo);
                        const uint32_t d = tmem + (uint32_t)(u * p.cin_per);
                        const bool acc = (it | j) != 0;
                        mma_bf16(d, alo, bhi, idesc, acc);
                        mma_bf16(d, ahi, blo, idesc, true);
                        mma_bf16(d, ahi, bhi, idesc, true);
                    }
                }
                mma_commit(&empty[s]);
                if (it == n_it - 1) mma_commit(accbar);
            }
            __syncwarp();
        }
    }
    // ------------------------------------------------------------------ epilogue: partial[cta][u][ci][row]
    // (row fastest: every store instruction of a warp is one 128-byte line; the reduce kernel maps
    // rows back to (group, co) -> filter tap)
    __syncwarp();
    float* out = p.partial + (size_t)blockIdx.x * p.nmma * g.Cin * kTile;
    if (n_it > 0) {
        mbar_wait(accbar, 0);
        tc_fence_after();
    }
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int u = 0; u < p.nmma; ++u)
        for (int c0 = 0; c0 < p.cin_per; c0 += 16) {
            float v[16];
            if (n_it > 0) tmem_ld16(trow + u * p.cin_per + c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (!(p.dbg & 2)) out[((size_t)u * g.Cin + ci0 + c0 + j) * kTile + tid] = n_it > 0 ? v[j] : 0.f;
        }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// dw / db = scale * sum over CTA partials / pack-block partials, in a fixed order (deterministic).
// Block = 32 consecutive partial elements x 8 part lanes (coalesced reads); element (u, ci, row) maps
// to dw[co][ci][tap] through the MMA -> tap table of the weight-gradient kernel.
struct S2ReduceMap {
    int nmma, Cin, Cout, G, rows_per;
    signed char mtap[9][4];
};

__global__ void __launch_bounds__(256) s2_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, int psize,
                                                               const S2ReduceMap mp, const float* __restrict__ db_partial,
                                                               int nblocks, float* __restrict__ dw, float* __restrict__ db,
                                                               float scale) {
    __shared__ float red[32][33];
    const int nb_w = psize / 32;   // psize = nmma * Cin * 128
    const bool seg1 = (int)blockIdx.x >= nb_w;
    // dw: 32 elements x 8 part lanes ; db: 8 channels x 32 part lanes (many more pack blocks than CTAs)
    const int wcols = seg1 ? 8 : 32, lanes = 256 / wcols;
    const float* src = seg1 ? db_partial : partial;
    const int n = seg1 ? mp.Cout : psize, rows = seg1 ? nblocks : nparts;
    const int col = threadIdx.x % wcols, rl = threadIdx.x / wcols;
    const int i = ((int)blockIdx.x - (seg1 ? nb_w : 0)) * wcols + col;
    float s = 0.f;
    if (i < n) {
        int r = rl;
        for (; r + 7 * lanes < rows; r += 8 * lanes) {   // eight independent loads in flight, fixed order
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(src + (size_t)(r + k * lanes) * n + i);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[k];
        }
        for (; r < rows; r += lanes) s += __ldg(src + (size_t)r * n + i);
    }
    red[rl][col] = s;
    __syncthreads();
    if (rl == 0 && i < n) {
        float t = 0.f;
        for (int k = 0; k < lanes; ++k) t += red[k][col];
        t *= scale;
        if (seg1) {
            db[i] = t;
        } else {
            const int row = i & (kTile - 1), ci = (i >> 7) % mp.Cin, u = (i >> 7) / mp.Cin;
            const int gi = row / mp.rows_per, co = row % mp.rows_per;
            const int tap = gi < mp.G ? mp.mtap[u][gi] : -1;
            if (tap >= 0 && co < mp.Cout) dw[((size_t)co * mp.Cin + ci) * 9 + tap] = t;
        }
    }
}

int next_pow2_cols(int c) {
    int p = 32;
    while (p < c) p <<= 1;
    return p;
}

template <class K>
int set_max_smem(K kernel, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return cnn_cuda_fail(e, name, __FILE__, __LINE__);
    return CNN_OK;
}

int attrs_once(int device) {
    std::lock_guard<std::recursive_mutex> lk(cnn_global_mutex());
    static bool done[16];
    if (device < 0 || device >= 16 || done[device]) return CNN_OK;
    if (int rc = set_max_smem(s2_gemm_kernel<false>, "s2_gemm_kernel<fwd>")) return rc;
    if (int rc = set_max_smem(s2_gemm_kernel<true>, "s2_gemm_kernel<dgrad>")) return rc;
    if (int rc = set_max_smem(s2_wgrad_kernel, "s2_wgrad_kernel")) return rc;
    done[device] = true;
    return CNN_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t px_bytes(const S2Geom& g) { return (size_t)3 * (g.Cin / 8) * 4 * g.RUNX * 16; }
size_t pd_bytes(const S2Geom& g) { return (size_t)2 * (g.Cout / 8) * g.RUND * 16; }

int launch_pack_x(cnn_ctx* ctx, const S2Geom& g, const float* x, uint4* px) {
    dim3 grid((unsigned)cdiv(g.RUNX, 256), (unsigned)(g.Cin / 8));
    CNN_LAUNCH(ctx, s2_pack_x_kernel, grid, 256, 0, x, px, g);
    return CNN_OK;
}

int launch_pack_d(cnn_ctx* ctx, const S2Geom& g, const float* d, uint4* pd, float* db_partial) {
    dim3 grid((unsigned)cdiv(g.RUND, 256), (unsigned)(g.Cout / 8));
    CNN_LAUNCH(ctx, s2_pack_d_kernel, grid, 256, 0, d, pd, db_partial, g);
    return CNN_OK;
}

int launch_gemm(cnn_ctx* ctx, const S2Geom& g, bool dgrad, const uint4* act, const uint4* wpk, const float* bias,
                const float* relu_y, float* dst, float* dst_relu, uint4* next_px = nullptr, bool dst_nhwc = false) {
    S2Gemm p{};
    p.act = act; p.wpk = wpk; p.bias = bias; p.relu_y = relu_y; p.dst = dst; p.dst_relu = dst_relu; p.g = g;
    p.dst_nhwc = dst_nhwc ? 1 : 0;
    if (next_px) {   // the following layer's input is this layer's [Cout][OH][OW] output
        const S2Geom ng = make_geom(g.B, g.Cout, g.OH, g.OW, 16);
        p.next_px = next_px; p.nx_HP = ng.HP; p.nx_PP = ng.PP; p.nx_ncg = g.Cout / 8; p.nx_RUNX = ng.RUNX;
    }
    p.KC = (dgrad ? g.Cout : g.Cin) / 16;
    p.N = dgrad ? g.Cin : g.Cout;
    p.acc_cols = 4 * p.N;   // dgrad: four patch cells ; forward: three filter-row accumulators + correction terms
    p.accbufs = 2 * p.acc_cols <= 512 ? 2 : 1;
    p.tmem_cols = next_pow2_cols(p.accbufs * p.acc_cols);
    if (const char* e = getenv("CNN_DBG_S2")) p.dbg = atoi(e);
    p.a_bytes = (uint32_t)(dgrad ? 4 : 24) * g.NPT * 16;
    p.b_bytes = (uint32_t)(dgrad ? 2 : 3) * 9 * 2 * p.N * 16;
    p.tiles = (int)(g.NPOS128 / kTile);
    // filters resident in shared memory when at least two activation stages still fit next to them
    const size_t budget = 227 * 1024 - kS2Header;
    const size_t w_all = (size_t)p.KC * p.b_bytes;
    p.wres = (w_all + 2 * (size_t)p.a_bytes <= budget && !getenv("CNN_DBG_S2_NOWRES")) ? 1 : 0;
    const size_t stage = (size_t)p.a_bytes + (p.wres ? 0 : p.b_bytes);
    CNN_REQUIRE((p.wres ? w_all : 0) + stage <= budget, "conv_s2: stage does not fit in shared memory");
    int ns = (int)((budget - (p.wres ? w_all : 0)) / stage);
    ns = std::max(1, std::min(ns, 8));
    // a ring deeper than all items of one CTA is wasted
    const int grid = std::min(p.tiles, ctx->sm_count);
    const int items = (p.tiles + grid - 1) / grid * p.KC;
    ns = std::min(ns, std::max(items, 1));
    p.nstage = ns;
    const size_t smem = kS2Header + (p.wres ? w_all : 0) + (size_t)ns * stage;
    if (int rc = attrs_once(ctx->device)) return rc;
    if (dgrad) { CNN_LAUNCH(ctx, s2_gemm_kernel<true>, grid, kS2GemmThreads, smem, p); }
    else { CNN_LAUNCH(ctx, s2_gemm_kernel<false>, grid, kS2GemmThreads, smem, p); }
    return CNN_OK;
}

}  // namespace

bool conv_s2_supported(const cnn_ctx* ctx, int Cin, int H, int W, int Cout, int k, int s) {
    (void)ctx;
    if (k != 3 || s != 2 || H < 3 || W < 3) return false;
    if (Cin % 16 || Cout % 16 || Cin > 256 || Cout > 128) return false;   // N <= 256 columns, M = Cout <= 128 rows (wgrad)
    if (4 * Cin > 512) return false;                                       // four dgrad accumulators in TMEM
    if ((W + 1) / 2 + 1 > 120) return false;                               // shift span stays small next to a tile
    return getenv("CNN_DBG_NOS2") == nullptr;
}

// ---- packed-buffer interface (the engine keeps P(x) from the forward pass for the weight gradient and
// packs delta once for both gradients) ----------------------------------------------------------------
void conv_s2_px_geom(int B, int Cin, int H, int W, int* HP, int* PP, long long* RUNX) {
    const S2Geom g = make_geom(B, Cin, H, W, 16);
    *HP = g.HP; *PP = g.PP; *RUNX = g.RUNX;
}
size_t conv_s2_px_bytes(int B, int Cin, int H, int W) { return align_up(px_bytes(make_geom(B, Cin, H, W, 16)), 256); }
size_t conv_s2_pd_bytes(int B, int Cout, int H, int W) { return align_up(pd_bytes(make_geom(B, 16, H, W, Cout)), 256); }
size_t conv_s2_dbp_bytes(int B, int Cout, int H, int W) {
    return align_up((size_t)cdiv(make_geom(B, 16, H, W, Cout).RUND, 256) * Cout * 4, 256);
}

int conv_s2_pack_x(cnn_ctx* ctx, const float* x, void* px, int B, int Cin, int H, int W) {
    return launch_pack_x(ctx, make_geom(B, Cin, H, W, 16), x, reinterpret_cast<uint4*>(px));
}

int conv_s2_pack_d(cnn_ctx* ctx, const float* delta, void* pd, float* dbp, int B, int Cout, int H, int W) {
    return launch_pack_d(ctx, make_geom(B, 16, H, W, Cout), delta, reinterpret_cast<uint4*>(pd), dbp);
}

size_t conv_s2_wpk_bytes(int Cin, int Cout, int dgrad) {
    return align_up((size_t)(dgrad ? Cout : Cin) / 16 * (dgrad ? 2 : 3) * 9 * 2 * (dgrad ? Cin : Cout) * 16, 256);
}

int conv_s2_pack_weights(cnn_ctx* ctx, const ConvS2PackJob* jobs, int n) {
    CNN_REQUIRE(n >= 0 && n <= 8, "conv_s2_pack_weights: at most 8 jobs per launch");
    if (n == 0) return CNN_OK;
    S2PackJobs j{};
    long long most = 0;
    for (int i = 0; i < n; ++i) {
        j.w[i] = jobs[i].w; j.out[i] = reinterpret_cast<uint4*>(jobs[i].out);
        j.Cin[i] = jobs[i].Cin; j.Cout[i] = jobs[i].Cout; j.dgrad[i] = jobs[i].dgrad;
        most = std::max<long long>(most, (long long)(jobs[i].dgrad ? jobs[i].Cout : jobs[i].Cin) / 16 * 9 * 2 *
                                             (jobs[i].dgrad ? jobs[i].Cin : jobs[i].Cout));
    }
    dim3 grid((unsigned)std::min<long long>(cdiv(most, 256), 64), (unsigned)n);
    CNN_LAUNCH(ctx, s2_pack_w_multi_kernel, grid, 256, 0, j);
    return CNN_OK;
}

int conv_s2_fwd_packed(cnn_ctx* ctx, const void* px, const float* w, const void* wpk_ready, const float* bias, float* y,
                       float* y_relu, int B, int Cin, int H, int W, int Cout, void* next_px) {
    const S2Geom g = make_geom(B, Cin, H, W, Cout);
    CNN_REQUIRE(!next_px || wpk_ready, "conv_s2: fused packing needs pre-packed filters");
    if (wpk_ready)
        return launch_gemm(ctx, g, false, reinterpret_cast<const uint4*>(px), reinterpret_cast<const uint4*>(wpk_ready), bias,
                           nullptr, y, y_relu, reinterpret_cast<uint4*>(next_px));
    uint8_t* scratch = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, (size_t)Cin / 16 * 3 * 9 * 2 * Cout * 16 + 256));
    CNN_REQUIRE(scratch, "scratch allocation failed");
    uint4* wpk = reinterpret_cast<uint4*>(align_up((uintptr_t)scratch, 256));
    CNN_LAUNCH(ctx, s2_pack_w_kernel, cdiv((long long)Cin / 16 * 9 * 2 * Cout, 256), 256, 0, w, wpk, Cin, Cout, 0);
    return launch_gemm(ctx, g, false, reinterpret_cast<const uint4*>(px), wpk, bias, nullptr, y, y_relu);
}

int conv_s2_dgrad_packed(cnn_ctx* ctx, const void* pd, const float* w, const void* wpk_ready, float* dx,
                         const float* relu_y, int B, int Cin, int H, int W, int Cout, bool dx_nhwc) {
    const S2Geom g = make_geom(B, Cin, H, W, Cout);
    CNN_REQUIRE(!(dx_nhwc && relu_y), "conv_s2_dgrad_packed: channel-last output has no fused ReLU mask");
    if (wpk_ready)
        return launch_gemm(ctx, g, true, reinterpret_cast<const uint4*>(pd), reinterpret_cast<const uint4*>(wpk_ready), nullptr,
                           relu_y, dx, nullptr, nullptr, dx_nhwc);
    uint8_t* scratch = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, (size_t)Cout / 16 * 2 * 9 * 2 * Cin * 16 + 256));
    CNN_REQUIRE(scratch, "scratch allocation failed");
    uint4* wpk = reinterpret_cast<uint4*>(align_up((uintptr_t)scratch, 256));
    CNN_LAUNCH(ctx, s2_pack_w_kernel, cdiv((long long)Cout / 16 * 9 * 2 * Cin, 256), 256, 0, w, wpk, Cin, Cout, 1);
    return launch_gemm(ctx, g, true, reinterpret_cast<const uint4*>(pd), wpk, nullptr, relu_y, dx, nullptr, nullptr, dx_nhwc);
}

int conv_s2_wgrad_packed(cnn_ctx* ctx, const void* px, const void* pd, const float* dbp, float* dw, float* db, int B,
                         int Cin, int H, int W, int Cout, float scale) {
    const S2Geom g = make_geom(B, Cin, H, W, Cout);
    S2Wgrad p{};
    p.g = g;
    // row groups: shifted delta copies stacked along the 128 MMA rows
    p.G = Cout <= 32 ? 4 : (Cout <= 64 ? 2 : 1);
    p.rows_per = kTile / p.G;
    const int shifts[4] = {0, 1, g.HP, g.HP + 1};
    int naddr, addr[4];
    if (p.G == 4) { naddr = 1; addr[0] = 0; for (int i = 0; i < 4; ++i) p.gshift[i] = shifts[i]; p.halo = 0; }
    else if (p.G == 2) { naddr = 2; addr[0] = 0; addr[1] = g.HP; p.gshift[0] = 0; p.gshift[1] = 1; p.halo = g.HP; }
    else { naddr = 4; for (int i = 0; i < 4; ++i) addr[i] = shifts[i]; p.gshift[0] = 0; p.halo = g.HP + 1; }
    // MMAs per K step: (plane, start-address shift) pairs that produce at least one real filter tap
    p.nmma = 0;
    for (int q = 0; q < 4; ++q)
        for (int ai = 0; ai < naddr; ++ai) {
            signed char taps[4] = {-1, -1, -1, -1};
            bool any = false;
            for (int gi = 0; gi < p.G; ++gi) {
                const int S = p.gshift[gi] + addr[ai];
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    if (((ky & 1) * 2 + (kx & 1)) == q && (ky >> 1) * g.HP + (kx >> 1) == S) { taps[gi] = (signed char)tap; any = true; }
                }
            }
            if (!any) continue;
            p.mq[p.nmma] = (signed char)q;
            p.maddr[p.nmma] = addr[ai];
            for (int gi = 0; gi < 4; ++gi) p.mtap[p.nmma][gi] = taps[gi];
            ++p.nmma;
        }
    // accumulators of cin_per columns each must fit the 512 TMEM columns
    p.cin_per = Cin;
    while (p.nmma * p.cin_per > 512) p.cin_per /= 2;
    CNN_REQUIRE(p.cin_per % 16 == 0, "conv_s2: unsupported channel split");
    const int nsplit = Cin / p.cin_per;
    p.tmem_cols = next_pow2_cols(p.nmma * p.cin_per);
    p.KT = 64;
    p.KTA = p.KT + p.halo;
    p.a_bytes = (uint32_t)2 * p.G * (p.rows_per / 8) * p.KTA * 16;
    p.b_bytes = (uint32_t)2 * 4 * (p.cin_per / 8) * p.KT * 16;
    const size_t stage = (size_t)p.a_bytes + p.b_bytes;
    int ns = 1;
    while (ns < 4 && (ns + 1) * stage + 128 <= 226 * 1024) ++ns;
    // two CTAs per SM when TMEM allows it and at least two stages each fit
    const bool two = p.tmem_cols <= 256 && 2 * (2 * stage + 128 + 1024) <= 227 * 1024;
    if (two) ns = std::min(ns, (int)((227 * 1024 / 2 - 1024 - 128) / stage));
    p.nstage = ns;
    const size_t smem = 128 + (size_t)ns * stage;
    CNN_REQUIRE(smem <= 227 * 1024, "conv_s2: weight-gradient stage does not fit in shared memory");
    if (const char* e = getenv("CNN_DBG_S2")) p.dbg = atoi(e);
    p.chunks = (int)((g.NPOS + p.KT - 1) / p.KT);   // a valid delta meets x of its own image: k < NPOS
    // CTAs: fill the machine, but keep the partial-sum traffic (one [Cout][Cin][9] block per CTA) small
    long long ctas = std::min<long long>(p.chunks, (long long)ctx->sm_count * (two ? 2 : 1) / nsplit);
    const long long cap = std::max<long long>(8, (long long)((size_t)(24 << 20) / ((size_t)p.nmma * Cin * kTile * 4)));
    ctas = std::max<long long>(1, std::min(ctas, cap));
    // accumulator-length cap (conv_tc.cu, kMaxAccPixels): at most 4096 pixels per TMEM accumulator, the
    // tensor core's truncating fp32 adds stay below ~3e-5; accuracy wins over partial-sum traffic
    ctas = std::max<long long>(ctas, ((long long)p.chunks * p.KT + 4095) / 4096);
    ctas = std::min<long long>(ctas, p.chunks);
    const int per = (p.chunks + (int)ctas - 1) / (int)ctas;
    ctas = (p.chunks + per - 1) / per;   // no empty CTA
    const unsigned pack_blocks = (unsigned)cdiv(g.RUND, 256);
    const int psize = p.nmma * Cin * kTile;
    uint8_t* scratch = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, (size_t)ctas * psize * 4 + 256));
    CNN_REQUIRE(scratch, "scratch allocation failed");
    float* partial = reinterpret_cast<float*>(align_up((uintptr_t)scratch, 256));
    p.px = reinterpret_cast<const uint4*>(px); p.pd = reinterpret_cast<const uint4*>(pd); p.partial = partial;
    if (int rc = attrs_once(ctx->device)) return rc;
    dim3 grid((unsigned)ctas, (unsigned)nsplit);
    CNN_LAUNCH(ctx, s2_wgrad_kernel, grid, kS2Threads, smem, p);
    S2ReduceMap mp{};
    mp.nmma = p.nmma; mp.Cin = Cin; mp.Cout = Cout; mp.G = p.G; mp.rows_per = p.rows_per;
    for (int u = 0; u < 9; ++u)
        for (int gi = 0; gi < 4; ++gi) mp.mtap[u][gi] = p.mtap[u][gi];
    CNN_LAUNCH(ctx, s2_wgrad_reduce_kernel, psize / 32 + cdiv(Cout, 8), 256, 0, partial, (int)ctas, psize, mp, dbp,
               (int)pack_blocks, dw, db, scale);
    return CNN_OK;
}

// ---- per-operator entry points (cnn_conv2d_*): pack into a side arena, then the packed kernels ----
namespace {
// packed operands of the stand-alone operator calls live in the context's side arena; the ctx scratch
// arena stays free for the filter blocks and partial sums of the packed kernels
uint8_t* op_arena(cnn_ctx* ctx, size_t bytes) { return reinterpret_cast<uint8_t*>(cnn_arena(ctx, bytes)); }
}  // namespace

int conv_fwd_s2(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, float* y_relu, int B,
                int Cin, int H, int W, int Cout) {
    uint8_t* px = op_arena(ctx, conv_s2_px_bytes(B, Cin, H, W));
    CNN_REQUIRE(px, "conv_s2: arena allocation failed");
    if (int rc = conv_s2_pack_x(ctx, x, px, B, Cin, H, W)) return rc;
    return conv_s2_fwd_packed(ctx, px, w, nullptr, bias, y, y_relu, B, Cin, H, W, Cout, nullptr);
}

int conv_dgrad_s2(cnn_ctx* ctx, const float* w, const float* delta, float* dx, const float* relu_y, int B, int Cin,
                  int H, int W, int Cout) {
    uint8_t* pd = op_arena(ctx, conv_s2_pd_bytes(B, Cout, H, W));
    CNN_REQUIRE(pd, "conv_s2: arena allocation failed");
    if (int rc = conv_s2_pack_d(ctx, delta, pd, nullptr, B, Cout, H, W)) return rc;
    return conv_s2_dgrad_packed(ctx, pd, w, nullptr, dx, relu_y, B, Cin, H, W, Cout);
}

int conv_wgrad_s2(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B, int Cin, int H, int W,
                  int Cout, float scale) {
    const size_t pxb = conv_s2_px_bytes(B, Cin, H, W), pdb = conv_s2_pd_bytes(B, Cout, H, W);
    uint8_t* a = op_arena(ctx, pxb + pdb + conv_s2_dbp_bytes(B, Cout, H, W));
    CNN_REQUIRE(a, "conv_s2: arena allocation failed");
    float* dbp = reinterpret_cast<float*>(a + pxb + pdb);
    if (int rc = conv_s2_pack_x(ctx, x, a, B, Cin, H, W)) return rc;
    if (int rc = conv_s2_pack_d(ctx, delta, a + pxb, dbp, B, Cout, H, W)) return rc;
    return conv_s2_wgrad_packed(ctx, a, a + pxb, dbp, dw, db, B, Cin, H, W, Cout, scale);
}
