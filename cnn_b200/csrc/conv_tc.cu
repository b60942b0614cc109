// conv_tc.cu -- Conv2D forward, input gradient and weight gradient as implicit GEMMs on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Numerics.  Blackwell has no IEEE-fp32 MMA and single-pass TF32 misses the 1e-4 parity bar
// (SURVEY §7 hard part 1).  Every fp32 operand is split in two (x = hi + lo) and each K-step
// issues three MMAs (lo*hi + hi*lo + hi*hi) into one fp32 TMEM accumulator:
//   TF32 mode (default): hi = rna_tf32(x), lo = x - hi (exact)  -> ~2^-22 per product, fp32-grade
//   BF16 mode          : hi = bf16(x), lo = bf16(x - hi)        -> ~2^-16 per product, 2x MMA rate
//
// Forward / input gradient share one persistent, warp-specialised "gather GEMM" kernel
//   D[m][n] = sum_k A[m][k] * Bw[n][k]
//   forward : m = output pixel (b,oy,ox), n = out channel, k = (ci,ky,kx) in the reference's
//             filter order; A[m][k] = x[b][ci][oy*s+ky][ox*s+kx]              (conv2d.cpp:69-92)
//   dgrad   : m = s x s input patch (b,py,px); one GEMM ("segment") per patch cell (pr,pc), each in
//             its own TMEM column range; n = in channel, k = (co, tap with ky%s==pr, kx%s==pc);
//             A[m][k] = delta[b][co][py-ky/s][px-kx/s] or 0 outside  (gather form of conv2d.cpp:192)
// Roles per CTA (416 threads, grid = resident CTAs, tiles strided over the grid):
//   warps 0-7   producers: gather one 128-row x 128-byte operand block per stage from NCHW global
//               memory (lane = pixel -> every load instruction is coalesced along W), split it and
//               store 16-byte chunks into the 128B-swizzled K-major tiles; fence.proxy.async; arrive
//   warp  12    one thread: TMA bulk copies (cp.async.bulk) of the pre-split, pre-swizzled filter
//               blocks into the stage, the tcgen05.mma chain, tcgen05.commit back to the producers
//   warps 8-11  epilogue: tcgen05.ld 32x32b (lane = row) from the double-buffered accumulator,
//               + bias, coalesced NCHW stores, while the next tile is already being gathered.
//
// Weight gradient: D[kidx][co] = sum_pixels A[kidx][p] * Bd[co][p] with both operands gathered
// (x taps and delta are pixel-contiguous in NCHW = K-major for this GEMM); an extra all-ones
// row of A yields the bias gradient; CTAs split the pixel range, partials are reduced in a
// fixed order (deterministic) by wgrad_reduce_kernel which also applies the 1/B scale.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kRows = 128;         // GEMM rows per tile == TMEM lanes
constexpr int kProducers = 256;    // warps 0-7
constexpr int kEpiWarp0 = 8;       // warps 8-11
constexpr int kMmaWarp = 12;
constexpr int kTmaWarp = 13;       // filter / row streamer
constexpr int kThreads = 14 * 32;
constexpr int kMaxStages = 6;
constexpr int kMaxAcc = 8;          // TMEM accumulator ring (tiles in flight between MMA and epilogue)
constexpr int kPadCode = 15;
constexpr long long kMaxAccPixels = 4096;   // reduction length per TMEM accumulator in the weight gradient

template <bool TF32>
struct Op {
    static constexpr int KBLK = TF32 ? 32 : 64;   // K elements per 128-byte row
    static constexpr int EPC = TF32 ? 4 : 8;      // elements per 16-byte chunk
    static constexpr int KSTEP = TF32 ? 8 : 16;   // K per tcgen05.mma
};

template <bool TF32>
__device__ __forceinline__ void mma_issue(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, bool acc) {
    if constexpr (TF32) mma_tf32(d, a, b, idesc, acc);
    else mma_bf16(d, a, b, idesc, acc);
}

// hi*hi + hi*lo + lo*hi over `ksteps` K-steps of one stage (small terms first)
template <bool TF32>
__device__ __forceinline__ void mma_block(uint32_t d, uint32_t sa_hi, uint32_t sa_lo, uint32_t sb_hi,
                                          uint32_t sb_lo, int ksteps, uint32_t idesc, bool first_acc) {
    for (int j = 0; j < ksteps; ++j) {
        const uint64_t ahi = smem_desc_k128(sa_hi + 32 * j), alo = smem_desc_k128(sa_lo + 32 * j);
        const uint64_t bhi = smem_desc_k128(sb_hi + 32 * j), blo = smem_desc_k128(sb_lo + 32 * j);
        mma_issue<TF32>(d, alo, bhi, idesc, first_acc || j != 0);
        mma_issue<TF32>(d, ahi, blo, idesc, true);
        mma_issue<TF32>(d, ahi, bhi, idesc, true);
    }
}

// split EPC gathered values into one hi and one lo 16-byte chunk
template <bool TF32>
__device__ __forceinline__ void split_chunk(const float* v, uint4& hi, uint4& lo) {
    if constexpr (TF32) {
        split_tf32(v[0], hi.x, lo.x);
        split_tf32(v[1], hi.y, lo.y);
        split_tf32(v[2], hi.z, lo.z);
        split_tf32(v[3], hi.w, lo.w);
    } else {
        split2(v[0], v[1], hi.x, lo.x);
        split2(v[2], v[3], hi.y, lo.y);
        split2(v[4], v[5], hi.z, lo.z);
        split2(v[6], v[7], hi.w, lo.w);
    }
}

// =============================================================================================
//                                forward / input gradient
// =============================================================================================
struct GatherGemm {
    const float* src;        // activations (x or delta), NCHW
    const int* table;        // forward: element offset per k; dgrad: (offset << 4) | position code (15 = pad)
    const uint8_t* packedB;  // [ntile][kb]{hi[Ntile][128B], lo[Ntile][128B]}, swizzled
    const float* bias;       // may be null
    float* dst;
    int K, KB;               // reduction length and K blocks
    int npos;                // dgrad: source positions (dy,dx) per delta channel (validity mask bits)
    signed char dy[9], dx[9];
    int GH, GW;              // row space: m -> (b, gy, gx)
    unsigned rows, mtiles;   // B*GH*GW, ceil(rows/128)
    int SC, SH, SW, sy, sx;  // source geometry
    // GEMM column n' = cell*ON + n, cell = pr*os + pc -> dst[b][n][gy*os+pr][gx*os+pc]
    int ON, OHt, OWt, os, Nreal;
    int Ntile;               // columns per CTA (multiple of 16, <= 256)
    int tmem_cols;           // power of two >= accbufs*Ntile
    int stages, accbufs;
    int bres;                // 1: all filter blocks of an n-tile stay resident in smem (loaded once per CTA)
    int single;              // 1: single-pass operands (hi * hi only; CNN_TC_BF16X1 = bf16 inputs, fp32 accumulate)
    int dbg;                 // experiment knobs (CNN_DBG_SKIP): 1 no epilogue stores, 2 no gathers, 4 no MMA
    long long* trace;        // CNN_DBG_TRACE: [3 roles][64 tiles][2] clock64 stamps of CTA 0
};

// Forward: Bw[n][k] = w[n][k] (k = ci*kk + tap is exactly the reference's filter memory order).
// Input gradient (patch cells folded into the GEMM columns): Bw[n' = cell*Cin + ci][k = co*npos + pos]
// = w[co][ci][dy*s+pr][dx*s+pc] if that tap exists, else 0.
struct PackInfo {
    int K, KB, npos, s, k;
    signed char dy[9], dx[9];
};

template <bool TF32>
__global__ void pack_filters_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, PackInfo pi,
                                    int dgrad, int Cin, int Cout, int Nreal, int Ntile, int ntiles) {
    using O = Op<TF32>;
    const int kk = pi.k * pi.k;
    const long long total = (long long)ntiles * pi.KB * Ntile * 8;  // one thread per 16-byte chunk
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(id & 7);
        long long t = id >> 3;
        const int nl = (int)(t % Ntile);
        t /= Ntile;
        const int kb = (int)(t % pi.KB);
        const int nt = (int)(t / pi.KB);
        const int n = nt * Ntile + nl;
        float v[O::EPC];
#pragma unroll
        for (int j = 0; j < O::EPC; ++j) {
            const int k = kb * O::KBLK + c * O::EPC + j;
            float val = 0.f;
            if (n < Nreal && k < pi.K) {
                if (!dgrad) {
                    val = w[(size_t)n * Cin * kk + k];
                } else {
                    const int cell = n / Cin, ci = n % Cin, co = k / pi.npos, pos = k % pi.npos;
                    const int ky = pi.dy[pos] * pi.s + cell / pi.s, kx = pi.dx[pos] * pi.s + cell % pi.s;
                    if (ky < pi.k && kx < pi.k) val = w[((size_t)co * Cin + ci) * kk + ky * pi.k + kx];
                }
            }
            v[j] = val;
        }
        uint4 hi, lo;
        split_chunk<TF32>(v, hi, lo);
        uint8_t* blk = out + ((size_t)nt * pi.KB + kb) * (size_t)(2 * Ntile * 128);
        *reinterpret_cast<uint4*>(blk + swz128(nl, c)) = hi;
        *reinterpret_cast<uint4*>(blk + (size_t)Ntile * 128 + swz128(nl, c)) = lo;
    }
}

struct SmemCarve {
    uint64_t* full;       // [stages]  producers -> MMA (one arrive per producer warp)
    uint64_t* bfull;      // [stages]  TMA filter block landed
    uint64_t* free_;      // [stages]  MMA done with the stage
    uint64_t* acc_full;   // [2]
    uint64_t* acc_empty;  // [2]       epilogue done with the accumulator (one arrive per warp)
    uint32_t* tmem_slot;
    int* rowtab;          // [128] wgrad row table
    uint8_t* ops;         // 1024-aligned operand stages
};

__device__ __forceinline__ SmemCarve carve(uint8_t* raw) {
    SmemCarve c;
    uint64_t* b = reinterpret_cast<uint64_t*>(raw);
    c.full = b;
    c.bfull = b + kMaxStages;
    c.free_ = b + 2 * kMaxStages;
    c.acc_full = b + 3 * kMaxStages;
    c.acc_empty = c.acc_full + kMaxAcc;
    c.tmem_slot = reinterpret_cast<uint32_t*>(c.acc_empty + kMaxAcc);
    c.rowtab = reinterpret_cast<int*>(raw + 512);
    const uint32_t base = smem_u32(raw);
    c.ops = raw + (((base + 1024 + 1023) & ~1023u) - base);
    return c;
}

constexpr int kProdWarps = kProducers / 32;

// MASKED: dgrad (taps may fall outside delta).  HOIST: the whole gather table of a thread fits in
// registers (KB == 1), so it is loaded once per kernel instead of once per tile.
template <bool TF32, bool MASKED, bool HOIST>
__global__ void __launch_bounds__(kThreads, 2) gather_gemm_ws(const GatherGemm g) {
    using O = Op<TF32>;
    extern __shared__ uint8_t smem_raw[];
    const SmemCarve sm = carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = g.stages, Ntile = g.Ntile, nt = blockIdx.y, KB = g.KB;
    const uint32_t a_bytes = kRows * 128, b_bytes = (uint32_t)Ntile * 128;
    const uint32_t stage_bytes = 2 * a_bytes + (g.bres ? 0u : 2 * b_bytes);
    const uint32_t tail_bytes = g.bres ? (uint32_t)KB * 2 * b_bytes : 0u;   // resident filters behind the stages

    if (warp == kMmaWarp) {
        tmem_alloc(sm.tmem_slot, (uint32_t)g.tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < S; ++i) {
                mbar_init(&sm.full[i], kProdWarps);
                mbar_init(&sm.bfull[i], 1);
                mbar_init(&sm.free_[i], 1);
            }
            for (int i = 0; i < kMaxAcc; ++i) {
                mbar_init(&sm.acc_full[i], 1);
                mbar_init(&sm.acc_empty[i], 4);
            }
            mbar_fence_init();
        }
    }
    {   // bias (or zeros) for the epilogue
        float* sb = reinterpret_cast<float*>(sm.ops + (size_t)S * stage_bytes + tail_bytes);
        for (int i = tid; i < g.ON; i += kThreads) sb[i] = g.bias ? g.bias[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    const unsigned my_tiles = (g.mtiles > blockIdx.x) ? (g.mtiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    float* s_bias = reinterpret_cast<float*>(sm.ops + (size_t)S * stage_bytes + tail_bytes);  // [ON]

    if (warp < kEpiWarp0) {
        // ------------------------------------------------------------------ producers
        // Thread (row, half) gathers the chunks 2j+half of its row for the K-steps j of a block.
        // The loop over (tile, kb) is flattened and register double-buffered: the gathers of block
        // i+1 are in flight while block i is split and stored.
        const int row = tid & 127, half = tid >> 7;
        constexpr int KSB = TF32 ? 4 : 2;        // K-steps per batch (16 gathered values per thread)
        constexpr int NBK = 4 / KSB;             // batches per K block
        static_assert(!HOIST || NBK == 1, "table hoisting needs one batch per K block");
        const unsigned n_it = my_tiles * (unsigned)(KB * NBK);
        // issue-side cursor
        unsigned tile_i = blockIdx.x;
        int kb_i = 0, bi_i = 0;
        const float* row_i = g.src;
        uint32_t vmask_i = 0;
        int safe_i = 0;
        // row coordinates (b, gy, gx) of this thread's row, advanced tile by tile without divisions:
        // the per-step row delta gridDim.x*128 is decomposed once into (db, dgy, dgx)
        int cb, cgy, cgx;
        {
            const unsigned m0 = blockIdx.x * kRows + row;
            cgx = (int)(m0 % (unsigned)g.GW);
            const unsigned t0 = m0 / (unsigned)g.GW;
            cgy = (int)(t0 % (unsigned)g.GH);
            cb = (int)(t0 / (unsigned)g.GH);
        }
        const unsigned dm = gridDim.x * kRows;
        const int dgx = (int)(dm % (unsigned)g.GW), dgy = (int)((dm / (unsigned)g.GW) % (unsigned)g.GH),
                  db = (int)((dm / (unsigned)g.GW) / (unsigned)g.GH);
        auto step_coords = [&]() {
            cgx += dgx;
            if (cgx >= g.GW) { cgx -= g.GW; ++cgy; }
            cgy += dgy;
            if (cgy >= g.GH) { cgy -= g.GH; ++cb; }
            cb += db;
        };
        auto decode = [&]() {   // uses (cb, cgy, cgx) of tile_i
            const unsigned m = tile_i * kRows + row;
            int base = 0;
            vmask_i = 0;
            if (m < g.rows) {
                base = cb * g.SC * g.SH * g.SW + (cgy * g.sy) * g.SW + cgx * g.sx;  // < 2^31 (host check)
                if constexpr (MASKED) {
                    for (int t2 = 0; t2 < g.npos; ++t2) {
                        const int yy = cgy - g.dy[t2], xx = cgx - g.dx[t2];
                        if (yy >= 0 && yy < g.SH && xx >= 0 && xx < g.SW) vmask_i |= 1u << t2;
                    }
                }
            }
            row_i = g.src + base;
            safe_i = -base;
        };
        int ereg[KSB][O::EPC];  // HOIST: this thread's table entries
        auto load_table = [&](int kb, int bi, int (&e)[KSB][O::EPC]) {
            const int* tb = g.table + kb * O::KBLK;
#pragma unroll
            for (int jj = 0; jj < KSB; ++jj) {
                const int j = bi * KSB + jj;
                const int4* t4 = reinterpret_cast<const int4*>(tb + (2 * j + half) * O::EPC);
#pragma unroll
                for (int q = 0; q < O::EPC / 4; ++q) {
                    const int4 t = __ldg(t4 + q);  // blocks are padded to KBLK entries: always readable
                    e[jj][4 * q] = t.x; e[jj][4 * q + 1] = t.y; e[jj][4 * q + 2] = t.z; e[jj][4 * q + 3] = t.w;
                }
            }
        };
        if constexpr (HOIST) load_table(0, 0, ereg);
        auto issue = [&](float (&v)[KSB][O::EPC]) {
            int eloc[KSB][O::EPC];
            if constexpr (!HOIST) load_table(kb_i, bi_i, eloc);
#pragma unroll
            for (int jj = 0; jj < KSB; ++jj)
#pragma unroll
                for (int q = 0; q < O::EPC; ++q) {
                    const int e = HOIST ? ereg[jj][q] : eloc[jj][q];
                    if (g.dbg & 2) { v[jj][q] = (float)e; continue; }
                    if constexpr (MASKED) {
                        // branch-free: out-of-range taps read a safe location and are zeroed afterwards
                        const bool ok = (vmask_i >> (e & 15)) & 1u;
                        const float x = __ldg(row_i + (ok ? (e >> 4) : safe_i));
                        v[jj][q] = ok ? x : 0.f;
                    } else {
                        // forward: every entry (K padding = offset 0) is a readable cell of this row's
                        // window; padded K columns meet zero filter columns, invalid rows are not stored
                        v[jj][q] = __ldg(row_i + e);
                    }
                }
            // advance the issue cursor
            if (++bi_i == NBK) {
                bi_i = 0;
                if (++kb_i == KB) {
                    kb_i = 0;
                    tile_i += gridDim.x;
                    step_coords();
                    if (tile_i < g.mtiles) decode();
                }
            }
        };
        // consume-side cursor
        int kb_c = 0, bi_c = 0;
        unsigned it = 0;  // K-block counter (stage ring position)
        uint32_t ps = 0, pph = 0;  // stage ring position / phase (division-free)
        auto consume = [&](float (&v)[KSB][O::EPC]) {
            const unsigned s = ps;
            uint8_t* stage = sm.ops + (size_t)s * stage_bytes;
            if (bi_c == 0 && it >= (unsigned)S) mbar_wait(&sm.free_[s], pph ^ 1);
            if (g.trace && blockIdx.x == 0 && tid == 0 && it < 64) g.trace[(0 * 64 + it) * 2] = clock64();
            if (g.trace && blockIdx.x == 0 && tid == 224 && it < 64) g.trace[(5 * 64 + it) * 2] = clock64();
            const int kvalid = min(O::KBLK, g.K - kb_c * O::KBLK);
            const int ksteps = (kvalid + O::KSTEP - 1) / O::KSTEP;
#pragma unroll
            for (int jj = 0; jj < KSB; ++jj) {
                const int j = bi_c * KSB + jj;
                if (j < ksteps && !(g.dbg & 64)) {
                    uint4 hi, lo;
                    split_chunk<TF32>(v[jj], hi, lo);
                    const uint32_t o = swz128(row, 2 * j + half);
                    *reinterpret_cast<uint4*>(stage + o) = hi;
                    *reinterpret_cast<uint4*>(stage + a_bytes + o) = lo;
                }
            }
            if (++bi_c == NBK) {
                bi_c = 0;
                if (++kb_c == KB) kb_c = 0;
                if (!(g.dbg & 16)) fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.full[s]);
                if (g.trace && blockIdx.x == 0 && tid == 0 && it < 64) g.trace[(0 * 64 + it) * 2 + 1] = clock64();
                if (g.trace && blockIdx.x == 0 && tid == 224 && it < 64) g.trace[(5 * 64 + it) * 2 + 1] = clock64();
                ++it;
                if (++ps == (uint32_t)S) { ps = 0; pph ^= 1; }
            }
        };
        if (n_it > 0) {
            float va[KSB][O::EPC], vb[KSB][O::EPC];
            decode();
            issue(va);
            for (unsigned i = 0;; i += 2) {
                if (i + 1 < n_it) issue(vb);
                consume(va);
                if (i + 1 >= n_it) break;
                if (i + 2 < n_it) issue(va);
                consume(vb);
                if (i + 2 >= n_it) break;
            }
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the (uniform) loop and one elected lane issues: with warp-uniform
        // control flow the descriptors stay in uniform registers and each tcgen05.mma is a single
        // UTCHMMA (under `lane == 0` the compiler wraps every MMA in an elect/branch loop).
        const uint32_t idesc = TF32 ? idesc_tf32(kRows, Ntile) : idesc_bf16(kRows, Ntile);
        const uint32_t ops_u32 = smem_u32(sm.ops);
        const uint32_t bres_u32 = ops_u32 + (uint32_t)S * stage_bytes;   // resident filter blocks
        const uint64_t desc_hi = smem_desc_k128(0);                      // all fields but the address
        const int KBm1_steps = ((min(O::KBLK, g.K - (KB - 1) * O::KBLK)) + O::KSTEP - 1) / O::KSTEP;
        if (g.bres) mbar_wait(&sm.bfull[0], 0);
        uint32_t s = 0, ph = 0, sa = ops_u32;
        uint32_t a = 0, aph = 0;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            if (ti >= (unsigned)g.accbufs) mbar_wait(&sm.acc_empty[a], aph ^ 1);
            const uint32_t d = tmem_base + a * Ntile;
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait(&sm.full[s], ph);
                uint32_t sb;
                if (g.bres) sb = bres_u32 + (uint32_t)kb * 2 * b_bytes;
                else { mbar_wait(&sm.bfull[s], ph); sb = sa + 2 * a_bytes; }
                tc_fence_after();
                const int ksteps = (kb == KB - 1) ? KBm1_steps : O::KBLK / O::KSTEP;
                if (elect_one()) {
                    if (!(g.dbg & 4)) {
                        // descriptor = constant high part | (address >> 4); one K-step = +32 bytes = +2
                        uint64_t ahi = desc_hi | ((sa & 0x3FFFFu) >> 4), alo = desc_hi | (((sa + a_bytes) & 0x3FFFFu) >> 4);
                        uint64_t bhi = desc_hi | ((sb & 0x3FFFFu) >> 4), blo = desc_hi | (((sb + b_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
                        for (int j = 0; j < O::KBLK / O::KSTEP; ++j) {
                            if (j < ksteps) {
                                if (g.single) {
                                    mma_issue<TF32>(d, ahi, bhi, idesc, (kb | j) != 0);
                                } else {
                                    mma_issue<TF32>(d, alo, bhi, idesc, (kb | j) != 0);
                                    mma_issue<TF32>(d, ahi, blo, idesc, true);
                                    mma_issue<TF32>(d, ahi, bhi, idesc, true);
                                }
                                ahi += 2; alo += 2; bhi += 2; blo += 2;
                            }
                        }
                    }
                    mma_commit(&sm.free_[s]);
                    if (kb == KB - 1) mma_commit(&sm.acc_full[a]);
                }
                __syncwarp();
                sa += stage_bytes;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; sa = ops_u32; }
            }
            if (++a == (uint32_t)g.accbufs) { a = 0; aph ^= 1; }
        }
    } else if (warp == kTmaWarp) {
        // ------------------------------------------------------------------ TMA filter streamer
        if (lane == 0) {
            const uint8_t* bsrc0 = g.packedB + (size_t)nt * KB * (size_t)(2 * b_bytes);
            if (g.bres) {   // the whole filter matrix of this n-tile stays in shared memory
                mbar_expect_tx(&sm.bfull[0], (uint32_t)KB * 2 * b_bytes);
                for (int kb = 0; kb < KB; ++kb)
                    tma_bulk_g2s(sm.ops + (size_t)S * stage_bytes + (size_t)kb * (2 * b_bytes),
                                 bsrc0 + (size_t)kb * (2 * b_bytes), 2 * b_bytes, &sm.bfull[0]);
            } else {
                uint32_t ts = 0, tph = 0, tkb = 0;
                const unsigned n_blk = my_tiles * (unsigned)KB;
                for (unsigned i = 0; i < n_blk; ++i) {
                    if (i >= (unsigned)S) mbar_wait(&sm.free_[ts], tph ^ 1);
                    mbar_expect_tx(&sm.bfull[ts], 2 * b_bytes);
                    tma_bulk_g2s(sm.ops + (size_t)ts * stage_bytes + 2 * a_bytes, bsrc0 + (size_t)tkb * (2 * b_bytes),
                                 2 * b_bytes, &sm.bfull[ts]);
                    if (++tkb == (uint32_t)KB) tkb = 0;
                    if (++ts == (uint32_t)S) { ts = 0; tph ^= 1; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int wq = warp - kEpiWarp0;
        const int row = wq * 32 + lane;
        const size_t oplane = (size_t)g.OHt * g.OWt;
        unsigned tile = blockIdx.x;
        int b, gy, gx;
        {
            const unsigned m0 = blockIdx.x * kRows + row;
            gx = (int)(m0 % (unsigned)g.GW);
            const unsigned t0 = m0 / (unsigned)g.GW;
            gy = (int)(t0 % (unsigned)g.GH);
            b = (int)(t0 / (unsigned)g.GH);
        }
        const unsigned dm = gridDim.x * kRows;
        const int dgx = (int)(dm % (unsigned)g.GW), dgy = (int)((dm / (unsigned)g.GW) % (unsigned)g.GH),
                  db = (int)((dm / (unsigned)g.GW) / (unsigned)g.GH);
        uint32_t a = 0, aph = 0;
        for (unsigned ti = 0; ti < my_tiles; ++ti, tile += gridDim.x) {
            const unsigned m = tile * kRows + row;
            const bool row_ok = m < g.rows;
            float* dimg = g.dst + (size_t)(row_ok ? b : 0) * g.ON * oplane;
            mbar_wait(&sm.acc_full[a], aph);
            if (g.trace && blockIdx.x == 0 && tid == kEpiWarp0 * 32 && ti < 64) g.trace[(2 * 64 + ti) * 2] = clock64();
            tc_fence_after();
            // column n' = cell*ON + n ; walk (cell, n) incrementally
            int n = (nt * Ntile) % g.ON;
            const int cell0 = (nt * Ntile) / g.ON;
            int pr = cell0 / g.os, pc = cell0 % g.os;
            for (int c0 = 0; c0 < Ntile; c0 += 16) {
                float v[16];
                if (!(g.dbg & 32)) tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + a * Ntile + c0, v);
                else
                    for (int j = 0; j < 16; ++j) v[j] = (float)j;
                if (g.trace && blockIdx.x == 0 && tid == kEpiWarp0 * 32 && ti < 64 && c0 == 0) g.trace[(3 * 64 + ti) * 2] = clock64();
                if (c0 + 16 >= Ntile) {
                    // the accumulator is in registers: hand the TMEM buffer back BEFORE the global
                    // stores (a tcgen05 fence after them would wait for the stores to drain)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.acc_empty[a]);
                    if (g.trace && blockIdx.x == 0 && tid == kEpiWarp0 * 32 && ti < 64) g.trace[(3 * 64 + ti) * 2 + 1] = clock64();
                }
                if (g.os == 1) {
                    // forward: column == channel; one pointer, stride = one output plane
                    const int nb = nt * Ntile + c0;
                    float* o = dimg + (size_t)nb * oplane + (size_t)gy * g.OWt + gx;
                    if (row_ok && !(g.dbg & 1)) {
                        if (nb + 16 <= g.ON) {   // full chunk: 4 LDS.128 of bias, pointer bumps, no predicates
                            const float4* b4 = reinterpret_cast<const float4*>(s_bias + nb);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 bb = b4[j4];
                                o[0] = v[4 * j4] + bb.x; o += oplane;
                                o[0] = v[4 * j4 + 1] + bb.y; o += oplane;
                                o[0] = v[4 * j4 + 2] + bb.z; o += oplane;
                                o[0] = v[4 * j4 + 3] + bb.w; o += oplane;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (nb + j < g.ON) o[(size_t)j * oplane] = v[j] + s_bias[nb + j];
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int oy = gy * g.os + pr, ox = gx * g.os + pc;
                        if (row_ok && nt * Ntile + c0 + j < g.Nreal && oy < g.OHt && ox < g.OWt && !(g.dbg & 1))
                            dimg[(size_t)n * oplane + (size_t)oy * g.OWt + ox] = v[j] + s_bias[n];
                        if (++n == g.ON) {
                            n = 0;
                            if (++pc == g.os) { pc = 0; ++pr; }
                        }
                    }
                }
            }
            if (g.trace && blockIdx.x == 0 && tid == kEpiWarp0 * 32 && ti < 64) g.trace[(2 * 64 + ti) * 2 + 1] = clock64();
            if (++a == (uint32_t)g.accbufs) { a = 0; aph ^= 1; }
            gx += dgx;
            if (gx >= g.GW) { gx -= g.GW; ++gy; }
            gy += dgy;
            if (gy >= g.GH) { gy -= g.GH; ++b; }
            b += db;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}


// ---------------------------------------------------------------------------------------------
// Row-staged forward / input gradient (TF32x3).  Same GEMM as gather_gemm_ws, but a tile is TR whole
// rows of the row space of ONE image (<= 128 pixels / patches), so everything it gathers lies in one
// contiguous run of source rows per source channel.  A TMA warp streams those runs into shared
// memory with cp.async.bulk (double-buffered, one tile ahead; copies start at the preceding 16-byte
// boundary and the element shift is folded into a per-tile offset table written next to the buffer),
// and the producers build the swizzled hi/lo tiles from shared memory: no global-load latency, no
// per-tile coordinate arithmetic (a thread's pixel inside the tile never changes), and the filters
// stay resident.  Used where a tile is reasonably full and two CTAs fit per SM; the global gather
// kernel above covers the rest.
struct RowsGather {
    GatherGemm g;            // geometry and tiling (g.table: static (ch << 20) | (off << 4) | pos)
    int TR, SCI;             // row-space rows per tile, tiles per image
    int seg;                 // bytes per staged channel segment (multiple of 16)
    int padt, padl;          // dgrad: delta rows / columns reached above / left of the tile
    int kext;                // source rows a row-space row reaches (forward: k, dgrad: padt + 1)
    int Kpad;                // KB * 32
    int nacc;                // 3: the split terms lo*hi, hi*lo, hi*hi accumulate in separate TMEM columns
    unsigned tiles;          // B * SCI
    long long src_bytes16;   // source tensor size rounded up to 16 bytes (copy clamp)
};

constexpr int kRgThreads = 14 * 32;   // 8 producer warps, 4 epilogue warps, MMA warp, TMA warp
constexpr int kRgTmaWarp = 13;

template <bool MASKED>
__global__ void __launch_bounds__(kRgThreads, 2) gather_rows_ws(const RowsGather rg) {
    using O = Op<true>;
    const GatherGemm& g = rg.g;
    extern __shared__ uint8_t smem_raw[];
    const SmemCarve sm = carve(smem_raw);
    uint64_t* raw_full = sm.bfull + 2;    // [2]
    uint64_t* raw_free = sm.bfull + 4;    // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = g.stages, Ntile = g.Ntile, KB = g.KB;
    auto WAIT = [&](uint64_t* bar, uint32_t parity) {
        if (g.dbg & 32) { while (!mbar_test(bar, parity)) {} }
        else mbar_wait(bar, parity);
    };
    const uint32_t a_bytes = kRows * 128, b_bytes = (uint32_t)Ntile * 128;
    const uint32_t stage_bytes = 2 * a_bytes;
    uint8_t* bres = sm.ops + (size_t)S * stage_bytes;                       // resident filters
    float* s_bias = reinterpret_cast<float*>(bres + (size_t)KB * 2 * b_bytes);
    const uint32_t bias_bytes = ((uint32_t)g.ON * 4 + 127) & ~127u;
    const uint32_t raw_bytes = (uint32_t)(g.SC * rg.seg);
    uint8_t* raw0 = reinterpret_cast<uint8_t*>(s_bias) + bias_bytes;
    int* etab = reinterpret_cast<int*>(raw0 + 2 * (size_t)raw_bytes);      // [2][Kpad]

    if (warp == kMmaWarp) {
        tmem_alloc(sm.tmem_slot, (uint32_t)g.tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < S; ++i) {
                mbar_init(&sm.full[i], kProdWarps);
                mbar_init(&sm.free_[i], 1);
            }
            mbar_init(&sm.bfull[0], 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&raw_full[i], 1);
                mbar_init(&raw_free[i], kProdWarps);
            }
            for (int i = 0; i < kMaxAcc; ++i) {
                mbar_init(&sm.acc_full[i], 1);
                mbar_init(&sm.acc_empty[i], 4);
            }
            mbar_fence_init();
        }
    }
    for (int i = tid; i < g.ON; i += kRgThreads) s_bias[i] = g.bias ? g.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;
    const unsigned my_tiles = (rg.tiles > blockIdx.x) ? (rg.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // tile -> (image, row group) without divisions in the loops
    const int step_b = (int)(gridDim.x / (unsigned)rg.SCI), step_gi = (int)(gridDim.x % (unsigned)rg.SCI);
    int tb = (int)(blockIdx.x / (unsigned)rg.SCI), tgi = (int)(blockIdx.x % (unsigned)rg.SCI);
    auto next_tile = [&]() {
        tgi += step_gi;
        if (tgi >= rg.SCI) { tgi -= rg.SCI; ++tb; }
        tb += step_b;
    };
    const int last_rows = g.GH - (rg.SCI - 1) * rg.TR;

    if (warp < kEpiWarp0) {
        // ------------------------------------------------------------------ producers (smem -> smem)
        const int row = tid & 127, half = tid >> 7;
        const bool in_tile = row < rg.TR * g.GW;
        const int oyl = in_tile ? row / g.GW : 0;
        const int gx = in_tile ? row - oyl * g.GW : 0;
        const int base_pix = (oyl * g.sy) * g.SW + gx * g.sx;
        uint32_t xmask = 0;
        if constexpr (MASKED) {
            for (int t = 0; t < g.npos; ++t) {
                const int xx = gx - g.dx[t];
                if (in_tile && xx >= 0 && xx < g.SW) xmask |= 1u << t;
            }
        }
        uint32_t ps = 0, pph = 0, it = 0;
        const int last_steps = ((g.K - (KB - 1) * O::KBLK) + O::KSTEP - 1) / O::KSTEP;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const uint32_t buf = ti & 1;
            const float* raw = reinterpret_cast<const float*>(raw0 + (size_t)buf * raw_bytes);
            const int* et = etab + buf * rg.Kpad;
            uint32_t vmask = xmask;
            if constexpr (MASKED) {
                const int nrows = (tgi == rg.SCI - 1) ? last_rows : rg.TR;
                const int gy = tgi * rg.TR + oyl;
                for (int t = 0; t < g.npos; ++t) {
                    const int yy = gy - g.dy[t];
                    if (!(yy >= 0 && yy < g.SH)) vmask &= ~(1u << t);
                }
                if (oyl >= nrows) vmask = 0;
            }
            WAIT(&raw_full[buf], (ti >> 1) & 1);
            if (g.trace && blockIdx.x == 0 && tid == 0 && ti < 64) g.trace[(1 * 64 + ti) * 2 + 0] = clock64();
            for (int kb = 0; kb < KB; ++kb, ++it) {
                uint8_t* stage = sm.ops + (size_t)ps * stage_bytes;
                if (it >= (uint32_t)S) WAIT(&sm.free_[ps], pph ^ 1);
            if (g.trace && blockIdx.x == 0 && tid == 0 && it < 64) g.trace[(0 * 64 + it) * 2 + 0] = clock64();
                const int ksteps = (kb == KB - 1) ? last_steps : O::KBLK / O::KSTEP;
                const int4* e4 = reinterpret_cast<const int4*>(et + kb * O::KBLK) + half;
#pragma unroll
                for (int j = 0; j < O::KBLK / O::KSTEP; ++j) {
                    if (j < ksteps && !(g.dbg & 2)) {
                        const int4 e = e4[2 * j];
                        const int ee[4] = {e.x, e.y, e.z, e.w};
                        float v[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if constexpr (MASKED) {
                                const bool ok = (vmask >> (ee[q] & 15)) & 1u;
                                const float x = raw[ok ? base_pix + (ee[q] >> 4) : 0];
                                v[q] = ok ? x : 0.f;
                            } else {
                                v[q] = raw[base_pix + (ee[q] >> 4)];
                            }
                        }
                        uint4 hi, lo;
                        split_chunk<true>(v, hi, lo);
                        const uint32_t o = swz128(row, 2 * j + half);
                        *reinterpret_cast<uint4*>(stage + o) = hi;
                        *reinterpret_cast<uint4*>(stage + a_bytes + o) = lo;
                    }
                }
                if (!(g.dbg & 16)) fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.full[ps]);
            if (g.trace && blockIdx.x == 0 && tid == 0 && it < 64) g.trace[(0 * 64 + it) * 2 + 1] = clock64();
                if (++ps == (uint32_t)S) { ps = 0; pph ^= 1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_free[buf]);
            if (g.trace && blockIdx.x == 0 && tid == 0 && ti < 64) g.trace[(1 * 64 + ti) * 2 + 1] = clock64();
            next_tile();
        }
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the (uniform) loop; one elected lane issues.  Keeping the control flow
        // warp-uniform lets descriptors live in uniform registers (see umma::elect_one).
        const uint32_t idesc = idesc_tf32(kRows, Ntile);
        const uint32_t ops_u32 = smem_u32(sm.ops);
        const uint32_t bres_u32 = smem_u32(bres);
        const uint64_t desc_hi = smem_desc_k128(0);
        const int last_steps = ((g.K - (KB - 1) * O::KBLK) + O::KSTEP - 1) / O::KSTEP;
        if (elect_one()) {
            mbar_expect_tx(&sm.bfull[0], (uint32_t)KB * 2 * b_bytes);
            for (int kb = 0; kb < KB; ++kb)
                tma_bulk_g2s(bres + (size_t)kb * (2 * b_bytes), g.packedB + (size_t)kb * (2 * b_bytes), 2 * b_bytes,
                             &sm.bfull[0]);
        }
        __syncwarp();
        WAIT(&sm.bfull[0], 0);
        uint32_t s = 0, ph = 0, sa = ops_u32, a = 0, aph = 0;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            if (ti >= (unsigned)g.accbufs) WAIT(&sm.acc_empty[a], aph ^ 1);
            if (g.trace && blockIdx.x == 0 && lane == 0 && ti < 64) g.trace[(2 * 64 + ti) * 2 + 0] = clock64();
            const uint32_t d = tmem_base + a * (uint32_t)(rg.nacc * Ntile);
            const uint32_t d1 = rg.nacc == 3 ? d + Ntile : d, d2 = rg.nacc == 3 ? d + 2 * Ntile : d;
            uint32_t sb = bres_u32;
            for (int kb = 0; kb < KB; ++kb) {
                WAIT(&sm.full[s], ph);
                if (g.trace && blockIdx.x == 0 && lane == 0 && (ti * KB + kb) < 64) g.trace[(5 * 64 + (ti * KB + kb)) * 2 + 0] = clock64();
                tc_fence_after();
                const int ksteps = (kb == KB - 1) ? last_steps : O::KBLK / O::KSTEP;
                if (elect_one()) {
                    uint64_t ahi = desc_hi | ((sa & 0x3FFFFu) >> 4), alo = desc_hi | (((sa + a_bytes) & 0x3FFFFu) >> 4);
                    uint64_t bhi = desc_hi | ((sb & 0x3FFFFu) >> 4), blo = desc_hi | (((sb + b_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
                    for (int j = 0; j < O::KBLK / O::KSTEP; ++j) {
                        if (j < ksteps && !(g.dbg & 4)) {
                            const bool acc = (kb | j) != 0;
                            mma_tf32(d, alo, bhi, idesc, acc);
                            mma_tf32(d1, ahi, blo, idesc, acc || rg.nacc != 3);
                            mma_tf32(d2, ahi, bhi, idesc, acc || rg.nacc != 3);
                            ahi += 2; alo += 2; bhi += 2; blo += 2;
                        }
                    }
                    mma_commit(&sm.free_[s]);
                    if (kb == KB - 1) mma_commit(&sm.acc_full[a]);
                }
                __syncwarp();
                if (g.trace && blockIdx.x == 0 && lane == 0 && (ti * KB + kb) < 64) g.trace[(5 * 64 + (ti * KB + kb)) * 2 + 1] = clock64();
                sa += stage_bytes;
                sb += 2 * b_bytes;
                if (++s == (uint32_t)S) { s = 0; ph ^= 1; sa = ops_u32; }
            }
            if (++a == (uint32_t)g.accbufs) { a = 0; aph ^= 1; }
        }
    } else if (warp == kRgTmaWarp) {
        // ------------------------------------------------------------------ TMA row streamer
        const long long plane = (long long)g.SH * g.SW;
        const int segf = rg.seg >> 2;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const uint32_t buf = ti & 1, use = ti >> 1;
            const int nrows = (tgi == rg.SCI - 1) ? last_rows : rg.TR;
            const int r_start = tgi * rg.TR * g.sy - rg.padt;           // first (virtual) source row
            const int skip = r_start < 0 ? -r_start : 0;
            const int r0 = r_start + skip;
            const int r1 = min(r_start + (nrows - 1) * g.sy + rg.kext, g.SH);
            const long long e00 = (long long)tb * g.SC * plane + (long long)r0 * g.SW;   // channel 0
            const long long nfl = (long long)(r1 - r0) * g.SW;
            uint8_t* raw = raw0 + (size_t)buf * raw_bytes;
            if (use > 0) WAIT(&raw_free[buf], (use - 1) & 1);
            if (g.trace && blockIdx.x == 0 && lane == 0 && ti < 64) g.trace[(4 * 64 + ti) * 2 + 0] = clock64();
            uint32_t mine = 0;
            if (nfl > 0 && !(g.dbg & 8)) {
                for (int ch = lane; ch < g.SC; ch += 32) {
                    const long long e0 = e00 + ch * plane;
                    const long long ea = e0 & ~3ll;
                    long long bytes = (((e0 - ea) + nfl) * 4 + 15) & ~15ll;
                    if (ea * 4 + bytes > rg.src_bytes16) bytes = rg.src_bytes16 - ea * 4;
                    tma_bulk_g2s(raw + (size_t)ch * rg.seg, g.src + ea, (uint32_t)bytes, &raw_full[buf]);
                    mine += (uint32_t)bytes;
                }
            }
            // this tile's gather offsets: segment base + alignment shift + static tap offset
            int* et = etab + buf * rg.Kpad;
            const int fix = -skip * g.SW - rg.padl;
            for (int k = lane; k < rg.Kpad; k += 32) {
                const int e = g.table[k];
                const int ch = e >> 20;
                const int sh = (int)((e00 + ch * plane) & 3);
                et[k] = ((ch * segf + sh + fix + ((e >> 4) & 0xFFFF)) * 16) | (e & 15);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&raw_full[buf], mine);
            if (g.trace && blockIdx.x == 0 && lane == 0 && ti < 64) g.trace[(4 * 64 + ti) * 2 + 1] = clock64();
            next_tile();
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        const int wq = warp - kEpiWarp0;
        const int row = wq * 32 + lane;
        const bool in_tile = row < rg.TR * g.GW;
        const int oyl = in_tile ? row / g.GW : 0;
        const int gx = in_tile ? row - oyl * g.GW : 0;
        const size_t oplane = (size_t)g.OHt * g.OWt;
        uint32_t a = 0, aph = 0;
        for (unsigned ti = 0; ti < my_tiles; ++ti) {
            const int nrows = (tgi == rg.SCI - 1) ? last_rows : rg.TR;
            const bool row_ok = in_tile && oyl < nrows;
            const int gy = tgi * rg.TR + oyl;
            float* dimg = g.dst + (size_t)tb * g.ON * oplane;
            WAIT(&sm.acc_full[a], aph);
            if (g.trace && blockIdx.x == 0 && tid == kEpiWarp0 * 32 && ti < 64) g.trace[(3 * 64 + ti) * 2 + 0] = clock64();
            tc_fence_after();
            int n = 0, pr = 0, pc = 0;
            for (int c0 = 0; c0 < Ntile; c0 += 16) {
                float v[16];
                const uint32_t ta = tmem_base + ((uint32_t)(wq * 32) << 16) + a * (uint32_t)(rg.nacc * Ntile) + c0;
                tmem_ld16(ta, v);
                if (rg.nacc == 3) {   // (lo*hi + hi*lo) + hi*hi
                    float v1[16], v2[16];
                    tmem_ld16(ta + Ntile, v1);
                    tmem_ld16(ta + 2 * Ntile, v2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = (v[j] + v1[j]) + v2[j];
                }
                if (c0 + 16 >= Ntile) {
                    // accumulator is in registers: release the TMEM buffer before the global stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sm.acc_empty[a]);
            if (g.trace && blockIdx.x == 0 && tid == kEpiWarp0 * 32 && ti < 64) g.trace[(3 * 64 + ti) * 2 + 1] = clock64();
                }
                if (g.os == 1) {
                    float* o = dimg + (size_t)c0 * oplane + (size_t)gy * g.OWt + gx;
                    if (row_ok && !(g.dbg & 1)) {
                        if (c0 + 16 <= g.ON) {
                            const float4* b4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
                            for (int j4 = 0; j4 < 4; ++j4) {
                                const float4 bb = b4[j4];
                                o[0] = v[4 * j4] + bb.x; o += oplane;
                                o[0] = v[4 * j4 + 1] + bb.y; o += oplane;
                                o[0] = v[4 * j4 + 2] + bb.z; o += oplane;
                                o[0] = v[4 * j4 + 3] + bb.w; o += oplane;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < g.ON) o[(size_t)j * oplane] = v[j] + s_bias[c0 + j];
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int oy = gy * g.os + pr, ox = gx * g.os + pc;
                        if (row_ok && c0 + j < g.Nreal && oy < g.OHt && ox < g.OWt && !(g.dbg & 1))
                            dimg[(size_t)n * oplane + (size_t)oy * g.OWt + ox] = v[j] + s_bias[n];
                        if (++n == g.ON) {
                            n = 0;
                            if (++pc == g.os) { pc = 0; ++pr; }
                        }
                    }
                }
            }
            if (++a == (uint32_t)g.accbufs) { a = 0; aph ^= 1; }
            next_tile();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// =============================================================================================
//                                      weight gradient
// =============================================================================================
struct WgradGemm {
    const float* x;
    const float* delta;
    const int* rowtab;     // [Mrows padded to 128]: >= 0 offset ci*H*W+ky*W+kx ; -1 ones row ; -2 unused
    float* partial;        // [splits][Mrows][Npad]
    int Mrows;             // Cin*k*k + 1 (last = all-ones row -> bias gradient)
    int Cin, H, W, Cout, OH, OW, s;
    unsigned P;            // B*OH*OW pixels
    unsigned nchunks, chunks_per_split;
    int Ntile, Npad, tmem_cols, stages;
    int single;            // 1: hi * hi only (CNN_TC_BF16X1)
};

template <bool TF32>
__global__ void __launch_bounds__(kThreads, 2) wgrad_ws(const WgradGemm g) {
    using O = Op<TF32>;
    constexpr int PPL = O::KBLK / 32;  // pixels per lane in one 128-byte row (1 tf32 / 2 bf16)
    extern __shared__ uint8_t smem_raw[];
    const SmemCarve sm = carve(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = g.stages, Ntile = g.Ntile;
    const int mt = blockIdx.x, split = blockIdx.y, nt = blockIdx.z;
    const uint32_t a_bytes = kRows * 128, b_bytes = (uint32_t)Ntile * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const unsigned q0 = split * g.chunks_per_split;
    const unsigned q1 = min(g.nchunks, q0 + g.chunks_per_split);
    int* s_rowtab = sm.rowtab;
    if (tid < kRows) s_rowtab[tid] = g.rowtab[mt * kRows + tid];

    if (warp == kMmaWarp) {
        tmem_alloc(sm.tmem_slot, (uint32_t)g.tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < S; ++i) {
                mbar_init(&sm.full[i], kProdWarps);
                mbar_init(&sm.free_[i], 1);
            }
            mbar_init(&sm.acc_full[0], 1);
            mbar_fence_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp < kEpiWarp0) {
        // ------------------------------------------------------------------ producers
        // Warp w owns rows w, w+8, ... of both operand tiles (so (row & 7) == w and the swizzled
        // position of a lane's word is loop-invariant); a lane owns PPL pixels of the chunk.
        // Row loads are issued in batches of G and register double-buffered across batches and
        // chunks: 2*G 128-byte lines in flight per warp.
        constexpr int G = 8;
        const int opl = g.OH * g.OW;
        const int plane = g.H * g.W;
        const int ones_r = (g.Mrows - 1) - mt * kRows;                 // row of the all-ones vector
        const int rows_x = max(0, min(kRows, ones_r));                 // x-tap rows of this M tile
        const bool own_ones = ones_r >= 0 && ones_r < kRows && (ones_r & 7) == warp;
        const int co0 = nt * Ntile;
        const int nB = max(0, min(Ntile, g.Cout - co0));               // real delta rows
        const int TA = rows_x > warp ? (rows_x - warp + 7) / 8 : 0;
        const int TB = nB > warp ? (nB - warp + 7) / 8 : 0;
        const int NBA = (TA + G - 1) / G, NBB = (TB + G - 1) / G;
        const int NB = max(1, NBA + NBB);                              // batches per chunk
        const uint32_t lane_off = (uint32_t)(warp * 128 + (((lane >> 2) ^ warp) << 4) + (lane & 3) * 4);
        // delta rows past Cout are never written by the loop: zero them once in every stage
        for (int n = nB + ((warp - nB) & 7); n < Ntile; n += 8)
            for (int st = 0; st < S; ++st) {
                uint8_t* t = sm.ops + (size_t)st * stage_bytes + 2 * a_bytes;
                const uint32_t o = (uint32_t)(n * 128 + (((lane >> 2) ^ (n & 7)) << 4) + (lane & 3) * 4);
                *reinterpret_cast<uint32_t*>(t + o) = 0u;
                *reinterpret_cast<uint32_t*>(t + b_bytes + o) = 0u;
            }
        struct Cursor { unsigned q; int batch; int xb[PPL], db[PPL]; bool ok[PPL]; };
        auto enter_chunk = [&](Cursor& c) {
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
                const unsigned p = c.q * O::KBLK + lane * PPL + i;
                c.ok[i] = p < g.P;
                const unsigned pp = c.ok[i] ? p : 0;   // masked pixels read pixel 0 and are zeroed
                const unsigned b = pp / (unsigned)opl, rem = pp % (unsigned)opl;
                const unsigned oy = rem / (unsigned)g.OW, ox = rem % (unsigned)g.OW;
                c.xb[i] = (int)(b * g.Cin * plane + (oy * g.s) * g.W + ox * g.s);
                c.db[i] = (int)(b * g.Cout * opl + rem) + co0 * opl;
            }
        };
        auto advance = [&](Cursor& c) -> bool {
            if (++c.batch < NB) return true;
            c.batch = 0;
            if (++c.q >= q1) return false;
            enter_chunk(c);
            return true;
        };
        auto issue = [&](const Cursor& c, float (&v)[G][PPL]) {
            if (c.batch < NBA) {
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const int i = c.batch * G + gi;
                    if (i < TA) {
                        const int e = s_rowtab[warp + 8 * i];
#pragma unroll
                        for (int k2 = 0; k2 < PPL; ++k2) {
                            const float t = __ldg(g.x + c.xb[k2] + e);
                            v[gi][k2] = c.ok[k2] ? t : 0.f;
                        }
                    }
                }
            } else {
                const int bb = c.batch - NBA;
#pragma unroll
                for (int gi = 0; gi < G; ++gi) {
                    const int i = bb * G + gi;
                    if (i < TB) {
#pragma unroll
                        for (int k2 = 0; k2 < PPL; ++k2) {
                            const float t = __ldg(g.delta + c.db[k2] + (warp + 8 * i) * opl);
                            v[gi][k2] = c.ok[k2] ? t : 0.f;
                        }
                    }
                }
            }
        };
        unsigned it = 0;
        uint32_t ps = 0, pph = 0;
        auto consume = [&](const Cursor& c, float (&v)[G][PPL]) {
            const unsigned s = ps;
            uint8_t* stage = sm.ops + (size_t)s * stage_bytes;
            if (c.batch == 0 && it >= (unsigned)S) mbar_wait(&sm.free_[s], pph ^ 1);
            const bool isA = c.batch < NBA;
            const int i0 = (isA ? c.batch : c.batch - NBA) * G;
            const int lim = isA ? TA : TB;
            uint8_t* tile = stage + (isA ? 0u : 2 * a_bytes) + lane_off;
            const uint32_t lo_off = isA ? a_bytes : b_bytes;
#pragma unroll
            for (int gi = 0; gi < G; ++gi) {
                const int i = i0 + gi;
                if (i < lim) {
                    uint32_t hi, lo;
                    if constexpr (TF32) split_tf32(v[gi][0], hi, lo);
                    else split2(v[gi][0], v[gi][PPL - 1], hi, lo);
                    *reinterpret_cast<uint32_t*>(tile + i * 1024) = hi;
                    *reinterpret_cast<uint32_t*>(tile + lo_off + i * 1024) = lo;
                }
            }
            if (c.batch == NB - 1) {
                if (own_ones) {  // all-ones row: its product with delta is the bias gradient
                    uint32_t hi;
                    if constexpr (TF32) hi = c.ok[0] ? 0x3F800000u : 0u;
                    else hi = (c.ok[0] ? 0x3F80u : 0u) | (c.ok[PPL - 1] ? 0x3F800000u : 0u);
                    uint8_t* t = stage + lane_off + (ones_r - warp) * 128;
                    *reinterpret_cast<uint32_t*>(t) = hi;
                    *reinterpret_cast<uint32_t*>(t + a_bytes) = 0u;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.full[s]);
                ++it;
                if (++ps == (uint32_t)S) { ps = 0; pph ^= 1; }
            }
        };
        if (q0 < q1) {
            float v0[G][PPL], v1[G][PPL];
            Cursor cur;
            cur.q = q0; cur.batch = 0;
            enter_chunk(cur);
            issue(cur, v0);
            while (true) {
                Cursor nxt = cur;
                const bool more1 = advance(nxt);
                if (more1) issue(nxt, v1);
                consume(cur, v0);
                if (!more1) break;
                cur = nxt;
                const bool more2 = advance(nxt);
                if (more2) issue(nxt, v0);
                consume(cur, v1);
                if (!more2) break;
                cur = nxt;
            }
        }
    } else if (warp == kMmaWarp) {
        // warp-uniform loop, one elected lane issues (see umma::elect_one)
        const uint32_t idesc = TF32 ? idesc_tf32(kRows, Ntile) : idesc_bf16(kRows, Ntile);
        const uint32_t ops_u32 = smem_u32(sm.ops);
        const uint64_t desc_hi = smem_desc_k128(0);
        uint32_t s = 0, ph = 0, sa = ops_u32;
        bool first = true;
        for (unsigned q = q0; q < q1; ++q) {
            mbar_wait(&sm.full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                uint64_t ahi = desc_hi | ((sa & 0x3FFFFu) >> 4), alo = desc_hi | (((sa + a_bytes) & 0x3FFFFu) >> 4);
                uint64_t bhi = desc_hi | (((sa + 2 * a_bytes) & 0x3FFFFu) >> 4);
                uint64_t blo = desc_hi | (((sa + 2 * a_bytes + b_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
                for (int j = 0; j < O::KBLK / O::KSTEP; ++j) {
                    if (g.single) {
                        mma_issue<TF32>(tmem_base, ahi, bhi, idesc, !(first && j == 0));
                    } else {
                        mma_issue<TF32>(tmem_base, alo, bhi, idesc, !(first && j == 0));
                        mma_issue<TF32>(tmem_base, ahi, blo, idesc, true);
                        mma_issue<TF32>(tmem_base, ahi, bhi, idesc, true);
                    }
                    ahi += 2; alo += 2; bhi += 2; blo += 2;
                }
                mma_commit(&sm.free_[s]);
            }
            __syncwarp();
            first = false;
            sa += stage_bytes;
            if (++s == (uint32_t)S) { s = 0; ph ^= 1; sa = ops_u32; }
        }
        if (elect_one()) mma_commit(&sm.acc_full[0]);
        __syncwarp();
    } else if (warp == kTmaWarp) {
        // unused in this kernel (the role layout is shared with gather_gemm_ws)
    } else {
        // ------------------------------------------------------------------ epilogue (once)
        const int wq = warp - kEpiWarp0;
        const int kidx = mt * kRows + wq * 32 + lane;
        mbar_wait(&sm.acc_full[0], 0);
        tc_fence_after();
        float* prow = g.partial + ((size_t)split * g.Mrows + kidx) * g.Npad + (size_t)nt * Ntile;
        for (int c0 = 0; c0 < Ntile; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + c0, v);
            if (kidx < g.Mrows) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// dw[co][kidx] = scale * sum_split partial[split][kidx][co] ; db[co] from the ones row.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                                           float* __restrict__ db, int Mrows, int Npad, int Cout,
                                                           int splits, float scale) {
    // block = 32 consecutive outputs x 8 split lanes; eight independent loads in flight per thread, lanes
    // meet in shared memory in a fixed order (deterministic).  With the accumulator-length cap there can
    // be thousands of splits: a one-thread-per-output loop would be pure load latency.
    __shared__ float red[8][33];
    const int total = Mrows * Cout;
    const int col = threadIdx.x & 31, sl = threadIdx.x >> 5;
    for (int base = blockIdx.x * 32; base < total; base += gridDim.x * 32) {
        const int id = base + col;
        const int co = id % Cout, kidx = id / Cout;  // consecutive threads -> consecutive partial columns
        float s = 0.f;
        if (id < total) {
            const float* src = partial + (size_t)kidx * Npad + co;
            const size_t stride = (size_t)Mrows * Npad;
            int sp = sl;
            for (; sp + 56 < splits; sp += 64) {
                float v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = __ldg(src + (size_t)(sp + 8 * k) * stride);
#pragma unroll
                for (int k = 0; k < 8; ++k) s += v[k];
            }
            for (; sp < splits; sp += 8) s += __ldg(src + (size_t)sp * stride);
        }
        red[sl][col] = s;
        __syncthreads();
        if (sl == 0 && id < total) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += red[k][col];
            t *= scale;
            if (kidx == Mrows - 1) db[co] = t;
            else dw[(size_t)co * (Mrows - 1) + kidx] = t;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Weight gradient, TMA row-staged variant (the default when the layer fits; wgrad_ws above is the
// general fallback).  The gather version exposes global-load latency: a warp keeps only two
// batches of row loads in flight.  Here the TMA bulk engine streams the RAW operand rows of a
// whole "super-chunk" (TR full output rows of one image, <= 128 pixels) into shared memory --
// per input channel one contiguous run of input rows, per output channel one contiguous run of
// delta -- double-buffered and one super-chunk ahead, with no registers or threads involved.
// The producer warps then build the swizzled hi/lo operand tiles from shared memory only.
// NCHW rows are only 4-byte aligned: every copy starts at the preceding 16-byte boundary and the
// consumer adds the per-segment element shift recorded next to the buffer.
struct WgradRows {
    const float* x;
    const float* delta;
    const int* rowtab;     // [mtiles*128]: (ci << 20) | (ky*W + kx) ; -1 ones row ; -2 unused
    float* partial;        // [splits][Mrows][Npad]
    int Mrows, Cin, H, W, Cout, OH, OW, s, k, B;
    int TR, SCI;           // output rows per super-chunk, super-chunks per image
    unsigned nsc, sc_per_split;
    int Ntile, Npad, tmem_cols, stages;
    int xseg, dseg;        // bytes per raw segment (multiples of 16)
    int nci_max;           // input channels an M tile can touch
    long long x_bytes16, d_bytes16;  // tensor sizes rounded up to 16 bytes (copy clamp)
    int single;            // 1: hi * hi only (CNN_TC_BF16X1)
};

// kRowsPW producer warps (smem -> smem tile builders) + one MMA warp + one TMA warp; warps 0-3 also
// run the epilogue.  8 producer warps / 2 CTAs per SM for thin layers (few operand rows), 16 / 1 else.

template <bool TF32, int kRowsPW, int kMinCtas>
__global__ void __launch_bounds__((kRowsPW + 2) * 32, kMinCtas) wgrad_rows_ws(const WgradRows g) {
    using O = Op<TF32>;
    constexpr int PPL = O::KBLK / 32;   // pixels per lane in one 128-byte operand row (1 tf32 / 2 bf16)
    extern __shared__ uint8_t smem_raw[];
    const SmemCarve sm = carve(smem_raw);
    uint64_t* raw_full = sm.bfull;        // [2]
    uint64_t* raw_free = sm.bfull + 2;    // [2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = g.stages, Ntile = g.Ntile, kk = g.k * g.k;
    const int mt = blockIdx.x, split = blockIdx.y, nt = blockIdx.z;
    const uint32_t a_bytes = kRows * 128, b_bytes = (uint32_t)Ntile * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const uint32_t raw_bytes = (uint32_t)(g.nci_max * g.xseg + Ntile * g.dseg);
    uint8_t* raw0 = sm.ops + (size_t)S * stage_bytes;
    int* shifts = reinterpret_cast<int*>(raw0 + 2 * (size_t)raw_bytes);   // [2][16 + Ntile]
    const int shift_stride = 16 + Ntile;
    const unsigned sc0 = split * g.sc_per_split, sc1 = min(g.nsc, sc0 + g.sc_per_split);
    const int ci_lo = (mt * kRows) / kk;
    const int ci_hi = min(g.Cin - 1, (mt * kRows + kRows - 1) / kk);
    const int nci = max(0, ci_hi - ci_lo + 1);
    const int co0 = nt * Ntile;
    const int nB = max(0, min(Ntile, g.Cout - co0));
    const int last_rows = g.OH - (g.SCI - 1) * g.TR;           // rows of an image's last super-chunk

    if (tid < kRows) sm.rowtab[tid] = g.rowtab[mt * kRows + tid];
    if (warp == kRowsPW) {
        tmem_alloc(sm.tmem_slot, (uint32_t)g.tmem_cols);
        if (lane == 0) {
            for (int i = 0; i < S; ++i) {
                mbar_init(&sm.full[i], kRowsPW);
                mbar_init(&sm.free_[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&raw_full[i], 1);
                mbar_init(&raw_free[i], kRowsPW);
            }
            mbar_init(&sm.acc_full[0], 1);
            mbar_fence_init();
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp < kRowsPW) {
        // ------------------------------------------------------------------ producers (smem -> smem)
        // warp w owns rows w, w+PW, ... (PW a multiple of 8, so (row & 7) == (w & 7) stays loop-invariant)
        const int ones_r = (g.Mrows - 1) - mt * kRows;
        const int rows_x = max(0, min(kRows, ones_r));
        const bool own_ones = ones_r >= 0 && ones_r < kRows && (ones_r % kRowsPW) == warp;
        const int TA = rows_x > warp ? (rows_x - warp + kRowsPW - 1) / kRowsPW : 0;
        const int TB = nB > warp ? (nB - warp + kRowsPW - 1) / kRowsPW : 0;
        const uint32_t lane_off = (uint32_t)(warp * 128 + (((lane >> 2) ^ (warp & 7)) << 4) + (lane & 3) * 4);
        for (int n = nB + ((warp - nB) & (kRowsPW - 1)); n < Ntile; n += kRowsPW)   // delta rows past Cout: zero once
            for (int st = 0; st < S; ++st) {
                uint8_t* t = sm.ops + (size_t)st * stage_bytes + 2 * a_bytes;
                const uint32_t o = (uint32_t)(n * 128 + (((lane >> 2) ^ (n & 7)) << 4) + (lane & 3) * 4);
                *reinterpret_cast<uint32_t*>(t + o) = 0u;
                *reinterpret_cast<uint32_t*>(t + b_bytes + o) = 0u;
            }
        // static part of this lane's row offsets (lane i <-> row warp + 8 i)
        int a_static = 0, a_ci = 0;
        if (lane < TA) {
            const int e = sm.rowtab[warp + kRowsPW * lane];
            a_ci = (e >> 20) - ci_lo;
            a_static = a_ci * (g.xseg >> 2) + (e & 0xFFFFF);
        }
        const int b_static = (warp + kRowsPW * lane) * (g.dseg >> 2);
        uint32_t ps = 0, pph = 0, it = 0;
        unsigned gi = sc0 % (unsigned)g.SCI;   // super-chunk index inside its image
        for (unsigned sc = sc0, i = 0; sc < sc1; ++sc, ++i) {
            const uint32_t buf = i & 1, use = i >> 1;
            const int nrows = (gi == (unsigned)g.SCI - 1) ? last_rows : g.TR;
            const int npx = nrows * g.OW, nkb = (npx + O::KBLK - 1) / O::KBLK;
            const float* rawx = reinterpret_cast<const float*>(raw0 + (size_t)buf * raw_bytes);
            const float* rawd = rawx + ((g.nci_max * g.xseg) >> 2);
            const int* sh = shifts + buf * shift_stride;
            mbar_wait(&raw_full[buf], use & 1);
            const int myA = a_static + ((lane < TA) ? sh[a_ci] : 0);
            const int myB = b_static + ((lane < TB) ? sh[16 + warp + kRowsPW * lane] : 0);
            for (int j = 0; j < nkb; ++j, ++it) {
                uint8_t* stage = sm.ops + (size_t)ps * stage_bytes;
                if (it >= (uint32_t)S) mbar_wait(&sm.free_[ps], pph ^ 1);
                int xo[PPL], po[PPL];
                bool ok[PPL];
#pragma unroll
                for (int q = 0; q < PPL; ++q) {
                    const int pl = j * O::KBLK + lane * PPL + q;
                    ok[q] = pl < npx;
                    const int oyl = ok[q] ? pl / g.OW : 0;
                    const int ox = ok[q] ? pl - oyl * g.OW : 0;
                    xo[q] = (oyl * g.s) * g.W + ox * g.s;
                    po[q] = ok[q] ? pl : 0;
                }
                uint8_t* ta = stage + lane_off;
#pragma unroll 4
                for (int r = 0; r < TA; ++r) {
                    const int off = __shfl_sync(0xffffffffu, myA, r);
                    float v[PPL];
#pragma unroll
                    for (int q = 0; q < PPL; ++q) v[q] = ok[q] ? rawx[off + xo[q]] : 0.f;
                    uint32_t hi, lo;
                    if constexpr (TF32) split_tf32(v[0], hi, lo);
                    else split2(v[0], v[PPL - 1], hi, lo);
                    *reinterpret_cast<uint32_t*>(ta + r * (kRowsPW * 128)) = hi;
                    *reinterpret_cast<uint32_t*>(ta + a_bytes + r * (kRowsPW * 128)) = lo;
                }
                uint8_t* tb = stage + 2 * a_bytes + lane_off;
#pragma unroll 4
                for (int r = 0; r < TB; ++r) {
                    const int off = __shfl_sync(0xffffffffu, myB, r);
                    float v[PPL];
#pragma unroll
                    for (int q = 0; q < PPL; ++q) v[q] = ok[q] ? rawd[off + po[q]] : 0.f;
                    uint32_t hi, lo;
                    if constexpr (TF32) split_tf32(v[0], hi, lo);
                    else split2(v[0], v[PPL - 1], hi, lo);
                    *reinterpret_cast<uint32_t*>(tb + r * (kRowsPW * 128)) = hi;
                    *reinterpret_cast<uint32_t*>(tb + b_bytes + r * (kRowsPW * 128)) = lo;
                }
                if (own_ones) {
                    uint8_t* t = stage + lane_off + (ones_r - warp) * 128;
                    uint32_t one;
                    if constexpr (TF32) one = ok[0] ? 0x3F800000u : 0u;
                    else one = (ok[0] ? 0x3F80u : 0u) | (ok[PPL - 1] ? 0x3F800000u : 0u);
                    *reinterpret_cast<uint32_t*>(t) = one;
                    *reinterpret_cast<uint32_t*>(t + a_bytes) = 0u;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.full[ps]);
                if (++ps == (uint32_t)S) { ps = 0; pph ^= 1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_free[buf]);
            if (++gi == (unsigned)g.SCI) gi = 0;
        }
        if (warp < 4) {
            // -------------------------------------------------------------- epilogue (once)
            const int kidx = mt * kRows + warp * 32 + lane;
            mbar_wait(&sm.acc_full[0], 0);
            tc_fence_after();
            float* prow = g.partial + ((size_t)split * g.Mrows + kidx) * g.Npad + (size_t)nt * Ntile;
            for (int c0 = 0; c0 < Ntile; c0 += 16) {
                float v[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
                if (kidx < g.Mrows) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(prow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
        }
    } else if (warp == kRowsPW) {
        // ------------------------------------------------------------------ MMA issuer
        // warp-uniform loop, one elected lane issues (see umma::elect_one)
        {
            const uint32_t idesc = TF32 ? idesc_tf32(kRows, Ntile) : idesc_bf16(kRows, Ntile);
            const uint32_t ops_u32 = smem_u32(sm.ops);
            const uint64_t desc_hi = smem_desc_k128(0);
            const int kb_full = (g.TR * g.OW + O::KBLK - 1) / O::KBLK, kb_last = (last_rows * g.OW + O::KBLK - 1) / O::KBLK;
            uint32_t s = 0, ph = 0, sa = ops_u32;
            bool first = true;
            unsigned gi = sc0 % (unsigned)g.SCI;
            for (unsigned sc = sc0; sc < sc1; ++sc) {
                const int nkb = (gi == (unsigned)g.SCI - 1) ? kb_last : kb_full;
                for (int j = 0; j < nkb; ++j) {
                    mbar_wait(&sm.full[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        uint64_t ahi = desc_hi | ((sa & 0x3FFFFu) >> 4), alo = desc_hi | (((sa + a_bytes) & 0x3FFFFu) >> 4);
                        uint64_t bhi = desc_hi | (((sa + 2 * a_bytes) & 0x3FFFFu) >> 4);
                        uint64_t blo = desc_hi | (((sa + 2 * a_bytes + b_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
                        for (int q = 0; q < O::KBLK / O::KSTEP; ++q) {
                            if (g.single) {
                                mma_issue<TF32>(tmem_base, ahi, bhi, idesc, !(first && q == 0));
                            } else {
                                mma_issue<TF32>(tmem_base, alo, bhi, idesc, !(first && q == 0));
                                mma_issue<TF32>(tmem_base, ahi, blo, idesc, true);
                                mma_issue<TF32>(tmem_base, ahi, bhi, idesc, true);
                            }
                            ahi += 2; alo += 2; bhi += 2; blo += 2;
                        }
                        mma_commit(&sm.free_[s]);
                        if (sc + 1 == sc1 && j + 1 == nkb) mma_commit(&sm.acc_full[0]);
                    }
                    __syncwarp();
                    first = false;
                    sa += stage_bytes;
                    if (++s == (uint32_t)S) { s = 0; ph ^= 1; sa = ops_u32; }
                }
                if (++gi == (unsigned)g.SCI) gi = 0;
            }
        }
    } else {
        // ------------------------------------------------------------------ TMA row streamer
        // All 32 lanes issue copies (lane <-> segment): the address arithmetic of a segment runs in
        // parallel and a super-chunk's 20-150 copies leave in a handful of instructions.  The
        // transaction count is posted after the copies (tx-count may go negative meanwhile).
        {
            unsigned gi = sc0 % (unsigned)g.SCI, b = sc0 / (unsigned)g.SCI;
            const int nseg = nci + nB;
            const long long xplane = (long long)g.H * g.W, dplane = (long long)g.OH * g.OW;
            for (unsigned sc = sc0, i = 0; sc < sc1; ++sc, ++i) {
                const uint32_t buf = i & 1, use = i >> 1;
                const int oy0 = (int)gi * g.TR;
                const int nrows = (gi == (unsigned)g.SCI - 1) ? last_rows : g.TR;
                uint8_t* rawx = raw0 + (size_t)buf * raw_bytes;
                uint8_t* rawd = rawx + (size_t)g.nci_max * g.xseg;
                int* sh = shifts + buf * shift_stride;
                if (use > 0) mbar_wait(&raw_free[buf], (use - 1) & 1);
                const long long xe = (long long)((nrows - 1) * g.s + g.k) * g.W;   // floats per input channel
                const long long de = (long long)nrows * g.OW;                      // floats per delta row
                const long long x0 = ((long long)b * g.Cin + ci_lo) * xplane + (long long)oy0 * g.s * g.W;
                const long long d0 = ((long long)b * g.Cout + co0) * dplane + (long long)oy0 * g.OW;
                uint32_t mine = 0;
                for (int sg = lane; sg < nseg; sg += 32) {
                    const bool isx = sg < nci;
                    const int q = isx ? sg : sg - nci;
                    const long long e0 = isx ? x0 + q * xplane : d0 + q * dplane;
                    const long long ea = e0 & ~3ll;
                    long long bytes = (((e0 - ea) + (isx ? xe : de)) * 4 + 15) & ~15ll;
                    const long long lim = isx ? g.x_bytes16 : g.d_bytes16;
                    if (ea * 4 + bytes > lim) bytes = lim - ea * 4;
                    sh[isx ? q : 16 + q] = (int)(e0 - ea);
                    tma_bulk_g2s(isx ? rawx + (size_t)q * g.xseg : rawd + (size_t)q * g.dseg,
                                 (isx ? g.x : g.delta) + ea, (uint32_t)bytes, &raw_full[buf]);
                    mine += (uint32_t)bytes;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
                __syncwarp();   // every lane's shift-table stores precede the releasing arrive
                if (lane == 0) mbar_expect_tx(&raw_full[buf], mine);
                if (++gi == (unsigned)g.SCI) { gi = 0; ++b; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kRowsPW) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// =============================================================================================
//                                          host side
// =============================================================================================
struct PlanKey {
    int kind, tf32, Cin, H, W, Cout, k, s;  // kind: 0 fwd, 1 dgrad, 2 wgrad
    bool operator<(const PlanKey& o) const {
        return std::tie(kind, tf32, Cin, H, W, Cout, k, s) <
               std::tie(o.kind, o.tf32, o.Cin, o.H, o.W, o.Cout, o.k, o.s);
    }
};

struct Plan {
    GatherGemm g{};
    PackInfo pack{};
    WgradGemm wg{};
    WgradRows wr{};
    bool rows_ok = false;    // the TMA row-staged weight-gradient kernel applies
    bool hoist = false;
    int* d_table = nullptr;  // device tables (owned; live as long as the process)
    int* d_table2 = nullptr;
    size_t rows_smem = 0;
    int rows_ctas = 1;
    bool rows_wide = false;
    RowsGather rg{};         // row-staged forward / input gradient
    bool rg_ok = false;
    size_t rg_smem = 0;
    int ntiles = 1, Nreal = 0, ctas_per_sm = 1;
    size_t packed_bytes = 0, smem = 0;
};

std::map<std::pair<int, PlanKey>, Plan>& plans() {
    static std::map<std::pair<int, PlanKey>, Plan> p;
    return p;
}

int next_pow2_cols(int c) {
    int p = 32;
    while (p < c) p <<= 1;
    return p;
}

constexpr size_t kSmemHeader = 1024 + 1024;  // barriers + wgrad row table + alignment slack
constexpr size_t kSmemMax = 227 * 1024;

// stages / residency: two CTAs per SM when three stages fit in half the shared memory
// Measured on B200 (tools/time_ops.py, CNN_DBG_STAGES / CNN_DBG_CTAS sweeps): these kernels are
// bound by exposed load latency, not by ring depth -- two resident CTAs per SM with a 2-deep ring
// beat one CTA with a 5-6-deep ring on every AlexNet-lite layer, three CTAs are worse again
// (register file: 416 threads x 72 registers).  So: 2 CTAs/SM whenever two stages fit in half
// of the shared memory, ring depth from what is left, at most 3.
void pick_stages(size_t stage_bytes, size_t fixed_bytes, int tmem_cols, int* stages, int* ctas, size_t* smem) {
    int S, c;
    const size_t head = kSmemHeader + fixed_bytes;
    const size_t half = kSmemMax / 2 - 1024;  // 1 KB per CTA is reserved by the system
    if (2 * stage_bytes + head <= half && tmem_cols <= 256) {
        c = 2;
        S = (int)((half - head) / stage_bytes);
        if (S > 3) S = 3;
    } else {
        c = 1;
        S = (int)((kSmemMax - 1024 - head) / stage_bytes);
        if (S > kMaxStages) S = kMaxStages;
    }
    if (S < 2) S = 2;
    *stages = S; *ctas = c; *smem = head + (size_t)S * stage_bytes;
}

template <class K>
int set_smem_attr(K kernel, const char* name) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
    if (e != cudaSuccess) return cnn_cuda_fail(e, name, __FILE__, __LINE__);
    return CNN_OK;
}

int upload_table(const std::vector<int>& table, int** d) {
    if (cudaMalloc(d, table.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(*d, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
        cnn_set_error("conv_tc: table upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        return CNN_ERR_CUDA;
    }
    return CNN_OK;
}

// forward (kind 0) / input gradient (kind 1): gather table + tiling, built once per geometry
int get_gather_plan(cnn_ctx* ctx, int dgrad, bool tf32, int Cin, int H, int W, int Cout, int k, int s,
                    Plan** out) {
    const PlanKey key{dgrad, tf32 ? 1 : 0, Cin, H, W, Cout, k, s};
    std::lock_guard<std::recursive_mutex> lk(cnn_global_mutex());   // plan cache is process-wide (map nodes are stable)
    auto& mp = plans();
    auto itp = mp.find({ctx->device, key});
    if (itp != mp.end()) { *out = &itp->second; return CNN_OK; }
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1, kk = k * k;
    const int KBLK = tf32 ? 32 : 64;
    Plan p;
    GatherGemm& g = p.g;
    PackInfo& pi = p.pack;
    std::vector<int> table;
    pi.s = s; pi.k = k; pi.npos = 0;
    if (!dgrad) {
        g.K = Cin * kk; g.npos = 0;
        for (int ci = 0; ci < Cin; ++ci)
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) table.push_back(ci * H * W + ky * W + kx);
        g.GH = OH; g.GW = OW; g.SC = Cin; g.SH = H; g.SW = W; g.sy = g.sx = s;
        g.ON = Cout; g.OHt = OH; g.OWt = OW; g.os = 1; g.Nreal = Cout;
    } else {
        // source positions (dy,dx) = (ky/s, kx/s) reached by any tap
        const int nd = (k - 1) / s + 1;
        int np = 0;
        for (int dy = 0; dy < nd; ++dy)
            for (int dx = 0; dx < nd; ++dx) { g.dy[np] = pi.dy[np] = (signed char)dy; g.dx[np] = pi.dx[np] = (signed char)dx; ++np; }
        g.npos = pi.npos = np;
        g.K = Cout * np;
        for (int co = 0; co < Cout; ++co)
            for (int t = 0; t < np; ++t) table.push_back(((co * OH * OW - g.dy[t] * OW - g.dx[t]) * 16) | t);
        g.GH = (H + s - 1) / s; g.GW = (W + s - 1) / s; g.SC = Cout; g.SH = OH; g.SW = OW; g.sy = g.sx = 1;
        g.ON = Cin; g.OHt = H; g.OWt = W; g.os = s; g.Nreal = s * s * Cin;
    }
    g.KB = (g.K + KBLK - 1) / KBLK;
    while ((int)table.size() < g.KB * KBLK) table.push_back(dgrad ? kPadCode : 0);
    pi.K = g.K; pi.KB = g.KB;
    p.Nreal = g.Nreal;
    // n tiling: two accumulator buffers (epilogue overlaps the next tile) when they fit in TMEM
    const int Npad = ((g.Nreal + 15) / 16) * 16;
    p.ntiles = (Npad + 255) / 256;
    g.Ntile = (((Npad + p.ntiles - 1) / p.ntiles) + 15) / 16 * 16;
    g.accbufs = (2 * g.Ntile <= 512) ? 2 : 1;
    g.tmem_cols = next_pow2_cols(g.accbufs * g.Ntile);
    p.packed_bytes = (size_t)p.ntiles * g.KB * 2 * g.Ntile * 128;
    p.hoist = (g.KB == 1) && tf32;
    // filters resident in shared memory (one TMA burst per CTA) when an n-tile's blocks fit in 64 KB
    const size_t bblk = 2 * (size_t)g.Ntile * 128, ball = (size_t)g.KB * bblk;
    g.bres = (ball <= 64 * 1024 && !getenv("CNN_DBG_NOBRES")) ? 1 : 0;
    const size_t stage_bytes = 2 * (size_t)kRows * 128 + (g.bres ? 0 : bblk);
    const size_t fixed = (g.bres ? ball : 0) + ((size_t)g.ON * sizeof(float) + 127) / 128 * 128;  // + bias
    pick_stages(stage_bytes, fixed, g.tmem_cols, &g.stages, &p.ctas_per_sm, &p.smem);
    if (const char* e = getenv("CNN_DBG_STAGES")) {  // experiment knobs, not part of the API
        g.stages = atoi(e);
        p.smem = kSmemHeader + fixed + (size_t)g.stages * stage_bytes;
    }
    if (const char* e = getenv("CNN_DBG_CTAS")) p.ctas_per_sm = atoi(e);
    if (const char* e = getenv("CNN_DBG_SKIP")) g.dbg = atoi(e);
    if (const char* e = getenv("CNN_DBG_ACC")) { g.accbufs = atoi(e); g.tmem_cols = next_pow2_cols(g.accbufs * g.Ntile); }
    if (const char* e = getenv("CNN_DBG_NOHOIST")) p.hoist = p.hoist && atoi(e) == 0;
    if (int rc = upload_table(table, &p.d_table)) return rc;
    g.table = p.d_table;
    int rc = CNN_OK;
    // ---- TMA row-staged variant: TR whole rows of the row space per tile
    if (tf32 && p.ntiles == 1 && g.bres && g.GW <= kRows && g.SC < 2048 && (k - 1) * g.SW + k < 65536 &&
        !getenv("CNN_DBG_NOROWSG")) {
        RowsGather& r = p.rg;
        const int nd = (k - 1) / s + 1;
        r.TR = std::max(1, std::min(g.GH, kRows / g.GW));
        r.SCI = (g.GH + r.TR - 1) / r.TR;
        r.padt = dgrad ? nd - 1 : 0;
        r.padl = dgrad ? nd - 1 : 0;
        r.kext = dgrad ? nd : k;
        r.Kpad = g.KB * KBLK;
        const int vrows = (r.TR - 1) * g.sy + r.kext;
        r.seg = (int)(((size_t)vrows * g.SW * 4 + 12 + 15) / 16 * 16);
        std::vector<int> st;
        if (!dgrad) {
            for (int ci = 0; ci < Cin; ++ci)
                for (int ky = 0; ky < k; ++ky)
                    for (int kx = 0; kx < k; ++kx) st.push_back((ci << 20) | ((ky * W + kx) << 4));
        } else {
            for (int co = 0; co < Cout; ++co)
                for (int t = 0; t < g.npos; ++t)
                    st.push_back((co << 20) | ((((r.padt - g.dy[t]) * g.SW + (r.padl - g.dx[t]))) << 4) | t);
        }
        while ((int)st.size() < r.Kpad) st.push_back(dgrad ? kPadCode : 0);
        const size_t rest = kSmemHeader + ball + ((size_t)g.ON * 4 + 127) / 128 * 128 + 2 * (size_t)g.SC * r.seg +
                            2 * (size_t)r.Kpad * 4 + 64;
        const size_t abytes = 2 * (size_t)kRows * 128;
        const size_t half = kSmemMax / 2 - 1024;
        const int fill = r.TR * g.GW;
        int min_fill = 96;
        if (const char* e = getenv("CNN_DBG_ROWSG_FILL")) min_fill = atoi(e);
        r.nacc = (g.accbufs * 3 * g.Ntile <= 256) ? 3 : 1;
        if (const char* e = getenv("CNN_DBG_NACC")) r.nacc = atoi(e);
        const int rg_cols = next_pow2_cols(g.accbufs * r.nacc * g.Ntile);
        if (rest + 2 * abytes <= half && fill >= min_fill && rg_cols <= 256) {
            int S = (int)((half - rest) / abytes);
            if (S > 3) S = 3;
            if (const char* e = getenv("CNN_DBG_STAGES")) S = atoi(e);
            r.g = g;
            r.g.stages = S;
            r.g.tmem_cols = rg_cols;
            p.rg_smem = rest + (size_t)S * abytes;
            if (int rc2 = upload_table(st, &p.d_table2)) return rc2;
            r.g.table = p.d_table2;
            rc = dgrad ? set_smem_attr(gather_rows_ws<true>, "gather_rows_ws<masked>")
                       : set_smem_attr(gather_rows_ws<false>, "gather_rows_ws");
            if (rc) return rc;
            p.rg_ok = true;
        }
    }
    if (tf32) {
        if (dgrad) rc = p.hoist ? set_smem_attr(gather_gemm_ws<true, true, true>, "gather_gemm_ws<tf32,masked,hoist>")
                                : set_smem_attr(gather_gemm_ws<true, true, false>, "gather_gemm_ws<tf32,masked>");
        else rc = p.hoist ? set_smem_attr(gather_gemm_ws<true, false, true>, "gather_gemm_ws<tf32,hoist>")
                          : set_smem_attr(gather_gemm_ws<true, false, false>, "gather_gemm_ws<tf32>");
    } else {
        rc = dgrad ? set_smem_attr(gather_gemm_ws<false, true, false>, "gather_gemm_ws<bf16,masked>")
                   : set_smem_attr(gather_gemm_ws<false, false, false>, "gather_gemm_ws<bf16>");
    }
    if (rc) return rc;
    auto ins = mp.emplace(std::make_pair(ctx->device, key), p);
    *out = &ins.first->second;
    return CNN_OK;
}

int run_gather_gemm(cnn_ctx* ctx, Plan* p, bool tf32, const float* w, const float* src, const float* bias,
                    float* dst, int B, int dgrad, int Cin, int Cout) {
    uint8_t* packed = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, p->packed_bytes + 1024));
    CNN_REQUIRE(packed, "scratch allocation failed");
    packed = reinterpret_cast<uint8_t*>(((uintptr_t)packed + 1023) & ~(uintptr_t)1023);
    GatherGemm g = p->g;
    const long long rows = (long long)B * g.GH * g.GW;
    CNN_REQUIRE(rows < (1ll << 31) - 256, "conv_tc: too many GEMM rows");
    CNN_REQUIRE((long long)B * g.SC * g.SH * g.SW < (1ll << 31), "conv_tc: source tensor too large");
    {
        const long long chunks = (long long)p->ntiles * g.KB * g.Ntile * 8;
        int grid = cdiv(chunks, 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        if (tf32) {
            CNN_LAUNCH(ctx, pack_filters_kernel<true>, grid, 256, 0, w, packed, p->pack, dgrad, Cin, Cout,
                       p->Nreal, g.Ntile, p->ntiles);
        } else {
            CNN_LAUNCH(ctx, pack_filters_kernel<false>, grid, 256, 0, w, packed, p->pack, dgrad, Cin, Cout,
                       p->Nreal, g.Ntile, p->ntiles);
        }
    }
    g.src = src; g.bias = bias; g.dst = dst; g.packedB = packed;
    g.single = ctx->tc_precision == CNN_TC_BF16X1 ? 1 : 0;
    static long long* d_trace = nullptr;
    const bool tracing = getenv("CNN_DBG_TRACE") != nullptr;
    if (tracing) {
        if (!d_trace) cudaMalloc(&d_trace, 8 * 64 * 2 * sizeof(long long));
        cudaMemset(d_trace, 0, 8 * 64 * 2 * sizeof(long long));
        g.trace = d_trace;
    }
    g.rows = (unsigned)rows;
    g.mtiles = (unsigned)((rows + kRows - 1) / kRows);
    if (p->rg_ok && tf32) {
        CNN_REQUIRE(((uintptr_t)src & 15) == 0, "conv_tc: source base pointer must be 16-byte aligned");
        RowsGather r = p->rg;
        r.g.src = src; r.g.bias = bias; r.g.dst = dst; r.g.packedB = packed;
        r.tiles = (unsigned)B * (unsigned)r.SCI;
        r.src_bytes16 = ((long long)B * g.SC * g.SH * g.SW * 4 + 15) & ~15ll;
        unsigned gr = (unsigned)ctx->sm_count * 2;
        if (gr > r.tiles) gr = r.tiles;
        r.g.trace = g.trace;
        if (dgrad) { CNN_LAUNCH(ctx, gather_rows_ws<true>, gr, kRgThreads, p->rg_smem, r); }
        else { CNN_LAUNCH(ctx, gather_rows_ws<false>, gr, kRgThreads, p->rg_smem, r); }
        if (tracing) {
            static long long h[8 * 64 * 2];
            cudaStreamSynchronize(ctx->stream);
            cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
            long long t0 = h[4 * 128];
            fprintf(stderr, "rows trace dgrad=%d KB=%d: T[freewait,arrive] P[rawfull, rawfree] | Pkb[freewait, fullarrive] Mkb[fullwait, commit] | M[accempty, accfull] E[accfull, accempty]\n", dgrad, g.KB);
            for (int i = 0; i < 20; ++i)
                fprintf(stderr, "%2d T %6lld %6lld P %6lld %6lld | Pkb %6lld %6lld Mkb %6lld %6lld | M %6lld %6lld E %6lld %6lld\n", i,
                        h[(4 * 64 + i) * 2] - t0, h[(4 * 64 + i) * 2 + 1] - t0, h[(1 * 64 + i) * 2] - t0, h[(1 * 64 + i) * 2 + 1] - t0,
                        h[(0 * 64 + i) * 2] - t0, h[(0 * 64 + i) * 2 + 1] - t0, h[(5 * 64 + i) * 2] - t0, h[(5 * 64 + i) * 2 + 1] - t0,
                        h[(2 * 64 + i) * 2] - t0, h[(2 * 64 + i) * 2 + 1] - t0, h[(3 * 64 + i) * 2] - t0, h[(3 * 64 + i) * 2 + 1] - t0);
        }
        return CNN_OK;
    }
    unsigned gx = (unsigned)(ctx->sm_count * p->ctas_per_sm) / (unsigned)p->ntiles;
    if (gx < 1) gx = 1;
    if (gx > g.mtiles) gx = g.mtiles;
    dim3 grid(gx, (unsigned)p->ntiles);
    if (tf32) {
        if (dgrad) {
            if (p->hoist) { CNN_LAUNCH(ctx, (gather_gemm_ws<true, true, true>), grid, kThreads, p->smem, g); }
            else { CNN_LAUNCH(ctx, (gather_gemm_ws<true, true, false>), grid, kThreads, p->smem, g); }
        } else {
            if (p->hoist) { CNN_LAUNCH(ctx, (gather_gemm_ws<true, false, true>), grid, kThreads, p->smem, g); }
            else { CNN_LAUNCH(ctx, (gather_gemm_ws<true, false, false>), grid, kThreads, p->smem, g); }
        }
    } else {
        if (dgrad) { CNN_LAUNCH(ctx, (gather_gemm_ws<false, true, false>), grid, kThreads, p->smem, g); }
        else { CNN_LAUNCH(ctx, (gather_gemm_ws<false, false, false>), grid, kThreads, p->smem, g); }
    }
    if (tracing) {
        long long h[8 * 64 * 2];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
        long long t0 = h[0] ? h[0] : h[1];
        fprintf(stderr, "trace dgrad=%d KB=%d (cycles rel. to first stamp) P:[free-wait done, arrived] M:[full done, committed] E:[acc_full done, released]\n", dgrad, g.KB);
        for (int i = 0; i < 24; ++i)
        {
            fprintf(stderr, "it %2d  P7 %7lld %7lld Mwait %7lld full %7lld | commit2 done %7lld acc_empty done %7lld\n", i, h[(320 + i) * 2] - t0, h[(320 + i) * 2 + 1] - t0, h[(384 + i) * 2] - t0, h[(384 + i) * 2 + 1] - t0, h[(448 + i) * 2] - t0, h[(448 + i) * 2 + 1] - t0);
            fprintf(stderr, "it %2d  P %7lld %7lld | M %7lld [mma %7lld %7lld] %7lld | E %7lld [ld %7lld rel %7lld] %7lld\n", i, h[i * 2] - t0, h[i * 2 + 1] - t0,
                    h[(64 + i) * 2] - t0, h[(256 + i) * 2] - t0, h[(256 + i) * 2 + 1] - t0, h[(64 + i) * 2 + 1] - t0, h[(128 + i) * 2] - t0, h[(192 + i) * 2] - t0, h[(192 + i) * 2 + 1] - t0, h[(128 + i) * 2 + 1] - t0);
        }
    }
    return CNN_OK;
}

int get_wgrad_plan(cnn_ctx* ctx, bool tf32, int Cin, int H, int W, int Cout, int k, int s, Plan** out) {
    const PlanKey key{2, tf32 ? 1 : 0, Cin, H, W, Cout, k, s};
    std::lock_guard<std::recursive_mutex> lk(cnn_global_mutex());
    auto& mp = plans();
    auto itp = mp.find({ctx->device, key});
    if (itp != mp.end()) { *out = &itp->second; return CNN_OK; }
    Plan p;
    WgradGemm& g = p.wg;
    const int kk = k * k;
    g.Mrows = Cin * kk + 1;
    g.Cin = Cin; g.H = H; g.W = W; g.Cout = Cout; g.OH = (H - k) / s + 1; g.OW = (W - k) / s + 1; g.s = s;
    const int mtiles = (g.Mrows + kRows - 1) / kRows;
    std::vector<int> table((size_t)mtiles * kRows, -2);
    for (int ci = 0; ci < Cin; ++ci)
        for (int ky = 0; ky < k; ++ky)
            for (int kx = 0; kx < k; ++kx) table[(size_t)ci * kk + ky * k + kx] = ci * H * W + ky * W + kx;
    table[(size_t)Cin * kk] = -1;
    const int Npad16 = ((Cout + 15) / 16) * 16;
    p.ntiles = (Npad16 + 255) / 256;
    g.Ntile = (((Npad16 + p.ntiles - 1) / p.ntiles) + 15) / 16 * 16;
    g.Npad = g.Ntile * p.ntiles;
    g.tmem_cols = next_pow2_cols(g.Ntile);
    const size_t stage_bytes = 2 * (size_t)kRows * 128 + 2 * (size_t)g.Ntile * 128;
    pick_stages(stage_bytes, 0, g.tmem_cols, &g.stages, &p.ctas_per_sm, &p.smem);
    if (const char* e = getenv("CNN_DBG_STAGES")) {  // experiment knobs, not part of the API
        g.stages = atoi(e);
        p.smem = kSmemHeader + (size_t)g.stages * stage_bytes;
    }
    if (const char* e = getenv("CNN_DBG_CTAS")) p.ctas_per_sm = atoi(e);
    if (int rc = upload_table(table, &p.d_table)) return rc;
    g.rowtab = p.d_table;
    if (int rc = tf32 ? set_smem_attr(wgrad_ws<true>, "wgrad_ws<tf32>") : set_smem_attr(wgrad_ws<false>, "wgrad_ws<bf16>"))
        return rc;
    // ---- TMA row-staged variant (TF32x3, 3x3 filters, output rows of at most 128 pixels)
    if (k == 3 && g.OW <= 128 && !getenv("CNN_DBG_NOROWS")) {
        WgradRows& r = p.wr;
        r.Mrows = g.Mrows; r.Cin = Cin; r.H = H; r.W = W; r.Cout = Cout; r.OH = g.OH; r.OW = g.OW; r.s = s; r.k = k;
        r.Ntile = g.Ntile; r.Npad = g.Npad; r.tmem_cols = g.tmem_cols;
        r.nci_max = std::min(Cin, 16);
        std::vector<int> rt((size_t)mtiles * kRows, -2);
        for (int ci = 0; ci < Cin; ++ci)
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) rt[(size_t)ci * kk + ky * k + kx] = (ci << 20) | (ky * W + kx);
        rt[(size_t)Cin * kk] = -1;
        const size_t stage_b = 2 * (size_t)kRows * 128 + 2 * (size_t)g.Ntile * 128;
        for (int TR = std::max(1, std::min(g.OH, 128 / g.OW)); TR >= 1 && !p.rows_ok; --TR) {
            r.TR = TR;
            r.SCI = (g.OH + TR - 1) / TR;
            r.xseg = (int)(((size_t)((TR - 1) * s + k) * W * 4 + 12 + 15) / 16 * 16);
            r.dseg = (int)(((size_t)TR * g.OW * 4 + 12 + 15) / 16 * 16);
            const size_t raw = (size_t)r.nci_max * r.xseg + (size_t)g.Ntile * r.dseg;
            const size_t fixed = 2 * raw + 2 * (16 + (size_t)g.Ntile) * sizeof(int) + 64;
            if (kSmemHeader + fixed + 2 * stage_b > kSmemMax - 1024) continue;
            if ((size_t)TR * g.OW * 4 < 256) break;   // tiny images: per-copy overhead dominates, keep the gather kernel
            int ctas = 1;
            size_t smem = 0;
            pick_stages(stage_b, fixed, g.tmem_cols, &r.stages, &ctas, &smem);
            p.rows_smem = smem;
            p.rows_ctas = ctas;
            p.rows_ok = true;
        }
        if (p.rows_ok) {
            if (int rc = upload_table(rt, &p.d_table2)) return rc;
            r.rowtab = p.d_table2;
            p.rows_wide = (r.Mrows + r.Ntile > 64) || p.rows_ctas < 2;   // 16 producer warps, one CTA per SM
            if (p.rows_wide) p.rows_ctas = 1;
            int rc;
            if (tf32) rc = p.rows_wide ? set_smem_attr(wgrad_rows_ws<true, 16, 1>, "wgrad_rows_ws<tf32,16>")
                                       : set_smem_attr(wgrad_rows_ws<true, 8, 2>, "wgrad_rows_ws<tf32,8>");
            else rc = p.rows_wide ? set_smem_attr(wgrad_rows_ws<false, 16, 1>, "wgrad_rows_ws<bf16,16>")
                                  : set_smem_attr(wgrad_rows_ws<false, 8, 2>, "wgrad_rows_ws<bf16,8>");
            if (rc) return rc;
        }
    }
    auto ins = mp.emplace(std::make_pair(ctx->device, key), p);
    *out = &ins.first->second;
    return CNN_OK;
}

}  // namespace

bool conv_tc_supported(int Cin, int Cout, int k, int s) {
    if (!(k == 1 || k == 3)) return false;
    if (s > 2 && k == 3) return false;      // dgrad segments are built for s in {1,2}
    if (s > k) return false;                // patches with uncovered cells: SIMT path
    if ((long long)Cin * k * k > (1 << 20) || (long long)Cout * k * k > (1 << 20)) return false;
    return true;
}

int conv_fwd_tc(cnn_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int Cin,
                int H, int W, int Cout, int k, int s) {
    Plan* p = nullptr;
    const bool tf32 = ctx->tc_precision != CNN_TC_BF16X3 && ctx->tc_precision != CNN_TC_BF16X1 && !getenv("CNN_DBG_BF16");
    if (int rc = get_gather_plan(ctx, 0, tf32, Cin, H, W, Cout, k, s, &p)) return rc;
    return run_gather_gemm(ctx, p, tf32, w, x, bias, y, B, 0, Cin, Cout);
}

int conv_dgrad_tc(cnn_ctx* ctx, const float* w, const float* delta, float* dx, int B, int Cin, int H, int W,
                  int Cout, int k, int s) {
    Plan* p = nullptr;
    const bool tf32 = ctx->tc_precision != CNN_TC_BF16X3 && ctx->tc_precision != CNN_TC_BF16X1;
    if (int rc = get_gather_plan(ctx, 1, tf32, Cin, H, W, Cout, k, s, &p)) return rc;
    return run_gather_gemm(ctx, p, tf32, w, delta, nullptr, dx, B, 1, Cin, Cout);
}

int conv_wgrad_tc(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B, int Cin,
                  int H, int W, int Cout, int k, int s, float scale) {
    Plan* p = nullptr;
    // TF32x3 unless the caller opted into MIXED / BF16X3 (BF16x3 weight gradient: half the K blocks and half
    // the accumulate steps, 2^-16 products)
    const bool tf32 = ctx->tc_precision == CNN_TC_TF32X3;
    if (int rc = get_wgrad_plan(ctx, tf32, Cin, H, W, Cout, k, s, &p)) return rc;
    if (p->rows_ok) {
        WgradRows r = p->wr;
        const int mtiles = (r.Mrows + kRows - 1) / kRows;
        r.B = B;
        r.nsc = (unsigned)B * (unsigned)r.SCI;
        long long want = (long long)ctx->sm_count * p->rows_ctas / ((long long)mtiles * p->ntiles);
        // accumulator-length cap: tcgen05 adds into its fp32 TMEM accumulator with truncation, a bias that
        // grows with the number of accumulate steps (measured 9e-5 at 15 k pixels per accumulator, 3.7e-4 at
        // 217 k: tools/vgg_parity_fullsize.py).  At most kMaxAccPixels pixels go into one accumulator; the
        // partials are then added in fp32 by wgrad_reduce_kernel.
        want = std::max(want, ((long long)r.nsc * r.TR * r.OW + kMaxAccPixels - 1) / kMaxAccPixels);
        if (want < 1) want = 1;
        if (want > r.nsc) want = r.nsc;
        r.sc_per_split = (unsigned)((r.nsc + want - 1) / want);
        const unsigned splits = (r.nsc + r.sc_per_split - 1) / r.sc_per_split;
        const size_t pbytes = sizeof(float) * (size_t)splits * r.Mrows * r.Npad;
        float* partial = cnn_scratch(ctx, pbytes + 64);
        CNN_REQUIRE(partial, "scratch allocation failed");
        partial = reinterpret_cast<float*>(((uintptr_t)partial + 15) & ~(uintptr_t)15);
        CNN_REQUIRE((((uintptr_t)x | (uintptr_t)delta) & 15) == 0, "conv_tc: operand base pointers must be 16-byte aligned");
        r.x = x; r.delta = delta; r.partial = partial;
        r.single = ctx->tc_precision == CNN_TC_BF16X1 ? 1 : 0;
        r.x_bytes16 = ((long long)B * Cin * H * W * 4 + 15) & ~15ll;
        r.d_bytes16 = ((long long)B * Cout * r.OH * r.OW * 4 + 15) & ~15ll;
        dim3 grid((unsigned)mtiles, splits, (unsigned)p->ntiles);
        if (tf32) {
            if (p->rows_wide) { CNN_LAUNCH(ctx, (wgrad_rows_ws<true, 16, 1>), grid, 18 * 32, p->rows_smem, r); }
            else { CNN_LAUNCH(ctx, (wgrad_rows_ws<true, 8, 2>), grid, 10 * 32, p->rows_smem, r); }
        } else {
            if (p->rows_wide) { CNN_LAUNCH(ctx, (wgrad_rows_ws<false, 16, 1>), grid, 18 * 32, p->rows_smem, r); }
            else { CNN_LAUNCH(ctx, (wgrad_rows_ws<false, 8, 2>), grid, 10 * 32, p->rows_smem, r); }
        }
        int rgrid = cdiv((long long)r.Mrows * Cout, 32);
        if (rgrid > ctx->sm_count * 16) rgrid = ctx->sm_count * 16;
        CNN_LAUNCH(ctx, wgrad_reduce_kernel, rgrid, 256, 0, partial, dw, db, r.Mrows, r.Npad, Cout, (int)splits, scale);
        return CNN_OK;
    }
    WgradGemm g = p->wg;
    const long long P = (long long)B * g.OH * g.OW;
    CNN_REQUIRE(P < (1ll << 31) - 256, "conv_tc: too many pixels");
    CNN_REQUIRE((long long)B * Cin * H * W < (1ll << 31) && P * Cout < (1ll << 31), "conv_tc: tensor too large");
    const int KBLK = tf32 ? 32 : 64;
    g.P = (unsigned)P;
    g.nchunks = (unsigned)((P + KBLK - 1) / KBLK);
    const int mtiles = (g.Mrows + kRows - 1) / kRows;
    // split the pixel range so that the grid fills the resident CTA slots (~2 waves at most)
    long long want = (long long)ctx->sm_count * p->ctas_per_sm / ((long long)mtiles * p->ntiles);
    want = std::max(want, (P + kMaxAccPixels - 1) / kMaxAccPixels);   // accumulator-length cap, see above
    if (want < 1) want = 1;
    if (want > g.nchunks) want = g.nchunks;
    g.chunks_per_split = (unsigned)((g.nchunks + want - 1) / want);
    const unsigned splits = (g.nchunks + g.chunks_per_split - 1) / g.chunks_per_split;
    const size_t pbytes = sizeof(float) * (size_t)splits * g.Mrows * g.Npad;
    float* partial = cnn_scratch(ctx, pbytes + 64);
    CNN_REQUIRE(partial, "scratch allocation failed");
    partial = reinterpret_cast<float*>(((uintptr_t)partial + 15) & ~(uintptr_t)15);
    g.x = x; g.delta = delta; g.partial = partial;
    g.single = ctx->tc_precision == CNN_TC_BF16X1 ? 1 : 0;
    dim3 grid((unsigned)mtiles, splits, (unsigned)p->ntiles);
    if (tf32) { CNN_LAUNCH(ctx, wgrad_ws<true>, grid, kThreads, p->smem, g); }
    else { CNN_LAUNCH(ctx, wgrad_ws<false>, grid, kThreads, p->smem, g); }
    int rgrid = cdiv((long long)g.Mrows * Cout, 32);
    if (rgrid > ctx->sm_count * 16) rgrid = ctx->sm_count * 16;
    CNN_LAUNCH(ctx, wgrad_reduce_kernel, rgrid, 256, 0, partial, dw, db, g.Mrows, g.Npad, Cout, (int)splits, scale);
    return CNN_OK;
}
