// conv_tc.cu -- Conv2D forward and input-gradient as implicit GEMMs on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a only.
//
// Numerics: Blackwell has no IEEE-fp32 MMA, and single-pass TF32 misses the 1e-4 parity bar
// (SURVEY §7 hard part 1).  Every fp32 operand is split into two bf16 values (x ~= hi + lo) and
// each K-step issues three bf16 MMAs (hi*hi + hi*lo + lo*hi) into the same fp32 TMEM
// accumulator: ~2^-16 relative per product, ~5e-6 normwise after accumulation.
//
// One kernel serves both directions ("gather GEMM"):
//   D[m][n] = sum_k A[m][k] * Bw[n][k]
//   forward : m = output pixel (b,oy,ox), n = out channel, k = (ci,ky,kx) in the reference's
//             filter order; A[m][k] = x[b][ci][oy*s+ky][ox*s+kx]             (conv2d.cpp:69-92)
//   dgrad   : m = s x s input patch (b,py,px), one GEMM ("segment") per cell (pr,pc) of the patch,
//             n = in channel, k = (co, tap with ky%s==pr, kx%s==pc);
//             A[m][k] = delta[b][co][py-ky/s][px-kx/s] or 0 outside    (gather form of conv2d.cpp:192)
// Data path per 128-row tile and 64-wide K block (2-stage ring):
//   * the filter block Bw (pre-split to bf16 hi/lo and pre-swizzled by pack_*_kernel) arrives by
//     ONE TMA bulk copy (cp.async.bulk -> mbarrier complete_tx),
//   * the 128 threads gather the activation rows from NCHW global memory (lane = pixel, so every
//     load instruction is coalesced along W), split them and store 16-byte chunks into the
//     128B-swizzled K-major A tiles,
//   * thread 0 issues the tcgen05.mma chain and tcgen05.commit's the stage back to the producers,
//   * epilogue: tcgen05.ld 32x32b (lane = row), + bias, coalesced NCHW stores.
// Several CTAs are resident per SM (stage memory scales with N), so gather, MMA and epilogue of
// different tiles overlap; the kernel is HBM/LSU-bound for the AlexNet-lite shapes.
#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "umma.cuh"

namespace {

using namespace umma;

constexpr int kRows = 128;   // GEMM rows per CTA == threads per CTA == TMEM lanes
constexpr int kBK = 64;      // K block: one 128-byte swizzle row of bf16
constexpr int kMaxSeg = 4;
constexpr int kPadCode = 15;

struct Segment {
    int K;          // valid K entries
    int KB;         // K blocks
    int ntaps;      // taps for the validity mask (0: rows are always in bounds)
    int tab_off;    // offset (entries) of this segment's gather table
    int b_off;      // offset (bytes) of this segment's packed filter blocks for n-tile 0
    int pr, pc;     // output cell inside the patch
    signed char dy[9], dx[9];
};

struct GatherGemm {
    const float* src;        // activations (x or delta), NCHW
    const int* table;        // per k: (offset << 4) | tap code
    const uint8_t* packedB;  // [seg][ntile][kb]{hi[Ntile][64], lo[Ntile][64]} bf16, swizzled
    const float* bias;       // may be null
    float* dst;
    int nseg;
    Segment seg[kMaxSeg];
    // row space: m -> (b, gy, gx)
    int GH, GW;
    long long rows;          // B*GH*GW
    // source geometry
    int SC, SH, SW, sy, sx;
    // output geometry: dst[b][n][gy*os+pr][gx*os+pc]
    int ON, OHt, OWt, os;
    int Ntile;               // columns per CTA (multiple of 16, <= 256)
    int tmem_cols;           // power of two >= nseg*Ntile
};

// --------------------------------------------------------------------------- pack kernels
// Forward: Bw[n][k] = w[n][k] (k = ci*kk + tap is exactly the reference's filter memory order).
// Input gradient: Bw[n=ci][k=(co,t)] = w[co][ci][tap_t].
// Output block (n-tile nt, k block kb): hi tile then lo tile, each [Ntile][64] bf16, swizzled.
struct PackSeg {
    int K, KB, ntaps, b_off;
    signed char tap[9];  // filter tap index ky*k+kx of tap t (dgrad)
};

__global__ void pack_filters_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, PackSeg sg,
                                    int dgrad, int Cin, int Cout, int kk, int Nreal, int Ntile,
                                    int ntiles) {
    // one thread per 16-byte chunk: (nt, kb, n, c)
    const long long total = (long long)ntiles * sg.KB * Ntile * 8;
    for (long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x; id < total;
         id += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(id & 7);
        long long t = id >> 3;
        const int nl = (int)(t % Ntile);
        t /= Ntile;
        const int kb = (int)(t % sg.KB);
        const int nt = (int)(t / sg.KB);
        const int n = nt * Ntile + nl;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = kb * kBK + c * 8 + j;
            float val = 0.f;
            if (n < Nreal && k < sg.K) {
                if (!dgrad) {
                    val = w[(size_t)n * Cin * kk + k];
                } else {
                    const int co = k / sg.ntaps, tt = k % sg.ntaps;
                    val = w[((size_t)co * Cin + n) * kk + sg.tap[tt]];
                }
            }
            v[j] = val;
        }
        uint4 hi, lo;
        split2(v[0], v[1], hi.x, lo.x);
        split2(v[2], v[3], hi.y, lo.y);
        split2(v[4], v[5], hi.z, lo.z);
        split2(v[6], v[7], hi.w, lo.w);
        uint8_t* blk = out + sg.b_off + ((size_t)nt * sg.KB + kb) * (size_t)(2 * Ntile * 128);
        *reinterpret_cast<uint4*>(blk + swz128(nl, c)) = hi;
        *reinterpret_cast<uint4*>(blk + (size_t)Ntile * 128 + swz128(nl, c)) = lo;
    }
}

// --------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(kRows)
gather_gemm_kernel(const GatherGemm g) {
    extern __shared__ uint8_t smem_raw[];
    // carve: [barriers | tmem slot | table stages] then 1024-aligned operand stages
    uint64_t* bar_free = reinterpret_cast<uint64_t*>(smem_raw);       // [2]
    uint64_t* bar_bfull = bar_free + 2;                               // [2]
    uint64_t* bar_acc = bar_free + 4;                                 // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_free + 5);
    int* tab = reinterpret_cast<int*>(smem_raw + 64);                 // [2][64]
    const uint32_t base_u32 = smem_u32(smem_raw);
    const uint32_t op_off = ((base_u32 + 64 + 512 + 1023) & ~1023u) - base_u32;
    uint8_t* ops = smem_raw + op_off;
    const int Ntile = g.Ntile;
    const uint32_t a_bytes = kRows * 128;              // one A tile (hi or lo)
    const uint32_t b_bytes = (uint32_t)Ntile * 128;    // one B tile (hi or lo)
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int nt = blockIdx.y;

    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
    if (tid == 0) {
        mbar_init(&bar_free[0], 1);
        mbar_init(&bar_free[1], 1);
        mbar_init(&bar_bfull[0], 1);
        mbar_init(&bar_bfull[1], 1);
        mbar_init(bar_acc, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- this thread's row
    const long long m = (long long)blockIdx.x * kRows + tid;
    const bool row_ok = m < g.rows;
    int b = 0, gy = 0, gx = 0;
    if (row_ok) {
        gx = (int)(m % g.GW);
        const long long t = m / g.GW;
        gy = (int)(t % g.GH);
        b = (int)(t / g.GH);
    }
    const float* src_row =
        g.src + (size_t)b * g.SC * g.SH * g.SW + (size_t)(gy * g.sy) * g.SW + (size_t)gx * g.sx;

    const uint32_t idesc = idesc_bf16(kRows, Ntile);
    int it = 0;  // global K-block iteration (stage ring position)
    for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const Segment& sg = g.seg[sgi];
        // validity mask over this segment's taps for this row (bit 0 always set for ntaps == 0)
        uint32_t vmask = 0;
        if (row_ok) {
            if (sg.ntaps == 0) vmask = 1;
            else
                for (int t = 0; t < sg.ntaps; ++t) {
                    const int yy = gy - sg.dy[t], xx = gx - sg.dx[t];
                    if (yy >= 0 && yy < g.SH && xx >= 0 && xx < g.SW) vmask |= 1u << t;
                }
        }
        const uint8_t* bsrc = g.packedB + sg.b_off + (size_t)nt * sg.KB * (2 * b_bytes);
        for (int kb = 0; kb < sg.KB; ++kb, ++it) {
            const int s = it & 1, u = it >> 1;
            uint8_t* stage = ops + (size_t)s * stage_bytes;
            if (it >= 2) mbar_wait(&bar_free[s], (uint32_t)((u - 1) & 1));
            if (tid == 0) {
                mbar_expect_tx(&bar_bfull[s], 2 * b_bytes);
                tma_bulk_g2s(stage + 2 * a_bytes, bsrc + (size_t)kb * (2 * b_bytes), 2 * b_bytes, &bar_bfull[s]);
            }
            if (tid < kBK) tab[s * kBK + tid] = g.table[sg.tab_off + kb * kBK + tid];
            __syncthreads();
            const int kvalid = min(kBK, sg.K - kb * kBK);
            const int ksteps = (kvalid + 15) >> 4;
            // ---- gather + split + swizzled store of this thread's row
            const int* tb = tab + s * kBK;
            for (int c = 0; c < 2 * ksteps; ++c) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int e = tb[c * 8 + j];
                    const bool ok = (vmask >> (e & 15)) & 1u;
                    v[j] = ok ? __ldg(src_row + (e >> 4)) : 0.f;
                }
                uint4 hi, lo;
                split2(v[0], v[1], hi.x, lo.x);
                split2(v[2], v[3], hi.y, lo.y);
                split2(v[4], v[5], hi.z, lo.z);
                split2(v[6], v[7], hi.w, lo.w);
                const uint32_t o = swz128(tid, c);
                *reinterpret_cast<uint4*>(stage + o) = hi;
                *reinterpret_cast<uint4*>(stage + a_bytes + o) = lo;
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                mbar_wait(&bar_bfull[s], (uint32_t)(u & 1));
                tc_fence_after();
                const uint32_t sa = smem_u32(stage);
                const uint32_t d = tmem_base + (uint32_t)(sgi * Ntile);
                for (int j = 0; j < ksteps; ++j) {
                    const uint64_t ahi = smem_desc_k128(sa + 32 * j);
                    const uint64_t alo = smem_desc_k128(sa + a_bytes + 32 * j);
                    const uint64_t bhi = smem_desc_k128(sa + 2 * a_bytes + 32 * j);
                    const uint64_t blo = smem_desc_k128(sa + 2 * a_bytes + b_bytes + 32 * j);
                    mma_bf16(d, alo, bhi, idesc, (kb | j) != 0);
                    mma_bf16(d, ahi, blo, idesc, true);
                    mma_bf16(d, ahi, bhi, idesc, true);
                }
                mma_commit(&bar_free[s]);
            }
        }
    }
    if (tid == 0) mma_commit(bar_acc);
    mbar_wait(bar_acc, 0);
    tc_fence_after();

    // ---- epilogue: lane = row; 16 columns per tcgen05.ld
    const size_t oplane = (size_t)g.OHt * g.OWt;
    for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const Segment& sg = g.seg[sgi];
        const int oy = gy * g.os + sg.pr, ox = gx * g.os + sg.pc;
        const bool st_ok = row_ok && oy < g.OHt && ox < g.OWt;
        float* drow = g.dst + (size_t)b * g.ON * oplane + (size_t)oy * g.OWt + ox;
        for (int c0 = 0; c0 < Ntile; c0 += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sgi * Ntile + c0), v);
            if (st_ok) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = nt * Ntile + c0 + j;
                    if (n < g.ON) drow[(size_t)n * oplane] = v[j] + (g.bias ? g.bias[n] : 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// --------------------------------------------------------------------------- host side
struct PlanKey {
    int dgrad, Cin, H, W, Cout, k, s;
    bool operator<(const PlanKey& o) const {
        return std::tie(dgrad, Cin, H, W, Cout, k, s) < std::tie(o.dgrad, o.Cin, o.H, o.W, o.Cout, o.k, o.s);
    }
};

struct Plan {
    GatherGemm g{};          // src/dst/bias/packedB/rows filled per call
    PackSeg pseg[kMaxSeg]{};
    int* d_table = nullptr;  // device gather tables (owned; lives as long as the process)
    int ntiles = 1, Nreal = 0;
    size_t packed_bytes = 0;
    size_t smem = 0;
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

// Builds (once per device and layer geometry) the gather tables and the tiling of a conv GEMM.
int get_plan(cnn_ctx* ctx, int dgrad, int Cin, int H, int W, int Cout, int k, int s, Plan** out) {
    const PlanKey key{dgrad, Cin, H, W, Cout, k, s};
    auto& mp = plans();
    auto itp = mp.find({ctx->device, key});
    if (itp != mp.end()) { *out = &itp->second; return CNN_OK; }
    const int OH = (H - k) / s + 1, OW = (W - k) / s + 1, kk = k * k;
    Plan p;
    GatherGemm& g = p.g;
    std::vector<int> table;
    const int Nreal = dgrad ? Cin : Cout;
    p.Nreal = Nreal;
    int nseg = dgrad ? s * s : 1;
    // n tiling: nseg * Ntile <= 512 TMEM columns, Ntile <= 256, multiple of 16
    int max_tile = 512 / nseg;
    if (max_tile > 256) max_tile = 256;
    max_tile = (max_tile / 16) * 16;
    const int Npad = ((Nreal + 15) / 16) * 16;
    p.ntiles = (Npad + max_tile - 1) / max_tile;
    int Ntile = (((Npad + p.ntiles - 1) / p.ntiles) + 15) / 16 * 16;
    g.Ntile = Ntile;
    g.nseg = nseg;
    g.tmem_cols = next_pow2_cols(nseg * Ntile);
    size_t boff = 0;
    for (int sgi = 0; sgi < nseg; ++sgi) {
        Segment& sg = g.seg[sgi];
        PackSeg& ps = p.pseg[sgi];
        sg.tab_off = (int)table.size();
        if (!dgrad) {
            sg.K = Cin * kk; sg.ntaps = 0; sg.pr = sg.pc = 0;
            for (int ci = 0; ci < Cin; ++ci)
                for (int ky = 0; ky < k; ++ky)
                    for (int kx = 0; kx < k; ++kx) table.push_back(((ci * H * W + ky * W + kx) << 4) | 0);
        } else {
            sg.pr = sgi / s; sg.pc = sgi % s;
            int nt = 0;
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx)
                    if (ky % s == sg.pr && kx % s == sg.pc) {
                        sg.dy[nt] = (signed char)(ky / s); sg.dx[nt] = (signed char)(kx / s);
                        ps.tap[nt] = (signed char)(ky * k + kx);
                        ++nt;
                    }
            sg.ntaps = nt;
            sg.K = Cout * nt;
            for (int co = 0; co < Cout; ++co)
                for (int t = 0; t < nt; ++t)
                    table.push_back(((co * OH * OW - sg.dy[t] * OW - sg.dx[t]) * 16) | t);
        }
        sg.KB = (sg.K + kBK - 1) / kBK;
        while ((int)table.size() < sg.tab_off + sg.KB * kBK) table.push_back(kPadCode);
        sg.b_off = (int)boff;
        ps.K = sg.K; ps.KB = sg.KB; ps.ntaps = sg.ntaps; ps.b_off = sg.b_off;
        boff += (size_t)p.ntiles * sg.KB * 2 * Ntile * 128;
    }
    p.packed_bytes = boff;
    if (!dgrad) {
        g.GH = OH; g.GW = OW; g.SC = Cin; g.SH = H; g.SW = W; g.sy = g.sx = s;
        g.ON = Cout; g.OHt = OH; g.OWt = OW; g.os = 1;
    } else {
        g.GH = (H + s - 1) / s; g.GW = (W + s - 1) / s; g.SC = Cout; g.SH = OH; g.SW = OW; g.sy = g.sx = 1;
        g.ON = Cin; g.OHt = H; g.OWt = W; g.os = s;
    }
    p.smem = 64 + 512 + 1024 + 2 * (size_t)(2 * kRows * 128 + 2 * Ntile * 128);
    if (cudaMalloc(&p.d_table, table.size() * sizeof(int)) != cudaSuccess ||
        cudaMemcpy(p.d_table, table.data(), table.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
        cnn_set_error("conv_tc: gather table upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        return CNN_ERR_CUDA;
    }
    g.table = p.d_table;
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return cnn_cuda_fail(e, "cudaFuncSetAttribute(gather_gemm_kernel)", __FILE__, __LINE__);
    auto ins = mp.emplace(std::make_pair(ctx->device, key), p);
    *out = &ins.first->second;
    return CNN_OK;
}

int run_gather_gemm(cnn_ctx* ctx, Plan* p, const float* w, const float* src, const float* bias, float* dst,
                    int B, int dgrad, int Cin, int Cout, int kk) {
    // a segment with zero taps (possible when s > k) leaves its cells at 0: handled by the caller
    uint8_t* packed = reinterpret_cast<uint8_t*>(cnn_scratch(ctx, p->packed_bytes + 1024));
    CNN_REQUIRE(packed, "scratch allocation failed");
    packed = reinterpret_cast<uint8_t*>(((uintptr_t)packed + 1023) & ~(uintptr_t)1023);
    GatherGemm g = p->g;
    for (int sgi = 0; sgi < g.nseg; ++sgi) {
        const PackSeg& ps = p->pseg[sgi];
        if (ps.KB == 0) continue;
        const long long chunks = (long long)p->ntiles * ps.KB * g.Ntile * 8;
        int grid = cdiv(chunks, 256);
        if (grid > ctx->sm_count * 8) grid = ctx->sm_count * 8;
        CNN_LAUNCH(ctx, pack_filters_kernel, grid, 256, 0, w, packed, ps, dgrad, Cin, Cout, kk, p->Nreal,
                   g.Ntile, p->ntiles);
    }
    g.src = src; g.bias = bias; g.dst = dst; g.packedB = packed;
    g.rows = (long long)B * g.GH * g.GW;
    dim3 grid((unsigned)((g.rows + kRows - 1) / kRows), (unsigned)p->ntiles);
    CNN_LAUNCH(ctx, gather_gemm_kernel, grid, kRows, p->smem, g);
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
    if (int rc = get_plan(ctx, 0, Cin, H, W, Cout, k, s, &p)) return rc;
    return run_gather_gemm(ctx, p, w, x, bias, y, B, 0, Cin, Cout, k * k);
}

int conv_dgrad_tc(cnn_ctx* ctx, const float* w, const float* delta, float* dx, int B, int Cin, int H, int W,
                  int Cout, int k, int s) {
    Plan* p = nullptr;
    if (int rc = get_plan(ctx, 1, Cin, H, W, Cout, k, s, &p)) return rc;
    return run_gather_gemm(ctx, p, w, delta, nullptr, dx, B, 1, Cin, Cout, k * k);
}

int conv_wgrad_tc(cnn_ctx* ctx, const float* x, const float* delta, float* dw, float* db, int B, int Cin,
                  int H, int W, int Cout, int k, int s, float scale) {
    return conv_wgrad_simt(ctx, x, delta, dw, db, B, Cin, H, W, Cout, k, s, scale);
}
