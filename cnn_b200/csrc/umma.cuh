// umma.cuh -- hand-written sm_100a plumbing for the tensor-core kernels: mbarriers, the TMA
// bulk-copy engine (cp.async.bulk, SASS UBLKCP), TMEM allocation, tcgen05.mma / commit / ld,
// shared-memory matrix descriptors, and the fp32 -> (bf16 hi, bf16 lo) operand split.
//
// Operand layout used everywhere here: K-major tiles with the 128-byte swizzle
// (UMMA LayoutType::SWIZZLE_128B): a tile is [rows][64 bf16] with 128-byte rows, 8-row groups
// 1024 bytes apart, tile base 1024-byte aligned; the 16-byte chunk c of row r lives at
//   r*128 + ((c ^ (r & 7)) << 4).
// One tcgen05.mma (kind::f16, bf16 in, fp32 accumulate) consumes K = 16 = 32 bytes of every row,
// so the k-th step of a 64-wide block starts 32*k bytes into the tile.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in HW, not in a spin loop
        : "memory");
    return ok != 0;
}
// non-blocking probe (no suspend)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// One lane of a converged warp.  Issuing tcgen05.mma / commit under this predicate (instead of
// `lane == 0`) lets the compiler keep descriptors in uniform registers and emit the UTCHMMA
// directly; under a divergent branch it wraps every MMA in an elect/branch "waterfall" loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- TMA bulk copy global -> shared, completion on an mbarrier ------------------------------
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- TMA tensor-map copy (cp.async.bulk.tensor, SASS UTMALDG): a 3-D box of a strided global tensor ->
// dense shared memory (row pitch = box width); elements of the box outside the tensor are zero-filled
// and still count in the transaction bytes.  The map is built on the host (cnn_tmap_encode_3d, ctx.cu)
// and passed as a __grid_constant__ kernel parameter.  dst must be 128-byte aligned.
__device__ __forceinline__ void tma_tensor3d_g2s(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_tensor4d_g2s(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_tensor5d_g2s(void* smem_dst, const void* tmap, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// L2 prefetch of a global range (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2(const void* gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM ------------------------------------------------------------------------------------
// Whole-warp calls. ncols: power of two in [32, 512]. The base address lands in *smem_slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: start>>4 [0,14), LBO>>4 [16,30) (unused
// for swizzled K-major, conventionally 1), SBO>>4 [32,46) = 1024 B between 8-row groups,
// version 1 at [46,48), layout type 2 (SWIZZLE_128B) at [61,64).
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor, kind::f16: D fp32 (bit 4), A and B bf16 (bits 7, 10), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// shared-memory matrix descriptor, no swizzle (layout type 0): start >> 4, LBO >> 4 at [16,30),
// SBO >> 4 at [32,46), version 1 at [46,48).
//   K-major : rows 16 B apart inside an 8-row core matrix, SBO = bytes between 8-row groups,
//             LBO = bytes between the two 16-byte K chunks of one MMA
//   MN-major: K rows 16 B apart inside a core matrix, SBO = bytes between 8-element MN chunks,
//             LBO = bytes between groups of 8 K rows
__device__ __forceinline__ uint64_t desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N) {   // both operands MN-major (bits 15, 16)
    return idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}

// kind::tf32: A and B fp32 words read as TF32 (format code 2), K = 8 per instruction.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32-bit, 16 consecutive columns per call ----------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// the same load without the wait: issue several, then one tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- fp32 -> bf16 hi + bf16 lo (x ~= hi + lo, 16 mantissa bits kept) ---------------------------
// Three bf16 MMAs hi*hi + hi*lo + lo*hi then reproduce the fp32 product to ~2^-16 relative
// (normwise error ~5e-6 after accumulation, SURVEY §7 hard part 1).
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    // cvt.rn.bf16x2.f32 d, x, y : x -> upper half, y -> lower half
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16);
    const float rb = b - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// ---- fp32 -> tf32 hi + fp32 lo (x == hi + lo exactly; lo has <= 13 significant bits) -----------
// hi is x truncated to TF32, lo = x - hi is exact in fp32 (<= 13 significant bits) and is itself read
// as TF32 by the tensor core (residual <= 2^-21 |x|).  hi*hi + hi*lo + lo*hi then differs from the fp32 product
// only by the dropped lo*lo term (<= 2^-22 relative): fp32-grade results from the tensor pipe.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;  // truncation to TF32 (2 instructions per element in total)
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
}

// x = hi + mid + lo with three bf16 pieces (8 significand bits each, 24 in total: the split is exact
// for normal fp32 values).  The forward pass multiplies with all three pieces (six MMAs per K step,
// dropped terms <= 2^-24): a ReLU / max-pool decision taken on the forward output must not flip
// against the fp32 reference, which a 16-bit split (~5e-6) would allow a few times per batch.
__device__ __forceinline__ void split3(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(rb), "f"(ra));
    const float sa = ra - __uint_as_float(mid << 16), sb = rb - __uint_as_float(mid & 0xFFFF0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(sb), "f"(sa));
}
__device__ __forceinline__ void split8x3(const float* v, uint4& hi, uint4& mid, uint4& lo) {
    split3(v[0], v[1], hi.x, mid.x, lo.x);
    split3(v[2], v[3], hi.y, mid.y, lo.y);
    split3(v[4], v[5], hi.z, mid.z, lo.z);
    split3(v[6], v[7], hi.w, mid.w, lo.w);
}

__device__ __forceinline__ uint32_t swz128(int row, int chunk16) {
    return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

}  // namespace umma
