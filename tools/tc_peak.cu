// tc_peak.cu -- tcgen05 issue-rate micro-benchmark ("measure TF32 with a tcgen05 micro-bench on the box").
// Every CTA (one per SM, or two with -c 2) keeps two operand tiles resident in shared memory and one elected
// thread issues M=128 x N=256 MMAs back to back into two alternating TMEM accumulators; no operand traffic, no
// epilogue -- the ceiling any tensor-core kernel of this library can reach for kind::f16 (bf16 in, fp32 accumulate,
// K = 16 per instruction) and kind::tf32 (K = 8).  Split-operand schemes divide it by their pass count
// (3: hi*hi + hi*lo + lo*hi; 6: three bf16 pieces, conv_s1.cu / conv_s2.cu forward).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Icnn_b200/csrc -o tools/_build/tc_peak tools/tc_peak.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace umma;

template <bool TF32>
__global__ void __launch_bounds__(128) peak_kernel(int iters, int N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 128-byte-swizzled K-major tiles: A 128 rows x 128 B, B N rows x 128 B; contents: small non-zero bit patterns
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < (128 + N) * 32; i += blockDim.x) w[i] = TF32 ? 0x3F800000u + (i * 2654435761u >> 12) : 0x3F803F80u + ((i * 2654435761u >> 20) & 0x007F007Fu);
    if (warp == 0) {
        tmem_alloc(&tslot, 512u);
        if (lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0) {
        const uint32_t idesc = TF32 ? idesc_tf32(128, N) : idesc_bf16(128, N);
        const uint64_t a0 = smem_desc_k128(smem_u32(smem)), b0 = smem_desc_k128(smem_u32(smem) + 128 * 128);
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {   // four K-steps of a 128-byte row, accumulators alternate
                    if (TF32) mma_tf32(tmem + (j & 1) * 256, a0 + 2 * j, b0 + 2 * j, idesc, it | (j >> 1));
                    else mma_bf16(tmem + (j & 1) * 256, a0 + 2 * j, b0 + 2 * j, idesc, it | (j >> 1));
                }
            }
            mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512u);
}

template <bool TF32>
double run(int ctas_per_sm, int iters, int N, int sms) {
    const size_t smem = (128 + N) * 128 + 1024;
    cudaFuncSetAttribute(peak_kernel<TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * ctas_per_sm;
    peak_kernel<TF32><<<grid, 128, smem>>>(iters / 10, N);   // warm-up
    double best = 0;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        peak_kernel<TF32><<<grid, 128, smem>>>(iters, N);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 128 * N * (TF32 ? 8 : 16) * 4.0 * iters * grid;
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return best;
}

// bf16, M = 128: the same issue loop over other operand layouts (timing only; operand contents are whatever the fill
// left).  layout 1 = no swizzle, K-major, core matrices of 8 rows x 16 B (what conv_s1 / conv_s2 stage: LBO between
// the two K chunks of an MMA, SBO = 128 B between 8-row groups), optionally with the A start address shifted by
// `shift` 16-byte rows (the kx taps); layout 2 = SWIZZLE_32B K-major (rows of 32 B, 256 B per 8-row atom).
__global__ void __launch_bounds__(128) layout_kernel(int iters, int N, int layout, int shift, int same) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < (160 + N) * 64 / 4; i += blockDim.x) w[i] = 0x3F803F80u + ((i * 2654435761u >> 20) & 0x007F007Fu);
    if (warp == 0) {
        tmem_alloc(&tslot, 512u);
        if (lane == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tslot;
    if (warp == 0) {
        const uint32_t idesc = idesc_bf16(128, N);
        const uint32_t sa = smem_u32(smem), sb = sa + 160 * 64;      // A: up to 160 rows x 4 K chunks of 16 B
        uint64_t a[4], b[4];
        for (int j = 0; j < 4; ++j) {
            if (layout == 1) {   // [k chunk][row][16 B]: chunk pairs (2j, 2j+1) -> j & 1, second pair reuses the first's bytes
                a[j] = desc_nosw(sa + (uint32_t)(j & 1) * 2 * 160 * 16 + (uint32_t)shift * 16, 160 * 16, 128);
                b[j] = desc_nosw(sb + (uint32_t)(j & 1) * 2 * N * 16, (uint32_t)N * 16, 128);
            } else {             // SWIZZLE_32B: one K step per 32-byte row, 8-row atoms of 256 B
                a[j] = (uint64_t)(((sa + (uint32_t)(j & 1) * 128 * 32) & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) |
                       ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
                b[j] = (uint64_t)(((sb + (uint32_t)(j & 1) * N * 32) & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) |
                       ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
            }
        }
        if (elect_one()) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 4; ++j) mma_bf16(tmem + (same ? 0 : (j & 1) * 256), a[j], b[j], idesc, same ? (it | j) : (it | (j >> 1)));
            }
            mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512u);
}

double run_layout(int iters, int N, int sms, int layout, int shift, int same = 0) {
    const size_t smem = (160 + N) * 64 + 1024;
    cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    layout_kernel<<<sms, 128, smem>>>(iters / 10, N, layout, shift, same);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        layout_kernel<<<sms, 128, smem>>>(iters, N, layout, shift, same);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::max(best, 2.0 * 128 * N * 16 * 4.0 * iters * sms / (ms * 1e-3) / 1e12);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return best;
}

int main(int argc, char** argv) {
    int sms = 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
    sms = prop.multiProcessorCount;
    const int iters = argc > 1 ? atoi(argv[1]) : 200000;
    printf("{\"device\": \"%s\", \"sms\": %d, \"iters\": %d", prop.name, sms, iters);
    for (int N : {256, 128, 64}) {
        // TMEM: two accumulators of 256 columns each are allocated whatever N is; two CTAs per SM cannot both hold 512 columns
        printf(", \"bf16_m128_n%d_tflops\": %.1f", N, run<false>(1, iters, N, sms));
        printf(", \"tf32_m128_n%d_tflops\": %.1f", N, run<true>(1, iters / 2, N, sms));
        fflush(stdout);
    }
    for (int N : {256, 128, 64}) {
        printf(", \"bf16_noswizzle_n%d_tflops\": %.1f", N, run_layout(iters / 2, N, sms, 1, 0));
        printf(", \"bf16_noswizzle_shift1_n%d_tflops\": %.1f", N, run_layout(iters / 2, N, sms, 1, 1));
        printf(", \"bf16_swizzle32_n%d_tflops\": %.1f", N, run_layout(iters / 2, N, sms, 2, 0));
        // every MMA accumulates into the same TMEM columns (a dependent chain) instead of two alternating accumulators
        printf(", \"bf16_noswizzle_same_acc_n%d_tflops\": %.1f", N, run_layout(iters / 2, N, sms, 1, 0, 1));
        fflush(stdout);
    }
    printf("}\n");
    return 0;
}
