// Microbenchmark (round 2): what bounds small-N tcgen05.mma?  tools/mma_rate.cu showed a flat ~47 cycles per
// instruction for every M in {64,128} and N <= 64 (so NOT the shared-memory operand port).  This one separates
//   (1) per-issuing-thread floor vs per-SM floor: W warps of one CTA each issue to their own accumulator;
//   (2) per-instruction floor vs per-SM-pipe floor: cta_group::2 (M = 256 over a CTA pair, one instruction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I disconet_b200/csrc tools/mma_rate2.cu -o tools/mma_rate2
#include "common.cuh"
#include <cstdio>
void disco_set_error(const char*, ...) {}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// ---------------- (1) several issuing warps in one CTA ---------------------------------------------------------
__global__ void __launch_bounds__(256) multi_issuer_kernel(int n, int nwarps, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 90 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = tbase;
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && warp < nwarps) {
        const uint32_t idesc = umma_idesc_f16(1, 128, n);
        const uint32_t a0 = (smem_u32(smem) + 1023u) & ~1023u, b0 = a0 + 48 * 1024;
        uint64_t da[8], db[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            da[i] = umma_desc_kmajor_noswizzle(a0 + warp * 4096 + (i & 3) * 16, 2912, 160);
            db[i] = umma_desc_kmajor_noswizzle(b0, (uint32_t)n * 16, 128);
        }
        const uint32_t tacc = td + (uint32_t)(warp * (512 / nwarps));
#pragma unroll 1
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) umma_f16(tacc, da[i], db[i], idesc, 1);
        }
        umma_commit(smem_u32(&bar[warp]));
        mbar_wait(smem_u32(&bar[warp]), 0);
    }
    __syncthreads();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(td, 512);
}

// ---------------- (2) CTA pair, cta_group::2 -------------------------------------------------------------------
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_kernel(int n, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 90 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t td = tbase;
    const uint32_t rank = cluster_ctarank();
    long long t0 = clock64();
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t idesc = umma_idesc_f16(1, 256, n);
        const uint32_t a0 = (smem_u32(smem) + 1023u) & ~1023u, b0 = a0 + 48 * 1024;
        uint64_t da[8], db[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            da[i] = umma_desc_kmajor_noswizzle(a0 + (i & 3) * 16, 2912, 160);
            db[i] = umma_desc_kmajor_noswizzle(b0, (uint32_t)(n / 2) * 16, 128);   // each CTA holds N/2 rows of B
        }
#pragma unroll 1
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) umma_f16_cg2((i & 1) ? td + 256 : td, da[i], db[i], idesc, 1);
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    if (threadIdx.x == 0) mbar_wait(smem_u32(&bar), 0);
    __syncthreads();
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(td), "r"(512) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(multi_issuer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 2048;
    for (int n : {16, 32, 64, 128}) for (int w : {1, 2, 4}) {
        if (w * n > 512) continue;
        long long h = 0;
        multi_issuer_kernel<<<148, 256, 100 * 1024>>>(n, w, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("cta_group::1 M 128 N %3d issuing warps %d : %7.1f cycles per MMA per SM (%7.1f per issuer; tensor ideal %d)%s\n", n, w,
               (double)h / (reps * w), (double)h / reps, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    for (int n : {32, 64, 128, 256}) {
        long long h = 0;
        pair_kernel<<<148, 128, 100 * 1024>>>(n, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("cta_group::2 M 256 N %3d one issuer per pair  : %7.1f cycles per MMA (= per 128 rows per SM; tensor ideal %d)%s\n", n,
               (double)h / reps, n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
