// Microbenchmark (round 2, fused sub-pixel mode): what does a bf16x3 "stacked pair" cost?  One thread issues, per pair,
// A_hi*[W_hi;W_lo] (N = 2n) and A_lo*W_hi (N = n) -- two instruction shapes alternating -- into 1, 2 or 4 accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I disconet_b200/csrc tools/mma_rate3.cu -o tools/mma_rate3
#include "common.cuh"
#include <cstdio>
void disco_set_error(const char*, ...) {}

// variant: 0 all N=2n, accumulators alternate | 1 all N=2n, ONE accumulator (dependent chain) | 2 pairs (2n, n), one accumulator
//          3 pairs, accumulator changes every pair (4 accumulators) | 4 like 3 but both MMAs of a pair use N=2n (no shape change)
//          5 pairs with shapes (2n, n) but the second MMA goes to ANOTHER accumulator (no same-accumulator back-to-back)
template <int V>
__global__ void __launch_bounds__(128) rate_kernel(int n, int reps, uint32_t a_sbo, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    for (int i = threadIdx.x; i < 90 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = tbase;
    if (threadIdx.x == 0) {
        const uint32_t id2 = umma_idesc_f16(1, 128, 2 * n), id1 = umma_idesc_f16(1, 128, n);
        const uint32_t a0 = (smem_u32(smem) + 1023u) & ~1023u, b0 = a0 + 48 * 1024;
        uint64_t da[8], dl[8], db[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            da[i] = umma_desc_kmajor_noswizzle(a0 + (i & 3) * 16, 2944, a_sbo);
            dl[i] = umma_desc_kmajor_noswizzle(a0 + 5888 + (i & 3) * 16, 2944, a_sbo);
            db[i] = umma_desc_kmajor_noswizzle(b0 + (i & 3) * 2048, (uint32_t)n * 32, 128);
        }
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < reps; r += 16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t acc4 = td + (uint32_t)(i & 3) * 128u;
                if (V == 0) { umma_f16(td + (0u) , da[i], db[i], id2, 1); umma_f16(td + 128u, dl[i], db[i], id2, 1); }
                if (V == 1) { umma_f16(td, da[i], db[i], id2, 1); umma_f16(td, dl[i], db[i], id2, 1); }
                if (V == 2) { umma_f16(td, da[i], db[i], id2, 1); umma_f16(td, dl[i], db[i], id1, 1); }
                if (V == 3) { umma_f16(acc4, da[i], db[i], id2, 1); umma_f16(acc4, dl[i], db[i], id1, 1); }
                if (V == 4) { umma_f16(acc4, da[i], db[i], id2, 1); umma_f16(acc4, dl[i], db[i], id2, 1); }
                if (V == 5) { umma_f16(acc4, da[i], db[i], id2, 1); umma_f16(td + (uint32_t)((i + 2) & 3) * 128u, dl[i], db[i], id1, 1); }
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(td, 512);
}

template <int V>
void run(int n, uint32_t sbo, long long* d, const char* what) {
    const int reps = 4096;
    cudaFuncSetAttribute(rate_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    rate_kernel<V><<<148, 128, 100 * 1024>>>(n, reps, sbo, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("n %2d sbo %3u  %-78s: %6.1f cycles/MMA %s\n", n, sbo, what, (double)h / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    for (int n : {32, 64}) for (uint32_t sbo : {160u, 288u}) {
        run<0>(n, sbo, d, "all N=2n, two accumulators alternating");
        run<1>(n, sbo, d, "all N=2n, ONE accumulator (dependent chain)");
        run<2>(n, sbo, d, "pairs (N=2n, N=n), one accumulator");
        run<3>(n, sbo, d, "pairs (N=2n, N=n), accumulator changes per pair (4)");
        run<4>(n, sbo, d, "pairs (N=2n, N=2n), accumulator changes per pair (4)");
        run<5>(n, sbo, d, "pairs (N=2n, N=n), second MMA of the pair to another accumulator");
    }
    return 0;
}
