// Microbenchmark: cycles per tcgen05.mma (M=128, kind::f16, bf16) for shared-memory operand layouts.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I disconet_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate
#include "common.cuh"
#include <cstdio>
void disco_set_error(const char*, ...) {}

struct Cfg { int m; int n; int layout; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int nacc; int reps; uint32_t a_step, b_step; int unrolled; };

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
    uint64_t d = umma_desc_kmajor_noswizzle(addr, lbo, sbo);
    d |= (uint64_t)layout << 61;
    return d;
}

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, long long* out) {
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
        const uint32_t idesc = umma_idesc_f16(1, c.m, c.n);
        const uint32_t a0 = (smem_u32(smem) + 1023u) & ~1023u, b0 = a0 + 48 * 1024;
        long long t0 = clock64();
        uint64_t da[8], db[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            da[i] = mkdesc(a0 + (i & 3) * c.a_step, c.a_lbo, c.a_sbo, c.layout);
            db[i] = mkdesc(b0 + (i & 3) * c.b_step, c.b_lbo, c.b_sbo, c.layout);
        }
        const uint32_t td1 = td + (c.nacc == 2 ? 256u : 0u);
        if (c.unrolled) {
#pragma unroll 1
            for (int r = 0; r < c.reps; r += 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) umma_f16((i & 1) ? td1 : td, da[i], db[i], idesc, 1);
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < c.reps; ++r) umma_f16((r & 1) ? td1 : td, da[r & 7], db[r & 7], idesc, 1);
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(td, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int reps = 2048;
    int ns[] = {32, 64, 128, 256};
    // round 2: M = 64 rows as well, and very small N -- separates "fixed ~47-cycle issue floor" from "shared-memory
    // operand port" (bytes read per MMA = (M + N) * 32 B at 128 B/clk)
    for (int grid : {148}) for (int nacc : {2}) for (int unr : {1}) for (int m : {128, 64}) for (int n : {16, 32, 64, 128, 256}) {
        struct { const char* name; Cfg c; } v[] = {
            // conv kernel's layout: A patch rows 160 B apart, planes 2912 B apart; B dense [chunk][n][8]
            {"noswz_conv ", {m, n, 0, 2912, 160, (uint32_t)n * 16, 128, nacc, reps, 16, 0, unr}},
            // canonical 128B swizzle, K-major, 64-element rows; k-step advances 32 B inside the atom
            {"swz128     ", {m, n, 2, 16, 1024, 16, 1024, nacc, reps, 32, 32, unr}},
        };
        for (auto& x : v) {
            long long h = 0;
            rate_kernel<<<grid, 128, 100 * 1024>>>(x.c, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("grid %3d nacc %d unr %d M %3d N %3d %s : %7.1f cycles/MMA (tensor ideal %d, smem-port %d)%s\n", grid, nacc, unr, m, n, x.name, (double)h / reps, n / 2, (m + n) / 4,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
