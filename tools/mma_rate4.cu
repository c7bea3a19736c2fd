// Microbenchmark (round 2): per-slot synchronisation cost around a stream of stacked bf16x3 pairs (N = 64 then N = 32, four
// accumulators), issued by an elected lane of a full warp, G MMAs per elected region:
//   W = 0 plain elect region | 1 + tcgen05.commit to an mbarrier after the group | 2 + an (already satisfied) mbarrier wait and
//   tcgen05.fence before the group | 3 like 2, with 12 more warps spinning on an mbarrier (the kernel's waiting roles)
//   4 like 2, and a second warp issues the same stream into other accumulators (dual issue)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I disconet_b200/csrc tools/mma_rate4.cu -o tools/mma_rate4
#include "common.cuh"
#include <cstdio>
void disco_set_error(const char*, ...) {}

template <int W, int G>
__global__ void __launch_bounds__(512) rate_kernel(int n, int groups, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[6];
    __shared__ uint32_t tbase;
    __shared__ uint32_t flag;
    for (int i = threadIdx.x; i < 90 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar[0]), 1); mbar_init(smem_u32(&bar[1]), 1); mbar_init(smem_u32(&bar[2]), 1); mbar_init(smem_u32(&bar[3]), 1); mbar_init(smem_u32(&bar[4]), 1); mbar_init(smem_u32(&bar[5]), 1);
        fence_mbar_init();
        flag = 0x7fffffffu;
        mbar_arrive(smem_u32(&bar[5]));      // phase 0 of bar[5] is complete from the start
    }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tbase), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t td = tbase;
    const int warp = threadIdx.x >> 5;
    const bool issuer = warp == 0 || (W == 4 && warp == 1);
    if (issuer) {
        const uint32_t id2 = umma_idesc_f16(1, 128, 2 * n), id1 = umma_idesc_f16(1, 128, n);
        const uint32_t a0 = ((smem_u32(smem) + 1023u) & ~1023u) + warp * 12288, b0 = a0 + 48 * 1024;
        uint64_t da[4], dl[4], db[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            da[i] = umma_desc_kmajor_noswizzle(a0 + i * 16, 2944, 160);
            dl[i] = umma_desc_kmajor_noswizzle(a0 + 5888 + i * 16, 2944, 160);
            db[i] = umma_desc_kmajor_noswizzle(b0 + i * 2048, (uint32_t)n * 32, 128);
        }
        const uint32_t tdw = td + (W == 4 ? warp * 256u : 0u);
        const uint32_t accs = (W == 4) ? 64u : 128u;
        const uint32_t cbar = smem_u32(&bar[warp]);       // commit target
        uint32_t par = 0;
        uint32_t par2[2] = {0u, 0u};
        long long t0 = clock64();
#pragma unroll 1
        for (int r = 0; r < groups; ++r) {
            if (W >= 2 && W <= 4) {
                if (r > 0) { mbar_wait(cbar, par); par ^= 1u; }   // previous group's commit: the pipe drains once per group
                tc_fence_after();
            }
            if (W == 5 || W == 7) mbar_wait(smem_u32(&bar[5]), 0);   // a phase that completed long ago (a weight slot that already landed)
            if (W == 5 || W == 6) tc_fence_after();
            if (W == 9 || W == 10) {      // poll a plain shared-memory word (written long ago) instead of an mbarrier
                uint32_t v;
                do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&flag)) : "memory"); } while (v < (uint32_t)r);
                if (W == 10) tc_fence_after();
            }
            if (W == 11) {     // mbarrier.test_wait (non-blocking form) on the long-complete barrier
                uint32_t ok;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(ok) : "r"(smem_u32(&bar[5])), "r"(0u) : "memory");
                } while (!ok);
            }
            if (W == 8) {      // wait on the commit of the group issued TWO groups ago (a ring with lookahead)
                if (r > 1) { mbar_wait(smem_u32(&bar[(r & 1)]), par2[r & 1]); par2[r & 1] ^= 1u; }
                tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
                for (int i = 0; i < G / 2; ++i) {
                    const uint32_t acc = tdw + (uint32_t)(i & 3) * accs;
                    umma_f16(acc, da[i & 3], db[i & 3], id2, 1);
                    umma_f16(acc, dl[i & 3], db[i & 3], id1, 1);
                }
                if (W >= 1 && W != 8) umma_commit(cbar);
                if (W == 8) umma_commit(smem_u32(&bar[r & 1]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&bar[2 + warp]));
        __syncwarp();
        mbar_wait(smem_u32(&bar[2 + warp]), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
    } else if (W == 3 && warp >= 4) {
        mbar_wait(smem_u32(&bar[2]), 0);     // 12 warps spin until the issuer is done
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(td, 512);
}

template <int W, int G>
void run(int n, long long* d, const char* what) {
    const int mmas = 8192;
    cudaFuncSetAttribute(rate_kernel<W, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    rate_kernel<W, G><<<148, 512, 100 * 1024>>>(n, mmas / G, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("n %2d  G %2d  %-70s: %6.1f cycles/MMA (per issuer) %s\n", n, G, what, (double)h / mmas, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

template <int G>
void all(int n, long long* d) {
    run<0, G>(n, d, "elect region per group");
    run<1, G>(n, d, "+ tcgen05.commit after the group");
    run<2, G>(n, d, "+ wait on the previous group's commit + fence before the group");
    run<3, G>(n, d, "+ 12 other warps spinning on an mbarrier");
    if (n <= 32) run<4, G>(n, d, "wait + commit, TWO issuing warps (own accumulators)");
    run<5, G>(n, d, "commit + wait on a LONG-complete mbarrier + fence before the group");
    run<6, G>(n, d, "commit + fence only");
    run<7, G>(n, d, "commit + long-complete wait only");
    run<8, G>(n, d, "commit + wait on the commit of the group before the previous one + fence");
    run<9, G>(n, d, "commit + poll a shared-memory WORD (ld.volatile.shared), no mbarrier");
    run<10, G>(n, d, "commit + poll a shared-memory word + tcgen05.fence");
    run<11, G>(n, d, "commit + mbarrier.test_wait (long-complete) only");
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    for (int n : {32, 64}) { all<8>(n, d); all<16>(n, d); all<24>(n, d); }
    return 0;
}
