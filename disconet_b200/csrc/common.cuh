// Shared device/host helpers for the DiscoNet B200 hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "disconet_b200 kernels are written for sm_100a only (-gencode arch=compute_100a,code=sm_100a)"
#endif

// ---------------------------------------------------------------------------------------------
// Error plumbing (C-ABI never throws; negative return code + thread-local text)
// ---------------------------------------------------------------------------------------------
#define DISCO_OK 0
#define DISCO_EINVAL (-1)
#define DISCO_ECUDA (-2)
#define DISCO_EARCH (-3)
#define DISCO_EKERNEL (-4)

void disco_set_error(const char* fmt, ...);

#define DISCO_CHECK_CUDA(expr)                                                         \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            disco_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
                            cudaGetErrorString(_e));                                   \
            return DISCO_ECUDA;                                                        \
        }                                                                              \
    } while (0)

#define DISCO_REQUIRE(cond, ...)                                                       \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            disco_set_error(__VA_ARGS__);                                              \
            return DISCO_EINVAL;                                                       \
        }                                                                              \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Split-bf16 ("bf16x3") helpers: v ~= hi + lo with hi = bf16(v), lo = bf16(v - hi)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float v) {
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
    hi = f32_to_bf16_bits(v);
    lo = f32_to_bf16_bits(v - bf16_bits_to_f32(hi));
}
__device__ __forceinline__ uint16_t f32_to_f16_bits(float v) { return __half_as_ushort(__float2half_rn(v)); }
__device__ __forceinline__ float f16_bits_to_f32(uint16_t b) { return __half2float(__ushort_as_half(b)); }

// element decode for activations: dtype 0 = fp16, 1 = bf16
__device__ __forceinline__ float act_to_f32(uint16_t b, int dtype) {
    return dtype ? bf16_bits_to_f32(b) : f16_bits_to_f32(b);
}
__device__ __forceinline__ uint16_t f32_to_act(float v, int dtype) {
    return dtype ? f32_to_bf16_bits(v) : f32_to_f16_bits(v);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier / cp.async / bulk copy / tcgen05)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a trapped kernel (sticky CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        if (mbar_try_wait(bar, parity)) return;
    }
    printf("disco_b200: mbarrier timeout (block %d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x, blockIdx.y,
           threadIdx.x, bar, parity);
    __trap();
}

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 16-byte async copy global->shared; src_bytes == 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Bulk (TMA engine, 1-D) global->shared copy completing on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// one lane of a converged warp (warp-uniform code around it stays on the uniform datapath)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, %1;\n\t"
        "@px mov.s32 %0, 1;\n\t}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers fp16 and bf16 inputs with fp32 accumulation
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with the two 64-bit shared-memory descriptors passed as (lo, hi) 32-bit halves: the hi halves
// (SBO, version) are loop constants and the lo halves advance by plain 32-bit adds in 16-byte units,
// which keeps the single-thread issue loop short.
__device__ __forceinline__ void umma_f16_parts(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit all prior async tcgen05 ops of this thread; arrives (count 1) on the mbarrier when they retire
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved" 8x16B core matrices):
// element (row r, k) lives at start + (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2   [bytes]
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswizzle(uint32_t saddr, uint32_t lbo_bytes,
                                                               uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version for sm_100
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// Instruction descriptor for kind::f16: fp32 accumulate, A/B K-major, M x N tile
__device__ __forceinline__ uint32_t umma_idesc_f16(int bf16_inputs, int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                               // D format: F32
    d |= (bf16_inputs ? 1u : 0u) << 7;          // A format
    d |= (bf16_inputs ? 1u : 0u) << 10;         // B format
    d |= (uint32_t)(N >> 3) << 17;              // N / 8
    d |= (uint32_t)(M >> 4) << 24;              // M / 16
    return d;
}
