// Weight gradient of the 3x3 / 1x1 convolutions on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators).
//
//   dW[co][tap][ci] = sum over output pixels p of  dz[p][co] * x[p (+) tap][ci]
//
// (autograd backward of every F.conv2d / conv3d of Backbone.encode/decode, the heads and PWF conv1_1 --
// CoDetModule.py:289-291 `loss.backward()`.)  The contraction index is the PIXEL, and both operands are
// pixel-major NHWC tensors, i.e. "MN-major" UMMA operands: the 16-byte rows [pixel][8 channels] that the
// forward kernel stages as K-major core matrices are, read with the a_major/b_major bits set, exactly the
// MN-major SWIZZLE_NONE canonical layout ((8 ch),(8 px, k)) : SBO = channel-group pitch, LBO = 8-pixel pitch.
// So the staging is the forward kernel's patch gather (one 18x10 / 33x17 / 128-pixel patch per channel block,
// the nine taps being nine start addresses into it, nearest-upsample / concat / zero padding as address
// arithmetic) plus a plain 16x8 tile of dz.
//
// Work item = (128-row block of c_out) x (NB-channel block of c_in, all taps) x (a contiguous range of pixel
// tiles, split-K).  The taps x NB fp32 accumulators (<= 432 of the 512 TMEM columns) stay resident over the
// whole pixel range; each CTA writes one partial [c_out block][tap][NB] which a second kernel reduces over the
// splits into PyTorch's OIHW parameter layout.
//   warps 0-3: producers (cp.async gather of the dz tile and the x patch, hi + lo planes), then the epilogue
//   warp  4  : TMEM allocation + MMA issue (3 passes: hi*hi + lo*hi + hi*lo, fp32 accumulate)
#include <stdlib.h>
#include "common.cuh"
#include "conv.h"
#include "train.h"

namespace {

constexpr int kProdThreads = 128;
constexpr int kIssuers = 2;     // MMA-issuing warps; each owns a disjoint set of taps (= of TMEM accumulators)
constexpr int kThreads = kProdThreads + 32 * kIssuers;
constexpr int kMaxStages = 4;
constexpr int kCtlBytes = 128;

struct WGeom {
    disco_wgrad_desc d;
    int mode;                 // 0: 3x3 s1, 1: 3x3 s2, 2: 1x1
    int c_in;                 // padded input channels
    int cob, mblks;           // c_out rows per item (<= 128), blocks
    int nb, nblks;            // c_in channels per item, blocks
    int chunks_a, chunks_b;   // cob/8, nb/8
    int plane_a, plane_b;     // bytes between 8-channel groups
    int parplane;             // stride-2: bytes between the even/odd column sub-planes
    int lbo_b;                // bytes between consecutive 8-pixel groups of the B patch (= tile row pitch)
    int PIX;                  // staged patch pixels
    int a_part, b_part, stage_bytes, stages;
    int tiles_h, tiles_w, n_tiles;
    long long total_pix;
    int splits;
    int tmem_cols;
    int smem_bytes;
    int sw;                   // 1: operand-swapped kernel (A = input patch, B = dz) for the skinny 3x3 stride-1 layers
    int kws;                  // sw: kw-shifted patch copies stacked along M (3 when 3*c_in <= 128, else 1)
};

struct __align__(8) WCtl {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad;
};
static_assert(sizeof(WCtl) <= kCtlBytes, "control block");

__device__ __forceinline__ uint32_t wgrad_idesc(int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                   // D: fp32
    d |= 1u << 7;                   // A: bf16
    d |= 1u << 10;                  // B: bf16
    d |= 1u << 15;                  // A is MN-major
    d |= 1u << 16;                  // B is MN-major
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) wgrad_tc_kernel(const WGeom g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    WCtl* ctl = reinterpret_cast<WCtl*>(smem_raw);
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t stage_base = smem_base + kCtlBytes;
    const disco_wgrad_desc& d = g.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int TAPS = (MODE == 2) ? 1 : 9;
    constexpr int STRIDE = (MODE == 1) ? 2 : 1;
    constexpr int PW = (MODE == 0) ? 10 : (MODE == 1) ? 17 : 128;

    // item decode: blockIdx.x = (split * mblks + mblk) * nblks + nblk
    const int nblk = blockIdx.x % g.nblks;
    const int mblk = (blockIdx.x / g.nblks) % g.mblks;
    const int split = blockIdx.x / (g.nblks * g.mblks);
    const int t0 = (int)((long long)split * g.n_tiles / g.splits);
    const int t1 = (int)((long long)(split + 1) * g.n_tiles / g.splits);
    const int co0 = mblk * g.cob, ci0 = nblk * g.nb;

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), kProdThreads);
            mbar_init(smem_u32(&ctl->empty[s]), TAPS > 1 ? kIssuers : 1);
        }
        mbar_init(smem_u32(&ctl->acc_full), TAPS > 1 ? kIssuers : 1);
        fence_mbar_init();
    }
    if (warp == 4) {
        tmem_alloc(smem_u32(&ctl->tmem_base), (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = ctl->tmem_base;

    if (warp < 4) {
        // =========================== producers ====================================================
        const uint16_t* dz = reinterpret_cast<const uint16_t*>(d.dz_hi);
        for (int t = t0; t < t1; ++t) {
            const int i = t - t0;
            const int slot = i % g.stages;
            mbar_wait(smem_u32(&ctl->empty[slot]), ((uint32_t)(i / g.stages) & 1u) ^ 1u);
            const uint32_t sa = stage_base + slot * g.stage_bytes;
            const uint32_t sb = sa + 2 * g.a_part;
            int img = 0, h0 = 0, w0 = 0;
            long long p0 = 0;
            if (MODE != 2) {
                const int per_img = g.tiles_h * g.tiles_w;
                img = t / per_img;
                const int rem = t - img * per_img;
                const int th = rem / g.tiles_w;
                h0 = th * 16;
                w0 = (rem - th * g.tiles_w) * 8;
            } else {
                p0 = (long long)t * 128;
            }
            // ---- A: dz tile [cob/8][128 px][8 ch], hi and lo planes --------------------------------
            for (int e = tid; e < 128 * g.chunks_a; e += kProdThreads) {
                const int m = e / g.chunks_a, chunk = e - m * g.chunks_a;
                long long pix;
                bool valid;
                if (MODE != 2) {
                    const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
                    valid = (oh < d.h_out) && (ow < d.w_out);
                    pix = ((long long)img * d.h_out + oh) * d.w_out + ow;
                } else {
                    pix = p0 + m;
                    valid = pix < g.total_pix;
                }
                const uint16_t* gp = valid ? dz + pix * d.c_out + co0 + chunk * 8 : dz;
                const uint32_t dst = sa + chunk * g.plane_a + m * 16;
                cp_async16(dst, gp, valid ? 16u : 0u);
                cp_async16(dst + g.a_part, valid ? gp + d.dz_lo_off : dz, valid ? 16u : 0u);
            }
            // ---- B: input patch [nb/8][PIX][8 ch], hi and lo planes ------------------------------------
            const int hi0 = h0 * STRIDE - 1, wi0 = w0 * STRIDE - 1;
            for (int e = tid; e < g.PIX * g.chunks_b; e += kProdThreads) {
                const int pi = e / g.chunks_b, chunk = e - pi * g.chunks_b;
                const int ch = ci0 + chunk * 8;
                const int sidx = (ch >= d.src_c[0]) ? 1 : 0;
                const uint16_t* src = reinterpret_cast<const uint16_t*>(d.src[sidx]);
                const int Cs = d.src_c[sidx];
                const int cl = sidx ? ch - d.src_c[0] : ch;
                const long long lo_off = d.src_lo_off[sidx];
                const int upm = d.src_up[sidx];
                const int up = upm ? 1 : 0;
                bool valid;
                const uint16_t* gp;
                uint32_t dst;
                if (MODE != 2) {
                    const int r = pi / PW, c = pi - r * PW;
                    const int hi = hi0 + r, wi = wi0 + c;
                    valid = (hi >= 0) && (hi < d.h_in) && (wi >= 0) && (wi < d.w_in) && !(upm == 2 && ((hi | wi) & 1));
                    const int Hs = d.h_in >> up, Ws = d.w_in >> up;
                    gp = src + (((long long)img * Hs + (hi >> up)) * Ws + (wi >> up)) * Cs + cl;
                    dst = sb + chunk * g.plane_b +
                          ((MODE == 0) ? (uint32_t)pi * 16u
                                       : (uint32_t)(c & 1) * g.parplane + (uint32_t)r * 144u + (uint32_t)(c >> 1) * 16u);
                } else {
                    const long long p = p0 + pi;
                    valid = p < g.total_pix;
                    gp = src + p * Cs + cl;
                    dst = sb + chunk * g.plane_b + (uint32_t)pi * 16u;
                }
                if (!valid) gp = src;
                cp_async16(dst, gp, valid ? 16u : 0u);
                cp_async16(dst + g.b_part, valid ? gp + lo_off : src, valid ? 16u : 0u);
            }
            cp_async_commit();
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(smem_u32(&ctl->full[slot]));
        }
        // =========================== epilogue: TMEM -> partial[split][co][tap][ci] ================
        mbar_wait(smem_u32(&ctl->acc_full), 0);
        tc_fence_after();
        const int co = co0 + warp * 32 + lane;           // TMEM lane == row of D == output channel
        const bool row_ok = (warp * 32 + lane < g.cob) && (co < d.c_out);
        float* prow = d.partial + ((long long)split * d.c_out + co) * TAPS * g.c_in + ci0;
        const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int tap = 0; tap < TAPS; ++tap) {
            for (int j = 0; j < g.nb; j += 16) {
                uint32_t v[16];
                tmem_ld16(t_lane + (uint32_t)(tap * g.nb + j), v);
                tmem_ld_wait();
                if (row_ok) {
                    float4* dst = reinterpret_cast<float4*>(prow + (long long)tap * g.c_in + j);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        dst[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                             __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                }
            }
        }
        tc_fence_before();
    } else {
        // =========================== MMA issuers ==================================================
        // One issuing thread sustains only ~80 cycles per small tcgen05.mma here (measured on the forward kernel,
        // profiles/r01_mma_issue_rate.txt); two warps issuing to DISJOINT accumulators (taps 0-4 / 5-8) reach the
        // shared-memory operand-port pace.  1x1 layers have a single accumulator: one issuer.
        const int iss = warp - 4;
        if (TAPS == 1 && iss > 0) goto done;
        const int tap_lo = (TAPS == 1) ? 0 : (iss == 0 ? 0 : 5);
        const int tap_hi = (TAPS == 1) ? 1 : (iss == 0 ? 5 : TAPS);
        const uint32_t idesc = wgrad_idesc(128, g.nb);
        // descriptor halves: lo = start>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version(1)<<14
        // MN-major SWIZZLE_NONE: LBO = pitch between 8-pixel (K) groups, SBO = pitch between 8-channel (M/N) groups
        const uint32_t a_lbo = 128u >> 4, a_sbo = (uint32_t)g.plane_a >> 4;
        const uint32_t b_lbo = (uint32_t)g.lbo_b >> 4, b_sbo = (uint32_t)g.plane_b >> 4;
        const uint32_t a_hi = a_sbo | (1u << 14), b_hi = b_sbo | (1u << 14);
        const int passes = d.passes;
        for (int t = t0; t < t1; ++t) {
            const int i = t - t0;
            const int slot = i % g.stages;
            mbar_wait(smem_u32(&ctl->full[slot]), (uint32_t)(i / g.stages) & 1u);
            tc_fence_after();
            const uint32_t sa = stage_base + slot * g.stage_bytes;
            const uint32_t sb = sa + 2 * g.a_part;
            if (elect_one()) {
#pragma unroll 1
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t a_lo0 = ((sa + (uint32_t)ks * 256u) >> 4) | (a_lbo << 16);
                    const uint32_t b_ks = sb + (uint32_t)(2 * ks) * (uint32_t)g.lbo_b;
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        if (tap < tap_lo || tap >= tap_hi) continue;
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t toff = (MODE == 0) ? (uint32_t)(kh * 10 + kw) * 16u
                                            : (MODE == 1) ? (uint32_t)(kw & 1) * (uint32_t)g.parplane + (uint32_t)(kh * 9 + (kw >> 1)) * 16u
                                                          : 0u;
                        const uint32_t b_lo0 = ((b_ks + toff) >> 4) | (b_lbo << 16);
                        const uint32_t td = tmem_d + (uint32_t)(tap * g.nb);
                        const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                        umma_f16_parts(td, a_lo0, a_hi, b_lo0, b_hi, idesc, acc);                                   // hi * hi
                        if (passes == 3) {
                            umma_f16_parts(td, a_lo0 + ((uint32_t)g.a_part >> 4), a_hi, b_lo0, b_hi, idesc, 1u);    // lo * hi
                            umma_f16_parts(td, a_lo0, a_hi, b_lo0 + ((uint32_t)g.b_part >> 4), b_hi, idesc, 1u);    // hi * lo
                        }
                    }
                }
                umma_commit(smem_u32(&ctl->empty[slot]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&ctl->acc_full));
        __syncwarp();
    }
done:
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_d, (uint32_t)g.tmem_cols);
}

// Operand-swapped variant for the skinny 3x3 stride-1 layers (c_out <= 64 at the 256^2 / 128^2 resolutions, where the
// kernel above is MMA-issue-bound: ~45 cycles per tcgen05.mma whatever its N, 216 of them per 128-pixel tile, with
// M = c_out padded to 128):
//   A = the input patch (M = channels of c_in), B = the dz tile (N = c_out), D[ci][co].
// When 3*c_in <= 128 the patch is staged THREE times, shifted by kw = 0,1,2 pixels, as consecutive channel-group planes,
// so one MMA covers a whole filter row: M = (kw, ci), one accumulator per kh, 72 MMAs per tile instead of 216.
// Otherwise (conv8_1: c_in = 96 > c_out = 32) plain swap: M = c_in, nine accumulators, 216 MMAs instead of 432.
__global__ void __launch_bounds__(kThreads, 1) wgrad_sw_kernel(const WGeom g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    WCtl* ctl = reinterpret_cast<WCtl*>(smem_raw);
    const uint32_t stage_base = smem_u32(smem_raw) + kCtlBytes;
    const disco_wgrad_desc& d = g.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = blockIdx.x;
    const int t0 = (int)((long long)split * g.n_tiles / g.splits);
    const int t1 = (int)((long long)(split + 1) * g.n_tiles / g.splits);
    const int KWS = g.kws;
    const int NACC = (KWS == 3) ? 3 : 9;          // accumulators: one per kh (stacked) or per tap

    if (tid == 0) {
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(smem_u32(&ctl->full[s]), kProdThreads);
            mbar_init(smem_u32(&ctl->empty[s]), kIssuers);
        }
        mbar_init(smem_u32(&ctl->acc_full), kIssuers);
        fence_mbar_init();
    }
    if (warp == 4) {
        tmem_alloc(smem_u32(&ctl->tmem_base), (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = ctl->tmem_base;
    // stage layout: [A hi | A lo | B hi | B lo];  A part = kws * chunks_b planes of plane_b, B part = chunks_a planes of plane_a
    const int a_part = KWS * g.chunks_b * g.plane_b, b_part = g.chunks_a * g.plane_a;

    if (warp < 4) {
        const uint16_t* dz = reinterpret_cast<const uint16_t*>(d.dz_hi);
        for (int t = t0; t < t1; ++t) {
            const int i = t - t0, slot = i % g.stages;
            mbar_wait(smem_u32(&ctl->empty[slot]), ((uint32_t)(i / g.stages) & 1u) ^ 1u);
            const uint32_t sa = stage_base + slot * g.stage_bytes;
            const uint32_t sb = sa + 2 * a_part;
            const int per_img = g.tiles_h * g.tiles_w;
            const int img = t / per_img;
            const int rem = t - img * per_img;
            const int th = rem / g.tiles_w;
            const int h0 = th * 16, w0 = (rem - th * g.tiles_w) * 8;
            // ---- B: dz tile [c_out/8][128 px][8 ch] ----
            for (int e = tid; e < 128 * g.chunks_a; e += kProdThreads) {
                const int m = e / g.chunks_a, chunk = e - m * g.chunks_a;
                const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
                const bool valid = (oh < d.h_out) && (ow < d.w_out);
                const long long pix = ((long long)img * d.h_out + oh) * d.w_out + ow;
                const uint16_t* gp = valid ? dz + pix * d.c_out + chunk * 8 : dz;
                const uint32_t dst = sb + chunk * g.plane_a + m * 16;
                cp_async16(dst, gp, valid ? 16u : 0u);
                cp_async16(dst + b_part, valid ? gp + d.dz_lo_off : dz, valid ? 16u : 0u);
            }
            // ---- A: input patch 18x10, kws shifted copies ----
            const int hi0 = h0 - 1, wi0 = w0 - 1;
            for (int e = tid; e < 180 * g.chunks_b; e += kProdThreads) {
                const int pi = e / g.chunks_b, chunk = e - pi * g.chunks_b;
                const int ch = chunk * 8;
                const int sidx = (ch >= d.src_c[0]) ? 1 : 0;
                const uint16_t* src = reinterpret_cast<const uint16_t*>(d.src[sidx]);
                const int Cs = d.src_c[sidx];
                const int cl = sidx ? ch - d.src_c[0] : ch;
                const long long lo_off = d.src_lo_off[sidx];
                const int upm = d.src_up[sidx], up = upm ? 1 : 0;
                const int r = pi / 10, c = pi - r * 10;
                const int hi = hi0 + r, wi = wi0 + c;
                const bool valid = (hi >= 0) && (hi < d.h_in) && (wi >= 0) && (wi < d.w_in) && !(upm == 2 && ((hi | wi) & 1));
                const int Hs = d.h_in >> up, Ws = d.w_in >> up;
                const uint16_t* gp = valid ? src + (((long long)img * Hs + (hi >> up)) * Ws + (wi >> up)) * Cs + cl : src;
                for (int kw = 0; kw < KWS; ++kw) {
                    if (pi < kw) continue;     // copy kw holds patch pixel (slot + kw)
                    const uint32_t dst = sa + (uint32_t)(kw * g.chunks_b + chunk) * g.plane_b + (uint32_t)(pi - kw) * 16u;
                    cp_async16(dst, gp, valid ? 16u : 0u);
                    cp_async16(dst + a_part, valid ? gp + lo_off : src, valid ? 16u : 0u);
                }
            }
            cp_async_commit();
            cp_async_wait<0>();
            fence_proxy_async_smem();
            mbar_arrive(smem_u32(&ctl->full[slot]));
        }
        // ---- epilogue: D[acc][m = (kw, ci)][co] -> partial[split][co][tap][ci] ----
        mbar_wait(smem_u32(&ctl->acc_full), 0);
        tc_fence_after();
        const int m = warp * 32 + lane;
        const int kw_m = (KWS == 3) ? m / g.c_in : 0;
        const int ci = (KWS == 3) ? m - kw_m * g.c_in : m;
        const bool row_ok = (KWS == 3) ? (m < 3 * g.c_in) : (m < g.c_in);
        const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
        float* pbase = d.partial + (long long)split * d.c_out * 9 * g.c_in;
        for (int a = 0; a < NACC; ++a) {
            const int tap = (KWS == 3) ? a * 3 + kw_m : a;
            for (int j = 0; j < d.c_out; j += 16) {
                uint32_t v[16];
                tmem_ld16(t_lane + (uint32_t)(a * d.c_out + j), v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        pbase[((long long)(j + q) * 9 + tap) * g.c_in + ci] = __uint_as_float(v[q]);
                }
            }
        }
        tc_fence_before();
    } else {
        // ---- MMA issuers: accumulators split between the two warps ----
        const int iss = warp - 4;
        const int a_lo_acc = (iss == 0) ? 0 : (NACC + 1) / 2, a_hi_acc = (iss == 0) ? (NACC + 1) / 2 : NACC;
        const uint32_t idesc = wgrad_idesc(128, d.c_out);
        const uint32_t a_lbo = 160u >> 4, a_sbo = (uint32_t)g.plane_b >> 4;     // A = patch: next tile row / next channel group
        const uint32_t b_lbo = 128u >> 4, b_sbo = (uint32_t)g.plane_a >> 4;     // B = dz tile
        const uint32_t a_hi = a_sbo | (1u << 14), b_hi = b_sbo | (1u << 14);
        for (int t = t0; t < t1; ++t) {
            const int i = t - t0, slot = i % g.stages;
            mbar_wait(smem_u32(&ctl->full[slot]), (uint32_t)(i / g.stages) & 1u);
            tc_fence_after();
            const uint32_t sa = stage_base + slot * g.stage_bytes;
            const uint32_t sb = sa + 2 * a_part;
            if (elect_one()) {
#pragma unroll 1
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t b_lo0 = ((sb + (uint32_t)ks * 256u) >> 4) | (b_lbo << 16);
                    const uint32_t a_ks = sa + (uint32_t)(2 * ks) * 160u;
#pragma unroll
                    for (int a = 0; a < 9; ++a) {            // straight-line issue: a rolled MMA loop serialises
                        if (a < a_lo_acc || a >= a_hi_acc) continue;
                        const int kh = (KWS == 3) ? a : a / 3, kw = (KWS == 3) ? 0 : a - (a / 3) * 3;
                        const uint32_t a_lo0 = ((a_ks + (uint32_t)(kh * 10 + kw) * 16u) >> 4) | (a_lbo << 16);
                        const uint32_t td = tmem_d + (uint32_t)(a * d.c_out);
                        const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                        umma_f16_parts(td, a_lo0, a_hi, b_lo0, b_hi, idesc, acc);                                  // hi * hi
                        if (d.passes == 3) {
                            umma_f16_parts(td, a_lo0 + ((uint32_t)a_part >> 4), a_hi, b_lo0, b_hi, idesc, 1u);     // lo * hi
                            umma_f16_parts(td, a_lo0, a_hi, b_lo0 + ((uint32_t)b_part >> 4), b_hi, idesc, 1u);     // hi * lo
                        }
                    }
                }
                umma_commit(smem_u32(&ctl->empty[slot]));
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&ctl->acc_full));
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_d, (uint32_t)g.tmem_cols);
}

// dw[co][ci][tap] = sum_s partial[s][co][tap][ci]   (PyTorch OIHW order; drops padded input channels)
__global__ void wgrad_reduce_kernel(const float* partial, int splits, int c_out, int taps, int c_in, int c_in_real,
                                    float* dw) {
    const long long total = (long long)c_out * c_in_real * taps;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int tap = (int)(idx % taps);
    const int ci = (int)((idx / taps) % c_in_real);
    const int co = (int)(idx / ((long long)taps * c_in_real));
    const long long per = (long long)c_out * taps * c_in;
    const float* p = partial + ((long long)co * taps + tap) * c_in + ci;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += p[s * per];
    dw[idx] = acc;
}

// CUDA-core validator: one thread per (co, tap, ci), straightforward loop over all pixels (tests only)
__device__ __forceinline__ float act_val(const uint16_t* p, long long lo_off) {
    return bf16_bits_to_f32(p[0]) + bf16_bits_to_f32(p[lo_off]);
}

__global__ void wgrad_ref_kernel(const disco_wgrad_desc d, int c_in) {
    const long long total = (long long)d.c_out * d.taps * d.c_in_real;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int tap = (int)(idx % d.taps);
    const int ci = (int)((idx / d.taps) % d.c_in_real);
    const int co = (int)(idx / ((long long)d.taps * d.c_in_real));
    const int k = d.taps == 9 ? 3 : 1, pad = d.taps == 9 ? 1 : 0;
    const int kh = tap / k, kw = tap - kh * k;
    const int sidx = ci >= d.src_c[0] ? 1 : 0;
    const int cl = sidx ? ci - d.src_c[0] : ci;
    const int upm = d.src_up[sidx], up = upm ? 1 : 0;
    const int Hs = d.h_in >> up, Ws = d.w_in >> up, Cs = d.src_c[sidx];
    const uint16_t* src = reinterpret_cast<const uint16_t*>(d.src[sidx]);
    const uint16_t* dz = reinterpret_cast<const uint16_t*>(d.dz_hi);
    double acc = 0.0;
    for (int n = 0; n < d.n; ++n)
        for (int oh = 0; oh < d.h_out; ++oh) {
            const int hi = oh * d.stride - pad + kh;
            if (hi < 0 || hi >= d.h_in) continue;
            for (int ow = 0; ow < d.w_out; ++ow) {
                const int wi = ow * d.stride - pad + kw;
                if (wi < 0 || wi >= d.w_in) continue;
                if (upm == 2 && ((hi | wi) & 1)) continue;
                const float x = act_val(src + (((long long)n * Hs + (hi >> up)) * Ws + (wi >> up)) * Cs + cl, d.src_lo_off[sidx]);
                const float g = act_val(dz + (((long long)n * d.h_out + oh) * d.w_out + ow) * d.c_out + co, d.dz_lo_off);
                acc += (double)x * (double)g;
            }
        }
    (void)c_in;
    d.dw[idx] = (float)acc;
}

int g_sms = 0;

int build_wgeom(const disco_wgrad_desc* d, WGeom* g) {
    DISCO_REQUIRE(d->taps == 9 || d->taps == 1, "wgrad: taps must be 9 or 1");
    DISCO_REQUIRE(d->stride == 1 || (d->stride == 2 && d->taps == 9), "wgrad: stride %d unsupported", d->stride);
    DISCO_REQUIRE(d->src[0] && d->dz_hi && d->dw, "wgrad: null tensor");
    DISCO_REQUIRE(d->src_c[0] > 0 && d->src_c[0] % 16 == 0 && d->src_c[1] % 16 == 0, "wgrad: source channels must be multiples of 16");
    DISCO_REQUIRE(d->c_out > 0 && d->c_out % 16 == 0 && (d->c_out <= 128 || d->c_out % 128 == 0), "wgrad: c_out %d unsupported", d->c_out);
    DISCO_REQUIRE(d->passes == 1 || d->passes == 3, "wgrad: passes must be 1 or 3");
    DISCO_REQUIRE(d->n > 0 && d->h_out > 0 && d->w_out > 0, "wgrad: empty input");
    if (g_sms == 0) {
        int dev = 0;
        DISCO_CHECK_CUDA(cudaGetDevice(&dev));
        DISCO_CHECK_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    g->d = *d;
    g->mode = d->taps == 1 ? 2 : (d->stride == 2 ? 1 : 0);
    g->c_in = d->src_c[0] + d->src_c[1];
    DISCO_REQUIRE(d->c_in_real > 0 && d->c_in_real <= g->c_in, "wgrad: c_in_real out of range");
    g->cob = d->c_out < 128 ? d->c_out : 128;
    g->mblks = (d->c_out + 127) / 128;
    // N block: taps * nb fp32 accumulator columns must fit the 512 TMEM columns; stride-2 patches are 3x larger
    int nb;
    if (g->mode == 2) nb = g->c_in <= 256 ? g->c_in : 256;
    else if (g->mode == 1) nb = 16;
    else nb = (g->c_in % 48 == 0) ? 48 : (g->c_in % 32 == 0) ? 32 : 16;
    DISCO_REQUIRE(g->c_in % nb == 0, "wgrad: c_in %d not a multiple of the channel block %d", g->c_in, nb);
    g->nb = nb;
    g->nblks = g->c_in / nb;
    g->chunks_a = g->cob / 8;
    g->chunks_b = nb / 8;
    g->plane_a = 128 * 16 + 16;
    if (g->mode == 0) { g->PIX = 180; g->parplane = 0; g->plane_b = 180 * 16 + 16; g->lbo_b = 160; }
    else if (g->mode == 1) { g->PIX = 33 * 17; g->parplane = 33 * 9 * 16; g->plane_b = 2 * g->parplane + 16; g->lbo_b = 288; }
    else { g->PIX = 128; g->parplane = 0; g->plane_b = 128 * 16 + 16; g->lbo_b = 128; }
    g->a_part = g->chunks_a * g->plane_a;
    g->b_part = g->chunks_b * g->plane_b;
    g->stage_bytes = ((2 * g->a_part + 2 * g->b_part + 127) / 128) * 128;
    // An M = 128 MMA reads 16 channel groups of A whatever cob is (rows >= cob produce accumulator rows that are
    // never read): that window, from the lo part of the last stage, has to stay inside the allocation.
    const int a_window = g->a_part + 16 * g->plane_a + 128;
    int stages = kMaxStages, smem = 0;
    for (; stages >= 1; --stages) {
        smem = kCtlBytes + stages * g->stage_bytes;
        const int need = kCtlBytes + (stages - 1) * g->stage_bytes + a_window;
        if (need > smem) smem = need;
        if (smem <= 227 * 1024) break;
    }
    DISCO_REQUIRE(stages >= 1, "wgrad: stage of %d bytes does not fit shared memory", g->stage_bytes);
    g->stages = stages;
    g->smem_bytes = smem;
    g->tiles_h = (d->h_out + 15) / 16;
    g->tiles_w = (d->w_out + 7) / 8;
    g->total_pix = (long long)d->n * d->h_out * d->w_out;
    const long long nt = (g->mode == 2) ? (g->total_pix + 127) / 128 : (long long)d->n * g->tiles_h * g->tiles_w;
    DISCO_REQUIRE(nt > 0 && nt < (1ll << 30), "wgrad: bad tile count");
    g->n_tiles = (int)nt;
    int cols = 32;
    while (cols < d->taps * nb) cols *= 2;
    DISCO_REQUIRE(cols <= 512, "wgrad: accumulators exceed TMEM");
    g->tmem_cols = cols;
    // operand-swapped kernel for the skinny 3x3 stride-1 layers (see wgrad_sw_kernel)
    g->sw = 0; g->kws = 1;
    {
        const char* e = getenv("DISCO_WGRAD_SW");
        const bool allow = !(e && e[0] == '0');
        const bool stack = 3 * g->c_in <= 128 && d->c_out <= 64;                      // M = (kw, ci), 3 accumulators
        const bool swap = g->c_in <= 128 && d->c_out <= 48 && g->c_in > d->c_out;     // M = ci, 9 accumulators
        if (allow && g->mode == 0 && (stack || swap)) {
            const int kws = stack ? 3 : 1;
            const int chunks_b = g->c_in / 8;           // the whole c_in in one item
            const int a_part = kws * chunks_b * g->plane_b, b_part = g->chunks_a * g->plane_a;
            const int stage_bytes = ((2 * a_part + 2 * b_part + 127) / 128) * 128;
            const int win = a_part + 16 * g->plane_b + 8 * 320 + 128;   // M = 128 reads 16 channel groups of the patch planes
            int st = kMaxStages, sm = 0;
            for (; st >= 1; --st) {
                sm = kCtlBytes + st * stage_bytes;
                const int need = kCtlBytes + (st - 1) * stage_bytes + win;
                if (need > sm) sm = need;
                if (sm <= 227 * 1024) break;
            }
            int c2 = 32;
            while (c2 < (kws == 3 ? 3 : 9) * d->c_out) c2 *= 2;
            if (st >= 1 && c2 <= 512) {
                g->sw = 1; g->kws = kws;
                g->chunks_b = chunks_b; g->nb = g->c_in; g->nblks = 1; g->mblks = 1;
                g->a_part = a_part; g->b_part = b_part; g->stage_bytes = stage_bytes;
                g->stages = st; g->smem_bytes = sm; g->tmem_cols = c2;
            }
        }
    }
    // split-K: aim at ~2 waves of CTAs, at least 2 pixel tiles per CTA
    const int items = g->mblks * g->nblks;
    int splits = (2 * g_sms + items - 1) / items;
    if (splits > g->n_tiles / 2) splits = g->n_tiles / 2;
    if (splits > 2 * g_sms) splits = 2 * g_sms;
    if (splits < 1) splits = 1;
    g->splits = splits;
    return DISCO_OK;
}

template <int MODE>
int launch_wgrad(const WGeom& g, cudaStream_t s) {
    static bool attr_set[64] = {false};
    int dev = 0;
    DISCO_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        DISCO_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    wgrad_tc_kernel<MODE><<<g.splits * g.mblks * g.nblks, kThreads, g.smem_bytes, s>>>(g);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

}  // namespace

int disco_wgrad_splits(const disco_wgrad_desc* d) {
    WGeom g;
    int rc = build_wgeom(d, &g);
    return rc < 0 ? rc : g.splits;
}

int disco_wgrad_tc_launch(const disco_wgrad_desc* d, void* stream) {
    WGeom g;
    int rc = build_wgeom(d, &g);
    if (rc < 0) return rc;
    DISCO_REQUIRE(d->partial, "wgrad: null partial workspace");
    DISCO_REQUIRE(d->splits >= g.splits, "wgrad: partial workspace sized for %d splits, need %d", d->splits, g.splits);
    cudaStream_t s = (cudaStream_t)stream;
    if (g.sw) {
        static bool attr_set[64] = {false};
        int dev = 0;
        DISCO_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev < 64 && !attr_set[dev]) {
            DISCO_CHECK_CUDA(cudaFuncSetAttribute(wgrad_sw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set[dev] = true;
        }
        wgrad_sw_kernel<<<g.splits, kThreads, g.smem_bytes, s>>>(g);
        DISCO_CHECK_CUDA(cudaGetLastError());
        rc = DISCO_OK;
    } else if (g.mode == 0) rc = launch_wgrad<0>(g, s);
    else if (g.mode == 1) rc = launch_wgrad<1>(g, s);
    else rc = launch_wgrad<2>(g, s);
    if (rc < 0) return rc;
    const long long total = (long long)d->c_out * d->c_in_real * d->taps;
    wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d->partial, g.splits, d->c_out, d->taps, g.c_in,
                                                                         d->c_in_real, d->dw);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_wgrad_ref_launch(const disco_wgrad_desc* d, void* stream) {
    DISCO_REQUIRE(d && d->src[0] && d->dz_hi && d->dw, "wgrad_ref: null tensor");
    const long long total = (long long)d->c_out * d->taps * d->c_in_real;
    wgrad_ref_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*d, d->src_c[0] + d->src_c[1]);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
