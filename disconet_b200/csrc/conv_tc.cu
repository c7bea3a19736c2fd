// Implicit-GEMM 3x3 / 1x1 convolution on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators).
//
// Replaces, for the DiscoNet hot path, every F.conv2d/conv3d + BatchNorm(eval) + ReLU of the reference
// (Backbone.py:102-136 encode, :173-237 decode incl. the nearest-x2 upsample + channel concat on the
// *load* side, DetModelBase.py:283-351 heads, DiscoNet.py:148 PWF conv1_1) -- BN is folded into the
// packed weights / bias on the host (disconet_b200/plan.py).
//
// One CTA computes a 128-pixel x block_n-channel output tile:
//   M = 128 output pixels = 16 rows x 8 cols of one image (or 128 consecutive pixels for 1x1)
//   N = block_n output channels (<= 256), K = taps * C_in, fp32 accumulation in TMEM.
//
// Data staging (this is the B200-specific part):
//   * A operand: the input patch needed by the tile (18x10 pixels for 3x3/s1, 33x17 for s2) is
//     gathered ONCE per channel block into shared memory in the UMMA "no-swizzle K-major" layout
//     [channel/8][pixel][8 ch] (16-byte core-matrix rows).  In that layout a filter tap (kh,kw) is
//     just a different 16-byte-aligned *start address* of the same patch, so the nine taps reuse one
//     staged patch: L2->SMEM traffic for activations is ~1.4x the input instead of 9x.
//     Stride-2 convs de-interleave even/odd columns into two sub-planes at gather time so the 8 rows
//     of a core matrix stay 16 B apart.  Nearest-upsample + concat are address arithmetic in the
//     gather (src_up / two sources), zero padding is cp.async zero-fill.
//   * B operand: host-packed weight images, one contiguous block per (channel block, tap), streamed
//     by the bulk-copy (TMA) engine (cp.async.bulk + mbarrier complete_tx).
//   * Warp roles: warps 0-3 gather A (cp.async) then run the epilogue (tcgen05.ld -> bias/ReLU ->
//     16-bit hi[/lo] or fp32 stores); warp 4 streams B; warp 5 allocates TMEM and issues the MMAs.
//
// precision DISCO_PREC_BF16X3 keeps activations/weights as bf16 hi+lo pairs and issues three MMAs
// per k-step (hi*hi + lo*hi + hi*lo) -> ~16 mantissa bits, which is what the <=1e-3 parity gate
// against the fp32 reference needs (plain fp16/bf16 operands measure 2.5e-3 / 2e-2).
#include "common.cuh"
#include "conv.h"

namespace {

constexpr int kProducerThreads = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct ConvGeom {
    disco_conv_desc d;
    int ncb, ncb0;            // K stages total / from source 0
    int chunks, chunk_shift;  // c_blk/8
    int nparts;               // 1 (fp16) | 2 (bf16 hi+lo)
    int PH, PW, PIX;          // staged patch (rows, cols, pixels)
    int plane, parplane;      // bytes
    int a_part_bytes, a_stage_bytes;
    int b_part_bytes, b_stage_bytes;
    int SA, SB;
    int sbo_a;
    int tiles_h, tiles_w;
    long long total_pix;      // n*h_out*w_out
    int tmem_cols;
    int smem_bytes;
};

struct __align__(8) SmemCtl {
    uint64_t a_full[kMaxStages];
    uint64_t a_empty[kMaxStages];
    uint64_t b_full[kMaxStages];
    uint64_t b_empty[kMaxStages];
    uint64_t acc_full;
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ uint4 pack_bf16_hi8(const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        w[i] = (uint32_t)f32_to_bf16_bits(v[2 * i]) | ((uint32_t)f32_to_bf16_bits(v[2 * i + 1]) << 16);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(kThreads) conv_tc_kernel(const ConvGeom g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw);
    const uint32_t smem_base = smem_u32(smem_raw);
    const uint32_t a_base = smem_base + 384;  // SmemCtl (272 B) rounded up to 128-byte multiple
    const uint32_t b_base = a_base + g.SA * g.a_stage_bytes;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const disco_conv_desc& d = g.d;

    // ---- tile coordinates --------------------------------------------------------------------
    const int m_tile = blockIdx.x;
    const int n_tile = blockIdx.y;
    int img = 0, h0 = 0, w0 = 0;
    long long p0 = 0;
    if (d.taps == 9) {
        const int per_img = g.tiles_h * g.tiles_w;
        img = m_tile / per_img;
        const int rem = m_tile - img * per_img;
        h0 = (rem / g.tiles_w) * 16;
        w0 = (rem % g.tiles_w) * 8;
    } else {
        p0 = (long long)m_tile * 128;
    }

    // ---- one-time setup ------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < g.SA; ++s) {
            mbar_init(smem_u32(&ctl->a_full[s]), kProducerThreads);
            mbar_init(smem_u32(&ctl->a_empty[s]), 1);
        }
        for (int s = 0; s < g.SB; ++s) {
            mbar_init(smem_u32(&ctl->b_full[s]), 1);
            mbar_init(smem_u32(&ctl->b_empty[s]), 1);
        }
        mbar_init(smem_u32(&ctl->acc_full), 1);
        fence_mbar_init();
    }
    if (warp == 5) {
        tmem_alloc(smem_u32(&ctl->tmem_base), (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = ctl->tmem_base;

    const int n_iters_b = g.ncb * d.taps;

    if (warp < 4) {
        // =========================== A producer (gather) ==========================================
        const int per_part = g.PIX << g.chunk_shift;
        for (int cb = 0; cb < g.ncb; ++cb) {
            const int sa = cb % g.SA;
            const uint32_t ph = (uint32_t)(cb / g.SA) & 1u;
            mbar_wait(smem_u32(&ctl->a_empty[sa]), ph ^ 1u);
            const int s = (cb < g.ncb0) ? 0 : 1;
            const int cbl = (s == 0) ? cb : cb - g.ncb0;
            const uint16_t* __restrict__ src = reinterpret_cast<const uint16_t*>(s ? d.src[1] : d.src[0]);
            const int Cs = s ? d.src_c[1] : d.src_c[0];
            const int up = s ? d.src_up[1] : d.src_up[0];
            const int Hs = d.h_in >> up, Ws = d.w_in >> up;
            const long long lo_off = s ? d.src_lo_off[1] : d.src_lo_off[0];
            const uint32_t stage = a_base + sa * g.a_stage_bytes;
            const int cofs = cbl * d.c_blk;
            for (int e = tid; e < per_part; e += kProducerThreads) {
                const int chunk = e & (g.chunks - 1);
                const int pix = e >> g.chunk_shift;
                long long goff;
                uint32_t soff;
                bool valid;
                if (d.taps == 9) {
                    const int r = pix / g.PW;
                    const int c = pix - r * g.PW;
                    const int hi = h0 * d.stride - 1 + r;
                    const int wi = w0 * d.stride - 1 + c;
                    valid = (hi >= 0) && (hi < d.h_in) && (wi >= 0) && (wi < d.w_in);
                    goff = (((long long)img * Hs + (hi >> up)) * Ws + (wi >> up)) * Cs + cofs + chunk * 8;
                    if (d.stride == 1) soff = (uint32_t)(r * 10 + c) * 16u;
                    else soff = (uint32_t)(c & 1) * g.parplane + (uint32_t)(r * 9 + (c >> 1)) * 16u;
                } else {
                    const long long p = p0 + pix;
                    valid = p < g.total_pix;
                    goff = p * Cs + cofs + chunk * 8;
                    soff = (uint32_t)pix * 16u;
                }
                const uint32_t dst = stage + chunk * g.plane + soff;
                const uint16_t* gp = valid ? (src + goff) : src;
                cp_async16(dst, gp, valid ? 16u : 0u);
                if (g.nparts == 2) cp_async16(dst + g.a_part_bytes, valid ? (gp + lo_off) : src, valid ? 16u : 0u);
            }
            cp_async_commit();
            if (cb >= 1) {  // keep one gather in flight while publishing the previous one
                cp_async_wait<1>();
                fence_proxy_async_smem();
                mbar_arrive(smem_u32(&ctl->a_full[(cb - 1) % g.SA]));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&ctl->a_full[(g.ncb - 1) % g.SA]));

        // =========================== epilogue =====================================================
        mbar_wait(smem_u32(&ctl->acc_full), 0);
        tc_fence_after();
        const int m = warp * 32 + lane;
        bool valid;
        long long pixel;
        if (d.taps == 9) {
            const int oh = h0 + (m >> 3), ow = w0 + (m & 7);
            valid = (oh < d.h_out) && (ow < d.w_out);
            pixel = ((long long)img * d.h_out + oh) * d.w_out + ow;
        } else {
            pixel = p0 + m;
            valid = pixel < g.total_pix;
        }
        const uint32_t t_lane = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int j = 0; j < d.block_n / 16; ++j) {
            uint32_t raw[16];
            tmem_ld16(t_lane + (uint32_t)(j * 16), raw);
            tmem_ld_wait();
            const int nb = n_tile * d.block_n + j * 16;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x = __uint_as_float(raw[i]) + __ldg(d.bias + nb + i);
                v[i] = d.relu ? fmaxf(x, 0.f) : x;
            }
            // stores are predicated per lane; the tcgen05.ld above must stay warp-convergent
            if (valid && d.out_mode == DISCO_OUT_ACT && nb < d.c_out) {
                uint16_t* oh_ = reinterpret_cast<uint16_t*>(d.out[0]) + pixel * d.c_out + nb;
                if (d.precision == DISCO_PREC_BF16X3) {
                    float lo[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) lo[i] = v[i] - bf16_bits_to_f32(f32_to_bf16_bits(v[i]));
                    reinterpret_cast<uint4*>(oh_)[0] = pack_bf16_hi8(v);
                    reinterpret_cast<uint4*>(oh_)[1] = pack_bf16_hi8(v + 8);
                    uint16_t* ol_ = oh_ + d.out_lo_off;
                    reinterpret_cast<uint4*>(ol_)[0] = pack_bf16_hi8(lo);
                    reinterpret_cast<uint4*>(ol_)[1] = pack_bf16_hi8(lo + 8);
                } else {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        w[i] = (uint32_t)f32_to_f16_bits(v[2 * i]) | ((uint32_t)f32_to_f16_bits(v[2 * i + 1]) << 16);
                    reinterpret_cast<uint4*>(oh_)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    reinterpret_cast<uint4*>(oh_)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
            } else if (valid && d.out_mode == DISCO_OUT_F32) {
                const int c1 = d.c_out - d.out_split;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = nb + 4 * q;
                    if (c >= d.c_out) continue;
                    float* dst = (c < d.out_split)
                                     ? reinterpret_cast<float*>(d.out[0]) + pixel * d.out_split + c
                                     : reinterpret_cast<float*>(d.out[1]) + pixel * c1 + (c - d.out_split);
                    *reinterpret_cast<float4*>(dst) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        }
    } else if (warp == 4) {
        // =========================== B loader (bulk copy engine) ===================================
        if (lane == 0) {
            const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.wpack) +
                                (size_t)n_tile * n_iters_b * g.b_stage_bytes;
            for (int it = 0; it < n_iters_b; ++it) {
                const int sb = it % g.SB;
                const uint32_t ph = (uint32_t)(it / g.SB) & 1u;
                mbar_wait(smem_u32(&ctl->b_empty[sb]), ph ^ 1u);
                const uint32_t bar = smem_u32(&ctl->b_full[sb]);
                mbar_arrive_expect_tx(bar, (uint32_t)g.b_stage_bytes);
                bulk_g2s(b_base + sb * g.b_stage_bytes, wp + (size_t)it * g.b_stage_bytes, (uint32_t)g.b_stage_bytes,
                         bar);
            }
        }
    } else {
        // =========================== MMA issuer ===================================================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(d.precision == DISCO_PREC_BF16X3, 128, d.block_n);
            const uint32_t lbo_b = (uint32_t)d.block_n * 16u;
            const int ksteps = d.c_blk / 16;
            uint32_t acc = 0;
            int it = 0;
            for (int cb = 0; cb < g.ncb; ++cb) {
                const int sa = cb % g.SA;
                mbar_wait(smem_u32(&ctl->a_full[sa]), (uint32_t)(cb / g.SA) & 1u);
                tc_fence_after();
                const uint32_t a_stage = a_base + sa * g.a_stage_bytes;
                for (int tap = 0; tap < d.taps; ++tap, ++it) {
                    const int sb = it % g.SB;
                    mbar_wait(smem_u32(&ctl->b_full[sb]), (uint32_t)(it / g.SB) & 1u);
                    tc_fence_after();
                    uint32_t a_tap = a_stage;
                    if (d.taps == 9) {
                        const int kh = tap / 3, kw = tap - kh * 3;
                        a_tap += (d.stride == 1) ? (uint32_t)(kh * 10 + kw) * 16u
                                                 : (uint32_t)(kw & 1) * g.parplane + (uint32_t)(kh * 9 + (kw >> 1)) * 16u;
                    }
                    const uint32_t b_stage = b_base + sb * g.b_stage_bytes;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t a_hi = a_tap + (uint32_t)(2 * ks) * g.plane;
                        const uint32_t b_hi = b_stage + (uint32_t)(2 * ks) * lbo_b;
                        const uint64_t da = umma_desc_kmajor_noswizzle(a_hi, (uint32_t)g.plane, (uint32_t)g.sbo_a);
                        const uint64_t db = umma_desc_kmajor_noswizzle(b_hi, lbo_b, 128u);
                        umma_f16(tmem_d, da, db, idesc, acc);
                        acc = 1;
                        if (g.nparts == 2) {
                            const uint64_t da_lo = umma_desc_kmajor_noswizzle(a_hi + g.a_part_bytes, (uint32_t)g.plane,
                                                                              (uint32_t)g.sbo_a);
                            const uint64_t db_lo = umma_desc_kmajor_noswizzle(b_hi + g.b_part_bytes, lbo_b, 128u);
                            umma_f16(tmem_d, da_lo, db, idesc, 1);
                            umma_f16(tmem_d, da, db_lo, idesc, 1);
                        }
                    }
                    umma_commit(smem_u32(&ctl->b_empty[sb]));
                }
                umma_commit(smem_u32(&ctl->a_empty[sa]));
            }
            umma_commit(smem_u32(&ctl->acc_full));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_d, (uint32_t)g.tmem_cols);
}

int build_geom(const disco_conv_desc* d, ConvGeom* g) {
    DISCO_REQUIRE(d->taps == 9 || d->taps == 1, "conv: taps must be 9 or 1 (got %d)", d->taps);
    DISCO_REQUIRE(d->stride == 1 || (d->stride == 2 && d->taps == 9), "conv: stride %d unsupported", d->stride);
    DISCO_REQUIRE(d->c_blk == 16 || d->c_blk == 32 || d->c_blk == 64, "conv: c_blk must be 16/32/64");
    DISCO_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "conv: bad block_n %d", d->block_n);
    DISCO_REQUIRE(d->src_c[0] > 0 && d->src_c[0] % d->c_blk == 0 && d->src_c[1] % d->c_blk == 0,
                  "conv: source channels (%d,%d) must be multiples of c_blk %d", d->src_c[0], d->src_c[1], d->c_blk);
    DISCO_REQUIRE(d->precision == DISCO_PREC_FP16 || d->precision == DISCO_PREC_BF16X3, "conv: bad precision");
    DISCO_REQUIRE(d->taps == 9 || (d->src_up[0] == 0 && d->src_up[1] == 0), "conv: 1x1 cannot upsample");
    DISCO_REQUIRE(d->n > 0 && d->h_in > 0 && d->w_in > 0, "conv: empty input");
    if (d->taps == 9) {
        DISCO_REQUIRE(d->h_out == (d->h_in - 1) / d->stride + 1 && d->w_out == (d->w_in - 1) / d->stride + 1,
                      "conv: output size %dx%d inconsistent with input %dx%d stride %d", d->h_out, d->w_out, d->h_in,
                      d->w_in, d->stride);
    } else {
        DISCO_REQUIRE(d->h_out == d->h_in && d->w_out == d->w_in, "conv1x1: size mismatch");
    }
    for (int s = 0; s < 2; ++s)
        if (d->src_c[s] && d->src_up[s])
            DISCO_REQUIRE(d->h_in % 2 == 0 && d->w_in % 2 == 0, "conv: upsampled source needs even H_in/W_in");
    if (d->out_mode == DISCO_OUT_ACT)
        DISCO_REQUIRE(d->c_out % 16 == 0, "conv: activation outputs need c_out %% 16 == 0");
    else
        DISCO_REQUIRE(d->out_split % 4 == 0 && (d->c_out - d->out_split) % 4 == 0 && d->out_split <= d->c_out,
                      "conv: fp32 output split must be a multiple of 4");

    g->d = *d;
    g->ncb0 = d->src_c[0] / d->c_blk;
    g->ncb = g->ncb0 + d->src_c[1] / d->c_blk;
    g->chunks = d->c_blk / 8;
    g->chunk_shift = (g->chunks == 2) ? 1 : (g->chunks == 4) ? 2 : 3;
    g->nparts = (d->precision == DISCO_PREC_BF16X3) ? 2 : 1;
    if (d->taps == 1) {
        g->PH = 1; g->PW = 128; g->PIX = 128; g->parplane = 0; g->sbo_a = 128;
    } else if (d->stride == 1) {
        g->PH = 18; g->PW = 10; g->PIX = 180; g->parplane = 0; g->sbo_a = 160;
    } else {
        g->PH = 33; g->PW = 17; g->PIX = 33 * 17; g->parplane = 33 * 9 * 16; g->sbo_a = 2 * 9 * 16;
    }
    int plane = (d->taps == 9 && d->stride == 2) ? 2 * g->parplane : g->PIX * 16;
    // pad so that the `chunks` 16-byte writes of one pixel land in distinct bank groups
    const int want = (128 / (g->chunks > 8 ? 8 : g->chunks)) % 128;
    while ((plane % 128) != want) plane += 16;
    g->plane = plane;
    g->a_part_bytes = g->chunks * plane;
    g->a_stage_bytes = ((g->nparts * g->a_part_bytes + 127) / 128) * 128;
    g->b_part_bytes = d->c_blk * d->block_n * 2;
    g->b_stage_bytes = g->nparts * g->b_part_bytes;
    g->SA = (g->ncb >= 2) ? 2 : 1;
    const int budget = 200 * 1024 - 384 - g->SA * g->a_stage_bytes;
    int sb = budget / g->b_stage_bytes;
    if (sb > kMaxStages) sb = kMaxStages;
    const int iters = g->ncb * d->taps;
    if (sb > iters) sb = iters;
    DISCO_REQUIRE(sb >= 1, "conv: tile does not fit shared memory (a_stage %d, b_stage %d)", g->a_stage_bytes,
                  g->b_stage_bytes);
    // small-tile layers: keep the footprint low enough for several CTAs per SM (epilogue/mainloop overlap)
    if (g->SA * g->a_stage_bytes + 4 * g->b_stage_bytes <= 70 * 1024 && sb > 4) sb = 4;
    g->SB = sb;
    g->smem_bytes = 384 + g->SA * g->a_stage_bytes + g->SB * g->b_stage_bytes;
    g->tiles_h = (d->h_out + 15) / 16;
    g->tiles_w = (d->w_out + 7) / 8;
    g->total_pix = (long long)d->n * d->h_out * d->w_out;
    int cols = 32;
    while (cols < d->block_n) cols *= 2;
    g->tmem_cols = cols;
    DISCO_REQUIRE((g->plane >> 4) < 16384 && (g->a_part_bytes + g->a_stage_bytes) < (1 << 18), "conv: descriptor range");
    return DISCO_OK;
}

}  // namespace

int disco_conv_tc_smem_bytes(const disco_conv_desc* d) {
    ConvGeom g;
    int rc = build_geom(d, &g);
    return rc < 0 ? rc : g.smem_bytes;
}

int disco_conv_tc_launch(const disco_conv_desc* d, void* stream) {
    ConvGeom g;
    int rc = build_geom(d, &g);
    if (rc < 0) return rc;
    static bool attr_set[64] = {false};
    int dev = 0;
    DISCO_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        DISCO_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    const long long m_tiles = (d->taps == 9) ? (long long)d->n * g.tiles_h * g.tiles_w : (g.total_pix + 127) / 128;
    const int n_tiles = (d->c_out + d->block_n - 1) / d->block_n;
    DISCO_REQUIRE(m_tiles > 0 && m_tiles < (1ll << 31), "conv: bad tile count");
    dim3 grid((unsigned)m_tiles, (unsigned)n_tiles);
    conv_tc_kernel<<<grid, kThreads, g.smem_bytes, (cudaStream_t)stream>>>(g);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
