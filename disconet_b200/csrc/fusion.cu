// DiscoGraph fusion block as ONE kernel (reference: ~560 launches of a Python triple loop per scene).
//
// For every scene b, ego agent i and BEV cell (y,x) of the collaboration layer:
//   * neighbour j's feature map is warped into the ego frame (affine_grid + bilinear grid_sample with
//     zeros padding, align_corners=False -- DetModelBase.py:139-169).  The reference flips H before and
//     after the warp (DetModelBase.py:67,91); here the flip is folded into the row index.
//   * PixelWeightedFusionSoftmax (DiscoNet.py:132-155) scores cat[ego, neighbour]: the 2C->128 1x1 conv
//     is linear and bias-free on the neighbour half, so it commutes with the bilinear warp; it is
//     applied once per agent on the tensor cores (conv_tc, `en` input) and only the 128->32->8->1 tail
//     runs here, one warp per cell with warp shuffles for the narrow layers.
//   * agent-axis softmax exp(w_k)/sum_k exp(w_k) over k = {ego, neighbours...} and the weighted feature
//     sum (DiscoNet.py:94-108) are accumulated on the fly; agents >= num_agent[b] keep their features
//     (local_com_mat_update is initialised as a copy, DiscoNet.py:52-57).
#include "common.cuh"
#include "conv.h"
#include "ops.h"

namespace {

constexpr int kWarps = 8;
constexpr int kCellsPerWarp = 4;
constexpr int kHid = 128, kH2 = 32, kH3 = 8;

// CPL channels per lane (C = 32*CPL): 4 -> one 8-byte load, 8 / 16 -> one / two 16-byte loads per plane
template <int CPL>
__device__ __forceinline__ void load_feat(const uint16_t* p, long long lo_off, int precision, float wgt,
                                          float (&acc)[CPL]) {
    constexpr int NW = CPL / 2;   // 32-bit words (2 channels each)
    uint32_t hw[NW], lw[NW];
    if (CPL == 4) {
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(p));
        hw[0] = h.x; hw[1] = h.y;
        if (precision == DISCO_PREC_BF16X3) {
            const uint2 l = __ldg(reinterpret_cast<const uint2*>(p + lo_off));
            lw[0] = l.x; lw[1] = l.y;
        }
    } else {
#pragma unroll
        for (int v = 0; v < CPL / 8; ++v) {
            const uint4 h = __ldg(reinterpret_cast<const uint4*>(p) + v);
            hw[4 * v] = h.x; hw[4 * v + 1] = h.y; hw[4 * v + 2] = h.z; hw[4 * v + 3] = h.w;
            if (precision == DISCO_PREC_BF16X3) {
                const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + lo_off) + v);
                lw[4 * v] = l.x; lw[4 * v + 1] = l.y; lw[4 * v + 2] = l.z; lw[4 * v + 3] = l.w;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        float a, b;
        if (precision == DISCO_PREC_BF16X3) {
            a = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
            b = __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
        } else {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[q]));
            a = f.x; b = f.y;
        }
        acc[2 * q] = fmaf(wgt, a, acc[2 * q]);
        acc[2 * q + 1] = fmaf(wgt, b, acc[2 * q + 1]);
    }
}

template <int CPL>
__global__ void __launch_bounds__(kWarps * 32, 3) fusion_kernel(const disco_fusion_desc d) {
    __shared__ __align__(16) float s_w2[kH2][kHid + 4];   // row pitch 132 floats: 16-byte aligned rows, LDS.128 by 8 lanes hits 32 banks
    __shared__ float s_b2[kH2];
    __shared__ float s_w3[kH3][kH2];
    __shared__ float s_b3[kH3];
    __shared__ float s_w4[kH3];
    __shared__ float s_b4;
    __shared__ __align__(16) float s_h1[kWarps][kHid];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!d.wpre) {
        for (int e = tid; e < kH2 * kHid; e += blockDim.x) s_w2[e / kHid][e % kHid] = d.w2[e];
        for (int e = tid; e < kH3 * kH2; e += blockDim.x) s_w3[e / kH2][e % kH2] = d.w3[e];
        if (tid < kH2) s_b2[tid] = d.b2[tid];
        if (tid < kH3) { s_b3[tid] = d.b3[tid]; s_w4[tid] = d.w4[tid]; }
        if (tid == 0) s_b4 = d.b4[0];
    }
    __syncthreads();

    const int C = d.C, h = d.h, w = d.w, A = d.A, B = d.B;
    const uint16_t* feat = reinterpret_cast<const uint16_t*>(d.feat_hi);
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out_hi);
    const long long cells = (long long)(d.row_end - d.row_begin) * h * w;
    const int cpl = CPL;  // channels per lane

    for (int t = 0; t < kCellsPerWarp; ++t) {
        const long long cell = ((long long)blockIdx.x * kWarps + warp) * kCellsPerWarp + t;
        if (cell >= cells) break;  // warp-uniform
        const int x = (int)(cell % w);
        const int y = (int)((cell / w) % h);
        const int n_i = d.row_begin + (int)(cell / ((long long)w * h));   // agent-major row of the ego map
        const int i = n_i / B, b = n_i - i * B;
        const int n_ag = d.num_agent[b];
        const long long row_i = (((long long)n_i * h + y) * w + x);
        const long long row_o = (((long long)(n_i - d.row_begin) * h + y) * w + x);

        float acc[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[c] = 0.f;
        float esum = 0.f;

        const bool passthrough = (i >= n_ag) || (d.outage && d.outage[b * A + i]);
        if (passthrough) {  // absent agent / communication outage: features pass through unchanged
            load_feat<CPL>(feat + row_i * C + lane * cpl, d.feat_lo_off, d.precision, 1.f, acc);
            esum = 1.f;
            if (d.weights && lane < A && i < n_ag)   // outage: the ego is its own (only) contributor
                d.weights[((((long long)b * A + i) * A + lane) * h + y) * w + x] = (lane == i) ? 1.f : 0.f;
        } else {
            // training mode: the PWF output maps were computed with per-pair batch statistics (fusion_train.cu)
            const bool pre = d.wpre != nullptr;
            const float4 e4 = pre ? make_float4(0.f, 0.f, 0.f, 0.f)
                                  : __ldg(reinterpret_cast<const float4*>(d.en + row_i * (2 * kHid)) + lane);
            const int yf = h - 1 - y;  // row in the H-flipped frame the reference warps in
            const float xb = (2.f * x + 1.f) / w - 1.f;
            const float yb = (2.f * yf + 1.f) / h - 1.f;
            for (int k = 0; k < n_ag; ++k) {
                // reference order: ego first, then neighbours j = 0..n-1 skipping i
                const int j = (k == 0) ? i : ((k - 1 < i) ? k - 1 : k);
                if (j != i && d.only_v2i && i != 0 && j != 0) continue;
                float nb[CPL];
#pragma unroll
                for (int c = 0; c < CPL; ++c) nb[c] = 0.f;
                float nn[4] = {0.f, 0.f, 0.f, 0.f};
                if (j == i) {
                    load_feat<CPL>(feat + row_i * C + lane * cpl, d.feat_lo_off, d.precision, 1.f, nb);
                    if (!pre) {
                        const float4 n4 = __ldg(reinterpret_cast<const float4*>(d.en + row_i * (2 * kHid) + kHid) + lane);
                        nn[0] = n4.x; nn[1] = n4.y; nn[2] = n4.z; nn[3] = n4.w;
                    }
                } else {
                    const double* T = d.trans + (((long long)b * A + j) * A + i) * 16;
                    const float m00 = (float)T[0], m01 = (float)T[1], m02 = (-(float)T[3]) * d.trans_scale;
                    const float m10 = (float)T[4], m11 = (float)T[5], m12 = (-(float)T[7]) * d.trans_scale;
                    const float gx = m00 * xb + m01 * yb + m02;
                    const float gy = m10 * xb + m11 * yb + m12;
                    const float ix = ((gx + 1.f) * w - 1.f) * 0.5f;
                    const float iy = ((gy + 1.f) * h - 1.f) * 0.5f;
                    const float fx = floorf(ix), fy = floorf(iy);
                    const float ax = ix - fx, ay = iy - fy;
                    // guard against inf/nan/huge coordinates before the int conversion
                    const bool finite = (fabsf(ix) < 1e6f) && (fabsf(iy) < 1e6f);
                    const int x0 = finite ? (int)fx : -10, y0 = finite ? (int)fy : -10;
                    const float tw[4] = {(1.f - ax) * (1.f - ay), ax * (1.f - ay), (1.f - ax) * ay, ax * ay};
#pragma unroll
                    for (int tp = 0; tp < 4; ++tp) {
                        const int xs = x0 + (tp & 1), ysf = y0 + (tp >> 1);
                        if (xs < 0 || xs >= w || ysf < 0 || ysf >= h) continue;  // zeros padding
                        const long long row_j = (((long long)(j * B + b) * h + (h - 1 - ysf)) * w + xs);
                        load_feat<CPL>(feat + row_j * C + lane * cpl, d.feat_lo_off, d.precision, tw[tp], nb);
                        if (pre) continue;
                        const float4 n4 =
                            __ldg(reinterpret_cast<const float4*>(d.en + row_j * (2 * kHid) + kHid) + lane);
                        nn[0] = fmaf(tw[tp], n4.x, nn[0]); nn[1] = fmaf(tw[tp], n4.y, nn[1]);
                        nn[2] = fmaf(tw[tp], n4.z, nn[2]); nn[3] = fmaf(tw[tp], n4.w, nn[3]);
                    }
                }
                float wk;
                if (pre) {
                    wk = __ldg(d.wpre + ((((long long)b * A + i) * A + j) * h + y) * w + x);
                } else {
                // ---- PWF tail: 128 -> 32 -> 8 -> 1 ------------------------------------------------
                float4 h1;
                h1.x = fmaxf(e4.x + nn[0], 0.f); h1.y = fmaxf(e4.y + nn[1], 0.f);
                h1.z = fmaxf(e4.z + nn[2], 0.f); h1.w = fmaxf(e4.w + nn[3], 0.f);
                __syncwarp();
                reinterpret_cast<float4*>(s_h1[warp])[lane] = h1;
                __syncwarp();
                float h2;
                {   // four independent accumulators: the 128-long FMA chain is otherwise pure latency
                    float q0 = s_b2[lane], q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 8
                    for (int c = 0; c < kHid; c += 4) {
                        const float4 hv = *reinterpret_cast<const float4*>(&s_h1[warp][c]);
                        const float4 wv = *reinterpret_cast<const float4*>(&s_w2[lane][c]);   // one LDS.128 instead of four LDS.32
                        q0 = fmaf(wv.x, hv.x, q0);
                        q1 = fmaf(wv.y, hv.y, q1);
                        q2 = fmaf(wv.z, hv.z, q2);
                        q3 = fmaf(wv.w, hv.w, q3);
                    }
                    h2 = (q0 + q1) + (q2 + q3);
                }
                h2 = fmaxf(h2, 0.f);
                {
                    // 32 -> 8: eight dot products over the lanes as a reduce-scatter (8 -> 4 -> 2 -> 1 values per lane, then two
                    // plain butterfly steps): 9 shuffles instead of 8 x 5; lane bits 4,3,2 then select the output unit q.
                    float v[kH3];
#pragma unroll
                    for (int q = 0; q < kH3; ++q) v[q] = s_w3[q][lane] * h2;
                    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
                    float u[4], t[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
                        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float send = b3 ? u[i] : u[i + 2], keep = b3 ? u[i + 2] : u[i];
                        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
                    float r = (b2 ? t[1] : t[0]) + __shfl_xor_sync(0xffffffffu, b2 ? t[0] : t[1], 4);
                    r += __shfl_xor_sync(0xffffffffu, r, 2);
                    r += __shfl_xor_sync(0xffffffffu, r, 1);
                    const int q = (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0);
                    // 8 -> 1: sum over q = sum over lane bits 4,3,2 (every q is replicated on the 4 lanes of its group)
                    float o = s_w4[q] * fmaxf(r + s_b3[q], 0.f);
                    o += __shfl_xor_sync(0xffffffffu, o, 4);
                    o += __shfl_xor_sync(0xffffffffu, o, 8);
                    o += __shfl_xor_sync(0xffffffffu, o, 16);
                    wk = fmaxf(o + s_b4, 0.f);
                }
                }
                const float ek = expf(wk);
                esum += ek;
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc[c] = fmaf(ek, nb[c], acc[c]);
                if (d.weights && lane == 0)
                    d.weights[((((long long)b * A + i) * A + j) * h + y) * w + x] = ek;
            }
            if (d.weights) {
                __syncwarp();
                if (lane < A) {
                    const long long wi = ((((long long)b * A + i) * A + lane) * h + y) * w + x;
                    const bool used = (lane < n_ag) && (lane == i || !(d.only_v2i && i != 0 && lane != 0));
                    d.weights[wi] = used ? d.weights[wi] / esum : 0.f;
                }
            }
        }
        // ---- normalise and store -------------------------------------------------------------------
        const float inv = passthrough ? 1.f : 1.f / esum;
        uint16_t* o = out + row_o * C + lane * cpl;
        uint32_t hi[CPL / 2], lo[CPL / 2];
#pragma unroll
        for (int q = 0; q < CPL / 2; ++q) {
            const float a = acc[2 * q] * inv, c2 = acc[2 * q + 1] * inv;
            if (d.precision == DISCO_PREC_BF16X3) {
                uint16_t h0, l0, h1_, l1;
                split_bf16(a, h0, l0);
                split_bf16(c2, h1_, l1);
                hi[q] = (uint32_t)h0 | ((uint32_t)h1_ << 16);
                lo[q] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            } else {
                hi[q] = (uint32_t)f32_to_f16_bits(a) | ((uint32_t)f32_to_f16_bits(c2) << 16);
                lo[q] = 0;
            }
        }
        if (CPL == 4) {
            *reinterpret_cast<uint2*>(o) = make_uint2(hi[0], hi[1]);
            if (d.precision == DISCO_PREC_BF16X3) *reinterpret_cast<uint2*>(o + d.out_lo_off) = make_uint2(lo[0], lo[1]);
        } else {
#pragma unroll
            for (int v = 0; v < CPL / 8; ++v) {
                reinterpret_cast<uint4*>(o)[v] = make_uint4(hi[4 * v], hi[4 * v + 1], hi[4 * v + 2], hi[4 * v + 3]);
                if (d.precision == DISCO_PREC_BF16X3)
                    reinterpret_cast<uint4*>(o + d.out_lo_off)[v] =
                        make_uint4(lo[4 * v], lo[4 * v + 1], lo[4 * v + 2], lo[4 * v + 3]);
            }
        }
    }
}

}  // namespace

int disco_fusion_launch(const disco_fusion_desc* d, void* stream) {
    DISCO_REQUIRE(d->feat_hi && d->out_hi && d->trans && d->num_agent, "fusion: null tensor");
    DISCO_REQUIRE(d->wpre || (d->en && d->w2 && d->b2 && d->w3 && d->b3 && d->w4 && d->b4), "fusion: null PWF weights");
    DISCO_REQUIRE(d->hid == kHid, "fusion: PWF hidden width must be %d (got %d)", kHid, d->hid);
    DISCO_REQUIRE(d->C == 128 || d->C == 256 || d->C == 512, "fusion: C must be 128, 256 or 512 (got %d)", d->C);
    DISCO_REQUIRE(d->A >= 1 && d->A <= 32 && d->B >= 1 && d->h > 0 && d->w > 0, "fusion: bad scene shape");
    DISCO_REQUIRE(d->row_begin >= 0 && d->row_begin < d->row_end && d->row_end <= d->A * d->B,
                  "fusion: bad ego row range [%d,%d)", d->row_begin, d->row_end);
    const long long cells = (long long)(d->row_end - d->row_begin) * d->h * d->w;
    const long long per_block = kWarps * kCellsPerWarp;
    const long long blocks = (cells + per_block - 1) / per_block;
    DISCO_REQUIRE(blocks < (1ll << 31), "fusion: too many cells");
    if (d->C == 128) fusion_kernel<4><<<(unsigned)blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(*d);
    else if (d->C == 256) fusion_kernel<8><<<(unsigned)blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(*d);
    else fusion_kernel<16><<<(unsigned)blocks, kWarps * 32, 0, (cudaStream_t)stream>>>(*d);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
