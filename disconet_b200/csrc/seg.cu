// Data-movement kernels of the BEV-segmentation U-Net (BASELINE config 5) around the conv kernel:
//   maxpool2            nn.MaxPool2d(2) in front of every Down block          (models/seg/SegModelBase.py:113-123)
//   upsample_bilinear2x nn.Upsample(scale_factor=2, bilinear, align_corners=True) of every Up block (:126-142); the
//                       channel concat [skip, up] that follows is address arithmetic in the conv gather
//   nhwc_to_nchw        layout of the logits the reference returns                                   (:145-151)
// All three are HBM-bound streaming kernels over NHWC activation buffers (hi/lo bf16 planes), 8 channels (16 bytes
// per plane) per thread.
#include "common.cuh"
#include "conv.h"
#include "ops.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void load_act8(const uint16_t* p, long long lo_off, int precision, float* v) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
    if (precision == DISCO_PREC_BF16X3) {
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + lo_off));
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[2 * q] = __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
            v[2 * q + 1] = __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[q]));
            v[2 * q] = f.x; v[2 * q + 1] = f.y;
        }
    }
}

__device__ __forceinline__ void store_act8(uint16_t* o, long long lo_off, int precision, const float* v) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (precision == DISCO_PREC_BF16X3) {
            uint16_t h0, l0, h1, l1;
            split_bf16(v[2 * q], h0, l0);
            split_bf16(v[2 * q + 1], h1, l1);
            hw[q] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            lw[q] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        } else {
            hw[q] = (uint32_t)f32_to_f16_bits(v[2 * q]) | ((uint32_t)f32_to_f16_bits(v[2 * q + 1]) << 16);
        }
    }
    *reinterpret_cast<uint4*>(o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    if (precision == DISCO_PREC_BF16X3) *reinterpret_cast<uint4*>(o + lo_off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

__global__ void __launch_bounds__(kThreads) maxpool2_kernel(const uint16_t* src, long long src_lo, uint16_t* dst, long long dst_lo,
                                                            int precision, int n, int h, int w, int c) {
    const int ho = h >> 1, wo = w >> 1, groups = c >> 3;
    const long long total = (long long)n * ho * wo * groups;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(e % groups);
        long long r = e / groups;
        const int x = (int)(r % wo); r /= wo;
        const int y = (int)(r % ho);
        const long long img = r / ho;
        const uint16_t* p = src + (((img * h + 2 * y) * w + 2 * x) * c + g * 8);
        float a[8], b[8], m[8];
        load_act8(p, src_lo, precision, m);
        load_act8(p + c, src_lo, precision, a);
        load_act8(p + (long long)w * c, src_lo, precision, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], fmaxf(a[i], b[i]));
        load_act8(p + (long long)w * c + c, src_lo, precision, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], a[i]);
        store_act8(dst + (((img * ho + y) * wo + x) * c + g * 8), dst_lo, precision, m);
    }
}

// out[oy, ox] = bilinear sample of src at (oy * (h-1)/(2h-1), ox * (w-1)/(2w-1))   (align_corners=True)
__global__ void __launch_bounds__(kThreads) upsample2x_kernel(const uint16_t* src, long long src_lo, uint16_t* dst, long long dst_lo,
                                                              int precision, int n, int h, int w, int c) {
    const int ho = 2 * h, wo = 2 * w, groups = c >> 3;
    const float rh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f;
    const float rw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
    const long long total = (long long)n * ho * wo * groups;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(e % groups);
        long long r = e / groups;
        const int ox = (int)(r % wo); r /= wo;
        const int oy = (int)(r % ho);
        const long long img = r / ho;
        const float sy = rh * oy, sx = rw * ox;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float ly = sy - y0, lx = sx - x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const uint16_t* base = src + img * h * w * c + g * 8;
        float v00[8], v01[8], v10[8], v11[8], o[8];
        load_act8(base + ((long long)y0 * w + x0) * c, src_lo, precision, v00);
        load_act8(base + ((long long)y0 * w + x1) * c, src_lo, precision, v01);
        load_act8(base + ((long long)y1 * w + x0) * c, src_lo, precision, v10);
        load_act8(base + ((long long)y1 * w + x1) * c, src_lo, precision, v11);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = hy * (hx * v00[i] + lx * v01[i]) + ly * (hx * v10[i] + lx * v11[i]);
        store_act8(dst + (((img * ho + oy) * wo + ox) * c + g * 8), dst_lo, precision, o);
    }
}

__global__ void __launch_bounds__(kThreads) nhwc_to_nchw_kernel(const float* src, int c_src, int c, long long hw, long long total,
                                                                float* dst) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long img = e / hw, p = e - img * hw;
        const float* s = src + e * c_src;
        for (int k = 0; k < c; ++k) dst[(img * c + k) * hw + p] = s[k];
    }
}

// Backward of MaxPool2d(2): the gradient of a pooled cell goes to the FIRST maximum of its 2x2 block in scan order
// (torch's argmax convention), zeros elsewhere.  x = the pool's input activation buffer; g fp32 [n, h/2, w/2, c].
__global__ void __launch_bounds__(kThreads) maxpool2_bwd_kernel(const uint16_t* x, long long x_lo, int precision, const float* g,
                                                                float* gx, int n, int h, int w, int c) {
    const int ho = h >> 1, wo = w >> 1, groups = c >> 3;
    const long long total = (long long)n * ho * wo * groups;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int gq = (int)(e % groups);
        long long r = e / groups;
        const int px = (int)(r % wo); r /= wo;
        const int py = (int)(r % ho);
        const long long img = r / ho;
        const long long base = ((img * h + 2 * py) * w + 2 * px) * c + gq * 8;
        const long long off[4] = {0, c, (long long)w * c, (long long)w * c + c};
        float v[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k) load_act8(x + base + off[k], x_lo, precision, v[k]);
        const float4* gp = reinterpret_cast<const float4*>(g + (((img * ho + py) * wo + px) * c + gq * 8));
        const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        float o[4][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int best = 0;
            float m = v[0][i];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (v[k][i] > m) { m = v[k][i]; best = k; }
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k][i] = (k == best) ? gg[i] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4* dst = reinterpret_cast<float4*>(gx + base + off[k]);
            dst[0] = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
            dst[1] = make_float4(o[k][4], o[k][5], o[k][6], o[k][7]);
        }
    }
}

// Backward of the bilinear x2 upsample (align_corners=True): transpose of upsample2x_kernel as a gather -- source cell
// (y, x) collects w_y * w_x * g_up[oy, ox] from the output rows / columns whose interpolation footprint contains it
// (the same float formulas as the forward, so the weights match bit for bit).  g_up fp32 [n, 2h, 2w, c] -> gs [n, h, w, c].
__global__ void __launch_bounds__(kThreads) upsample2x_bwd_kernel(const float* g_up, float* gs, int n, int h, int w, int c) {
    const int ho = 2 * h, wo = 2 * w, groups = c >> 3;
    const float rh = ho > 1 ? (float)(h - 1) / (float)(ho - 1) : 0.f;
    const float rw = wo > 1 ? (float)(w - 1) / (float)(wo - 1) : 0.f;
    const long long total = (long long)n * h * w * groups;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int gq = (int)(e % groups);
        long long r = e / groups;
        const int x = (int)(r % w); r /= w;
        const int y = (int)(r % h);
        const long long img = r / h;
        // candidate output rows / columns: sy = rh * oy in [y - 1, y + 1]
        const int oy_lo = max(0, 2 * y - 3), oy_hi = min(ho - 1, 2 * y + 3);
        const int ox_lo = max(0, 2 * x - 3), ox_hi = min(wo - 1, 2 * x + 3);
        float wy[7], wx[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            wy[k] = 0.f; wx[k] = 0.f;
            const int oy = oy_lo + k;
            if (oy <= oy_hi) {
                const float sy = rh * oy;
                const int y0 = (int)sy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
                const float ly = sy - y0;
                if (y0 == y) wy[k] += 1.f - ly;
                if (y1 == y) wy[k] += ly;
            }
            const int ox = ox_lo + k;
            if (ox <= ox_hi) {
                const float sx = rw * ox;
                const int x0 = (int)sx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
                const float lx = sx - x0;
                if (x0 == x) wx[k] += 1.f - lx;
                if (x1 == x) wx[k] += lx;
            }
        }
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int a = 0; a < 7; ++a) {
            if (wy[a] == 0.f) continue;
            for (int b = 0; b < 7; ++b) {
                const float wgt = wy[a] * wx[b];
                if (wgt == 0.f) continue;
                const float4* p = reinterpret_cast<const float4*>(g_up + (((img * ho + oy_lo + a) * wo + ox_lo + b) * c + gq * 8));
                const float4 u0 = __ldg(p), u1 = __ldg(p + 1);
                acc[0] = fmaf(wgt, u0.x, acc[0]); acc[1] = fmaf(wgt, u0.y, acc[1]); acc[2] = fmaf(wgt, u0.z, acc[2]); acc[3] = fmaf(wgt, u0.w, acc[3]);
                acc[4] = fmaf(wgt, u1.x, acc[4]); acc[5] = fmaf(wgt, u1.y, acc[5]); acc[6] = fmaf(wgt, u1.z, acc[6]); acc[7] = fmaf(wgt, u1.w, acc[7]);
            }
        }
        float4* dst = reinterpret_cast<float4*>(gs + (((img * h + y) * w + x) * c + gq * 8));
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

unsigned blocks_for(long long total) {
    long long b = (total + kThreads - 1) / kThreads;
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

int disco_maxpool2_launch(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision, int n,
                          int h, int w, int c, void* stream) {
    DISCO_REQUIRE(src_hi && dst_hi && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "maxpool2: bad arguments");
    maxpool2_kernel<<<blocks_for((long long)n * (h / 2) * (w / 2) * (c / 8)), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint16_t*>(src_hi), src_lo_off, reinterpret_cast<uint16_t*>(dst_hi), dst_lo_off, precision, n, h, w, c);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_upsample_bilinear2x_launch(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision,
                                     int n, int h, int w, int c, void* stream) {
    DISCO_REQUIRE(src_hi && dst_hi && n > 0 && h > 0 && w > 0 && c % 8 == 0, "upsample_bilinear2x: bad arguments");
    upsample2x_kernel<<<blocks_for((long long)n * (2 * h) * (2 * w) * (c / 8)), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint16_t*>(src_hi), src_lo_off, reinterpret_cast<uint16_t*>(dst_hi), dst_lo_off, precision, n, h, w, c);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_nhwc_to_nchw_launch(const float* src, int n, int h, int w, int c_src, int c, float* dst, void* stream) {
    DISCO_REQUIRE(src && dst && n > 0 && h > 0 && w > 0 && c > 0 && c <= c_src, "nhwc_to_nchw: bad arguments");
    const long long hw = (long long)h * w, total = hw * n;
    nhwc_to_nchw_kernel<<<blocks_for(total), kThreads, 0, (cudaStream_t)stream>>>(src, c_src, c, hw, total, dst);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_maxpool2_backward_launch(const void* x_hi, long long x_lo_off, int precision, const float* g, float* gx, int n, int h, int w,
                                   int c, void* stream) {
    DISCO_REQUIRE(x_hi && g && gx && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "maxpool2_backward: bad arguments");
    maxpool2_bwd_kernel<<<blocks_for((long long)n * (h / 2) * (w / 2) * (c / 8)), kThreads, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint16_t*>(x_hi), x_lo_off, precision, g, gx, n, h, w, c);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_upsample_bilinear2x_backward_launch(const float* g_up, float* gs, int n, int h, int w, int c, void* stream) {
    DISCO_REQUIRE(g_up && gs && n > 0 && h > 1 && w > 1 && c % 8 == 0, "upsample_bilinear2x_backward: bad arguments");
    upsample2x_bwd_kernel<<<blocks_for((long long)n * h * w * (c / 8)), kThreads, 0, (cudaStream_t)stream>>>(g_up, gs, n, h, w, c);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
