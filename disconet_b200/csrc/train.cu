// Training-mode BatchNorm (batch statistics) forward/backward and gradient re-packing kernels.
//
// Replaces, for the DiscoNet hot path in train() mode, the F.batch_norm + F.relu calls of Backbone.encode /
// decode (Backbone.py:102-136,173-237), the heads (DetModelBase.py:283-351) and their autograd backward
// (CoDetModule.py:289-291).  All kernels are HBM-bound streaming passes over fp32 NHWC conv outputs:
//   forward  : stats (read z) + apply (read z, write hi/lo bf16)          = 12 B/element
//   backward : reduce (read z, g) + apply (read z, g, write hi/lo bf16)   = 20 B/element
// Per-channel sums are accumulated in fp32 per thread over <= ~150 pixels, then in double precision across
// threads/blocks, so the batch variance E[z^2] - mean^2 does not lose bits to cancellation.
#include "common.cuh"
#include "train.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxC = 512;

// 8 consecutive channels [c, c+8) of one gradient source at pixel `pix` (2x2 block sum for pooled sources)
__device__ __forceinline__ void grad_src_add8(const disco_grad_src& s, long long pix, int h, int w, int c, float* g) {
    if (!s.pool) {
        const float4* p = reinterpret_cast<const float4*>(s.ptr + pix * s.c_total + s.c_off + c);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        g[0] += a.x; g[1] += a.y; g[2] += a.z; g[3] += a.w;
        g[4] += b.x; g[5] += b.y; g[6] += b.z; g[7] += b.w;
        return;
    }
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const long long n = pix / ((long long)w * h);
    const long long base = ((n * (2 * h) + 2 * y) * (2 * w) + 2 * x) * s.c_total + s.c_off + c;
    const long long row = (long long)(2 * w) * s.c_total;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4* p = reinterpret_cast<const float4*>(s.ptr + base + (k >> 1) * row + (k & 1) * s.c_total);
        const float4 a = __ldg(p), b = __ldg(p + 1);
        g[0] += a.x; g[1] += a.y; g[2] += a.z; g[3] += a.w;
        g[4] += b.x; g[5] += b.y; g[6] += b.z; g[7] += b.w;
    }
}

// Per-channel reduction over all pixels.  Thread t owns the 8-channel group t % (C/8) and the pixel lane t / (C/8):
// 16-byte loads, 8 + 8 fp32 accumulators per thread, two pixels in flight per iteration; then a shared-memory
// reduction over the block's pixel lanes and one double-precision atomic per channel and block.
// MODE 0: sum z, sum z^2.   MODE 1: sum g, sum g*xhat with g = (sum of gradient sources) * [relu gate].
template <int MODE>
__global__ void __launch_bounds__(kThreads) bn_reduce_kernel(const disco_bn_desc d, long long M) {
    __shared__ float s_a[kThreads * 8 + 8], s_b[kThreads * 8 + 8];
    const int C = d.c, groups = C >> 3;
    const int t = threadIdx.x;
    const int lanes = kThreads / groups;            // pixel lanes per block (groups <= 64)
    const int lane = t / groups, c = (t - lane * groups) * 8;
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 0.f; b[i] = 0.f; }
    float mean[8], rstd[8], gam[8], bet[8];
    if (MODE == 1 && lane < lanes) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { mean[i] = d.stats[c + i]; rstd[i] = d.stats[C + c + i]; gam[i] = d.gamma[c + i]; bet[i] = d.beta[c + i]; }
    }
    if (lane < lanes) {
        constexpr int kUnroll = 2;
        const long long stride = (long long)gridDim.x * lanes;
        for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < M; p0 += stride * kUnroll) {
            float z[kUnroll][8], g[kUnroll][8];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long p = p0 + u * stride;
                const bool ok = p < M;
#pragma unroll
                for (int i = 0; i < 8; ++i) { z[u][i] = 0.f; g[u][i] = 0.f; }
                if (ok) {
                    const float4* zp = reinterpret_cast<const float4*>(d.z + p * C + c);
                    const float4 z0 = __ldg(zp), z1 = __ldg(zp + 1);
                    z[u][0] = z0.x; z[u][1] = z0.y; z[u][2] = z0.z; z[u][3] = z0.w;
                    z[u][4] = z1.x; z[u][5] = z1.y; z[u][6] = z1.z; z[u][7] = z1.w;
                    if (MODE == 1)
                        for (int s = 0; s < d.n_g; ++s) grad_src_add8(d.g[s], p, d.h, d.w, c, g[u]);
                }
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) { a[i] += z[u][i]; b[i] = fmaf(z[u][i], z[u][i], b[i]); }
                } else if (ok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float xh = (z[u][i] - mean[i]) * rstd[i];
                        float gi = g[u][i];
                        if (d.relu && fmaf(xh, gam[i], bet[i]) <= 0.f) gi = 0.f;
                        a[i] += gi;
                        b[i] = fmaf(gi, xh, b[i]);
                    }
                }
            }
        }
    }
    // [lane][channel] in shared memory (channel fastest: the final sum over lanes reads conflict-free)
    if (lane < lanes) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s_a[lane * C + c + i] = a[i]; s_b[lane * C + c + i] = b[i]; }
    }
    __syncthreads();
    for (int ch = t; ch < C; ch += kThreads) {
        double sa = 0.0, sb = 0.0;
        for (int l = 0; l < lanes; ++l) { sa += (double)s_a[l * C + ch]; sb += (double)s_b[l * C + ch]; }
        atomicAdd(d.sums + ch, sa);
        atomicAdd(d.sums + C + ch, sb);
    }
}

__global__ void bn_finalize_kernel(const disco_bn_desc d, long long M) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < d.c) {
        const double mean = d.sums[c] / (double)M;
        double var = d.sums[d.c + c] / (double)M - mean * mean;
        if (var < 0.0) var = 0.0;
        d.stats[c] = (float)mean;
        d.stats[d.c + c] = (float)(1.0 / sqrt(var + (double)d.eps));
        if (d.running_mean) {
            const double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
            d.running_mean[c] = (float)((1.0 - d.momentum) * (double)d.running_mean[c] + d.momentum * mean);
            d.running_var[c] = (float)((1.0 - d.momentum) * (double)d.running_var[c] + d.momentum * unb);
        }
    }
    if (c == 0 && d.num_batches_tracked) *d.num_batches_tracked += 1;
}

__device__ __forceinline__ void store_act8(uint16_t* o, long long lo_off, const float* v) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint16_t h0, l0, h1, l1;
        split_bf16(v[2 * q], h0, l0);
        split_bf16(v[2 * q + 1], h1, l1);
        hw[q] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        lw[q] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    *reinterpret_cast<uint4*>(o) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    *reinterpret_cast<uint4*>(o + lo_off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

// y = relu(z * scale + shift), 8 channels per thread
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const disco_bn_desc d, long long M) {
    __shared__ float s_scale[kMaxC], s_shift[kMaxC];
    const int C = d.c;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float sc = d.gamma[c] * d.stats[C + c];
        s_scale[c] = sc;
        s_shift[c] = d.beta[c] - d.stats[c] * sc;
    }
    __syncthreads();
    const int groups = C >> 3;
    const long long total = M * groups;
    uint16_t* out = reinterpret_cast<uint16_t*>(d.out_hi);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long p = e / groups;
        const int c = (int)(e - p * groups) * 8;
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(d.z + p * C + c));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(d.z + p * C + c + 4));
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float y = fmaf(zz[i], s_scale[c + i], s_shift[c + i]);
            v[i] = d.relu ? fmaxf(y, 0.f) : y;
        }
        store_act8(out + p * C + c, d.out_lo_off, v);
    }
}

// dz = gamma*rstd * (g - mean(g) - xhat*mean(g*xhat)), 8 channels per thread; block 0 also writes dgamma/dbeta
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const disco_bn_desc d, long long M) {
    __shared__ float s_mean[kMaxC], s_rstd[kMaxC], s_gam[kMaxC], s_bet[kMaxC], s_mg[kMaxC], s_mgx[kMaxC];
    const int C = d.c;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        s_mean[c] = d.stats[c];
        s_rstd[c] = d.stats[C + c];
        s_gam[c] = d.gamma[c];
        s_bet[c] = d.beta[c];
        s_mg[c] = (float)(d.sums[c] / (double)M);
        s_mgx[c] = (float)(d.sums[C + c] / (double)M);
        if (blockIdx.x == 0) {
            if (d.dbeta) d.dbeta[c] = (float)d.sums[c];
            if (d.dgamma) d.dgamma[c] = (float)d.sums[C + c];
        }
    }
    __syncthreads();
    const int groups = C >> 3;
    const long long total = M * groups;
    uint16_t* out = reinterpret_cast<uint16_t*>(d.dz_hi);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long p = e / groups;
        const int c = (int)(e - p * groups) * 8;
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(d.z + p * C + c));
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(d.z + p * C + c + 4));
        const float zz[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < d.n_g; ++s) grad_src_add8(d.g[s], p, d.h, d.w, c, g);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xh = (zz[i] - s_mean[c + i]) * s_rstd[c + i];
            float gi = g[i];
            if (d.relu && fmaf(xh, s_gam[c + i], s_bet[c + i]) <= 0.f) gi = 0.f;
            v[i] = s_gam[c + i] * s_rstd[c + i] * (gi - s_mg[c + i] - xh * s_mgx[c + i]);
        }
        store_act8(out + p * C + c, d.dz_lo_off, v);
    }
}

__global__ void __launch_bounds__(kThreads) grad_pack_kernel(const float* a, int ca, const float* b, int cb, long long M,
                                                             uint16_t* out, long long lo_off) {
    const int C = ca + cb;
    const int groups = C >> 2;   // 4 channels per thread (ca, cb multiples of 4)
    const long long total = M * groups;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long p = e / groups;
        const int c = (int)(e - p * groups) * 4;
        const float4 v = (c < ca) ? __ldg(reinterpret_cast<const float4*>(a + p * ca + c))
                                  : __ldg(reinterpret_cast<const float4*>(b + p * cb + (c - ca)));
        uint16_t h[4], l[4];
        split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]);
        split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
        uint16_t* o = out + p * C + c;
        *reinterpret_cast<uint2*>(o) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        *reinterpret_cast<uint2*>(o + lo_off) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

__global__ void __launch_bounds__(kThreads) channel_sum_kernel(const float* src, long long M, int C, double* sums) {
    __shared__ float s_a[kThreads];
    const int lanes = kThreads / C;   // C <= 256
    const int t = threadIdx.x;
    const int lane = t / C, c = t - lane * C;
    float a = 0.f;
    if (lane < lanes) {
        const long long stride = (long long)gridDim.x * lanes;
        for (long long p0 = (long long)blockIdx.x * lanes + lane; p0 < M; p0 += stride * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (p0 + u * stride < M) ? __ldg(src + (p0 + u * stride) * C + c) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) a += v[u];
        }
    }
    s_a[t] = a;
    __syncthreads();
    if (t < C) {
        double sa = 0.0;
        for (int l = 0; l < lanes; ++l) sa += (double)s_a[l * C + t];
        atomicAdd(sums + t, sa);
    }
}

__global__ void sums_to_f32_kernel(const double* sums, int n, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)sums[i];
}

// fp32 NCHW -> NHWC through a 32x32 shared-memory tile per (n, channel block, pixel block)
__global__ void nchw_to_nhwc_kernel(const float* src, int C, long long HW, float* dst) {
    __shared__ float tile[32][33];
    const long long n = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    const long long p0 = (long long)blockIdx.x * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long long p = p0 + tx;
        tile[j][tx] = (c < C && p < HW) ? __ldg(src + (n * C + c) * HW + p) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const long long p = p0 + j;
        const int c = c0 + tx;
        if (c < C && p < HW) dst[(n * HW + p) * C + c] = tile[tx][j];
    }
}

__global__ void add_f32_kernel(float* dst, const float* a, const float* b, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = a[i] + b[i];
}

// one thread per (n_tile, k block, tap, 8-channel chunk, n): 8 K-consecutive values, hi and lo planes
__global__ void __launch_bounds__(kThreads) pack_weights_kernel(const disco_pack_desc d, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = d.c_blk >> 3, ncb = d.k_pad / d.c_blk;
    if (idx < (long long)d.n_tiles * d.block_n && d.bias) {
        const int nn = (int)idx;
        d.bias[nn] = (d.bias_src && nn < d.n_real) ? d.bias_src[nn] : 0.f;
    }
    if (idx >= total) return;
    long long r = idx;
    const int n = (int)(r % d.block_n); r /= d.block_n;
    const int chunk = (int)(r % chunks); r /= chunks;
    const int tap = (int)(r % d.taps); r /= d.taps;
    const int cb = (int)(r % ncb);
    const int n_tile = (int)(r / ncb);
    const int nn = n_tile * d.block_n + n;
    const int k_real = d.transpose ? d.co_src : d.ci_src;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int kc = cb * d.c_blk + chunk * 8 + e;
        float x = 0.f;
        if (nn < d.n_real && kc < k_real)
            x = d.transpose ? d.w[((long long)kc * d.ci_src + d.c0 + nn) * d.taps + (d.taps - 1 - tap)]
                            : d.w[((long long)nn * d.ci_src + kc) * d.taps + tap];
        v[e] = x;
    }
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint16_t h0, l0, h1, l1;
        split_bf16(v[2 * q], h0, l0);
        split_bf16(v[2 * q + 1], h1, l1);
        hw[q] = (uint32_t)h0 | ((uint32_t)h1 << 16);
        lw[q] = (uint32_t)l0 | ((uint32_t)l1 << 16);
    }
    // element offsets (in 8-element rows): [n_tile][cb][tap] block of 2*chunks*block_n rows
    const long long blk = (((long long)n_tile * ncb + cb) * d.taps + tap) * (2LL * chunks * d.block_n);
    long long row_hi, row_lo;
    if (d.stacked) {
        row_hi = blk + ((long long)chunk * 2 + 0) * d.block_n + n;
        row_lo = blk + ((long long)chunk * 2 + 1) * d.block_n + n;
    } else {
        row_hi = blk + (long long)chunk * d.block_n + n;
        row_lo = blk + ((long long)chunks + chunk) * d.block_n + n;
    }
    uint4* out = reinterpret_cast<uint4*>(d.wpack);
    out[row_hi] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    out[row_lo] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}

int grid_for(long long work_items, int per_block, int cap_blocks = 148 * 8) {
    long long b = (work_items + per_block - 1) / per_block;
    const long long cap = cap_blocks;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

int check_bn(const disco_bn_desc* d) {
    DISCO_REQUIRE(d && d->z && d->gamma && d->beta && d->sums && d->stats, "bn: null tensor");
    DISCO_REQUIRE(d->c >= 8 && d->c % 8 == 0 && d->c <= kMaxC, "bn: channels %d unsupported", d->c);
    DISCO_REQUIRE(kThreads % (d->c / 8) == 0, "bn: channels/8 = %d must divide %d", d->c / 8, kThreads);
    DISCO_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0, "bn: empty input");
    return DISCO_OK;
}

}  // namespace

int disco_bn_train_forward_launch(const disco_bn_desc* d, void* stream) {
    int rc = check_bn(d);
    if (rc < 0) return rc;
    DISCO_REQUIRE(d->out_hi, "bn forward: null output");
    cudaStream_t s = (cudaStream_t)stream;
    const long long M = (long long)d->n * d->h * d->w;
    DISCO_CHECK_CUDA(cudaMemsetAsync(d->sums, 0, sizeof(double) * 2 * d->c, s));
    const int lanes = kThreads / (d->c / 8);
    bn_reduce_kernel<0><<<grid_for(M, lanes * 4, 148 * 4), kThreads, 0, s>>>(*d, M);
    bn_finalize_kernel<<<(d->c + 127) / 128, 128, 0, s>>>(*d, M);
    bn_apply_kernel<<<grid_for(M * (d->c / 8), kThreads * 4), kThreads, 0, s>>>(*d, M);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_bn_train_backward_launch(const disco_bn_desc* d, void* stream) {
    int rc = check_bn(d);
    if (rc < 0) return rc;
    DISCO_REQUIRE(d->dz_hi && d->n_g >= 1 && d->n_g <= 3, "bn backward: null output / bad source count");
    for (int i = 0; i < d->n_g; ++i) {
        DISCO_REQUIRE(d->g[i].ptr && d->g[i].c_total % 4 == 0 && d->g[i].c_off % 4 == 0, "bn backward: bad gradient source %d", i);
        DISCO_REQUIRE(d->g[i].c_off + d->c <= d->g[i].c_total, "bn backward: gradient source %d too narrow", i);
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long M = (long long)d->n * d->h * d->w;
    DISCO_CHECK_CUDA(cudaMemsetAsync(d->sums, 0, sizeof(double) * 2 * d->c, s));
    const int lanes = kThreads / (d->c / 8);
    bn_reduce_kernel<1><<<grid_for(M, lanes * 4, 148 * 4), kThreads, 0, s>>>(*d, M);
    bn_bwd_apply_kernel<<<grid_for(M * (d->c / 8), kThreads * 4), kThreads, 0, s>>>(*d, M);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_pack_weights_launch(const disco_pack_desc* d, void* stream) {
    DISCO_REQUIRE(d && d->w && d->wpack, "pack_weights: null tensor");
    DISCO_REQUIRE(d->c_blk == 16 || d->c_blk == 32 || d->c_blk == 64, "pack_weights: c_blk must be 16/32/64");
    DISCO_REQUIRE(d->k_pad > 0 && d->k_pad % d->c_blk == 0 && d->block_n % 16 == 0 && d->n_tiles >= 1 && d->taps >= 1,
                  "pack_weights: bad geometry");
    DISCO_REQUIRE(d->n_real <= d->n_tiles * d->block_n, "pack_weights: n_real exceeds the padded rows");
    DISCO_REQUIRE(!d->transpose || d->c0 + d->n_real <= d->ci_src, "pack_weights: slice out of range");
    const long long total = (long long)d->n_tiles * (d->k_pad / d->c_blk) * d->taps * (d->c_blk / 8) * d->block_n;
    const long long blocks = (total + kThreads - 1) / kThreads;
    DISCO_REQUIRE(blocks < (1ll << 31), "pack_weights: too large");
    pack_weights_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(*d, total);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_grad_pack_launch(const float* a, int ca, const float* b, int cb, long long n_pix, void* out_hi,
                           long long out_lo_off, void* stream) {
    DISCO_REQUIRE(a && out_hi && ca > 0 && ca % 4 == 0 && cb % 4 == 0 && (cb == 0 || b) && n_pix > 0, "grad_pack: bad arguments");
    grad_pack_kernel<<<grid_for(n_pix * ((ca + cb) / 4), kThreads * 4), kThreads, 0, (cudaStream_t)stream>>>(
        a, ca, b, cb, n_pix, reinterpret_cast<uint16_t*>(out_hi), out_lo_off);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_channel_sum_launch(const float* src, long long n_pix, int c, double* sums, float* out, void* stream) {
    DISCO_REQUIRE(src && sums && out && c > 0 && c <= kThreads && n_pix > 0, "channel_sum: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    DISCO_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * c, s));
    channel_sum_kernel<<<grid_for(n_pix, (kThreads / c) * 16, 148 * 4), kThreads, 0, s>>>(src, n_pix, c, sums);
    sums_to_f32_kernel<<<(c + 127) / 128, 128, 0, s>>>(sums, c, out);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_nchw_to_nhwc_launch(const float* src, int n, int c, int h, int w, float* dst, void* stream) {
    DISCO_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, "nchw_to_nhwc: bad arguments");
    const long long HW = (long long)h * w;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((c + 31) / 32), (unsigned)n);
    nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, c, HW, dst);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_add_f32_launch(float* dst, const float* a, const float* b, long long n, void* stream) {
    DISCO_REQUIRE(dst && a && b && n > 0, "add_f32: bad arguments");
    add_f32_kernel<<<grid_for(n, kThreads * 4), kThreads, 0, (cudaStream_t)stream>>>(dst, a, b, n);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// KD loss term: nn.KLDivLoss(mean over elements)(log_softmax(student, C), softmax(teacher, C)) of
// two NCHW fp32 maps, forward value and gradient wrt the student in ONE pass pair (FaFModule.get_kd_loss,
// CoDetModule.py:334-382, does permute + reshape + log_softmax + softmax + kl_div per map and autograd walks them
// back: ~15 full passes over the maps; this is 4 reads + 1 write per element).
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kKdLanes = kThreads / 32;   // channel lanes per pixel (one warp per lane: loads stay coalesced along pixels)

// Block = 32 consecutive pixels x 8 channel lanes; lane l owns channels l, l+8, ...  Two passes over the block's
// [32 px, C] slab: online max / sum-of-exponentials per lane combined through shared memory, then the KL terms and the
// gradient.  (One thread per pixel was latency-bound on the 32x32 maps: 20 K threads walking 256 channels each.)
__global__ void __launch_bounds__(kThreads) kd_kl_kernel(const float* __restrict__ s, const float* __restrict__ t, int C, long long hw,
                                                         long long total, double* loss_sum, float* __restrict__ grad, float gscale) {
    __shared__ float s_ms[kKdLanes][32], s_zs[kKdLanes][32], s_mt[kKdLanes][32], s_zt[kKdLanes][32];
    __shared__ double s_part[kKdLanes];
    const int px = threadIdx.x & 31, lane = threadIdx.x >> 5;
    double local = 0.0;
    for (long long e0 = (long long)blockIdx.x * 32; e0 < total; e0 += (long long)gridDim.x * 32) {
        const long long e = e0 + px;
        const bool ok = e < total;
        const long long img = ok ? e / hw : 0, p = ok ? e - img * hw : 0;
        const float* sp = s + img * C * hw + p;
        const float* tp = t + img * C * hw + p;
        float ms = -INFINITY, mt = -INFINITY, zs = 0.f, zt = 0.f;
        if (ok) {
            for (int c = lane; c < C; c += kKdLanes) {
                const float a = __ldg(sp + (long long)c * hw), b = __ldg(tp + (long long)c * hw);
                if (a > ms) { zs *= expf(ms - a); ms = a; }
                zs += expf(a - ms);
                if (b > mt) { zt *= expf(mt - b); mt = b; }
                zt += expf(b - mt);
            }
        }
        s_ms[lane][px] = ms; s_zs[lane][px] = zs; s_mt[lane][px] = mt; s_zt[lane][px] = zt;
        __syncthreads();
        float Ms = -INFINITY, Mt = -INFINITY;
#pragma unroll
        for (int l = 0; l < kKdLanes; ++l) { Ms = fmaxf(Ms, s_ms[l][px]); Mt = fmaxf(Mt, s_mt[l][px]); }
        float Zs = 0.f, Zt = 0.f;
#pragma unroll
        for (int l = 0; l < kKdLanes; ++l) {
            if (s_zs[l][px] > 0.f) Zs += s_zs[l][px] * expf(s_ms[l][px] - Ms);
            if (s_zt[l][px] > 0.f) Zt += s_zt[l][px] * expf(s_mt[l][px] - Mt);
        }
        __syncthreads();   // shared statistics are rewritten by the next slab
        if (ok) {
            const float lzs = logf(Zs), lzt = logf(Zt), izs = 1.f / Zs, izt = 1.f / Zt;
            float acc = 0.f;
            for (int c = lane; c < C; c += kKdLanes) {
                const float a = __ldg(sp + (long long)c * hw) - Ms, b = __ldg(tp + (long long)c * hw) - Mt;
                const float pt = expf(b) * izt;
                if (pt > 0.f) acc = fmaf(pt, (b - lzt) - (a - lzs), acc);
                if (grad) grad[img * C * hw + (long long)c * hw + p] = (expf(a) * izs - pt) * gscale;
            }
            local += (double)acc;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (px == 0) s_part[lane] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < kKdLanes; ++i) a += s_part[i];
        atomicAdd(loss_sum, a);
    }
}

}  // namespace

int disco_kd_kl_launch(const float* student, const float* teacher, int n, int c, long long hw, double* loss_sum, float* grad,
                       float grad_scale, void* stream) {
    DISCO_REQUIRE(student && teacher && loss_sum && n > 0 && c > 0 && hw > 0, "kd_kl: bad arguments");
    const long long total = (long long)n * hw;
    kd_kl_kernel<<<grid_for(total, 32, 148 * 16), kThreads, 0, (cudaStream_t)stream>>>(student, teacher, c, hw, total, loss_sum,
                                                                                           grad, grad_scale);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Softmax focal classification loss (SoftmaxFocalClassificationLoss._compute_loss, utils/loss.py:
// 322-394, with _softmax_cross_entropy_with_logits :213-219), per anchor with K <= 8 classes:
//   idx = argmax(target), ce = -log_softmax(logits)[idx], p = softmax(logits)
//   p_t[c] = t[c] p[c] + (1 - t[c]) (1 - p[c]),   alpha_w = (t[0] == 1) ? 1 - alpha : alpha
//   loss[c] = (1 - p_t[c])^gamma * alpha_w * ce * t[c]
// forward writes loss[anchor][c]; backward writes dlogits[anchor][j] = sum_c gout[c] * dloss[c]/dlogit[j].
// One thread per anchor, 2K loads + K stores (the reference chain is ~15 elementwise ops + CrossEntropy + autograd).
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxCls = 8;

template <bool BWD>
__global__ void __launch_bounds__(kThreads) focal_kernel(const float* __restrict__ logits, const float* __restrict__ target, int K,
                                                         long long n_anchor, float gamma, float alpha, int use_alpha,
                                                         const float* __restrict__ gout, long long gout_stride, float* __restrict__ out) {
    for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < n_anchor; a += (long long)gridDim.x * blockDim.x) {
        float z[kMaxCls], t[kMaxCls], p[kMaxCls];
        float m = -INFINITY;
        int idx = 0;
        float tmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c) {
            if (c >= K) break;
            z[c] = __ldg(logits + a * K + c);
            t[c] = __ldg(target + a * K + c);
            m = fmaxf(m, z[c]);
            if (t[c] > tmax) { tmax = t[c]; idx = c; }   // first maximum, like torch.max(dim)[1]
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c) {
            if (c >= K) break;
            p[c] = expf(z[c] - m);
            s += p[c];
        }
        const float inv = 1.f / s, ls = logf(s);
        float zi = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxCls; ++c) {
            if (c >= K) break;
            p[c] *= inv;
            if (c == idx) zi = z[c];
        }
        const float ce = -(zi - m - ls);
        const float aw = use_alpha ? ((t[0] == 1.f) ? 1.f - alpha : alpha) : 1.f;
        if (!BWD) {
#pragma unroll
            for (int c = 0; c < kMaxCls; ++c) {
                if (c >= K) break;
                const float pt = t[c] * p[c] + (1.f - t[c]) * (1.f - p[c]);
                const float mod = (gamma != 0.f) ? powf(fmaxf(1.f - pt, 0.f), gamma) : 1.f;
                out[a * K + c] = mod * aw * ce * t[c];
            }
        } else {
            float dz[kMaxCls];
#pragma unroll
            for (int j = 0; j < kMaxCls; ++j) dz[j] = 0.f;
#pragma unroll
            for (int c = 0; c < kMaxCls; ++c) {
                if (c >= K) break;
                const float g = gout[(a * K + c) * gout_stride] * aw * t[c];
                if (g == 0.f) continue;
                const float pt = t[c] * p[c] + (1.f - t[c]) * (1.f - p[c]);
                const float om = fmaxf(1.f - pt, 0.f);
                const float mod = (gamma != 0.f) ? powf(om, gamma) : 1.f;
                // d mod / d p_t = -gamma (1-p_t)^(gamma-1);  d p_t / d z_j = (2 t_c - 1) p_c (delta_cj - p_j)
                const float dmod = (gamma != 0.f) ? -gamma * powf(om, gamma - 1.f) * (2.f * t[c] - 1.f) * p[c] : 0.f;
#pragma unroll
                for (int j = 0; j < kMaxCls; ++j) {
                    if (j >= K) break;
                    const float dpt = dmod * ((c == j ? 1.f : 0.f) - p[j]);
                    const float dce = p[j] - (j == idx ? 1.f : 0.f);
                    dz[j] += g * (dpt * ce + mod * dce);
                }
            }
#pragma unroll
            for (int j = 0; j < kMaxCls; ++j) {
                if (j >= K) break;
                out[a * K + j] = dz[j];
            }
        }
    }
}

}  // namespace

int disco_focal_loss_launch(const float* logits, const float* target, int k, long long n_anchor, float gamma, float alpha,
                            int use_alpha, const float* grad_out, long long grad_out_stride, float* out, void* stream) {
    DISCO_REQUIRE(logits && target && out && k >= 1 && k <= kMaxCls && n_anchor > 0, "focal_loss: bad arguments (1 <= classes <= 8)");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = grid_for(n_anchor, kThreads, 148 * 16);
    if (grad_out) focal_kernel<true><<<grid, kThreads, 0, s>>>(logits, target, k, n_anchor, gamma, alpha, use_alpha, grad_out, grad_out_stride, out);
    else focal_kernel<false><<<grid, kThreads, 0, s>>>(logits, target, k, n_anchor, gamma, alpha, use_alpha, nullptr, 0, out);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
