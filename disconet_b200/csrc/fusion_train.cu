// DiscoGraph fusion block in TRAINING mode: PixelWeightedFusionSoftmax with per-call batch statistics, and the
// backward of the whole block.
//
// Reference semantics (DiscoNet.py:59-113,132-155 in train() mode): the PWF MLP is called once per (scene b, ego i,
// k-th neighbour) on a [1, 2C, h, w] tensor, so each of its three BatchNorms normalises with the statistics of
// THAT call's h*w pixels and updates its running statistics once per call, sequentially in (b, i, k) order.
//
//   pwf_fwd_pass_kernel<P> four launches (statistics of layer 1 -> 2 -> 3 -> output map), CTA = (128-pixel chunk,
//                          pair (b, i, j)); every pass recomputes the cheap chain from the tensor-core product `en`
//                          and adds its partial sums to the pair's double-precision sums with atomics
//   pwf_running_kernel     sequential EMA of the per-pair statistics in the reference's call order
//   (softmax over k + weighted sum: the eval fusion kernel reading the precomputed maps, fusion.cu `wpre`)
//   fusion_combine_bwd     one warp per (ego, cell): softmax / weighted-sum backward, bilinear-warp transpose
//                          (scatter-add into the feature gradient), gradient wrt the PWF output maps
//   pwf_bwd_pass_kernel<P> BatchNorm(train) backward needs per-pair sums of each layer's gradient, hence four
//                          recompute launches (layer 3 -> 2 -> 1 -> input), same CTA grid; weight gradients of the
//                          128->32->8->1 tail accumulate in registers / shared memory, then one atomic per CTA;
//                          the gradient wrt `en` (ego half direct, neighbour half through the warp transpose) goes
//                          back to the tensor cores (conv1_1 data + weight gradient).
// As in the eval kernel the conv1_1 neighbour half commutes with the bilinear warp and the H flips are folded
// into the row index.
#include "common.cuh"
#include "conv.h"
#include "train.h"

namespace {

constexpr int kHid = 128, kH2 = 32, kH3 = 8;
constexpr int kStatC = kHid + kH2 + kH3;   // 168 BatchNorm channels per pair
// dparams layout
constexpr int kDG1 = 0, kDBE1 = 128, kDW2 = 256, kDG2 = 4352, kDBE2 = 4384, kDW3 = 4416, kDG3 = 4672, kDBE3 = 4680,
              kDW4 = 4688;   // dW4[8] followed by db4 (4696); 4697 floats in total

struct Taps {
    int n;
    long long row[4];
    float w[4];
};

// Sampling taps of neighbour j's map for ego i at output cell (y, x) (unflipped frame); j == i: the cell itself.
__device__ __forceinline__ Taps pair_taps(const disco_pwf_train_desc& d, int b, int i, int j, int y, int x) {
    Taps t;
    const int h = d.h, w = d.w, B = d.B, A = d.A;
    if (j == i) {
        t.n = 1;
        t.row[0] = ((long long)(i * B + b) * h + y) * w + x;
        t.w[0] = 1.f;
        return t;
    }
    const int yf = h - 1 - y;
    const float xb = (2.f * x + 1.f) / w - 1.f;
    const float yb = (2.f * yf + 1.f) / h - 1.f;
    const double* T = d.trans + (((long long)b * A + j) * A + i) * 16;
    const float m00 = (float)T[0], m01 = (float)T[1], m02 = (-(float)T[3]) * d.trans_scale;
    const float m10 = (float)T[4], m11 = (float)T[5], m12 = (-(float)T[7]) * d.trans_scale;
    const float gx = m00 * xb + m01 * yb + m02;
    const float gy = m10 * xb + m11 * yb + m12;
    const float ix = ((gx + 1.f) * w - 1.f) * 0.5f;
    const float iy = ((gy + 1.f) * h - 1.f) * 0.5f;
    const float fx = floorf(ix), fy = floorf(iy);
    const float ax = ix - fx, ay = iy - fy;
    const bool finite = (fabsf(ix) < 1e6f) && (fabsf(iy) < 1e6f);
    const int x0 = finite ? (int)fx : -10, y0 = finite ? (int)fy : -10;
    const float tw[4] = {(1.f - ax) * (1.f - ay), ax * (1.f - ay), (1.f - ax) * ay, ax * ay};
    t.n = 0;
#pragma unroll
    for (int tp = 0; tp < 4; ++tp) {
        const int xs = x0 + (tp & 1), ysf = y0 + (tp >> 1);
        if (xs < 0 || xs >= w || ysf < 0 || ysf >= h) continue;   // zeros padding
        t.row[t.n] = ((long long)(j * B + b) * h + (h - 1 - ysf)) * w + xs;
        t.w[t.n] = tw[tp];
        ++t.n;
    }
    return t;
}

__device__ __forceinline__ bool pair_used(const disco_pwf_train_desc& d, int b, int i, int j) {
    const int n_ag = d.num_agent[b];
    if (i >= n_ag || j >= n_ag) return false;
    if (d.outage && d.outage[b * d.A + i]) return false;
    if (j != i && d.only_v2i && i != 0 && j != 0) return false;
    return true;
}

// z1 (pre-BN output of conv1_1 on cat[ego, warped neighbour]) for this lane's 4 channels 4*lane..4*lane+3
__device__ __forceinline__ float4 pair_z1(const disco_pwf_train_desc& d, long long row_i, const Taps& t, int lane) {
    float4 z = __ldg(reinterpret_cast<const float4*>(d.en + row_i * (2 * kHid)) + lane);
    for (int k = 0; k < t.n; ++k) {
        const float4 n4 = __ldg(reinterpret_cast<const float4*>(d.en + t.row[k] * (2 * kHid) + kHid) + lane);
        z.x = fmaf(t.w[k], n4.x, z.x); z.y = fmaf(t.w[k], n4.y, z.y);
        z.z = fmaf(t.w[k], n4.z, z.z); z.w = fmaf(t.w[k], n4.w, z.w);
    }
    return z;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------------------
// PWF forward / backward: one kernel launch per pass, CTA = (pixel chunk, pair)
// ------------------------------------------------------------------------------------------------------------
constexpr int kPW = 8;            // warps per CTA
constexpr int kChunkPix = 128;    // pixels per CTA (16 per warp)

struct PairSmem {
    float w2[kH2][kHid + 1];
    float w3[kH3][kH2];
    float b2[kH2], b3[kH3], w4[kH3], b4;
    float gam[kStatC], bet[kStatC];       // BN gamma/beta of the three layers, concatenated 128 | 32 | 8
    float mu[kStatC], rs[kStatC];         // per-pair batch mean / rstd
    float mg[kStatC], mgx[kStatC];        // backward: per-pair mean(g), mean(g * xhat)
    alignas(16) float h1[kPW][kHid];      // a1
    alignas(16) float xh1[kPW][kHid];     // xhat1
    alignas(16) float dz1[kPW][kHid];
    float dz2[kPW][kH2];
    float red[kPW][2 * kHid];             // cross-warp scratch
    float dw2[kH2 * kHid];                // CTA accumulators of the tail weight gradients
    float dw3[kH3 * kH2];
    float dw4[kH3 + 1];                   // dW4[8], db4
};

// parameters + the per-pair statistics of the layers already normalised (from the double sums of earlier passes)
__device__ void pair_prologue(PairSmem& s, const disco_pwf_train_desc& d, int pair, int n_stat_layers, int n_grad_layers) {
    const int tid = threadIdx.x, HW = d.h * d.w;
    for (int e = tid; e < kH2 * kHid; e += blockDim.x) { s.w2[e / kHid][e % kHid] = d.w2[e]; s.dw2[e] = 0.f; }
    for (int e = tid; e < kH3 * kH2; e += blockDim.x) { s.w3[e / kH2][e % kH2] = d.w3[e]; s.dw3[e] = 0.f; }
    if (tid < kH2) s.b2[tid] = d.b2[tid];
    if (tid < kH3) { s.b3[tid] = d.b3[tid]; s.w4[tid] = d.w4[tid]; }
    if (tid <= kH3) s.dw4[tid] = 0.f;
    if (tid == 0) s.b4 = d.b4[0];
    if (tid < kHid) { s.gam[tid] = d.g1[tid]; s.bet[tid] = d.be1[tid]; }
    if (tid < kH2) { s.gam[kHid + tid] = d.g2[tid]; s.bet[kHid + tid] = d.be2[tid]; }
    if (tid < kH3) { s.gam[kHid + kH2 + tid] = d.g3[tid]; s.bet[kHid + kH2 + tid] = d.be3[tid]; }
    if (tid < kStatC) {
        const int layer = tid < kHid ? 0 : (tid < kHid + kH2 ? 1 : 2);
        float mu = 0.f, rs = 1.f;
        if (layer < n_stat_layers) {
            const double* ps = d.psum + (long long)pair * 2 * kStatC;
            const double mean = ps[tid] / HW;
            double var = ps[kStatC + tid] / HW - mean * mean;
            if (var < 0.0) var = 0.0;
            mu = (float)mean;
            rs = (float)(1.0 / sqrt(var + (double)d.eps));
        }
        s.mu[tid] = mu; s.rs[tid] = rs;
        float mg = 0.f, mgx = 0.f;
        if (2 - layer < n_grad_layers) {       // gradient sums exist for layers 3 (after pass 0), 2, 1
            const float* gs = d.gsum + (long long)pair * 2 * kStatC;
            mg = gs[tid] / HW; mgx = gs[kStatC + tid] / HW;
        }
        s.mg[tid] = mg; s.mgx[tid] = mgx;
    }
    __syncthreads();
}

// PASS 0/1/2: accumulate the sum / sum of squares of layer 1/2/3's pre-BN activations; PASS 3: write the output map
template <int PASS>
__global__ void __launch_bounds__(kPW * 32) pwf_fwd_pass_kernel(const disco_pwf_train_desc d) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PairSmem& s = *reinterpret_cast<PairSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = d.A, h = d.h, w = d.w, HW = h * w;
    const int pair = blockIdx.y;
    const int j = pair % A, i = (pair / A) % A, b = pair / (A * A);
    const int p_begin = blockIdx.x * kChunkPix, p_end = min(HW, p_begin + kChunkPix);
    float* wl = d.wlogit + (long long)pair * HW;
    if (!pair_used(d, b, i, j)) {
        if (PASS == 3) for (int p = p_begin + tid; p < p_end; p += blockDim.x) wl[p] = 0.f;
        return;
    }
    pair_prologue(s, d, pair, PASS, 0);
    const int n_i = i * d.B + b;
    float sa[8], sq[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { sa[q] = 0.f; sq[q] = 0.f; }
    for (int p = p_begin + warp; p < p_end; p += kPW) {
        const int y = p / w, x = p - y * w;
        const long long row_i = ((long long)n_i * h + y) * w + x;
        const Taps t = pair_taps(d, b, i, j, y, x);
        const float4 z4 = pair_z1(d, row_i, t, lane);
        const float z1[4] = {z4.x, z4.y, z4.z, z4.w};
        if (PASS == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { sa[q] += z1[q]; sq[q] = fmaf(z1[q], z1[q], sq[q]); }
            continue;
        }
        float a1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 4 * lane + q;
            a1[q] = fmaxf(fmaf((z1[q] - s.mu[c]) * s.rs[c], s.gam[c], s.bet[c]), 0.f);
        }
        __syncwarp();
        reinterpret_cast<float4*>(s.h1[warp])[lane] = make_float4(a1[0], a1[1], a1[2], a1[3]);
        __syncwarp();
        float z2;
        {   // four independent accumulators: the 128-long FMA chain is otherwise pure latency
            float q0 = s.b2[lane], q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 8
            for (int c = 0; c < kHid; c += 4) {
                const float4 hv = *reinterpret_cast<const float4*>(&s.h1[warp][c]);
                q0 = fmaf(s.w2[lane][c], hv.x, q0);
                q1 = fmaf(s.w2[lane][c + 1], hv.y, q1);
                q2 = fmaf(s.w2[lane][c + 2], hv.z, q2);
                q3 = fmaf(s.w2[lane][c + 3], hv.w, q3);
            }
            z2 = (q0 + q1) + (q2 + q3);
        }
        if (PASS == 1) {
            sa[0] += z2; sq[0] = fmaf(z2, z2, sq[0]);
            continue;
        }
        const int c2 = kHid + lane;
        const float a2 = fmaxf(fmaf((z2 - s.mu[c2]) * s.rs[c2], s.gam[c2], s.bet[c2]), 0.f);
        float z3[kH3];
#pragma unroll
        for (int q = 0; q < kH3; ++q) z3[q] = warp_sum(s.w3[q][lane] * a2) + s.b3[q];
        if (PASS == 2) {
#pragma unroll
            for (int q = 0; q < kH3; ++q) { sa[q] += z3[q]; sq[q] = fmaf(z3[q], z3[q], sq[q]); }
            continue;
        }
        float z4o = s.b4;
#pragma unroll
        for (int q = 0; q < kH3; ++q) {
            const int c3 = kHid + kH2 + q;
            z4o = fmaf(s.w4[q], fmaxf(fmaf((z3[q] - s.mu[c3]) * s.rs[c3], s.gam[c3], s.bet[c3]), 0.f), z4o);
        }
        if (lane == 0) wl[p] = fmaxf(z4o, 0.f);
    }
    if (PASS == 3) return;
    // ---- CTA reduction -> double atomics on the pair's sums ------------------------------------------
    int c_base, nch;
    if (PASS == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { s.red[warp][4 * lane + q] = sa[q]; s.red[warp][kHid + 4 * lane + q] = sq[q]; }
        c_base = 0; nch = kHid;
    } else if (PASS == 1) {
        s.red[warp][lane] = sa[0]; s.red[warp][kHid + lane] = sq[0];
        c_base = kHid; nch = kH2;
    } else {
        if (lane < kH3) {
            float va = 0.f, vq = 0.f;
#pragma unroll
            for (int q = 0; q < kH3; ++q) if (lane == q) { va = sa[q]; vq = sq[q]; }
            s.red[warp][lane] = va; s.red[warp][kHid + lane] = vq;
        }
        c_base = kHid + kH2; nch = kH3;
    }
    __syncthreads();
    if (tid < nch) {
        double a = 0.0, q2 = 0.0;
        for (int wp = 0; wp < kPW; ++wp) { a += (double)s.red[wp][tid]; q2 += (double)s.red[wp][kHid + tid]; }
        double* ps = d.psum + (long long)pair * 2 * kStatC;
        atomicAdd(ps + c_base + tid, a);
        atomicAdd(ps + kStatC + c_base + tid, q2);
    }
}

// Sequential running-statistics update in the reference's call order (b, ego i, neighbour list [i, j != i ...])
__global__ void pwf_running_kernel(const disco_pwf_train_desc d) {
    const int c = threadIdx.x;
    if (c >= kStatC) return;
    float* rm; float* rv;
    int cl;
    if (c < kHid) { rm = d.rm1; rv = d.rv1; cl = c; }
    else if (c < kHid + kH2) { rm = d.rm2; rv = d.rv2; cl = c - kHid; }
    else { rm = d.rm3; rv = d.rv3; cl = c - kHid - kH2; }
    const int A = d.A, HW = d.h * d.w;
    const double unb = HW > 1 ? (double)HW / (double)(HW - 1) : 1.0;
    float m = rm[cl], v = rv[cl];
    long long calls = 0;
    for (int b = 0; b < d.B; ++b)
        for (int i = 0; i < A; ++i)
            for (int k = 0; k < A; ++k) {
                const int j = (k == 0) ? i : ((k - 1 < i) ? k - 1 : k);
                if (!pair_used(d, b, i, j)) continue;
                const double* ps = d.psum + ((long long)(b * A + i) * A + j) * 2 * kStatC;
                const double mean = ps[c] / HW;
                double var = ps[kStatC + c] / HW - mean * mean;
                if (var < 0.0) var = 0.0;
                m = (1.f - d.momentum) * m + d.momentum * (float)mean;
                v = (1.f - d.momentum) * v + d.momentum * (float)(var * unb);
                ++calls;
            }
    rm[cl] = m;
    rv[cl] = v;
    if (c == 0) { *d.nbt1 += calls; *d.nbt2 += calls; *d.nbt3 += calls; }
}

// PASS 0: sums of layer 3's gradient (+ dW4, db4); 1: layer 2 (+ dW3); 2: layer 1 (+ dW2); 3: gradient wrt `en`
template <int PASS>
__global__ void __launch_bounds__(kPW * 32, PASS == 2 ? 1 : 2) pwf_bwd_pass_kernel(const disco_pwf_train_desc d) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    PairSmem& s = *reinterpret_cast<PairSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = d.A, h = d.h, w = d.w, HW = h * w;
    const int pair = blockIdx.y;
    const int j = pair % A, i = (pair / A) % A, b = pair / (A * A);
    if (!pair_used(d, b, i, j)) return;
    const int p_begin = blockIdx.x * kChunkPix, p_end = min(HW, p_begin + kChunkPix);
    pair_prologue(s, d, pair, 3, PASS);
    const int n_i = i * d.B + b;
    const float* dwl = d.dwlogit + (long long)pair * HW;

    float s_g[8], s_gx[8];          // pass 0: layer 3 (8 ch, uniform); pass 1: [0] layer 2 (ch = lane); pass 2: [0..3] layer 1 (ch = lane + 32q)
#pragma unroll
    for (int q = 0; q < 8; ++q) { s_g[q] = 0.f; s_gx[q] = 0.f; }
    float acc_w4[kH3 + 1];          // pass 0: dW4, db4
#pragma unroll
    for (int q = 0; q <= kH3; ++q) acc_w4[q] = 0.f;
    float acc_w3[kH3];              // pass 1: dW3[q][lane]
#pragma unroll
    for (int q = 0; q < kH3; ++q) acc_w3[q] = 0.f;
    float acc_w2[(PASS == 2) ? kH2 : 1][4];   // pass 2: dW2[o][lane + 32 q]
    if (PASS == 2) {
#pragma unroll
        for (int o = 0; o < kH2; ++o)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc_w2[o][q] = 0.f;
    }

    for (int p = p_begin + warp; p < p_end; p += kPW) {
        const int y = p / w, x = p - y * w;
        const long long row_i = ((long long)n_i * h + y) * w + x;
        const Taps t = pair_taps(d, b, i, j, y, x);
        // ---- forward recompute ----------------------------------------------------------------
        const float4 z4v = pair_z1(d, row_i, t, lane);
        const float z1[4] = {z4v.x, z4v.y, z4v.z, z4v.w};
        float a1[4], xh1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 4 * lane + q;
            xh1[q] = (z1[q] - s.mu[c]) * s.rs[c];
            a1[q] = fmaxf(fmaf(xh1[q], s.gam[c], s.bet[c]), 0.f);
        }
        __syncwarp();
        reinterpret_cast<float4*>(s.h1[warp])[lane] = make_float4(a1[0], a1[1], a1[2], a1[3]);
        reinterpret_cast<float4*>(s.xh1[warp])[lane] = make_float4(xh1[0], xh1[1], xh1[2], xh1[3]);
        __syncwarp();
        float z2;
        {   // four independent accumulators: the 128-long FMA chain is otherwise pure latency
            float q0 = s.b2[lane], q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 8
            for (int c = 0; c < kHid; c += 4) {
                const float4 hv = *reinterpret_cast<const float4*>(&s.h1[warp][c]);
                q0 = fmaf(s.w2[lane][c], hv.x, q0);
                q1 = fmaf(s.w2[lane][c + 1], hv.y, q1);
                q2 = fmaf(s.w2[lane][c + 2], hv.z, q2);
                q3 = fmaf(s.w2[lane][c + 3], hv.w, q3);
            }
            z2 = (q0 + q1) + (q2 + q3);
        }
        const int c2 = kHid + lane;
        const float xh2 = (z2 - s.mu[c2]) * s.rs[c2];
        const float y2 = fmaf(xh2, s.gam[c2], s.bet[c2]);
        const float a2 = fmaxf(y2, 0.f);
        float xh3[kH3], a3[kH3];
        bool on3[kH3];
        float z4o = s.b4;
#pragma unroll
        for (int q = 0; q < kH3; ++q) {
            const int c3 = kHid + kH2 + q;
            const float z3 = warp_sum(s.w3[q][lane] * a2) + s.b3[q];
            xh3[q] = (z3 - s.mu[c3]) * s.rs[c3];
            const float y3 = fmaf(xh3[q], s.gam[c3], s.bet[c3]);
            on3[q] = y3 > 0.f;
            a3[q] = fmaxf(y3, 0.f);
            z4o = fmaf(s.w4[q], a3[q], z4o);
        }
        // ---- backward chain -------------------------------------------------------------------
        const float dz4 = (z4o > 0.f) ? __ldg(dwl + p) : 0.f;
        float g3[kH3];
#pragma unroll
        for (int q = 0; q < kH3; ++q) g3[q] = on3[q] ? s.w4[q] * dz4 : 0.f;
        if (PASS == 0) {
#pragma unroll
            for (int q = 0; q < kH3; ++q) {
                acc_w4[q] = fmaf(dz4, a3[q], acc_w4[q]);
                s_g[q] += g3[q];
                s_gx[q] = fmaf(g3[q], xh3[q], s_gx[q]);
            }
            acc_w4[kH3] += dz4;
            continue;
        }
        float da2 = 0.f;
#pragma unroll
        for (int q = 0; q < kH3; ++q) {
            const int c3 = kHid + kH2 + q;
            const float dz3 = s.gam[c3] * s.rs[c3] * (g3[q] - s.mg[c3] - xh3[q] * s.mgx[c3]);
            if (PASS == 1) acc_w3[q] = fmaf(dz3, a2, acc_w3[q]);
            da2 = fmaf(s.w3[q][lane], dz3, da2);
        }
        const float g2 = (y2 > 0.f) ? da2 : 0.f;
        if (PASS == 1) {
            s_g[0] += g2;
            s_gx[0] = fmaf(g2, xh2, s_gx[0]);
            continue;
        }
        const float dz2 = s.gam[c2] * s.rs[c2] * (g2 - s.mg[c2] - xh2 * s.mgx[c2]);
        __syncwarp();
        s.dz2[warp][lane] = dz2;
        __syncwarp();
        // channels of layer 1 owned in this part: c = lane + 32 q  (conflict-free reads of w2 rows / h1 / xh1)
        float a1c[4], xh1c[4], da1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) { a1c[q] = s.h1[warp][lane + 32 * q]; xh1c[q] = s.xh1[warp][lane + 32 * q]; }
#pragma unroll
        for (int o = 0; o < kH2; ++o) {
            const float dzo = s.dz2[warp][o];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (PASS == 2) acc_w2[o][q] = fmaf(dzo, a1c[q], acc_w2[o][q]);
                da1[q] = fmaf(s.w2[o][lane + 32 * q], dzo, da1[q]);
            }
        }
        float g1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) g1[q] = (a1c[q] > 0.f) ? da1[q] : 0.f;
        if (PASS == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) { s_g[q] += g1[q]; s_gx[q] = fmaf(g1[q], xh1c[q], s_gx[q]); }
            continue;
        }
        // pass 3: dz1 -> gradient wrt en (ego half direct, neighbour half through the warp transpose)
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = lane + 32 * q;
            s.dz1[warp][c] = s.gam[c] * s.rs[c] * (g1[q] - s.mg[c] - xh1c[q] * s.mgx[c]);
        }
        __syncwarp();
        const float4 dz = reinterpret_cast<const float4*>(s.dz1[warp])[lane];
        atomicAdd(reinterpret_cast<float4*>(d.den + row_i * (2 * kHid)) + lane, dz);
        for (int k = 0; k < t.n; ++k)
            atomicAdd(reinterpret_cast<float4*>(d.den + t.row[k] * (2 * kHid) + kHid) + lane,
                      make_float4(t.w[k] * dz.x, t.w[k] * dz.y, t.w[k] * dz.z, t.w[k] * dz.w));
    }
    if (PASS == 3) return;
    // ---- CTA reduction: gradient sums of this pass's layer -> gsum[pair]; weight gradients -> dparams ------
    int c_base, nch;
    if (PASS == 0) {
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < kH3; ++q) { s.red[warp][q] = s_g[q]; s.red[warp][kHid + q] = s_gx[q]; }
#pragma unroll
            for (int q = 0; q <= kH3; ++q) atomicAdd(&s.dw4[q], acc_w4[q]);
        }
        c_base = kHid + kH2; nch = kH3;
    } else if (PASS == 1) {
        s.red[warp][lane] = s_g[0]; s.red[warp][kHid + lane] = s_gx[0];
#pragma unroll
        for (int q = 0; q < kH3; ++q) atomicAdd(&s.dw3[q * kH2 + lane], acc_w3[q]);
        c_base = kHid; nch = kH2;
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) { s.red[warp][lane + 32 * q] = s_g[q]; s.red[warp][kHid + lane + 32 * q] = s_gx[q]; }
        if (PASS == 2) {
#pragma unroll
            for (int o = 0; o < kH2; ++o)
#pragma unroll
                for (int q = 0; q < 4; ++q) atomicAdd(&s.dw2[o * kHid + lane + 32 * q], acc_w2[o][q]);
        }
        c_base = 0; nch = kHid;
    }
    __syncthreads();
    if (tid < nch) {
        float a = 0.f, bx = 0.f;
        for (int wp = 0; wp < kPW; ++wp) { a += s.red[wp][tid]; bx += s.red[wp][kHid + tid]; }
        float* gs = d.gsum + (long long)pair * 2 * kStatC;
        atomicAdd(gs + c_base + tid, a);
        atomicAdd(gs + kStatC + c_base + tid, bx);
    }
    if (PASS == 0) { if (tid <= kH3) atomicAdd(d.dparams + kDW4 + tid, s.dw4[tid]); }
    else if (PASS == 1) { for (int e = tid; e < kH3 * kH2; e += blockDim.x) atomicAdd(d.dparams + kDW3 + e, s.dw3[e]); }
    else { for (int e = tid; e < kH2 * kHid; e += blockDim.x) atomicAdd(d.dparams + kDW2 + e, s.dw2[e]); }
}

// dgamma / dbeta of the three PWF BatchNorms = sums over the used pairs of the per-pair gradient sums
__global__ void pwf_bwd_finish_kernel(const disco_pwf_train_desc d) {
    const int c = threadIdx.x;
    if (c >= kStatC) return;
    float dg = 0.f, db = 0.f;
    const int A = d.A;
    for (int pair = 0; pair < d.B * A * A; ++pair) {
        const int j = pair % A, i = (pair / A) % A, b = pair / (A * A);
        if (!pair_used(d, b, i, j)) continue;
        const float* gs = d.gsum + (long long)pair * 2 * kStatC;
        db += gs[c];
        dg += gs[kStatC + c];
    }
    int og, ob, cl;
    if (c < kHid) { og = kDG1; ob = kDBE1; cl = c; }
    else if (c < kHid + kH2) { og = kDG2; ob = kDBE2; cl = c - kHid; }
    else { og = kDG3; ob = kDBE3; cl = c - kHid - kH2; }
    d.dparams[og + cl] = dg;
    d.dparams[ob + cl] = db;
}

// ------------------------------------------------------------------------------------------------------------
// backward of softmax-over-agents + weighted sum + warp (one warp per (ego row, cell))
// ------------------------------------------------------------------------------------------------------------
template <int CPL>
__device__ __forceinline__ void load_feat_f32(const uint16_t* p, long long lo_off, float (&v)[CPL]) {
#pragma unroll
    for (int u = 0; u < CPL / 4; ++u) {
        const uint2 hh = __ldg(reinterpret_cast<const uint2*>(p) + u);
        const uint2 ll = __ldg(reinterpret_cast<const uint2*>(p + lo_off) + u);
        v[4 * u] = __uint_as_float(hh.x << 16) + __uint_as_float(ll.x << 16);
        v[4 * u + 1] = __uint_as_float(hh.x & 0xffff0000u) + __uint_as_float(ll.x & 0xffff0000u);
        v[4 * u + 2] = __uint_as_float(hh.y << 16) + __uint_as_float(ll.y << 16);
        v[4 * u + 3] = __uint_as_float(hh.y & 0xffff0000u) + __uint_as_float(ll.y & 0xffff0000u);
    }
}

template <int CPL>
__device__ __forceinline__ void atomic_axpy(float* dst, float a, const float (&g)[CPL]) {
#pragma unroll
    for (int u = 0; u < CPL / 4; ++u)
        atomicAdd(reinterpret_cast<float4*>(dst) + u, make_float4(a * g[4 * u], a * g[4 * u + 1], a * g[4 * u + 2], a * g[4 * u + 3]));
}

constexpr int kCombWarps = 8;

template <int CPL>
__global__ void __launch_bounds__(kCombWarps * 32) fusion_combine_bwd_kernel(const disco_pwf_train_desc d) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = d.C, h = d.h, w = d.w, A = d.A, B = d.B, HW = h * w;
    const long long cell = (long long)blockIdx.x * kCombWarps + warp;
    if (cell >= (long long)A * B * HW) return;
    const int x = (int)(cell % w), y = (int)((cell / w) % h);
    const int n_i = (int)(cell / HW);
    const int i = n_i / B, b = n_i - i * B;
    const int n_ag = d.num_agent[b];
    const long long row_i = ((long long)n_i * h + y) * w + x;
    const uint16_t* feat = reinterpret_cast<const uint16_t*>(d.feat_hi);
    float g[CPL];
    {
        const float4* gp = reinterpret_cast<const float4*>(d.dfused + row_i * C + lane * CPL);
#pragma unroll
        for (int u = 0; u < CPL / 4; ++u) {
            const float4 v = __ldg(gp + u);
            g[4 * u] = v.x; g[4 * u + 1] = v.y; g[4 * u + 2] = v.z; g[4 * u + 3] = v.w;
        }
    }
    const bool passthrough = (i >= n_ag) || (d.outage && d.outage[b * A + i]);
    if (passthrough) {
        atomic_axpy<CPL>(d.dfeat + row_i * C + lane * CPL, 1.f, g);
        if (lane < A) d.dwlogit[(((long long)b * A + i) * A + lane) * HW + y * w + x] = 0.f;
        return;
    }
    // pass 1: e_k = exp(w_k), dot_k = <dfused, nb_k>; lane j keeps the values of neighbour id j
    float my_e = 0.f, my_dot = 0.f, my_wl = 0.f;
    float esum = 0.f;
    for (int j = 0; j < A; ++j) {
        if (!pair_used(d, b, i, j)) continue;
        const Taps t = pair_taps(d, b, i, j, y, x);
        float dot = 0.f;
        for (int k = 0; k < t.n; ++k) {
            float nb[CPL];
            load_feat_f32<CPL>(feat + t.row[k] * C + lane * CPL, d.feat_lo_off, nb);
            float part = 0.f;
#pragma unroll
            for (int c = 0; c < CPL; ++c) part = fmaf(nb[c], g[c], part);
            dot = fmaf(t.w[k], part, dot);
        }
        dot = warp_sum(dot);
        const float wl = d.wlogit[(((long long)b * A + i) * A + j) * HW + y * w + x];
        const float e = expf(wl);
        esum += e;
        if (lane == j) { my_e = e; my_dot = dot; my_wl = wl; }
    }
    const float inv = 1.f / esum;
    const float my_a = my_e * inv;                       // softmax weight of neighbour id `lane` (0 if unused)
    const float S = warp_sum(my_a * my_dot);
    if (lane < A) {
        // d/dw_k of sum_m a_m nb_m = a_k (dot_k - S); the PWF output ReLU gate is applied here
        const float dl = (my_wl > 0.f) ? my_a * (my_dot - S) : 0.f;
        d.dwlogit[(((long long)b * A + i) * A + lane) * HW + y * w + x] = dl;
    }
    // pass 2: d nb_k = a_k * dfused, scattered through the transpose of the bilinear warp
    for (int j = 0; j < A; ++j) {
        const float a = __shfl_sync(0xffffffffu, my_a, j);
        if (!pair_used(d, b, i, j)) continue;
        const Taps t = pair_taps(d, b, i, j, y, x);
        for (int k = 0; k < t.n; ++k) atomic_axpy<CPL>(d.dfeat + t.row[k] * C + lane * CPL, a * t.w[k], g);
    }
}

int check_desc(const disco_pwf_train_desc* d) {
    DISCO_REQUIRE(d && d->en && d->trans && d->num_agent && d->psum && d->wlogit, "pwf_train: null tensor");
    DISCO_REQUIRE(d->g1 && d->be1 && d->w2 && d->b2 && d->g2 && d->be2 && d->w3 && d->b3 && d->g3 && d->be3 && d->w4 && d->b4,
                  "pwf_train: null parameter");
    DISCO_REQUIRE(d->hid == kHid, "pwf_train: hidden width must be %d", kHid);
    DISCO_REQUIRE(d->A >= 1 && d->A <= 32 && d->B >= 1 && d->h > 0 && d->w > 0, "pwf_train: bad scene shape");
    return DISCO_OK;
}

// once per (kernel, device): keeps attribute calls out of CUDA-graph captures of later steps
template <typename K>
int set_smem_attr(K kernel) {
    static const void* seen[16 * 8];     // (kernel, device) pairs already configured (8 kernels here)
    static int n_seen = 0;
    int dev = 0;
    DISCO_CHECK_CUDA(cudaGetDevice(&dev));
    const void* key = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(kernel) ^ ((uintptr_t)(dev + 1) << 48));
    for (int i = 0; i < n_seen; ++i)
        if (seen[i] == key) return DISCO_OK;
    DISCO_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairSmem)));
    if (n_seen < 16 * 8) seen[n_seen++] = key;
    return DISCO_OK;
}

}  // namespace

int disco_pwf_train_forward_launch(const disco_pwf_train_desc* d, void* stream) {
    int rc = check_desc(d);
    if (rc < 0) return rc;
    DISCO_REQUIRE(d->rm1 && d->rv1 && d->rm2 && d->rv2 && d->rm3 && d->rv3 && d->nbt1 && d->nbt2 && d->nbt3,
                  "pwf_train forward: null running statistics");
    cudaStream_t s = (cudaStream_t)stream;
    const int pairs = d->B * d->A * d->A, HW = d->h * d->w;
    if ((rc = set_smem_attr(pwf_fwd_pass_kernel<0>)) < 0 || (rc = set_smem_attr(pwf_fwd_pass_kernel<1>)) < 0 ||
        (rc = set_smem_attr(pwf_fwd_pass_kernel<2>)) < 0 || (rc = set_smem_attr(pwf_fwd_pass_kernel<3>)) < 0) return rc;
    DISCO_CHECK_CUDA(cudaMemsetAsync(d->psum, 0, sizeof(double) * 2 * kStatC * pairs, s));
    const dim3 grid((HW + kChunkPix - 1) / kChunkPix, pairs);
    pwf_fwd_pass_kernel<0><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_fwd_pass_kernel<1><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_fwd_pass_kernel<2><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_fwd_pass_kernel<3><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_running_kernel<<<1, 192, 0, s>>>(*d);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_fusion_combine_backward_launch(const disco_pwf_train_desc* d, void* stream) {
    int rc = check_desc(d);
    if (rc < 0) return rc;
    DISCO_REQUIRE(d->feat_hi && d->dfused && d->dwlogit && d->dfeat, "fusion backward: null tensor");
    DISCO_REQUIRE(d->C == 128 || d->C == 256 || d->C == 512, "fusion backward: C must be 128, 256 or 512");
    const long long cells = (long long)d->A * d->B * d->h * d->w;
    const unsigned blocks = (unsigned)((cells + kCombWarps - 1) / kCombWarps);
    cudaStream_t s = (cudaStream_t)stream;
    if (d->C == 128) fusion_combine_bwd_kernel<4><<<blocks, kCombWarps * 32, 0, s>>>(*d);
    else if (d->C == 256) fusion_combine_bwd_kernel<8><<<blocks, kCombWarps * 32, 0, s>>>(*d);
    else fusion_combine_bwd_kernel<16><<<blocks, kCombWarps * 32, 0, s>>>(*d);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_pwf_train_backward_launch(const disco_pwf_train_desc* d, void* stream) {
    int rc = check_desc(d);
    if (rc < 0) return rc;
    DISCO_REQUIRE(d->dwlogit && d->den && d->dparams && d->gsum, "pwf_train backward: null tensor");
    cudaStream_t s = (cudaStream_t)stream;
    const int pairs = d->B * d->A * d->A, HW = d->h * d->w;
    if ((rc = set_smem_attr(pwf_bwd_pass_kernel<0>)) < 0 || (rc = set_smem_attr(pwf_bwd_pass_kernel<1>)) < 0 ||
        (rc = set_smem_attr(pwf_bwd_pass_kernel<2>)) < 0 || (rc = set_smem_attr(pwf_bwd_pass_kernel<3>)) < 0) return rc;
    DISCO_CHECK_CUDA(cudaMemsetAsync(d->gsum, 0, sizeof(float) * 2 * kStatC * pairs, s));
    const dim3 grid((HW + kChunkPix - 1) / kChunkPix, pairs);
    pwf_bwd_pass_kernel<0><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_bwd_pass_kernel<1><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_bwd_pass_kernel<2><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_bwd_pass_kernel<3><<<grid, kPW * 32, sizeof(PairSmem), s>>>(*d);
    pwf_bwd_finish_kernel<<<1, 192, 0, s>>>(*d);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
