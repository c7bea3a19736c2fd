// Detection post-processing and regression loss on the device.
//
//   nms_rotated  : the reference's `non_max_suppression` (utils/postprocess.py:72-115) -- keep `scores > 0.7`, order by
//                  descending score, then greedily pick the best box and drop every remaining box whose rotated-polygon
//                  IoU with it exceeds the threshold (0.01 at both call sites: detection_util.py:349-351 `apply_nms_det`
//                  and :962-964 `late_fusion`).  The reference builds shapely polygons on the host and runs an O(n^2)
//                  Python loop per agent; here: one block sorts the candidates of an agent (bitonic, 64-bit keys), a
//                  grid computes the upper-triangular "IoU > thr" bit matrix with an exact float64 convex-quad clip
//                  (Sutherland-Hodgman + shoelace; same operation order as oracle/post_oracle.py, no FMA contraction),
//                  and one warp per agent replays the greedy scan over the bit rows.
//   corner_loss  : `FaFModule.corner_loss` (utils/CoDetModule.py:80-105): decode prediction and target of every
//                  assigned anchor (bev_box_decode_torch, detection_util.py:376-400), four rotated corners each
//                  (center_to_corner_box2d_torch :403-479), sum of corner distances / N -- value and gradient wrt the
//                  regression map in ONE launch (the reference gathers, decodes and back-propagates through ~40 torch ops).
#include "common.cuh"
#include "conv.h"
#include "ops.h"

namespace {

// ------------------------------------------------------------------------------------------------------------
// exact-order float64 helpers (no FMA contraction: every product and sum is rounded like numpy does it)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
// cross(b - a, p - a)
__device__ __forceinline__ double cross3(double ax, double ay, double bx, double by, double px, double py) {
    return dsub(dmul(dsub(bx, ax), dsub(py, ay)), dmul(dsub(by, ay), dsub(px, ax)));
}

// shoelace: 0.5 * sum(x_i * y_{i+1} - x_{i+1} * y_i), summed in index order
__device__ __forceinline__ double poly_area_signed(const double* x, const double* y, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        s = dadd(s, dsub(dmul(x[i], y[j]), dmul(x[j], y[i])));
    }
    return dmul(0.5, s);
}

// Intersection area of two convex quads (vertices in either orientation).  Sutherland-Hodgman: clip P by the four edges
// of Q (both made counter-clockwise first).  Inside test `cross >= 0`; an edge crossing is s + t (e - s) with
// t = ds / (ds - de).  Must stay in lock-step with oracle/post_oracle.py::quad_intersection_area.
// The two vertex lists (<= 8 vertices each) are indexed dynamically; they live in SHARED memory, interleaved over the 64
// threads of the CTA (element k of thread t at scratch[k * 64 + t]: conflict-free) -- as per-thread stack arrays they went to
// local memory and the kernel moved 2.5 GB through L2 per 80 agents (round-2 ncu capture).
constexpr int kClipStride = 64;
struct ClipScratch {
    double* base;   // &scratch[threadIdx.x]; two vertex lists (x[8], y[8] each) at stride kClipStride
};

__device__ double quad_intersection_area(const double* P, const double* Q, const ClipScratch& w) {
    double px[4], py[4], qx[4], qy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { px[i] = P[2 * i]; py[i] = P[2 * i + 1]; qx[i] = Q[2 * i]; qy[i] = Q[2 * i + 1]; }
    if (poly_area_signed(px, py, 4) < 0.0) {
        double t;
        t = px[1]; px[1] = px[3]; px[3] = t; t = py[1]; py[1] = py[3]; py[3] = t;
    }
    if (poly_area_signed(qx, qy, 4) < 0.0) {
        double t;
        t = qx[1]; qx[1] = qx[3]; qx[3] = t; t = qy[1]; qy[1] = qy[3]; qy[3] = t;
    }
    int n = 4;
    // two vertex lists in the thread's shared-memory scratch, used alternately as source and destination of a clip pass
    double* cur = w.base;                          // x at [i], y at [8 + i] (stride kClipStride)
    double* nxt = w.base + 16 * kClipStride;
#pragma unroll
    for (int i = 0; i < 4; ++i) { cur[i * kClipStride] = px[i]; cur[(8 + i) * kClipStride] = py[i]; }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        if (n <= 0) break;
        const double ax = qx[e], ay = qy[e], bx = qx[(e + 1) & 3], by = qy[(e + 1) & 3];
        double sjx_c = 0.0, sjy_c = 0.0, de_c = 0.0;
        int m = 0;
        // (the end vertex of step i is the start vertex of step i + 1: its coordinates and its cross product are carried over
        //  instead of being re-read and re-evaluated -- same expressions on the same inputs, so the same bits)
        double six = cur[0], siy = cur[8 * kClipStride];
        double ds = cross3(ax, ay, bx, by, six, siy);
        const double s0x = six, s0y = siy, d0 = ds;
        for (int i = 0; i < n; ++i, six = sjx_c, siy = sjy_c, ds = de_c) {
            const bool last = i + 1 == n;
            const double sjx = last ? s0x : cur[(i + 1) * kClipStride], sjy = last ? s0y : cur[(8 + i + 1) * kClipStride];
            const double de = last ? d0 : cross3(ax, ay, bx, by, sjx, sjy);
            sjx_c = sjx; sjy_c = sjy; de_c = de;
            const bool in_s = ds >= 0.0, in_e = de >= 0.0;
            if (in_s && m < 8) { nxt[m * kClipStride] = six; nxt[(8 + m) * kClipStride] = siy; ++m; }
            if (in_s != in_e && m < 8) {
                const double t = ds / dsub(ds, de);
                nxt[m * kClipStride] = dadd(six, dmul(t, dsub(sjx, six)));
                nxt[(8 + m) * kClipStride] = dadd(siy, dmul(t, dsub(sjy, siy)));
                ++m;
            }
        }
        n = m;
        double* tmp = cur; cur = nxt; nxt = tmp;
    }
    if (n < 3) return 0.0;
    double s = 0.0;   // shoelace in index order (poly_area_signed)
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        s = dadd(s, dsub(dmul(cur[i * kClipStride], cur[(8 + j) * kClipStride]), dmul(cur[j * kClipStride], cur[(8 + i) * kClipStride])));
    }
    return fabs(dmul(0.5, s));
}

__device__ __forceinline__ double quad_area(const double* P) {
    double x[4], y[4];
    for (int i = 0; i < 4; ++i) { x[i] = P[2 * i]; y[i] = P[2 * i + 1]; }
    return fabs(poly_area_signed(x, y, 4));
}

// iou > thr, with the decision the oracle makes: inter / (area_p + area_q - inter); a zero union gives NaN -> false.
__device__ __forceinline__ bool quad_iou_above(const double* P, const double* Q, double thr, const ClipScratch& w) {
    // (the caller has already rejected pairs whose axis-aligned bounds do not touch: exact-zero intersection, never > thr)
    const double inter = quad_intersection_area(P, Q, w);
    const double uni = dsub(dadd(quad_area(P), quad_area(Q)), inter);
    const double iou = inter / uni;
    return iou > thr;   // NaN compares false
}

// ------------------------------------------------------------------------------------------------------------
// 1. per-agent candidate sort: key = score bits | anchor id | slot, descending (score desc, ties: larger id first, which is
//    what `scores.argsort()[::-1]` yields for a stable argsort)
// ------------------------------------------------------------------------------------------------------------
template <typename CT>
__global__ void __launch_bounds__(1024) nms_sort_kernel(const CT* __restrict__ corners, const float* __restrict__ scores,
                                                        const int* __restrict__ ids, const int* __restrict__ count, int cap,
                                                        int kmax, float score_thresh, double* __restrict__ s_corners,
                                                        float* __restrict__ s_scores, int* __restrict__ s_ids,
                                                        int* __restrict__ s_slot, int* __restrict__ s_count) {
    extern __shared__ unsigned long long keys[];
    const int a = blockIdx.x, tid = threadIdx.x;
    int k_in = count ? count[a] : cap;
    if (k_in > cap) k_in = cap;
    int P = 32;
    while (P < k_in) P <<= 1;
    for (int i = tid; i < P; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < k_in) {
            const float sc = scores[(long long)a * cap + i];
            if (sc > score_thresh) {
                const unsigned id = ids ? (unsigned)ids[(long long)a * cap + i] : (unsigned)i;
                key = ((unsigned long long)__float_as_uint(sc) << 32) | ((unsigned long long)(id & 0x7FFFFu) << 13) |
                      (unsigned long long)(i & 0x1FFF);
            }
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < P / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));   // index with bit `stride` clear
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long x = keys[lo], y = keys[hi];
                if ((x < y) == desc) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    }
    // number of valid (score > thresh) candidates = non-zero keys (scores > 0 have non-zero bit patterns)
    __shared__ int n_valid;
    if (tid == 0) n_valid = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < P; i += blockDim.x) local += keys[i] != 0ull;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((tid & 31) == 0 && local) atomicAdd(&n_valid, local);
    __syncthreads();
    int K = n_valid;
    if (K > kmax) K = kmax;
    if (tid == 0) s_count[a] = n_valid;   // (> kmax reports the overflow to the host)
    for (int r = tid; r < K; r += blockDim.x) {
        const unsigned long long key = keys[r];
        const int slot = (int)(key & 0x1FFF);
        const long long src = (long long)a * cap + slot, dst = (long long)a * kmax + r;
        s_scores[dst] = __uint_as_float((unsigned)(key >> 32));
        s_ids[dst] = ids ? ids[src] : slot;
        s_slot[dst] = slot;
#pragma unroll
        for (int q = 0; q < 8; ++q) s_corners[dst * 8 + q] = (double)corners[src * 8 + q];
    }
}

// ------------------------------------------------------------------------------------------------------------
// 2. upper-triangular "IoU > thr" bit matrix: mask[a][i][j / 64] bit (j % 64), j > i.  kMaskBlocks CTAs per agent walk the
//    (row block, column block) pairs of the K x K upper triangle (K is only known on the device, so the grid cannot be sized to
//    it: round 2's first version launched 32 x 32 x n mostly-empty CTAs and spent 0.8 ms per 80 agents on their launch slots).
// ------------------------------------------------------------------------------------------------------------
constexpr int kMaskBlocks = 96;   // >= the 91 block pairs of ~800 candidates: one pair per CTA in the common case
// (Measured and rejected in round 2: dealing the (agent, block pair) items of ALL agents round-robin over one 1184-CTA grid -- K varies
//  between agents, mean ~700, largest 2048 -- 317 -> 346 us; a separating-axis pre-check before the clip -- the candidates that
//  pass the bounds test are the six anchors of neighbouring cells and nearly all truly intersect -- 346 -> 375 us.)

__global__ void __launch_bounds__(64) nms_mask_kernel(const double* __restrict__ s_corners, const int* __restrict__ s_count,
                                                      int kmax, int words, double thr,
                                                      unsigned long long* __restrict__ mask) {
    const int a = blockIdx.y, t = threadIdx.x;
    int K = s_count[a];
    if (K > kmax) K = kmax;
    const int nb = (K + 63) / 64;                 // 64-box blocks per side
    const int n_pairs = nb * (nb + 1) / 2;        // upper triangle incl. the diagonal blocks
    __shared__ double cq[64][8];
    __shared__ double qbox[64][4];                // column boxes' axis-aligned bounds (xmin, xmax, ymin, ymax)
    __shared__ double clip[32 * kClipStride];     // Sutherland-Hodgman vertex lists of the 64 threads (16 KB)
    const ClipScratch scratch{clip + t};
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        // pair -> (rb, cb), rb <= cb: rows of the triangle have nb, nb-1, ... entries
        int rb = 0, rem = pair;
        while (rem >= nb - rb) { rem -= nb - rb; ++rb; }
        const int cb = rb + rem;
        const int j0 = cb * 64;
        __syncthreads();
        if (j0 + t < K) {
            double xl, xh, yl, yh;
#pragma unroll
            for (int q = 0; q < 8; ++q) cq[t][q] = s_corners[((long long)a * kmax + j0 + t) * 8 + q];
            xl = xh = cq[t][0]; yl = yh = cq[t][1];
#pragma unroll
            for (int q = 1; q < 4; ++q) {
                xl = fmin(xl, cq[t][2 * q]); xh = fmax(xh, cq[t][2 * q]);
                yl = fmin(yl, cq[t][2 * q + 1]); yh = fmax(yh, cq[t][2 * q + 1]);
            }
            qbox[t][0] = xl; qbox[t][1] = xh; qbox[t][2] = yl; qbox[t][3] = yh;
        }
        __syncthreads();
        const int i = rb * 64 + t;
        if (i >= K) continue;
        double P[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) P[q] = s_corners[((long long)a * kmax + i) * 8 + q];
        double pl = P[0], ph = P[0], pb = P[1], pt = P[1];
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            pl = fmin(pl, P[2 * q]); ph = fmax(ph, P[2 * q]); pb = fmin(pb, P[2 * q + 1]); pt = fmax(pt, P[2 * q + 1]);
        }
        // phase 1: which column boxes can intersect at all (axis-aligned bounds touch; otherwise IoU is 0 or 0/0, never > thr)
        unsigned long long hits = 0ull;
        const int jn = min(64, K - j0);
        for (int jj = 0; jj < jn; ++jj) {
            if (j0 + jj <= i) continue;
            if (!(ph < qbox[jj][0] || qbox[jj][1] < pl || pt < qbox[jj][2] || qbox[jj][3] < pb)) hits |= 1ull << jj;
        }
        // phase 2: every lane clips ITS OWN next hit in each trip, so the warp runs max-popcount trips of the expensive float64
        // clip instead of one (mostly idle) trip per column that any lane hits (~6 % of the pairs overlap: 8-10 trips, not ~55)
        unsigned long long bits = 0ull;
        while (hits) {
            const int jj = __ffsll((long long)hits) - 1;
            hits &= hits - 1;
            if (quad_iou_above(P, cq[jj], thr, scratch)) bits |= 1ull << jj;
        }
        mask[((long long)a * kmax + i) * words + cb] = bits;
    }
}

// ------------------------------------------------------------------------------------------------------------
// 3. greedy scan (postprocess.py:97-112), one CTA per agent, 64 candidates (one chunk) at a time.  The sequential part must never
//    wait on global memory: warps 1-7 copy the 64 rows of chunk c + 1 into the other half of a shared-memory double buffer while
//    warp 0 resolves chunk c -- the 64 x 64 diagonal block decides which of its boxes survive (a register chain of test + OR whose
//    shared-memory loads do not depend on the chain, so they pipeline), then the survivors' rows are OR-ed into the removed set,
//    lane = word.  The kept list is written afterwards by all threads from the per-chunk survivor masks.
//    (Round 2, first version: one warp, rows read from global memory with one dependent ~600-cycle load per kept box: 180 us per
//    80 agents of ~700 candidates, the largest agent -- 2048 candidates -- setting the time.)
// ------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;

__global__ void __launch_bounds__(kScanThreads) nms_scan_kernel(const unsigned long long* __restrict__ mask, const int* __restrict__ s_count,
                                                                const int* __restrict__ s_slot, int kmax, int words,
                                                                int* __restrict__ keep, int* __restrict__ n_keep) {
    extern __shared__ unsigned long long scan_smem[];    // [2][64][words] row buffers, [words] survivor masks, [words + 1] prefix counts
    unsigned long long* kept_s = scan_smem + 2 * 64 * words;
    int* prefix = reinterpret_cast<int*>(kept_s + words);
    const int a = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int K = s_count[a];
    if (K > kmax) K = kmax;
    const int kw = (K + 63) / 64;                        // <= 128 (kmax <= 8192)
    const unsigned long long* grow = mask + (long long)a * kmax * words;
    // rows of chunk c, words c .. kw-1 (the words left of the diagonal are never written by nms_mask_kernel, nor read here)
    auto load_chunk = [&](int c, int t0, int nt) {
        unsigned long long* buf = scan_smem + (c & 1) * 64 * words;
        const int i0 = c * 64, cn = min(64, K - i0), nw = kw - c;
        for (int idx = t0; idx < cn * nw; idx += nt) {
            const int r = idx / nw, w = c + idx - r * nw;
            buf[r * words + w] = grow[(long long)(i0 + r) * words + w];
        }
    };
    if (kw > 0) load_chunk(0, tid, kScanThreads);
    __syncthreads();
    unsigned long long rem[4] = {0ull, 0ull, 0ull, 0ull};   // warp 0, lane l: removed bits of words l, l + 32, l + 64, l + 96
    int nk = 0;
    for (int c = 0; c < kw; ++c) {
        if (warp > 0) {
            if (c + 1 < kw) load_chunk(c + 1, tid - 32, kScanThreads - 32);
        } else {
            const unsigned long long* buf = scan_smem + (c & 1) * 64 * words;
            const int cn = min(64, K - c * 64);
            unsigned long long mine = rem[0];
#pragma unroll
            for (int k = 1; k < 4; ++k) mine = ((c >> 5) == k) ? rem[k] : mine;
            unsigned long long remc = __shfl_sync(0xffffffffu, mine, c & 31), kept = 0ull;
            const unsigned long long* dr = buf + c;      // diagonal word of row r: dr[r * words]
#pragma unroll 8
            for (int r = 0; r < cn; ++r) {               // every lane runs the same chain (no shuffles inside)
                const unsigned long long d = dr[r * words];
                const bool alive = !((remc >> r) & 1ull);
                kept |= alive ? (1ull << r) : 0ull;
                remc |= alive ? d : 0ull;
            }
            if (lane == 0) { kept_s[c] = kept; prefix[c] = nk; }
            nk += __popcll(kept);
#pragma unroll
            for (int k = 0; k < 4; ++k) {                // fold the survivors' rows into the later words (lane + 32 k = word)
                const int w = lane + 32 * k;
                if (w > c && w < kw) {
                    unsigned long long m = kept, acc = 0ull;
                    while (m) {
                        const int r = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        acc |= buf[r * words + w];
                    }
                    rem[k] |= acc;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) n_keep[a] = nk;
    // kept list in pick order: candidate i is the (prefix[chunk] + rank inside the chunk)-th survivor
    for (int i = tid; i < K; i += kScanThreads) {
        const unsigned long long kc = kept_s[i >> 6];
        if ((kc >> (i & 63)) & 1ull) {
            const int pos = prefix[i >> 6] + __popcll(kc & ((1ull << (i & 63)) - 1ull));
            keep[(long long)a * kmax + pos] = s_slot[(long long)a * kmax + i];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// corner loss: value + gradient, one thread per (anchor, time step)
// ------------------------------------------------------------------------------------------------------------
struct Box6 { float x, y, w, h, s, c; };

__device__ __forceinline__ Box6 decode6(const float* p, const float* q) {
    Box6 b;
    b.h = q[3] / expf(p[3]);
    b.w = q[2] / expf(p[2]);
    b.x = q[0] - b.w * p[0];
    b.y = q[1] - b.h * p[1];
    b.s = q[4] * p[5] + q[5] * p[4];
    b.c = q[5] * p[5] - q[4] * p[4];
    return b;
}

__global__ void __launch_bounds__(256) corner_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          const float* __restrict__ anchors, const unsigned char* __restrict__ mask,
                                                          long long n_entries, int t_len, float inv_n, double* __restrict__ loss_sum,
                                                          float* __restrict__ grad) {
    __shared__ double warp_sum[8];
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double mine = 0.0;
    if (e < n_entries) {
        float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (mask[e]) {
            const float* p = pred + e * 6;
            const float* q = anchors + (e / t_len) * 6;
            const Box6 bp = decode6(p, q), bt = decode6(target + e * 6, q);
            const float nx[4] = {-0.5f, 0.5f, 0.5f, -0.5f}, ny[4] = {0.5f, 0.5f, -0.5f, -0.5f};
            float Gx = 0.f, Gy = 0.f, Gw = 0.f, Gh = 0.f, Gs = 0.f, Gc = 0.f, l = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float px = bp.w * nx[k], py = bp.h * ny[k];
                const float cx = px * bp.c + py * bp.s + bp.x, cy = -px * bp.s + py * bp.c + bp.y;
                const float qx = bt.w * nx[k], qy = bt.h * ny[k];
                const float tx = qx * bt.c + qy * bt.s + bt.x, ty = -qx * bt.s + qy * bt.c + bt.y;
                const float dx = cx - tx, dy = cy - ty;
                const float d = sqrtf(dx * dx + dy * dy);
                l += d;
                const float gx = d > 0.f ? dx / d : 0.f, gy = d > 0.f ? dy / d : 0.f;   // torch.norm's subgradient at 0
                Gx += gx; Gy += gy;
                Gw += nx[k] * (gx * bp.c - gy * bp.s);
                Gh += ny[k] * (gx * bp.s + gy * bp.c);
                Gc += gx * px + gy * py;
                Gs += gx * py - gy * px;
            }
            mine = (double)l;
            g[0] = -bp.w * Gx;
            g[1] = -bp.h * Gy;
            g[2] = -bp.w * (Gw - p[0] * Gx);
            g[3] = -bp.h * (Gh - p[1] * Gy);
            g[4] = Gs * q[5] - Gc * q[4];
            g[5] = Gs * q[4] + Gc * q[5];
        }
        if (grad) {
#pragma unroll
            for (int i = 0; i < 6; ++i) grad[e * 6 + i] = g[i] * inv_n;
        }
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += warp_sum[w];
        if (s != 0.0) atomicAdd(loss_sum, s);
    }
}

}  // namespace

size_t disco_nms_workspace_bytes_impl(int n, int kmax) {
    const size_t words = (size_t)(kmax + 63) / 64;
    size_t b = 0;
    b += (size_t)n * kmax * 8 * sizeof(double);   // sorted corners
    b += (size_t)n * kmax * sizeof(float);        // sorted scores
    b += (size_t)n * kmax * sizeof(int) * 2;      // sorted ids, slots
    b += (size_t)n * sizeof(int) * 2;             // valid count (+pad)
    b += (size_t)n * kmax * words * 8;            // bit matrix
    return b + 256;
}

int disco_nms_rotated_launch(const void* corners, int corners_f64, const float* scores, const int* ids, const int* count, int n,
                             int cap, int kmax, float score_thresh, double iou_thresh, void* workspace, size_t workspace_bytes,
                             int* keep, int* n_keep, int* n_valid, void* stream) {
    DISCO_REQUIRE(corners && scores && workspace && keep && n_keep, "nms: null tensor");
    DISCO_REQUIRE(n > 0 && cap > 0 && cap <= 8192 && kmax > 0 && kmax <= cap, "nms: need 0 < kmax <= cap <= 8192 (got %d, %d)", kmax, cap);
    DISCO_REQUIRE(workspace_bytes >= disco_nms_workspace_bytes_impl(n, kmax), "nms: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const int words = (kmax + 63) / 64;
    uint8_t* w = reinterpret_cast<uint8_t*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    double* s_corners = reinterpret_cast<double*>(w); w += (size_t)n * kmax * 8 * sizeof(double);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(w); w += (size_t)n * kmax * words * 8;
    float* s_scores = reinterpret_cast<float*>(w); w += (size_t)n * kmax * sizeof(float);
    int* s_ids = reinterpret_cast<int*>(w); w += (size_t)n * kmax * sizeof(int);
    int* s_slot = reinterpret_cast<int*>(w); w += (size_t)n * kmax * sizeof(int);
    int* s_count = reinterpret_cast<int*>(w);
    int P = 32;
    while (P < cap) P <<= 1;
    const size_t sort_smem = (size_t)P * 8;
    static bool attr_set[64] = {false};
    int dev = 0;
    DISCO_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        DISCO_CHECK_CUDA(cudaFuncSetAttribute(nms_sort_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        DISCO_CHECK_CUDA(cudaFuncSetAttribute(nms_sort_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set[dev] = true;
    }
    if (corners_f64)
        nms_sort_kernel<double><<<n, 1024, sort_smem, s>>>((const double*)corners, scores, ids, count, cap, kmax, score_thresh,
                                                           s_corners, s_scores, s_ids, s_slot, s_count);
    else
        nms_sort_kernel<float><<<n, 1024, sort_smem, s>>>((const float*)corners, scores, ids, count, cap, kmax, score_thresh,
                                                          s_corners, s_scores, s_ids, s_slot, s_count);
    DISCO_CHECK_CUDA(cudaGetLastError());
    dim3 grid(kMaskBlocks, n);
    nms_mask_kernel<<<grid, 64, 0, s>>>(s_corners, s_count, kmax, words, iou_thresh, mask);
    DISCO_CHECK_CUDA(cudaGetLastError());
    {
        static bool scan_attr[64] = {false};
        if (dev < 64 && !scan_attr[dev]) {
            DISCO_CHECK_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            scan_attr[dev] = true;
        }
        const size_t scan_smem_bytes = (size_t)(2 * 64 + 1) * words * 8 + (size_t)(words + 1) * 4;   // <= 133 KB at kmax = 8192
        nms_scan_kernel<<<n, kScanThreads, scan_smem_bytes, s>>>(mask, s_count, s_slot, kmax, words, keep, n_keep);
    }
    DISCO_CHECK_CUDA(cudaGetLastError());
    if (n_valid) DISCO_CHECK_CUDA(cudaMemcpyAsync(n_valid, s_count, sizeof(int) * n, cudaMemcpyDeviceToDevice, s));
    return DISCO_OK;
}

int disco_corner_loss_launch(const float* pred, const float* target, const float* anchors, const unsigned char* mask,
                             long long n_entries, int t_len, float inv_n, double* loss_sum, float* grad, void* stream) {
    DISCO_REQUIRE(pred && target && anchors && mask && loss_sum, "corner_loss: null tensor");
    DISCO_REQUIRE(n_entries > 0 && t_len > 0, "corner_loss: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    DISCO_CHECK_CUDA(cudaMemsetAsync(loss_sum, 0, sizeof(double), s));
    const long long blocks = (n_entries + 255) / 256;
    DISCO_REQUIRE(blocks < (1ll << 31), "corner_loss: too many anchors");
    corner_loss_kernel<<<(unsigned)blocks, 256, 0, s>>>(pred, target, anchors, mask, n_entries, t_len, inv_n, loss_sum, grad);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
