// HBM-bound data-format kernels either side of the conv stack.
//   bev_pack      : dense BEV fp32 [N,1,H,W,Z] (what the reference DataLoader hands to the model,
//                   V2XSimDet.py:293-302 / DiscoNet.py:42) -> NHWC 16-bit with Z padded to 16
//   act_unpack    : NHWC 16-bit hi[/lo] -> fp32 NCHW (the layout the KD outputs are returned in)
//   voxelize      : LiDAR points -> occupancy bitmap -> lexicographically sorted unique voxel indices
//                   (data_util.py:625-717 voxelize_occupy; bit-exact incl. the float64 floor-divide)
//   bev_scatter   : voxel indices -> dense BEV incl. np.rot90(.,3) (V2XSimDet.py:293-302)
#include "common.cuh"
#include "conv.h"
#include "ops.h"

namespace {

__global__ void bev_pack_kernel(const float* __restrict__ bev, long long n_pix, int z, uint16_t* __restrict__ out_hi,
                                long long lo_off, int precision, int* __restrict__ lo_nonzero) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const float* src = bev + p * z;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c0 = 2 * i, c1 = 2 * i + 1;
        const float v0 = (c0 < z) ? __ldg(src + c0) : 0.f;
        const float v1 = (c1 < z) ? __ldg(src + c1) : 0.f;
        if (precision == DISCO_PREC_BF16X3) {
            uint16_t h0, l0, h1, l1;
            split_bf16(v0, h0, l0);
            split_bf16(v1, h1, l1);
            hi[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
            lo[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
        } else {
            hi[i] = (uint32_t)f32_to_f16_bits(v0) | ((uint32_t)f32_to_f16_bits(v1) << 16);
            lo[i] = 0;
        }
    }
    uint4* oh = reinterpret_cast<uint4*>(out_hi + p * 16);
    oh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    oh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    if (precision == DISCO_PREC_BF16X3) {
        uint4* ol = reinterpret_cast<uint4*>(out_hi + lo_off + p * 16);
        ol[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        ol[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        // occupancy inputs are exact in bf16 (0/1): tell the first conv when it may skip the lo plane altogether
        if (lo_nonzero && ((lo[0] | lo[1] | lo[2] | lo[3] | lo[4] | lo[5] | lo[6] | lo[7]) & 0x7FFF7FFFu)) atomicOr(lo_nonzero, 1);
    }
}

// [N, HW, C] 16-bit -> [N, C, HW] fp32 through a 32x33 shared tile (coalesced on both sides)
__global__ void act_unpack_nchw_kernel(const uint16_t* __restrict__ act, long long lo_off, int precision, int hw, int c,
                                       float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, ch = c0 + tx;
        float v = 0.f;
        if (p < hw && ch < c) {
            const uint16_t* a = act + ((long long)n * hw + p) * c + ch;
            v = (precision == DISCO_PREC_BF16X3) ? bf16_bits_to_f32(a[0]) + bf16_bits_to_f32(a[lo_off])
                                                 : f16_bits_to_f32(a[0]);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int ch = c0 + r, p = p0 + tx;
        if (p < hw && ch < c) out[((long long)n * c + ch) * hw + p] = tile[tx][r];
    }
}

// ---------------------------------------------------------------------------------------------
// voxelize_occupy
// ---------------------------------------------------------------------------------------------
__global__ void voxel_mark_kernel(const float* __restrict__ pts, int n_points, int stride, double x0, double x1,
                                  double y0, double y1, double z0, double z1, double vx, double vy, double vz, int dx,
                                  int dy, int dz, unsigned int* __restrict__ bitmap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const float* p = pts + (long long)i * stride;
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    // strict bounds, evaluated in float64 exactly like numpy's f64-vs-f32 comparison (data_util.py:657-664)
    if (!(x0 < x && x < x1 && y0 < y && y < y1 && z0 < z && z < z1)) return;
    // np.floor(pts / voxel_size) runs in float64 (f32 array / python-float tuple); IEEE division, no fast-math
    const int ix = (int)floor(__ddiv_rn(x, vx)) - (int)floor(__ddiv_rn(x0, vx));
    const int iy = (int)floor(__ddiv_rn(y, vy)) - (int)floor(__ddiv_rn(y0, vy));
    const int iz = (int)floor(__ddiv_rn(z, vz)) - (int)floor(__ddiv_rn(z0, vz));
    if (ix < 0 || ix >= dx || iy < 0 || iy >= dy || iz < 0 || iz >= dz) return;
    const unsigned int key = ((unsigned)ix * dy + iy) * dz + iz;  // ascending key == lexsort (x, y, z) order
    atomicOr(bitmap + (key >> 5), 1u << (key & 31));
}

// one block: ordered compaction of the set bits -> [M,3] int32 indices (sorted unique by construction)
__global__ void __launch_bounds__(1024) voxel_compact_kernel(const unsigned int* __restrict__ bitmap, int n_words,
                                                             int dy, int dz, int n_bits, int* __restrict__ out_idx,
                                                             int* __restrict__ n_out) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    // words are processed in rounds of 1024 consecutive words so that global order == key order
    for (int base = 0; base < n_words; base += 1024) {
        const int wi = base + tid;
        const unsigned int word = (wi < n_words) ? bitmap[wi] : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int v = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            warp_sums[lane] = v;  // inclusive over warps
        }
        __syncthreads();
        const int offset = carry + (wid ? warp_sums[wid - 1] : 0) + incl - cnt;
        if (out_idx) {
            unsigned int wbits = word;
            int o = offset;
            while (wbits) {
                const int b = __ffs(wbits) - 1;
                wbits &= wbits - 1;
                const int key = wi * 32 + b;
                if (key < n_bits) {
                    const int iz = key % dz, t = key / dz;
                    out_idx[3 * o + 0] = t / dy;
                    out_idx[3 * o + 1] = t % dy;
                    out_idx[3 * o + 2] = iz;
                    ++o;
                }
            }
        }
        __syncthreads();
        if (tid == 0) carry += warp_sums[31];
        __syncthreads();
    }
    if (tid == 0) *n_out = carry;
}

// ---- batched form: S sweeps in three launches (mark, per-block bit counts, ordered multi-block compaction) ----------------
__global__ void voxel_mark_batched_kernel(const float* __restrict__ pts, const int* __restrict__ n_points, int p_max, int stride,
                                          double x0, double x1, double y0, double y1, double z0, double z1, double vx, double vy,
                                          double vz, int dx, int dy, int dz, int n_words, unsigned int* __restrict__ bitmap) {
    const int s = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p_max || i >= n_points[s]) return;
    const float* p = pts + ((long long)s * p_max + i) * stride;
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    if (!(x0 < x && x < x1 && y0 < y && y < y1 && z0 < z && z < z1)) return;
    const int ix = (int)floor(__ddiv_rn(x, vx)) - (int)floor(__ddiv_rn(x0, vx));
    const int iy = (int)floor(__ddiv_rn(y, vy)) - (int)floor(__ddiv_rn(y0, vy));
    const int iz = (int)floor(__ddiv_rn(z, vz)) - (int)floor(__ddiv_rn(z0, vz));
    if (ix < 0 || ix >= dx || iy < 0 || iy >= dy || iz < 0 || iz >= dz) return;
    const unsigned int key = ((unsigned)ix * dy + iy) * dz + iz;
    atomicOr(bitmap + (long long)s * n_words + (key >> 5), 1u << (key & 31));
}

// block (b, s) counts the set bits of words [1024 b, 1024 (b+1)) of sweep s
__global__ void __launch_bounds__(1024) voxel_count_kernel(const unsigned int* __restrict__ bitmap, int n_words, int n_blocks,
                                                           int* __restrict__ block_count) {
    __shared__ int warp_sums[32];
    const int s = blockIdx.y, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wi = b * 1024 + tid;
    int cnt = (wi < n_words) ? __popc(bitmap[(long long)s * n_words + wi]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) warp_sums[wid] = cnt;
    __syncthreads();
    if (wid == 0) {
        int v = warp_sums[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) block_count[s * n_blocks + b] = v;
    }
}

// block (b, s): offset = bits in the blocks before it; ordered compaction of its 1024 words -> indices [s, offset..]; pad rows -1
__global__ void __launch_bounds__(1024) voxel_compact_batched_kernel(const unsigned int* __restrict__ bitmap, int n_words, int n_blocks,
                                                                     const int* __restrict__ block_count, int dy, int dz, int n_bits,
                                                                     int m_max, int* __restrict__ out_idx, int* __restrict__ n_out) {
    __shared__ int warp_sums[32];
    __shared__ int base;
    const int s = blockIdx.y, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        int off = 0;
        for (int k = 0; k < b; ++k) off += block_count[s * n_blocks + k];
        base = off;
        if (b == n_blocks - 1) n_out[s] = off + block_count[s * n_blocks + b];
    }
    const int wi = b * 1024 + tid;
    const unsigned int word = (wi < n_words) ? bitmap[(long long)s * n_words + wi] : 0u;
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int v = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        warp_sums[lane] = v;
    }
    __syncthreads();
    int o = base + (wid ? warp_sums[wid - 1] : 0) + incl - cnt;
    unsigned int wbits = word;
    while (wbits) {
        const int bit = __ffs(wbits) - 1;
        wbits &= wbits - 1;
        const int key = wi * 32 + bit;
        if (key < n_bits && o < m_max) {
            int* dst = out_idx + ((long long)s * m_max + o) * 3;
            const int iz = key % dz, t = key / dz;
            dst[0] = t / dy; dst[1] = t % dy; dst[2] = iz;
        }
        ++o;
    }
}

__global__ void voxel_dense_kernel(const unsigned int* __restrict__ bitmap, int n_bits, float* __restrict__ dense) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_bits) dense[i] = ((bitmap[i >> 5] >> (i & 31)) & 1u) ? 1.f : 0.f;
}

// bev[r = y, c = X-1-x, z] = 1  (np.rot90(vox, 3) of vox[x, y, z]); optional packed 16-channel activation copy
__global__ void bev_scatter_kernel(const int* __restrict__ idx, int n, int dx, int dy, int dz, float* __restrict__ bev,
                                   uint16_t* __restrict__ act, int act_c, uint16_t one_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = idx[3 * i], y = idx[3 * i + 1], z = idx[3 * i + 2];
    if (x < 0 || x >= dx || y < 0 || y >= dy || z < 0 || z >= dz) return;
    const long long pix = (long long)y * dx + (dx - 1 - x);
    if (bev) bev[pix * dz + z] = 1.f;
    if (act) act[pix * act_c + z] = one_bits;
}

// Batched dataset scatter straight into the encoder's input activation: indices [n, m_max, 3] (x, y, z; rows >= count[a] are
// ignored) -> act[a, y, X-1-x, z] = 1 in the 16-channel NHWC input buffer (hi plane; the lo plane of an exact 0/1 tensor is 0).
__global__ void bev_scatter_batched_kernel(const int* __restrict__ idx, const int* __restrict__ count, int m_max, int dx, int dy,
                                           int dz, uint16_t* __restrict__ act, int act_c, uint16_t one_bits) {
    const int a = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m_max || i >= count[a]) return;
    const int* p = idx + ((long long)a * m_max + i) * 3;
    const int x = p[0], y = p[1], z = p[2];
    if (x < 0 || x >= dx || y < 0 || y >= dy || z < 0 || z >= dz) return;
    const long long pix = ((long long)a * dy + y) * dx + (dx - 1 - x);
    act[pix * act_c + z] = one_bits;
}

}  // namespace

int disco_bev_scatter_batched_launch(const int* voxel_indices, const int* counts, int n, int m_max, const int* dims, void* act_hi,
                                     long long act_lo_off, int act_c, int precision, int* lo_nonzero, void* stream) {
    DISCO_REQUIRE(dims && act_hi && counts, "bev_scatter_batched: null argument");
    DISCO_REQUIRE(n > 0 && m_max >= 0 && (m_max == 0 || voxel_indices), "bev_scatter_batched: bad indices");
    DISCO_REQUIRE(act_c >= dims[2], "bev_scatter_batched: act_c %d < z dim %d", act_c, dims[2]);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t plane = (size_t)n * dims[0] * dims[1] * act_c * 2;
    if (lo_nonzero) DISCO_CHECK_CUDA(cudaMemsetAsync(lo_nonzero, 0, sizeof(int), s));   // the lo plane of a 0/1 tensor is all zero
    DISCO_CHECK_CUDA(cudaMemsetAsync(act_hi, 0, plane, s));
    if (precision == DISCO_PREC_BF16X3) DISCO_CHECK_CUDA(cudaMemsetAsync((uint16_t*)act_hi + act_lo_off, 0, plane, s));
    if (m_max > 0) {
        const uint16_t one = (precision == DISCO_PREC_BF16X3) ? 0x3F80 : 0x3C00;
        dim3 grid((m_max + 255) / 256, n);
        bev_scatter_batched_kernel<<<grid, 256, 0, s>>>(voxel_indices, counts, m_max, dims[0], dims[1], dims[2], (uint16_t*)act_hi,
                                                        act_c, one);
        DISCO_CHECK_CUDA(cudaGetLastError());
    }
    return DISCO_OK;
}

int disco_bev_pack_launch(const float* bev, long long n_pix, int z, void* out_hi, long long out_lo_off, int precision,
                          int* lo_nonzero, void* stream) {
    DISCO_REQUIRE(bev && out_hi && n_pix > 0 && z > 0 && z <= 16, "bev_pack: bad arguments (z=%d)", z);
    const int threads = 256;
    const long long blocks = (n_pix + threads - 1) / threads;
    if (lo_nonzero) DISCO_CHECK_CUDA(cudaMemsetAsync(lo_nonzero, 0, sizeof(int), (cudaStream_t)stream));
    bev_pack_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(bev, n_pix, z, (uint16_t*)out_hi, out_lo_off,
                                                                             precision, lo_nonzero);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_act_unpack_nchw_launch(const void* act_hi, long long lo_off, int precision, int n, int h, int w, int c,
                                 float* out_nchw, void* stream) {
    DISCO_REQUIRE(act_hi && out_nchw && n > 0 && h > 0 && w > 0 && c > 0, "act_unpack: bad arguments");
    const int hw = h * w;
    dim3 grid((hw + 31) / 32, (c + 31) / 32, n), block(32, 8);
    act_unpack_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint16_t*)act_hi, lo_off, precision, hw, c,
                                                                      out_nchw);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_voxelize_launch(const float* points, int n_points, int point_stride, const double* extents,
                          const double* voxel_size, const int* dims, unsigned int* bitmap, int* voxel_indices,
                          int* n_voxels, float* dense, void* stream) {
    DISCO_REQUIRE(extents && voxel_size && dims && bitmap && n_voxels, "voxelize: null argument");
    DISCO_REQUIRE(n_points >= 0 && (n_points == 0 || points), "voxelize: bad points");
    DISCO_REQUIRE(point_stride >= 3, "voxelize: points need >= 3 columns (got %d)", point_stride);
    const long long n_bits = (long long)dims[0] * dims[1] * dims[2];
    DISCO_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && n_bits < (1ll << 30), "voxelize: bad dims");
    const int n_words = (int)((n_bits + 31) / 32);
    cudaStream_t s = (cudaStream_t)stream;
    DISCO_CHECK_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)n_words * 4, s));
    if (n_points > 0) {
        voxel_mark_kernel<<<(n_points + 255) / 256, 256, 0, s>>>(points, n_points, point_stride, extents[0], extents[1],
                                                                 extents[2], extents[3], extents[4], extents[5],
                                                                 voxel_size[0], voxel_size[1], voxel_size[2], dims[0],
                                                                 dims[1], dims[2], bitmap);
        DISCO_CHECK_CUDA(cudaGetLastError());
    }
    voxel_compact_kernel<<<1, 1024, 0, s>>>(bitmap, n_words, dims[1], dims[2], (int)n_bits, voxel_indices, n_voxels);
    DISCO_CHECK_CUDA(cudaGetLastError());
    if (dense) {
        voxel_dense_kernel<<<(unsigned)((n_bits + 255) / 256), 256, 0, s>>>(bitmap, (int)n_bits, dense);
        DISCO_CHECK_CUDA(cudaGetLastError());
    }
    return DISCO_OK;
}

int disco_voxelize_batched_launch(const float* points, const int* n_points, int n_sweeps, int p_max, int point_stride,
                                  const double* extents, const double* voxel_size, const int* dims, unsigned int* bitmap,
                                  int* block_count, int* voxel_indices, int m_max, int* n_voxels, void* stream) {
    DISCO_REQUIRE(extents && voxel_size && dims && bitmap && block_count && n_voxels && n_points && voxel_indices, "voxelize_batched: null argument");
    DISCO_REQUIRE(n_sweeps > 0 && p_max >= 0 && (p_max == 0 || points) && m_max > 0, "voxelize_batched: bad sizes");
    DISCO_REQUIRE(point_stride >= 3, "voxelize_batched: points need >= 3 columns (got %d)", point_stride);
    const long long n_bits = (long long)dims[0] * dims[1] * dims[2];
    DISCO_REQUIRE(dims[0] > 0 && dims[1] > 0 && dims[2] > 0 && n_bits < (1ll << 30), "voxelize_batched: bad dims");
    const int n_words = (int)((n_bits + 31) / 32), n_blocks = (n_words + 1023) / 1024;
    cudaStream_t s = (cudaStream_t)stream;
    DISCO_CHECK_CUDA(cudaMemsetAsync(bitmap, 0, (size_t)n_sweeps * n_words * 4, s));
    DISCO_CHECK_CUDA(cudaMemsetAsync(voxel_indices, 0xFF, (size_t)n_sweeps * m_max * 3 * sizeof(int), s));   // padding rows = -1
    if (p_max > 0) {
        dim3 grid((p_max + 255) / 256, n_sweeps);
        voxel_mark_batched_kernel<<<grid, 256, 0, s>>>(points, n_points, p_max, point_stride, extents[0], extents[1], extents[2],
                                                       extents[3], extents[4], extents[5], voxel_size[0], voxel_size[1], voxel_size[2],
                                                       dims[0], dims[1], dims[2], n_words, bitmap);
        DISCO_CHECK_CUDA(cudaGetLastError());
    }
    dim3 grid2(n_blocks, n_sweeps);
    voxel_count_kernel<<<grid2, 1024, 0, s>>>(bitmap, n_words, n_blocks, block_count);
    DISCO_CHECK_CUDA(cudaGetLastError());
    voxel_compact_batched_kernel<<<grid2, 1024, 0, s>>>(bitmap, n_words, n_blocks, block_count, dims[1], dims[2], (int)n_bits, m_max,
                                                        voxel_indices, n_voxels);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

int disco_bev_scatter_launch(const int* voxel_indices, int n_voxels, const int* dims, float* bev_f32, void* act_hi,
                             int act_c, int precision, void* stream) {
    DISCO_REQUIRE(dims && (bev_f32 || act_hi), "bev_scatter: null argument");
    DISCO_REQUIRE(n_voxels >= 0 && (n_voxels == 0 || voxel_indices), "bev_scatter: bad indices");
    DISCO_REQUIRE(!act_hi || act_c >= dims[2], "bev_scatter: act_c %d < z dim %d", act_c, dims[2]);
    cudaStream_t s = (cudaStream_t)stream;
    const long long n_pix = (long long)dims[0] * dims[1];
    if (bev_f32) DISCO_CHECK_CUDA(cudaMemsetAsync(bev_f32, 0, (size_t)n_pix * dims[2] * 4, s));
    if (act_hi) DISCO_CHECK_CUDA(cudaMemsetAsync(act_hi, 0, (size_t)n_pix * act_c * 2, s));
    if (n_voxels > 0) {
        const uint16_t one = (precision == DISCO_PREC_BF16X3) ? 0x3F80 : 0x3C00;
        bev_scatter_kernel<<<(n_voxels + 255) / 256, 256, 0, s>>>(voxel_indices, n_voxels, dims[0], dims[1], dims[2],
                                                                  bev_f32, (uint16_t*)act_hi, act_c, one);
        DISCO_CHECK_CUDA(cudaGetLastError());
    }
    return DISCO_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Detection candidates: the per-anchor half of apply_nms_det (utils/detection_util.py:256-373)
// on the device -- foreground probability softmax(cls)[1], threshold (postprocess.py:85: scores > 0.7), box decode
// (bev_box_decode_torch :376-400) and the four rotated corners (obj_util.py:271-359) -- compacted with one atomic per
// candidate.  The reference copies ALL scores and ALL decoded boxes of every agent to the host (11 MB per agent) and
// filters there; this writes only the survivors (a few hundred boxes).  Order is restored by a sort on the scores.
// ---------------------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) det_candidates_kernel(const float* __restrict__ loc, const float* __restrict__ cls,
                                                             const float* __restrict__ anchors, long long anchors_per_agent,
                                                             long long anchor_agent_stride, int n_agents, float thresh,
                                                             int max_cand, int* __restrict__ count, float* __restrict__ corners,
                                                             float* __restrict__ scores, int* __restrict__ index) {
    const long long total = anchors_per_agent * n_agents;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const float2 z = __ldg(reinterpret_cast<const float2*>(cls) + e);
        const float m = fmaxf(z.x, z.y);
        const float e0 = expf(z.x - m), e1 = expf(z.y - m);
        const float score = e1 / (e0 + e1);
        if (!(score > thresh)) continue;
        const int agent = (int)(e / anchors_per_agent);
        const long long a = e - agent * anchors_per_agent;
        const int slot = atomicAdd(count + agent, 1);
        if (slot >= max_cand) continue;
        const float* p = loc + e * 6;
        const float* q = anchors + agent * anchor_agent_stride + a * 6;
        const float h = q[3] / expf(p[3]), w = q[2] / expf(p[2]);
        const float x = q[0] - w * p[0], y = q[1] - h * p[1];
        const float s = q[4] * p[5] + q[5] * p[4], c = q[5] * p[5] - q[4] * p[4];
        const float nx[4] = {-0.5f, 0.5f, 0.5f, -0.5f}, ny[4] = {0.5f, 0.5f, -0.5f, -0.5f};
        float* o = corners + ((long long)agent * max_cand + slot) * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float px = w * nx[k], py = h * ny[k];
            o[2 * k] = px * c + py * s + x;
            o[2 * k + 1] = -px * s + py * c + y;
        }
        scores[(long long)agent * max_cand + slot] = score;
        index[(long long)agent * max_cand + slot] = (int)a;
    }
}

}  // namespace

int disco_det_candidates_launch(const float* loc, const float* cls, const float* anchors, long long anchors_per_agent,
                                long long anchor_agent_stride, int n_agents, float thresh, int max_cand, int* count,
                                float* corners, float* scores, int* index, void* stream) {
    DISCO_REQUIRE(loc && cls && anchors && count && corners && scores && index, "det_candidates: null tensor");
    DISCO_REQUIRE(anchors_per_agent > 0 && n_agents > 0 && max_cand > 0, "det_candidates: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    DISCO_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * n_agents, s));
    long long blocks = (anchors_per_agent * n_agents + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    det_candidates_kernel<<<(unsigned)blocks, 256, 0, s>>>(loc, cls, anchors, anchors_per_agent, anchor_agent_stride, n_agents, thresh,
                                                           max_cand, count, corners, scores, index);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
