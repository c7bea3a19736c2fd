// Persistent, warp-specialised implicit-GEMM 3x3 / 1x1 convolution on the 5th-gen tensor cores
// (tcgen05.mma, TMEM accumulators) for sm_100a.
//
// Replaces, for the DiscoNet hot path, every F.conv2d/conv3d + BatchNorm(eval) + ReLU of the reference
// (Backbone.py:102-136 encode, :173-237 decode incl. the nearest-x2 upsample + channel concat on the
// *load* side, DetModelBase.py:283-351 heads, DiscoNet.py:148 PWF conv1_1) -- BN is folded into the
// packed weights / bias on the host (disconet_b200/plan.py).
//
// Work item = (n_tile, m_tile): MSUB x 128 output pixels (MSUB sub-tiles of 16 rows x 8 cols of one
// image, side by side; or 128 consecutive pixels each for 1x1) x block_n output channels,
// K = taps * C_in, fp32 accumulation in TMEM.  One CTA per SM loops over its items (static
// round-robin), so barrier set-up, TMEM allocation and pipeline fill are paid once per launch and the
// gather of item i+1 / epilogue of item i-1 overlap the MMAs of item i:
//
//   warps 0-3  epilogue: tcgen05.ld accumulator -> bias/ReLU -> 16-bit hi[/lo] or fp32 16-byte stores,
//              then release the accumulator buffer (TMEM is double buffered: 2 x MSUB x block_n cols)
//   warps 4-7  A producers, one *stage per warp* in flight (4 independent cp.async streams):
//              the input patch of a sub-tile (18x10 px for 3x3/s1, 33x17 for s2) is gathered ONCE per
//              channel block into shared memory in the UMMA "no-swizzle K-major" layout
//              [channel/8][pixel][8 ch] (16-byte core-matrix rows).  In that layout a filter tap (kh,kw)
//              is just a different 16-byte-aligned descriptor start address, so the nine taps reuse one
//              staged patch (L2->SMEM activation traffic ~1.4x the input instead of 9x).  Stride-2
//              convs de-interleave even/odd columns into two sub-planes; nearest-upsample + concat are
//              address arithmetic (src_up / two sources); zero padding is cp.async zero-fill.
//   warp 8     B loader (bulk-copy/TMA engine, cp.async.bulk + mbarrier complete_tx): host-packed
//              weight images, one contiguous block per (channel block, tap).  If the whole weight set
//              of the n_tile fits next to the A stages it is loaded ONCE and stays resident
//              ("stationary": all C_out<=64 layers); otherwise it is streamed through a ring, and
//              MSUB=2 halves that stream per MAC.
//   warp 9     TMEM allocation + single-thread MMA issue; tcgen05.commit releases smem stages and
//              publishes accumulators.
//
// precision DISCO_PREC_BF16X3 keeps activations/weights as bf16 hi+lo pairs and issues three MMAs
// per k-step (hi*hi + lo*hi + hi*lo) -> ~16 mantissa bits, which is what the <=1e-3 parity gate
// against the fp32 reference needs (plain fp16 operands measure 2-3e-3).
#include <stdlib.h>
#include <cuda.h>
#include "common.cuh"
#include "conv.h"

namespace {

// ---- TMA (cp.async.bulk.tensor) plumbing: the A operand of every non-upsampled source is fetched by the TMA engine as one
// box per (8-channel chunk, precision part) -- {8 ch, patch width, patch height, 1 image} of the NHWC tensor, zero-filled
// outside the image (= the conv's zero padding) -- straight into the UMMA no-swizzle K-major plane [pixel][8 ch].  One
// elected lane issues 4..16 instructions per K stage instead of ~1500 warp instructions of per-lane cp.async address math
// (round-2 ncu source view: the four producer warps were the busiest warps of the C_out <= 64 layers).  A nearest-x2
// upsampled source is fetched at its native (half) resolution into a small scratch box and expanded shared->shared.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// NHWC 16-bit tensor [N, H, W, C] -> 4-D map (innermost first: C, W, H, N); box {8, box_w, box_h, 1}; estride_w = 2 loads
// every second column (stride-2 convs de-interleave even / odd columns into two planes)
bool make_tmap4(CUtensorMap* m, const void* base, int C, int W, int H, int N, int box_w, int box_h, int estride_w) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {8, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, (cuuint32_t)estride_w, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// [pixels, C] view for the 1x1 convs; box {8, 128}
bool make_tmap2(CUtensorMap* m, const void* base, int C, long long pixels) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {8, 128};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
constexpr int kScrBox = 1024;   // scratch bytes per (chunk, part) low-resolution box (10 x 6 pixels x 16 B = 960, 128-B aligned)

constexpr int kEpiWarps = 8, kProdWarps = 4;   // two epilogue groups of 4 warps, each owning alternate items
constexpr int kMmaWarps = 2;   // two independent issue streams (each owns alternate items) when weights are stationary
constexpr int kThreads = (kEpiWarps + kProdWarps + 1 + kMmaWarps + 1) * 32;  // 512
constexpr int kWarpB = kEpiWarps + kProdWarps, kWarpMma = kWarpB + 1;
constexpr int kWarpChain = kWarpMma + kMmaWarps;   // issues the chained 1x1 MMAs (heads)
constexpr int kBias2Off = 384;                     // float index of the chain bias inside the bias region
constexpr int kMaxAcc = 4;
constexpr int kMaxStages = 8;
constexpr int kCtlBytes = 512;
constexpr int kBiasBytes = 2048;   // bias staged in shared memory (<= 512 output channels)
constexpr int kStageRow = 80;      // epilogue staging: per pixel 16 ch hi (32 B) + lo (32 B) + 16 B pad (bank spread)
constexpr int kStagePerWarp = 32 * kStageRow;
constexpr int kStageBytes = kEpiWarps * kStagePerWarp;   // 20 KB

// n / d for 0 <= n < 2^31 without the ~40-instruction integer division sequence (Granlund-Montgomery, the CUTLASS FastDivmod
// construction): the per-item tile decode sits on the critical path of single warps (round-2 trace: ~1000 cycles of divisions
// per item in the MMA and epilogue roles of the C_out <= 64 layers).
struct FastDiv {
    uint32_t mul, shr;
    int d;
};
FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = d;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    int l = 0;
    while ((1ll << l) < d) ++l;
    const int p = 31 + l;
    f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1) / (uint64_t)d);
    f.shr = (uint32_t)(p - 32);
    return f;
}
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
    return f.d <= 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr);
}

struct alignas(64) ConvGeom {
    CUtensorMap tmap[4];      // [source][part]: TMA views of the conv sources (valid when use_tma)
    int use_tma;              // A operand through the TMA engine (sources that are not zero-stuffed)
    int scratch_warp;         // bytes of low-resolution scratch per producer warp (nearest-upsampled sources), else 0
    int scr_bufs;             // 2: double-buffered (next box prefetched; 16-channel stages only: a 32-channel box pair would cost 64 KB), else 1
    int scratch_total;
    disco_conv_desc d;
    int ncb, ncb0;            // K stages total / from source 0
    int chunks, chunk_shift;  // c_blk/8
    int nparts;               // 1 (fp16) | 2 (bf16 hi+lo)
    int PIX;                  // staged patch pixels per sub-tile
    int plane, parplane;      // bytes
    int a_part_bytes, a_stage_bytes;
    int b_part_bytes, b_stage_bytes;
    int SA, SB;
    int sbo_a;
    int msub;                 // 128-pixel sub-tiles per item (1|2)
    int nacc;                 // accumulator buffers in TMEM (1|2|4)
    int nmma;                 // MMA issuing warps (2 when stationary: the single-thread issue stream, ~80 cycles per
                              // tcgen05.mma here, is the bottleneck of the small-N layers; the SMEM operand port allows ~40)
    int acc_stride;           // TMEM columns per accumulator (block_n rounded up to 32)
    int nprod;                // producer warps per A ring: min(4 / nmma, SAr) (a ring with fewer slots than independent
                              // producers would let one warp lap another through the parity alias)
    int SAr;                  // slots per A ring = SA / nrings
    int nrings;               // A rings: 2 = one private ring per issuer (issuers split ITEMS, stationary weights),
                              //          1 = one ring (single issuer, or issuers split the two SUB-TILES of each item)
    int by_sub;               // 1: streamed weights, MSUB = 2, N <= 64: issuer m handles sub-tile m of every item; both
                              //    read the same B stage (b_empty / acc_full count 2)
    int stationary;           // weights resident in smem
    int w_bytes;              // stationary: bytes of one n_tile's weights
    int tiles_h, tiles_w;
    int m_tiles, n_tiles, items;
    FastDiv fd_m_tiles, fd_per_img, fd_tiles_w;
    int w_iters;              // weight blocks per N tile (= K stages x taps; sub-pixel mode: 4 per source-0 stage, 9 per source-1 stage)
    int plane0, a_part0;      // sub-pixel mode: chunk / part strides of a SOURCE-0 stage (18 x 10 low-res patch)
    long long total_pix;      // n*h_out*w_out
    int tmem_cols;
    int smem_bytes;
    int grid;
    int chain;                // chained 1x1 conv active
    int chain_bn;             // its N tile (chain_c_out rounded up to 16)
    int a2_bytes, w2_bytes;   // per-group A2 staging / resident W2 image
    int acc2_col, acc2_stride;// TMEM columns of the chain accumulators (one per epilogue group)
    int direct;               // epilogue stores straight from registers (no shared-memory staging)
    int dbg;                  // debug (DISCO_CONV_DBG, MODE 4 timing experiments; results are WRONG when set): 1 no weight copies,
                              // 2 no A copies, 4 no output stores, 8 no MMAs
    long long* trace;         // debug: per-role clock64 stamps of CTA 0 (DISCO_CONV_TRACE), else null
};

struct __align__(8) SmemCtl {
    uint64_t a_full[kMaxStages];
    uint64_t a_empty[kMaxStages];
    uint64_t b_full[kMaxStages];
    uint64_t b_empty[kMaxStages];
    uint64_t acc_full[kMaxAcc];
    uint64_t acc_empty[kMaxAcc];
    uint64_t w_full;
    uint64_t w2_full;
    uint64_t a2_full[2];
    uint64_t acc2_full[2];
    uint64_t scr_full[8];     // low-resolution scratch box [producer warp p][buffer 0|1] has landed (TMA complete_tx)
    uint32_t tmem_base;
    uint32_t pad;
};
static_assert(sizeof(SmemCtl) <= kCtlBytes, "control block");

// 16 fp32 values -> bf16 hi (two uint4) and bf16 lo = bf16(v - hi) (two uint4) with the packed two-at-a-time converter: 6
// instructions per value pair instead of ~10 (round-2 trace: the bias/ReLU/split arithmetic of one 16-channel chunk took ~830
// cycles of an epilogue warp, the longest leg of the C_out <= 64 layers' epilogue).  Same round-to-nearest-even results.
__device__ __forceinline__ void split_pack16(const float* v, uint4& h0, uint4& h1, uint4& l0, uint4& l1) {
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t p, q;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[2 * i + 1]), "f"(v[2 * i]));   // upper half <- first source
        const float le = v[2 * i] - __uint_as_float(p << 16), lo_ = v[2 * i + 1] - __uint_as_float(p & 0xffff0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(q) : "f"(lo_), "f"(le));
        hw[i] = p; lw[i] = q;
    }
    h0 = make_uint4(hw[0], hw[1], hw[2], hw[3]); h1 = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    l0 = make_uint4(lw[0], lw[1], lw[2], lw[3]); l1 = make_uint4(lw[4], lw[5], lw[6], lw[7]);
}

struct Item {
    int n_tile, img, h0, w0;
    long long p0;
};

// output pixel of tile row m (0..127) of sub-tile `sub`
template <int MODE>
__device__ __forceinline__ void out_coords(const ConvGeom& g, const Item& it, int sub, int m, int& oh, int& ow) {
    oh = it.h0 + (m >> 3);
    ow = it.w0 + sub * 8 + (m & 7);
    if (MODE == 3) { oh = 2 * oh + g.d.sub_py; ow = 2 * ow + g.d.sub_px; }
    if (MODE == 4) { oh = 2 * (it.h0 + (m >> 3)) + (sub >> 1); ow = 2 * (it.w0 + (m & 7)) + (sub & 1); }   // sub = class 2*py + px
}

template <int MODE>
__device__ __forceinline__ Item decode_item(const ConvGeom& g, int item) {
    Item it;
    it.n_tile = fast_div(item, g.fd_m_tiles);  // n-major: concurrently running CTAs stream the same weights
    const int m_tile = item - it.n_tile * g.m_tiles;
    it.img = 0; it.h0 = 0; it.w0 = 0; it.p0 = 0;
    if (MODE != 2) {
        const int per_img = g.tiles_h * g.tiles_w;
        it.img = fast_div(m_tile, g.fd_per_img);
        const int rem = m_tile - it.img * per_img;
        const int th = fast_div(rem, g.fd_tiles_w);
        it.h0 = th * 16;
        it.w0 = (rem - th * g.tiles_w) * 8 * g.msub;
    } else {
        it.p0 = (long long)m_tile * 128 * g.msub;
    }
    return it;
}

// MODE 0: 3x3 stride 1 (patch 18x10) | MODE 1: 3x3 stride 2 (patch 33x17) | MODE 2: 1x1
// MODE 3: one output-parity class (sub_py, sub_px) of a 3x3 conv over cat(nearest_up2(src0), src1) ("sub-pixel" decomposition):
//         tiles of 16 x 8 class pixels (output (2a + py, 2b + px)); source 0 is read at its native half resolution as an
//         18 x 10 patch with only the 2 x 2 taps the class touches (weights pre-summed on the host), source 1 exactly like a
//         stride-2 conv whose input origin is shifted by (py, px)
// MODE 4: ALL FOUR output-parity classes of such a conv in one item ("fused sub-pixel", d.subpix == 2): an item is a 16 x 8 tile of
//         LOW-RES positions (a, b) = 512 output pixels; the four class accumulators (class 2*py + px -> output (2a + py, 2b + px))
//         sit side by side in TMEM and share every staged operand: a source-0 stage is the 18 x 10 low-res patch (4 pre-summed
//         taps per class = 16 class-taps, issued as 12 MMA pairs, instead of 4 tiles x 9 taps = 36), a source-1 stage the two
//         column-parity planes of the 34 x 18 full-resolution window (class (py, px), tap (kh, kw) reads plane (px + kw) & 1 at
//         row py + kh, column (px + kw) >> 1: 36 class-taps issued as 24 pairs, the window staged once instead of once per
//         class).  Weights are streamed in slots of nine units
//         (unit = [W_hi | W_lo] rows of one tap): source 0 [channel block][low-res tap row ty][py][chunk][(px, tx)] (+ one pad
//         unit), source 1 [channel block][kh][chunk][kw = 2, 1, 0] -- chunk-major inside a group, so two neighbouring units form
//         ONE B operand of twice the rows (the px-merged MMAs of the issuer).
template <int MODE, int KSTEPS, int PASSES>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvGeom g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem_raw);
    const uint32_t smem_base = smem_u32(smem_raw);
    // (chained layers never use the OUT_ACT store staging, so its 20 KB go to the A stages)
    const uint32_t scr_base = smem_base + kCtlBytes + kBiasBytes + (g.chain ? 0 : kStageBytes);
    const uint32_t a_base = scr_base + (uint32_t)g.scratch_total;
    const float* s_bias = reinterpret_cast<const float*>(smem_raw + kCtlBytes);
    const uint32_t b_base = a_base + g.SA * g.a_stage_bytes;
    const uint32_t w2_base = b_base + (uint32_t)g.w_bytes;          // chain only (stationary layers)
    const uint32_t a2_base = w2_base + (uint32_t)g.w2_bytes;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const disco_conv_desc& d = g.d;
    constexpr int PW = (MODE == 0 || MODE == 3 || MODE == 4) ? 10 : (MODE == 1) ? 17 : 128;
    constexpr int TAPS = (MODE == 2) ? 1 : 9;
    constexpr int STRIDE = (MODE == 1) ? 2 : 1;
    constexpr bool SPLIT = PASSES != 1;     // bf16 hi+lo operands
    constexpr bool STACKED = PASSES == 2;   // B image = [chunk][hi rows | lo rows][8]: hi*hi and hi*lo in ONE MMA of N = 2*block_n
    const int acc_cols = STACKED ? 2 * g.d.block_n : g.d.block_n;   // TMEM columns written per accumulator
    (void)acc_cols;

    // ---- one-time setup ------------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < g.SA; ++s) {
            mbar_init(smem_u32(&ctl->a_full[s]), 1);   // one arrival per stage: the TMA issuer's expect_tx, or lane 0 of a cp.async warp
            mbar_init(smem_u32(&ctl->a_empty[s]), MODE == 4 ? 2 : 1);      // MODE 4: both issuers read every stage
        }
        for (int s = 0; s < g.SB; ++s) {
            mbar_init(smem_u32(&ctl->b_full[s]), 1);
            mbar_init(smem_u32(&ctl->b_empty[s]), (g.by_sub || MODE == 4) ? 2 : 1);
        }
        for (int s = 0; s < kMaxAcc; ++s) {
            mbar_init(smem_u32(&ctl->acc_full[s]), (g.by_sub || MODE == 4) ? 2 : 1);
            mbar_init(smem_u32(&ctl->acc_empty[s]), (MODE == 4 ? 8 : 4) * 32);   // one epilogue group (4 warps) drains a buffer (MODE 4: both groups, two classes each)
        }
        mbar_init(smem_u32(&ctl->w_full), 1);
        mbar_init(smem_u32(&ctl->w2_full), 1);
        for (int s = 0; s < 8; ++s) mbar_init(smem_u32(&ctl->scr_full[s]), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&ctl->a2_full[s]), 4 * 32);
            mbar_init(smem_u32(&ctl->acc2_full[s]), 1);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < g.n_tiles * g.d.block_n; i += kThreads)
        reinterpret_cast<float*>(smem_raw + kCtlBytes)[i] = g.d.bias[i];
    if (g.chain)
        for (int i = tid; i < g.chain_bn; i += kThreads)
            reinterpret_cast<float*>(smem_raw + kCtlBytes)[kBias2Off + i] = g.d.chain_bias[i];
    if (warp == kWarpMma) {
        tmem_alloc(smem_u32(&ctl->tmem_base), (uint32_t)g.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = ctl->tmem_base;
    const int stages_per_item = g.ncb * g.msub;
    // exact 0/1 occupancy input (first encoder conv): the lo plane is identically zero -> neither fetched nor multiplied
    const bool skip_lo = SPLIT && g.d.src_lo_nonzero != nullptr && *g.d.src_lo_nonzero == 0;
    const int a_parts = skip_lo ? 1 : g.nparts;
    // trace layout: [role 0..3][item 0..63][4 stamps]; roles: 0 epilogue, 1 producer warp 0, 2 MMA, 3 B loader
#define TRACE(role, it_, k) do { if (g.trace && blockIdx.x == 0 && (it_) < 64 && lane == 0) g.trace[((role) * 64 + (it_)) * 4 + (k)] = clock64(); } while (0)

    if (warp < kEpiWarps) {
        // =========================== epilogue =====================================================
        int iacc = 0;
        int buf = 0;               // == iacc % nacc, acc_par == (iacc / nacc) & 1 (kept by counters: no divisions per item)
        uint32_t acc_par = 0;
        const int egrp = warp >> 2;   // the epilogue of one item is latency-bound (~3000 cycles); two groups overlap two items
        for (int item = blockIdx.x; item < g.items; item += gridDim.x, ++iacc, buf = (buf + 1 == g.nacc) ? 0 : buf + 1, acc_par ^= (buf == 0)) {
            if (MODE != 4 && (iacc & 1) != egrp) continue;
            const Item it = decode_item<MODE>(g, item);
            if (warp == 0) TRACE(0, iacc, 0);
            mbar_wait(smem_u32(&ctl->acc_full[buf]), acc_par);
            tc_fence_after();
            if (warp == 0) TRACE(0, iacc, 1);
            const int m = (warp & 3) * 32 + lane;   // TMEM lane == pixel row of the tile; warp w may touch lanes 32*(w%4)..+31
            const int nsub = (MODE == 4) ? 4 : g.msub;                   // accumulators per buffer
            const int sub_b = (MODE == 4) ? 2 * egrp : 0, sub_e = (MODE == 4) ? 2 * egrp + 2 : g.msub;
            for (int sub = sub_b; sub < sub_e; ++sub) {
                bool valid;
                long long pixel;
                if (MODE != 2) {
                    int oh, ow;
                    out_coords<MODE>(g, it, sub, m, oh, ow);
                    valid = (oh < d.h_out) && (ow < d.w_out);
                    pixel = ((long long)it.img * d.h_out + oh) * d.w_out + ow;
                } else {
                    pixel = it.p0 + sub * 128 + m;
                    valid = pixel < g.total_pix;
                }
                const uint32_t t_lane = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) +
                                        (uint32_t)((buf * nsub + sub) * g.acc_stride);
                const int nchunks = d.block_n / 16;
                // staged stores: pass r writes pixel r*16 + (lane & 15), 16-byte piece lane >> 4 -- element offsets / validity per pass
                // depend on the item only, not on the channel chunk
                long long st_off[2];
                bool st_ok[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int m2 = (warp & 3) * 32 + r * 16 + (lane & 15);
                    long long pix2;
                    if (MODE != 2) {
                        int oh, ow;
                        out_coords<MODE>(g, it, sub, m2, oh, ow);
                        st_ok[r] = (oh < d.h_out) && (ow < d.w_out);
                        pix2 = ((long long)it.img * d.h_out + oh) * d.w_out + ow;
                    } else {
                        pix2 = it.p0 + sub * 128 + m2;
                        st_ok[r] = pix2 < g.total_pix;
                    }
                    st_off[r] = pix2 * d.c_out + (lane >> 4) * 8;
                }
                uint32_t nxt[16], nxt2[16];
                tmem_ld16(t_lane, nxt);
                if (STACKED) tmem_ld16(t_lane + (uint32_t)d.block_n, nxt2);
                for (int j = 0; j < nchunks; ++j) {
                    uint32_t raw[16];
                    tmem_ld_wait();
                    if (warp == 0 && sub == 0 && j == 0) TRACE(0, iacc, 3);
                    if (warp == 0 && sub == 0 && j == 1) TRACE(3, iacc, 3);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        raw[i] = STACKED ? __float_as_uint(__uint_as_float(nxt[i]) + __uint_as_float(nxt2[i])) : nxt[i];
                    if (j + 1 < nchunks) {   // software pipeline: next chunk's TMEM read overlaps this chunk's math/stores
                        tmem_ld16(t_lane + (uint32_t)((j + 1) * 16), nxt);
                        if (STACKED) tmem_ld16(t_lane + (uint32_t)(d.block_n + (j + 1) * 16), nxt2);
                    }
                    const int nb = it.n_tile * d.block_n + j * 16;
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + nb + 4 * q);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float x = __uint_as_float(raw[4 * q + i]) + bb[i];
                            v[4 * q + i] = d.relu ? fmaxf(x, 0.f) : x;
                        }
                    }
                    if (g.chain) {
                        // chained 1x1: the ReLU'd tile becomes the A operand of a second MMA -- write bf16 hi/lo
                        // straight into the UMMA K-major layout [part][channel/8][pixel][8 ch] of this group's A2 buffer
                        uint4 ph0, ph1, pl0, pl1;
                        split_pack16(v, ph0, ph1, pl0, pl1);
                        uint8_t* a2 = smem_raw + (a2_base - smem_base) + egrp * g.a2_bytes;
                        const int part_b = (d.block_n / 8) * 2048;
                        uint4* dsth = reinterpret_cast<uint4*>(a2 + (2 * j) * 2048 + m * 16);
                        dsth[0] = ph0;
                        dsth[128] = ph1;              // next channel chunk: +2048 B
                        uint4* dstl = reinterpret_cast<uint4*>(a2 + part_b + (2 * j) * 2048 + m * 16);
                        dstl[0] = pl0;
                        dstl[128] = pl1;
                    } else if (g.direct && d.out_mode == DISCO_OUT_ACT && SPLIT) {
                        // Register-direct stores (no shared-memory staging): a lane holds one pixel's 16 channels = one 32-byte
                        // sector per plane.  Lane pairs swap halves (one shuffle per plane) so that every store instruction writes
                        // BOTH halves of a sector -- a lane storing its own two halves in two instructions writes half sectors
                        // (measured: slower than staging).
                        const bool odd = lane & 1;
                        bool val_e, val_o;
                        long long pix_e, pix_o;
                        {
                            const int me = m & ~1, mo = m | 1;
                            if (MODE != 2) {
                                int oh, ow;
                                out_coords<MODE>(g, it, sub, me, oh, ow);
                                val_e = (oh < d.h_out) && (ow < d.w_out);
                                pix_e = ((long long)it.img * d.h_out + oh) * d.w_out + ow;
                                out_coords<MODE>(g, it, sub, mo, oh, ow);
                                val_o = (oh < d.h_out) && (ow < d.w_out);
                                pix_o = ((long long)it.img * d.h_out + oh) * d.w_out + ow;
                            } else {
                                pix_e = it.p0 + sub * 128 + me; val_e = pix_e < g.total_pix;
                                pix_o = it.p0 + sub * 128 + mo; val_o = pix_o < g.total_pix;
                            }
                        }
                        uint4 h0, h1, l0, l1;
                        split_pack16(v, h0, h1, l0, l1);
                        uint4 sh = odd ? h0 : h1, sl = odd ? l0 : l1;       // even lanes give away their upper half, odd lanes their lower half
                        sh.x = __shfl_xor_sync(0xffffffffu, sh.x, 1); sh.y = __shfl_xor_sync(0xffffffffu, sh.y, 1);
                        sh.z = __shfl_xor_sync(0xffffffffu, sh.z, 1); sh.w = __shfl_xor_sync(0xffffffffu, sh.w, 1);
                        sl.x = __shfl_xor_sync(0xffffffffu, sl.x, 1); sl.y = __shfl_xor_sync(0xffffffffu, sl.y, 1);
                        sl.z = __shfl_xor_sync(0xffffffffu, sl.z, 1); sl.w = __shfl_xor_sync(0xffffffffu, sl.w, 1);
                        if (nb < d.c_out && !(g.dbg & 4)) {
                            uint16_t* oe = reinterpret_cast<uint16_t*>(d.out[0]) + pix_e * d.c_out + nb + (odd ? 8 : 0);
                            uint16_t* oo = reinterpret_cast<uint16_t*>(d.out[0]) + pix_o * d.c_out + nb + (odd ? 8 : 0);
                            if (val_e) {      // the even lane's pixel: own lower half | partner's copy of its upper half
                                *reinterpret_cast<uint4*>(oe) = odd ? sh : h0;
                                *reinterpret_cast<uint4*>(oe + d.out_lo_off) = odd ? sl : l0;
                            }
                            if (val_o) {      // the odd lane's pixel
                                *reinterpret_cast<uint4*>(oo) = odd ? h1 : sh;
                                *reinterpret_cast<uint4*>(oo + d.out_lo_off) = odd ? l1 : sl;
                            }
                        }
                    } else if (d.out_mode == DISCO_OUT_ACT) {
                        // Stage the warp's 32 pixels x 16 channels in shared memory, then store row-wise: each
                        // store instruction writes 8 neighbouring pixels x 32 B (full sectors; one contiguous 512 B
                        // run per two chunks when c_out == 32) instead of 32 half-filled sectors 2*c_out bytes apart.
                        uint8_t* st = smem_raw + kCtlBytes + kBiasBytes + warp * kStagePerWarp;
                        uint4 h0v, h1v, l0v, l1v;
                        if (SPLIT) {
                            split_pack16(v, h0v, h1v, l0v, l1v);
                        } else {
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                w[i] = (uint32_t)f32_to_f16_bits(v[2 * i]) | ((uint32_t)f32_to_f16_bits(v[2 * i + 1]) << 16);
                            h0v = make_uint4(w[0], w[1], w[2], w[3]); h1v = make_uint4(w[4], w[5], w[6], w[7]);
                            l0v = h0v; l1v = h1v;
                        }
                        if (warp == 0 && sub == 0 && j == 0) TRACE(3, iacc, 0);
                        __syncwarp();   // previous chunk's read-back is done
                        uint4* row = reinterpret_cast<uint4*>(st + lane * kStageRow);
                        row[0] = h0v; row[1] = h1v;
                        if (SPLIT) { row[2] = l0v; row[3] = l1v; }
                        __syncwarp();
                        if (warp == 0 && sub == 0 && j == 0) TRACE(3, iacc, 1);
                        if (nb < d.c_out) {
#pragma unroll
                            for (int r = 0; r < 2; ++r) {   // 2 passes x 16 pixels x 2 pieces of 16 B
                                // lanes 0-15 read piece 0 of 16 pixels, lanes 16-31 piece 1: every quarter-warp of the LDS.128 reads
                                // 8 rows at the 80-byte row pitch = 8 distinct 4-bank groups (the (lane >> 1, lane & 1) mapping was a
                                // 2-way conflict: round-2 source view, 2x the ideal wavefronts on these four loads)
                                const int px = r * 16 + (lane & 15), piece = lane >> 4;
                                if (st_ok[r]) {
                                    const uint4* srow = reinterpret_cast<const uint4*>(st + px * kStageRow);
                                    uint16_t* o = reinterpret_cast<uint16_t*>(d.out[0]) + st_off[r] + nb;
                                    *reinterpret_cast<uint4*>(o) = srow[piece];
                                    if (SPLIT) *reinterpret_cast<uint4*>(o + d.out_lo_off) = srow[2 + piece];
                                }
                            }
                        }
                        if (warp == 0 && sub == 0 && j == 0) TRACE(3, iacc, 2);
                    } else if (valid) {
                        const int c1 = d.c_out - d.out_split;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = nb + 4 * q;
                            if (c >= d.c_out) continue;
                            float* dst = (c < d.out_split)
                                             ? reinterpret_cast<float*>(d.out[0]) + pixel * d.out_split + c
                                             : reinterpret_cast<float*>(d.out[1]) + pixel * c1 + (c - d.out_split);
                            *reinterpret_cast<float4*>(dst) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            if (warp == 0) TRACE(0, iacc, 2);
            mbar_arrive(smem_u32(&ctl->acc_empty[buf]));
            if (g.chain) {
                // hand the A2 tile to the chain issuer, then drain its accumulator: bias2, fp32 NHWC stores
                fence_proxy_async_smem();
                mbar_arrive(smem_u32(&ctl->a2_full[egrp]));
                mbar_wait(smem_u32(&ctl->acc2_full[egrp]), (uint32_t)(iacc >> 1) & 1u);
                tc_fence_after();
                const int mrow = (warp & 3) * 32 + lane;
                int oh, ow;
                out_coords<MODE>(g, it, 0, mrow, oh, ow);
                const bool ok = (MODE != 2) ? ((oh < d.h_out) && (ow < d.w_out)) : (it.p0 + mrow < g.total_pix);
                const long long px = (MODE != 2) ? ((long long)it.img * d.h_out + oh) * d.w_out + ow : it.p0 + mrow;
                const uint32_t t2 = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) +
                                    (uint32_t)(g.acc2_col + egrp * g.acc2_stride);
                const int c1 = d.chain_c_out - d.out_split;
                const int nch2 = g.chain_bn / 16;
                uint32_t n0[16], n1[16];
                tmem_ld16(t2, n0);
                tmem_ld16(t2 + (uint32_t)g.chain_bn, n1);                  // stacked: columns [N2, 2*N2) = A2_hi * W2_lo
                for (int j = 0; j < nch2; ++j) {
                    float r0[16];
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) r0[i] = __uint_as_float(n0[i]) + __uint_as_float(n1[i]);
                    if (j + 1 < nch2) {      // next chunk's TMEM reads overlap this chunk's bias add and stores
                        tmem_ld16(t2 + (uint32_t)((j + 1) * 16), n0);
                        tmem_ld16(t2 + (uint32_t)(g.chain_bn + (j + 1) * 16), n1);
                    }
                    if (ok) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c = j * 16 + 4 * q;
                            if (c >= d.chain_c_out) continue;
                            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + kBias2Off + c);
                            float o4[4] = {r0[4 * q] + b4.x, r0[4 * q + 1] + b4.y, r0[4 * q + 2] + b4.z, r0[4 * q + 3] + b4.w};
                            if (d.chain_relu) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) o4[i] = fmaxf(o4[i], 0.f);
                            }
                            float* dst = (c < d.out_split)
                                             ? reinterpret_cast<float*>(d.out[0]) + px * d.out_split + c
                                             : reinterpret_cast<float*>(d.out[1]) + px * c1 + (c - d.out_split);
                            *reinterpret_cast<float4*>(dst) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                        }
                    }
                }
                tc_fence_before();
            }
        }
    } else if (warp < kEpiWarps + kProdWarps) {
        // =========================== A producers: warp p owns stages p, p+4, p+8, ... ===============
        const int pw = warp - kEpiWarps;
        const int per_part = g.PIX << g.chunk_shift;
        // A-stage rings: with two MMA issuers every issuer owns a private ring (slots [r*SAr, (r+1)*SAr)) fed with
        // the stages of "its" items only -- an mbarrier ring is only safe with ONE in-order consumer (a second
        // consumer could be a full wrap ahead and pass the parity wait on a stale phase).  Producer warp pw serves
        // ring pw % nmma and, inside it, the stages q with q % nprod == pw / nmma (nprod <= SAr, same argument).
        const int ring = pw % g.nrings, pr = pw / g.nrings;
        // nearest-upsampled sources: the low-resolution box of this warp's NEXT upsampled stage is fetched into the other half of
        // its scratch while the current one is expanded (and while the warp waits for that stage's ring slot), so the L2 latency
        // of the box is off the stage's critical path
        uint32_t scr_par[2] = {0u, 0u};
        int up_n = 0;          // upsampled stages handled so far (buffer = up_n & 1)
        bool up_pf = false;    // the box of the stage about to be handled is already in flight
        const uint32_t scr_half = (g.scr_bufs == 2) ? (uint32_t)g.scratch_warp >> 1 : 0u;
        auto issue_scratch = [&](int item_x, int s_x, int buf) {     // lane 0 only
            const Item ix = decode_item<MODE>(g, item_x);
            const int cb_x = s_x / g.msub, sub_x = s_x - cb_x * g.msub;
            const int sidx_x = (cb_x < g.ncb0) ? 0 : 1;
            const int cofs_x = (sidx_x ? cb_x - g.ncb0 : cb_x) * d.c_blk;
            const int hi_x = ix.h0 - 1, wi_x = ix.w0 + sub_x * 8 - 1;
            const uint32_t sbar = smem_u32(&ctl->scr_full[pw * 2 + buf]);
            const uint32_t dst = scr_base + (uint32_t)pw * g.scratch_warp + (uint32_t)buf * scr_half;
            mbar_arrive_expect_tx(sbar, (uint32_t)(g.nparts * g.chunks) * 960u);
            for (int part = 0; part < g.nparts; ++part)
                for (int chunk = 0; chunk < g.chunks; ++chunk)
                    tma_load_4d(dst + (uint32_t)(part * g.chunks + chunk) * kScrBox, &g.tmap[sidx_x * 2 + part], cofs_x + chunk * 8,
                                wi_x >> 1, hi_x >> 1, ix.img, sbar);
        };
        int iacc_p = 0;
        for (int item = blockIdx.x; item < g.items; item += gridDim.x, ++iacc_p) {
            if (pr >= g.nprod) break;
            if ((iacc_p % g.nrings) != ring) continue;
            const Item it = decode_item<MODE>(g, item);
            const int q0 = (iacc_p / g.nrings) * stages_per_item;   // ring-local index of this item's first stage
            const int s0 = (pr - (q0 % g.nprod) + g.nprod) % g.nprod;
            for (int s = s0; s < stages_per_item; s += g.nprod) {
                const int q = q0 + s;
                const int cb = s / g.msub, sub = s - cb * g.msub;
                const int sa = ring * g.SAr + q % g.SAr;
                if (pw == 0) TRACE(1, q / g.nprod, 0);
                mbar_wait(smem_u32(&ctl->a_empty[sa]), ((uint32_t)(q / g.SAr) & 1u) ^ 1u);
                if (pw == 0) TRACE(1, q / g.nprod, 1);
                const int sidx = (cb < g.ncb0) ? 0 : 1;
                const int cbl = sidx ? cb - g.ncb0 : cb;
                const uint16_t* __restrict__ src = reinterpret_cast<const uint16_t*>(sidx ? d.src[1] : d.src[0]);
                const int Cs = sidx ? d.src_c[1] : d.src_c[0];
                const int upm = sidx ? d.src_up[1] : d.src_up[0];
                const int up = upm ? 1 : 0;          // 1: nearest x2 upsample, 2: zero-stuffed x2 (transposed stride-2 conv)
                const bool stuff = upm == 2;
                const int Hs = d.h_in >> up, Ws = d.w_in >> up;
                const long long lo_off = sidx ? d.src_lo_off[1] : d.src_lo_off[0];
                const uint32_t stage = a_base + sa * g.a_stage_bytes;
                const int cofs = cbl * d.c_blk;
                const int hi0 = it.h0 * STRIDE - 1, wi0 = (it.w0 + sub * 8) * STRIDE - 1;
                if (MODE == 3) {
                    // ---- sub-pixel class: source 0 at native half resolution (18 x 10 box), source 1 as the two stride-2
                    //      column-parity planes of a 33 x 17 window whose origin is shifted by (py, px) ----
                    if (lane == 0) {
                        const uint32_t bar = smem_u32(&ctl->a_full[sa]);
                        const int a0 = it.h0, b0 = it.w0 + sub * 8;
                        if (sidx == 0) {
                            mbar_arrive_expect_tx(bar, (uint32_t)(a_parts * g.chunks) * 2880u);
                            for (int part = 0; part < a_parts; ++part)
                                for (int chunk = 0; chunk < g.chunks; ++chunk)
                                    tma_load_4d(stage + (uint32_t)part * g.a_part0 + (uint32_t)chunk * g.plane0, &g.tmap[part],
                                                cofs + chunk * 8, b0 - 1, a0 - 1, it.img, bar);
                        } else {
                            mbar_arrive_expect_tx(bar, (uint32_t)(a_parts * g.chunks) * 9504u);
                            const int hs = 2 * a0 - 1 + d.sub_py, wsx = 2 * b0 - 1 + d.sub_px;
                            for (int part = 0; part < a_parts; ++part)
                                for (int chunk = 0; chunk < g.chunks; ++chunk) {
                                    const uint32_t dst = stage + (uint32_t)part * g.a_part_bytes + (uint32_t)chunk * g.plane;
                                    tma_load_4d(dst, &g.tmap[2 + part], cofs + chunk * 8, wsx, hs, it.img, bar);
                                    tma_load_4d(dst + g.parplane, &g.tmap[2 + part], cofs + chunk * 8, wsx + 1, hs, it.img, bar);
                                }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                if (MODE == 4) {
                    // ---- fused sub-pixel item: source 0 = 18 x 10 low-res patch, source 1 = column-parity planes (34 rows x 9) of the
                    //      34 x 18 full-resolution window whose origin is (2*a0 - 1, 2*b0 - 1) ----
                    if (lane == 0) {
                        const uint32_t bar = smem_u32(&ctl->a_full[sa]);
                        const int a0 = it.h0, b0 = it.w0;
                        if (g.dbg & 2) {
                            mbar_arrive(bar);
                        } else if (sidx == 0) {
                            mbar_arrive_expect_tx(bar, (uint32_t)(g.nparts * g.chunks) * 2880u);
                            for (int part = 0; part < g.nparts; ++part)
                                for (int chunk = 0; chunk < g.chunks; ++chunk)
                                    tma_load_4d(stage + (uint32_t)part * g.a_part0 + (uint32_t)chunk * g.plane0, &g.tmap[part],
                                                cofs + chunk * 8, b0 - 1, a0 - 1, it.img, bar);
                        } else {
                            mbar_arrive_expect_tx(bar, (uint32_t)(g.nparts * g.chunks) * (2u * 34u * 9u * 16u));
                            const int hs = 2 * a0 - 1, wsx = 2 * b0 - 1;
                            for (int part = 0; part < g.nparts; ++part)
                                for (int chunk = 0; chunk < g.chunks; ++chunk) {
                                    const uint32_t dst = stage + (uint32_t)part * g.a_part_bytes + (uint32_t)chunk * g.plane;
                                    tma_load_4d(dst, &g.tmap[2 + part], cofs + chunk * 8, wsx, hs, it.img, bar);
                                    tma_load_4d(dst + g.parplane, &g.tmap[2 + part], cofs + chunk * 8, wsx + 1, hs, it.img, bar);
                                }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                if (g.use_tma && upm == 0) {
                    // ---- TMA: one box per (chunk, part); out-of-image pixels arrive as zeros (= the conv padding) ----
                    if (lane == 0) {
                        const uint32_t bar = smem_u32(&ctl->a_full[sa]);
                        constexpr uint32_t kBox = (MODE == 0) ? 18u * 10u * 16u : (MODE == 1) ? 2u * 33u * 9u * 16u : 128u * 16u;
                        mbar_arrive_expect_tx(bar, (uint32_t)(a_parts * g.chunks) * kBox);
                        for (int part = 0; part < a_parts; ++part) {
                            const CUtensorMap* tm = &g.tmap[sidx * 2 + part];
                            for (int chunk = 0; chunk < g.chunks; ++chunk) {
                                const uint32_t dst = stage + (uint32_t)part * g.a_part_bytes + (uint32_t)chunk * g.plane;
                                const int c0 = cofs + chunk * 8;
                                if (MODE == 0) {
                                    tma_load_4d(dst, tm, c0, wi0, hi0, it.img, bar);
                                } else if (MODE == 1) {
                                    tma_load_4d(dst, tm, c0, wi0, hi0, it.img, bar);                    // even patch columns
                                    tma_load_4d(dst + g.parplane, tm, c0, wi0 + 1, hi0, it.img, bar);   // odd patch columns
                                } else {
                                    tma_load_2d(dst, tm, c0, (int)(it.p0 + sub * 128), bar);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    continue;
                }
                if (g.use_tma && upm == 1 && MODE == 0) {
                    // ---- nearest-x2 upsampled source: TMA the 10 x 6 low-resolution box, expand shared -> shared ------
                    const int buf = (g.scr_bufs == 2) ? (up_n & 1) : 0;
                    const uint32_t scr = scr_base + (uint32_t)pw * g.scratch_warp + (uint32_t)buf * scr_half;
                    const int nbox = g.nparts * g.chunks;
                    // this warp's next upsampled stage (same item or a later one of its ring)
                    int it2 = item, ia2 = iacc_p, s2 = s + g.nprod;
                    bool found = false;
                    for (int guard = 0; guard < 32 && !found && g.scr_bufs == 2; ++guard) {
                        if (s2 >= stages_per_item) {
                            do { it2 += gridDim.x; ++ia2; } while (it2 < g.items && (ia2 % g.nrings) != ring);
                            if (it2 >= g.items) break;
                            const int q02 = (ia2 / g.nrings) * stages_per_item;
                            s2 = (pr - (q02 % g.nprod) + g.nprod) % g.nprod;
                            if (s2 >= stages_per_item) continue;
                        }
                        const int cb2 = s2 / g.msub;
                        if (((cb2 < g.ncb0) ? d.src_up[0] : d.src_up[1]) == 1) found = true;
                        else s2 += g.nprod;
                    }
                    if (lane == 0) {
                        if (!up_pf) issue_scratch(item, s, buf);
                        if (found) issue_scratch(it2, s2, buf ^ 1);
                    }
                    up_pf = found;
                    mbar_wait(smem_u32(&ctl->scr_full[pw * 2 + buf]), scr_par[buf]);
                    scr_par[buf] ^= 1u;
                    ++up_n;
                    // patch pixel (r, c) <- low-res pixel ((r + 1) >> 1, (c + 1) >> 1) of the box (hi0, wi0 are odd)
                    for (int idx = lane; idx < nbox * 180; idx += 32) {
                        const int pc = idx / 180, rem = idx - pc * 180;
                        const int r = rem / 10, c = rem - r * 10;
                        const int part = pc >> g.chunk_shift, chunk = pc & (g.chunks - 1);
                        const uint4 v = lds128(scr + (uint32_t)pc * kScrBox + (uint32_t)((((r + 1) >> 1) * 6 + ((c + 1) >> 1)) * 16));
                        sts128(stage + (uint32_t)part * g.a_part_bytes + (uint32_t)chunk * g.plane + (uint32_t)(r * 160 + c * 16), v);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&ctl->a_full[sa]));
                    continue;
                }
                if (MODE != 2 && MODE != 3 && MODE != 4) {
                    // Row-wise gather: the (column, chunk) a lane handles is the same for every patch row, so
                    // its shared/global offsets are computed once per stage and each row only adds its base.
                    constexpr int PH = (MODE == 0) ? 18 : 33;
                    constexpr int ROWPITCH = (MODE == 0) ? 160 : 144;
                    constexpr int NJ = (MODE == 0) ? 3 : 5;     // ceil(PW * max chunks / 32)
                    const int per_row = PW << g.chunk_shift;
                    uint32_t soff[NJ];
                    int goff[NJ];
                    bool act[NJ], vcol[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int e = lane + 32 * j;
                        const int c = e >> g.chunk_shift, chunk = e & (g.chunks - 1);
                        const int wi = wi0 + c;
                        act[j] = e < per_row;
                        vcol[j] = (wi >= 0) && (wi < d.w_in) && !(stuff && (wi & 1));
                        soff[j] = (uint32_t)chunk * g.plane +
                                  ((MODE == 0) ? (uint32_t)c * 16u
                                               : (uint32_t)(c & 1) * g.parplane + (uint32_t)(c >> 1) * 16u);
                        goff[j] = (wi >> up) * Cs + cofs + chunk * 8;
                    }
                    const long long img_base = (long long)it.img * Hs;
                    const long long row_stride = (long long)Ws * Cs;
#pragma unroll 2
                    for (int r = 0; r < PH; ++r) {
                        const int hi = hi0 + r;
                        const bool rv = (hi >= 0) && (hi < d.h_in) && !(stuff && (hi & 1));
                        const uint16_t* rowp = src + (img_base + (hi >> up)) * row_stride;
                        const uint32_t drow = stage + (uint32_t)r * ROWPITCH;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (!act[j]) continue;
                            const bool valid = rv && vcol[j];
                            const uint16_t* gp = valid ? rowp + goff[j] : src;
                            cp_async16(drow + soff[j], gp, valid ? 16u : 0u);
                            if (SPLIT) cp_async16(drow + soff[j] + g.a_part_bytes, valid ? gp + lo_off : src, valid ? 16u : 0u);
                        }
                    }
                } else if (MODE == 2) {
                    const long long pbase = it.p0 + sub * 128;
                    for (int e = lane; e < per_part; e += 32) {
                        const int chunk = e & (g.chunks - 1);
                        const int pix = e >> g.chunk_shift;
                        const long long p = pbase + pix;
                        const bool valid = p < g.total_pix;
                        const uint32_t dst = stage + chunk * g.plane + (uint32_t)pix * 16u;
                        const uint16_t* gp = valid ? (src + p * Cs + cofs + chunk * 8) : src;
                        cp_async16(dst, gp, valid ? 16u : 0u);
                        if (SPLIT) cp_async16(dst + g.a_part_bytes, valid ? (gp + lo_off) : src, valid ? 16u : 0u);
                    }
                }
                cp_async_commit();
                if (pw == 0) TRACE(1, q / g.nprod, 2);
                cp_async_wait<0>();
                fence_proxy_async_smem();
                if (pw == 0) TRACE(1, q / g.nprod, 3);
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ctl->a_full[sa]));
            }
        }
    } else if (warp == kWarpB) {
        // =========================== B loader (bulk copy engine) ===================================
        if (lane == 0) {
            const int iters_per_tile = g.w_iters;
            if (g.chain) {
                const uint32_t bar2 = smem_u32(&ctl->w2_full);
                mbar_arrive_expect_tx(bar2, (uint32_t)g.w2_bytes);
                bulk_g2s(w2_base, d.chain_wpack, (uint32_t)g.w2_bytes, bar2);
            }
            if (g.stationary) {
                // n_tiles == 1: the whole packed weight set becomes resident
                const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.wpack);
                const uint32_t bar = smem_u32(&ctl->w_full);
                mbar_arrive_expect_tx(bar, (uint32_t)g.w_bytes);
                for (int off = 0; off < g.w_bytes; off += 32768) {
                    const int n = (g.w_bytes - off < 32768) ? g.w_bytes - off : 32768;
                    bulk_g2s(b_base + off, wp + off, (uint32_t)n, bar);
                }
            } else {
                int sb = 0;              // ring slot and wrap parity by counters (no divisions per slot)
                uint32_t sb_par = 0;
                for (int item = blockIdx.x; item < g.items; item += gridDim.x) {
                    const int n_tile = fast_div(item, g.fd_m_tiles);
                    const uint8_t* wp = reinterpret_cast<const uint8_t*>(d.wpack) +
                                        (size_t)n_tile * iters_per_tile * g.b_stage_bytes;
                    for (int t = 0; t < iters_per_tile; ++t) {
                        mbar_wait(smem_u32(&ctl->b_empty[sb]), sb_par ^ 1u);
                        const uint32_t bar = smem_u32(&ctl->b_full[sb]);
                        if (MODE == 4 && (g.dbg & 1)) {
                            mbar_arrive(bar);
                        } else {
                            mbar_arrive_expect_tx(bar, (uint32_t)g.b_stage_bytes);
                            bulk_g2s(b_base + sb * g.b_stage_bytes, wp + (size_t)t * g.b_stage_bytes,
                                     (uint32_t)g.b_stage_bytes, bar);
                        }
                        if (++sb == g.SB) { sb = 0; sb_par ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == kWarpChain) {
        // =========================== chained 1x1 issuer ==========================================
        if (g.chain) {
            const uint32_t idesc1 = umma_idesc_f16(1, 128, g.chain_bn);        // A2_lo * W2_hi
            const uint32_t idesc2 = umma_idesc_f16(1, 128, 2 * g.chain_bn);    // A2_hi * [W2_hi; W2_lo]
            const uint32_t a_hi = (128u >> 4) | (1u << 14);                      // SBO 128 B
            const uint32_t b_hi = (128u >> 4) | (1u << 14);
            const uint32_t lbo_b16 = 2u * (uint32_t)g.chain_bn;                  // chunk stride of W2: 2 parts * N2 rows * 16 B
            const uint32_t a_lo_c = (2048u >> 4) << 16;                          // LBO: next channel chunk
            const uint32_t b_lo_c = lbo_b16 << 16;
            const uint32_t a_part16 = ((uint32_t)(g.d.block_n / 8) * 2048u) >> 4;
            const int ks2 = g.d.block_n / 16;
            mbar_wait(smem_u32(&ctl->w2_full), 0);
            int iacc = 0;
            for (int item = blockIdx.x; item < g.items; item += gridDim.x, ++iacc) {
                const int grp = iacc & 1;
                mbar_wait(smem_u32(&ctl->a2_full[grp]), (uint32_t)(iacc >> 1) & 1u);
                tc_fence_after();
                const uint32_t td = tmem_d + (uint32_t)(g.acc2_col + grp * g.acc2_stride);
                const uint32_t a16 = ((a2_base + (uint32_t)(grp * g.a2_bytes)) >> 4) + a_lo_c;
                const uint32_t b16 = (w2_base >> 4) + b_lo_c;
                if (elect_one()) {
                    for (int ks = 0; ks < ks2; ++ks) {
                        const uint32_t alo = a16 + (uint32_t)(2 * ks) * (2048u >> 4);
                        const uint32_t blo = b16 + (uint32_t)(2 * ks) * lbo_b16;
                        umma_f16_parts(td, alo, a_hi, blo, b_hi, idesc2, ks > 0 ? 1u : 0u);
                        umma_f16_parts(td, alo + a_part16, a_hi, blo, b_hi, idesc1, 1u);
                    }
                    umma_commit(smem_u32(&ctl->acc2_full[grp]));
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== MMA issuer ===================================================
        // The MMA stream must be straight-line code with distinct descriptor registers: a rolled loop
        // serialises on the uniform-register hazards at ~100-150 cycles per tcgen05.mma, an unrolled one
        // issues every ~47 cycles (tools/mma_rate.cu, profiles/r01_mma_issue_rate.txt).  So taps, k-steps
        // and the three split passes are compile-time unrolled (KSTEPS, SPLIT template parameters) and
        // every descriptor is base + precomputed offset.  The whole warp runs the warp-uniform control
        // flow; one elected lane issues.
        // (Measured in round 2: running the whole role on ONE lane -- `if (lane == 0)` around the item loop -- is 10-30 % SLOWER
        // than warp-uniform control flow with an elected issuing lane: divergent code loses the uniform datapath.)
        {
            const uint32_t idesc = umma_idesc_f16(d.precision == DISCO_PREC_BF16X3, 128, d.block_n);
            const uint32_t idesc2 = umma_idesc_f16(1, 128, 2 * d.block_n);   // STACKED: [W_hi; W_lo]
            const uint32_t idesc4 = (MODE == 4) ? umma_idesc_f16(1, 128, 4 * d.block_n) : 0u;   // two classes' stacked units
            const uint32_t idesc3 = (MODE == 4) ? umma_idesc_f16(1, 128, 3 * d.block_n) : 0u;   // their lo pass: [W_hi; W_lo; W_hi]
            (void)idesc4; (void)idesc3;
            // descriptor halves (see umma_desc_kmajor_noswizzle): lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version
            const uint32_t lbo_b16 = (uint32_t)d.block_n * (STACKED ? 2u : 1u);   // rows per chunk * 16 B, >> 4
            const uint32_t a_hi = ((uint32_t)g.sbo_a >> 4) | (1u << 14);
            const uint32_t b_hi = (128u >> 4) | (1u << 14);
            const uint32_t a_lo_c = ((uint32_t)g.plane >> 4) << 16;
            const uint32_t b_lo_c = lbo_b16 << 16;
            const uint32_t a_part16 = (uint32_t)g.a_part_bytes >> 4;
            const uint32_t b_part16 = STACKED ? (uint32_t)d.block_n : ((uint32_t)g.b_part_bytes >> 4);
            const uint32_t a_base16 = (a_base >> 4) + a_lo_c, b_base16 = (b_base >> 4) + b_lo_c;  // LBO folded in
            const uint32_t a_stage16 = (uint32_t)g.a_stage_bytes >> 4, b_stage16 = (uint32_t)g.b_stage_bytes >> 4;
            const uint32_t bar_a_full = smem_u32(&ctl->a_full[0]), bar_a_empty = smem_u32(&ctl->a_empty[0]);
            const uint32_t bar_b_full = smem_u32(&ctl->b_full[0]), bar_b_empty = smem_u32(&ctl->b_empty[0]);
            uint32_t a_ks[KSTEPS], b_ks[KSTEPS];
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                a_ks[ks] = ((uint32_t)(2 * ks) * (uint32_t)g.plane) >> 4;   // two channel chunks per k-step
                b_ks[ks] = (uint32_t)(2 * ks) * lbo_b16;
            }
            const uint32_t par16 = (uint32_t)g.parplane >> 4;
            // sub-pixel mode: the 2 x 2 low-res taps (of the 3 x 3 window) class (py, px) touches -- rows {py, py+1} x cols {px, px+1}
            const uint32_t sub_mask0 = (MODE == 3) ? ((3u << d.sub_px) | ((3u << d.sub_px) << 3)) << (3 * d.sub_py) : 0u;
            (void)sub_mask0;
            const int mw = warp - kWarpMma;              // issuing warp index; owns items with iacc % nmma == mw
            int iacc = 0;
            int sb_slot = 0;                             // B ring position (streamed weights: single issuer)
            uint32_t sb_phase = 0;
            if (g.stationary && mw < g.nmma) {
                mbar_wait(smem_u32(&ctl->w_full), 0);
                tc_fence_after();
            }
            // ring positions are kept by counters (no integer divisions on this warp's critical path):
            //   buf / acc_par : accumulator buffer iacc % nacc and the parity of iacc / nacc (advance with EVERY item)
            //   turn          : iacc % nmma (which issuer owns the item when the issuers split items)
            //   rs / rs_par   : ring-local A slot of the next stage this issuer consumes and its wrap parity
            int buf = 0, turn = 0, rs = 0;
            uint32_t acc_par = 0, rs_par = 0;
            const int ring0 = (g.nrings == 2) ? mw * g.SAr : 0;   // private ring when issuers split items
            const int sub_lo = g.by_sub ? mw : 0, sub_hi = g.by_sub ? mw + 1 : g.msub;
            for (int item = blockIdx.x; item < g.items; item += gridDim.x, ++iacc, buf = (buf + 1 == g.nacc) ? 0 : buf + 1,
                     acc_par ^= (buf == 0), turn = (turn + 1 == g.nmma) ? 0 : turn + 1) {
                if (mw >= g.nmma || (MODE != 4 && !g.by_sub && turn != mw)) continue;
                if (mw == 0) TRACE(2, iacc, 0);
                uint32_t wblk = 0;   // sub-pixel mode, resident weights: running weight-block index inside the N tile
                (void)wblk;
                mbar_wait(smem_u32(&ctl->acc_empty[buf]), acc_par ^ 1u);
                tc_fence_after();
                if (mw == 0) TRACE(2, iacc, 1);
                const uint32_t td0 = tmem_d + (uint32_t)(buf * ((MODE == 4) ? 4 : g.msub) * g.acc_stride);
                const uint32_t td1 = td0 + (uint32_t)g.acc_stride;
                for (int cb = 0; cb < g.ncb; ++cb) {
                    // wait for the MSUB patches of this channel block
                    const int slot0 = ring0 + rs;
                    const uint32_t sa_phase = rs_par;
                    if (sub_lo == 0) mbar_wait(bar_a_full + 8u * slot0, sa_phase);
                    const uint32_t a0_16 = a_base16 + (uint32_t)slot0 * a_stage16;
                    int slot1 = slot0;
                    uint32_t a1_16 = a0_16;
                    if (g.msub == 2) {
                        slot1 = slot0 + 1;
                        uint32_t phase1 = sa_phase;
                        if (slot1 == ring0 + g.SAr) { slot1 = ring0; phase1 ^= 1u; }
                        if (sub_hi == 2) mbar_wait(bar_a_full + 8u * slot1, phase1);
                        a1_16 = a_base16 + (uint32_t)slot1 * a_stage16;
                    }
                    tc_fence_after();
                    if (cb == 0 && mw == 0) TRACE(2, iacc, 2);
                    const uint32_t first = (cb > 0) ? 1u : 0u;   // accumulate flag of the first MMA of the item
                    if (MODE == 4) {
                        // ---- fused sub-pixel item: four class accumulators (td0 + cls * acc_stride) share this stage.  TWO issuers:
                        //      issuer `mw` owns the classes of output-row parity py = mw.  (tools/mma_rate4.cu: a satisfied mbarrier
                        //      wait costs the issuing thread ~200 cycles and tcgen05.mma issue is serial in the thread, ~45 cycles
                        //      each, so one issuer loses every wait; two issuers hide each other's waits until the shared-memory
                        //      operand port, 44 cycles per MMA of this mix, is the limit.)  Weight slots hold nine blocks of
                        //      64 * block_n bytes and are read by both issuers. ----
                        //      px-merged MMAs: the classes (py, 0) and (py, 1) read the SAME A window for neighbouring taps (class px = 0 with
                        //      tap column t + 1 and class px = 1 with tap column t), and their accumulators are adjacent in TMEM, so
                        //      one N = 4 * block_n instruction with the two classes' weight units stacked -- the units of a group are
                        //      laid out so that the pair is contiguous -- replaces two N = 2 * block_n ones (lo pass: N = 3 * block_n,
                        //      whose middle block adds the exact A_lo * W_lo term to one class).  Operand-port bytes per item -19 %.
                        const bool s1 = cb >= g.ncb0;
                        const uint32_t n2 = 2u * (uint32_t)d.block_n;            // rows of one weight unit [W_hi | W_lo] = its size in 16-byte units
                        const uint32_t accs = (uint32_t)g.acc_stride;            // == n2 (checked on the host): class accumulators are contiguous
                        const uint32_t py = (uint32_t)mw;
                        const uint32_t tdp = td0 + 2u * py * accs;
                        const uint32_t bs = (b_base >> 4) + (uint32_t)sb_slot * b_stage16;
                        if (!s1) {
                            const uint32_t ahi = (160u >> 4) | (1u << 14);
                            const uint32_t part16 = (uint32_t)g.a_part0 >> 4;
                            const uint32_t sa0 = (a_base >> 4) + (uint32_t)slot0 * a_stage16 + (((uint32_t)g.plane0 >> 4) << 16) + py * 10u;
#pragma unroll
                            for (int ty = 0; ty < 2; ++ty) {      // slot = low-res tap row ty: [py][chunk][(px0,tx0) (px0,tx1) (px1,tx0) (px1,tx1)] (+ pad)
                                mbar_wait(bar_b_full + 8u * sb_slot, sb_phase);
                                tc_fence_after();
                                const uint32_t bg = (b_base >> 4) + (uint32_t)sb_slot * b_stage16 + py * 8u * n2 + ((4u * n2) << 16);
                                if (elect_one()) {
                                    if (!(g.dbg & 8)) {
                                        const uint32_t a16 = sa0 + (uint32_t)(ty * 10);
                                        const uint32_t f = (ty == 0) ? first : 1u;
                                        umma_f16_parts(tdp, a16, ahi, bg, b_hi, idesc2, f);                         // column 0: class px = 0
                                        umma_f16_parts(tdp, a16 + part16, ahi, bg, b_hi, idesc, 1u);
                                        umma_f16_parts(tdp + accs, a16 + 2u, ahi, bg + 3u * n2, b_hi, idesc2, f);   // column 2: class px = 1
                                        umma_f16_parts(tdp + accs, a16 + 2u + part16, ahi, bg + 3u * n2, b_hi, idesc, 1u);
                                        umma_f16_parts(tdp, a16 + 1u, ahi, bg + n2, b_hi, idesc4, 1u);              // column 1: both classes
                                        umma_f16_parts(tdp, a16 + 1u + part16, ahi, bg + n2, b_hi, idesc3, 1u);
                                    }
                                    umma_commit(bar_b_empty + 8u * sb_slot);
                                }
                                __syncwarp();
                                if (++sb_slot == g.SB) { sb_slot = 0; sb_phase ^= 1u; }
                            }
                        } else {
                            mbar_wait(bar_b_full + 8u * sb_slot, sb_phase);      // slot = [kh][chunk][kw = 2, 1, 0] of this channel block
                            tc_fence_after();
                            const uint32_t a0p = a0_16 + py * 9u;
                            if (elect_one()) {
                                if (!(g.dbg & 8))
#pragma unroll
                                for (int kh = 0; kh < 3; ++kh) {
                                    const uint32_t bg = bs + (uint32_t)kh * 6u * n2 + ((3u * n2) << 16);
                                    const uint32_t ar = a0p + (uint32_t)(kh * 9);
                                    // window column c = px + kw: plane c & 1, column index c >> 1
                                    umma_f16_parts(tdp, ar, a_hi, bg + 2u * n2, b_hi, idesc2, 1u);                              // c = 0: px 0, kw 0
                                    umma_f16_parts(tdp, ar + a_part16, a_hi, bg + 2u * n2, b_hi, idesc, 1u);
                                    umma_f16_parts(tdp + accs, ar + par16 + 1u, a_hi, bg, b_hi, idesc2, 1u);                    // c = 3: px 1, kw 2
                                    umma_f16_parts(tdp + accs, ar + par16 + 1u + a_part16, a_hi, bg, b_hi, idesc, 1u);
                                    umma_f16_parts(tdp, ar + par16, a_hi, bg + n2, b_hi, idesc4, 1u);                           // c = 1: [W(kh,1); W(kh,0)]
                                    umma_f16_parts(tdp, ar + par16 + a_part16, a_hi, bg + n2, b_hi, idesc3, 1u);
                                    umma_f16_parts(tdp, ar + 1u, a_hi, bg, b_hi, idesc4, 1u);                                   // c = 2: [W(kh,2); W(kh,1)]
                                    umma_f16_parts(tdp, ar + 1u + a_part16, a_hi, bg, b_hi, idesc3, 1u);
                                }
                                umma_commit(bar_b_empty + 8u * sb_slot);
                            }
                            __syncwarp();
                            if (++sb_slot == g.SB) { sb_slot = 0; sb_phase ^= 1u; }
                        }
                    } else if (MODE == 3) {
                        // ---- sub-pixel class: 4 pre-summed taps on a source-0 stage (18 x 10 low-res patch), all 9 taps on a
                        //      source-1 stage (stride-2 parity planes); weight blocks are consumed in packed order ----
                        const bool s1 = cb >= g.ncb0;
                        const uint32_t ahi = s1 ? a_hi : ((160u >> 4) | (1u << 14));
                        const uint32_t lbo16 = s1 ? a_lo_c : (((uint32_t)g.plane0 >> 4) << 16);
                        const uint32_t part16 = s1 ? a_part16 : ((uint32_t)g.a_part0 >> 4);
                        const uint32_t sa0 = (a_base >> 4) + (uint32_t)slot0 * a_stage16 + lbo16;
                        const uint32_t sa1 = (a_base >> 4) + (uint32_t)slot1 * a_stage16 + lbo16;
                        const uint32_t mask = s1 ? 0x1FFu : sub_mask0;
                        uint32_t acc = first;
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap) {
                            if (!((mask >> tap) & 1u)) continue;
                            uint32_t b16;
                            if (g.stationary) {
                                b16 = b_base16 + wblk * b_stage16;
                            } else {
                                mbar_wait(bar_b_full + 8u * sb_slot, sb_phase);
                                tc_fence_after();
                                b16 = b_base16 + (uint32_t)sb_slot * b_stage16;
                            }
                            const int kh = tap / 3, kw = tap - kh * 3;
                            const uint32_t toff = s1 ? ((uint32_t)(kw & 1) * par16 + (uint32_t)(kh * 9 + (kw >> 1))) : (uint32_t)(kh * 10 + kw);
                            if (elect_one()) {
#pragma unroll
                                for (int sub = 0; sub < 2; ++sub) {
                                    if (sub >= sub_lo && sub < sub_hi) {
                                        const uint32_t td = sub ? td1 : td0;
                                        const uint32_t a16 = (sub ? sa1 : sa0) + toff;
                                        if (STACKED) {
                                            umma_f16_parts(td, a16, ahi, b16, b_hi, idesc2, acc);
                                            if (!skip_lo) umma_f16_parts(td, a16 + part16, ahi, b16, b_hi, idesc, 1u);
                                        } else {
                                            umma_f16_parts(td, a16, ahi, b16, b_hi, idesc, acc);
                                            if (SPLIT) {
                                                if (!skip_lo) umma_f16_parts(td, a16 + part16, ahi, b16, b_hi, idesc, 1u);
                                                umma_f16_parts(td, a16, ahi, b16 + b_part16, b_hi, idesc, 1u);
                                            }
                                        }
                                    }
                                }
                                if (!g.stationary) umma_commit(bar_b_empty + 8u * sb_slot);
                            }
                            __syncwarp();
                            acc = 1u;
                            ++wblk;
                            if (!g.stationary) {
                                if (++sb_slot == g.SB) { sb_slot = 0; sb_phase ^= 1u; }
                            }
                        }
                    } else if (g.stationary) {
                        // Resident weights: nothing to wait for between taps, so ONE elected region issues the whole channel
                        // block as a branch-free run of tcgen05.mma with descriptor = base + precomputed offset (round-2 ncu
                        // source view: the per-tap ring / elect / constant-bank code of the generic loop below cost ~45 scalar
                        // instructions per 4 MMAs, i.e. ~124 cycles per MMA and issuer against a hardware floor of ~50).
                        const uint32_t b_cb16 = b_base16 + (uint32_t)(cb * TAPS) * b_stage16;
                        if (elect_one()) {
#pragma unroll
                            for (int sub = 0; sub < 2; ++sub) {
                                if (sub >= sub_lo && sub < sub_hi) {
                                    const uint32_t td = sub ? td1 : td0;
                                    const uint32_t a16s = sub ? a1_16 : a0_16;
#pragma unroll
                                    for (int tap = 0; tap < TAPS; ++tap) {
                                        const int kh = tap / 3, kw = tap - kh * 3;
                                        const uint32_t toff = (MODE == 0) ? (uint32_t)(kh * 10 + kw)
                                                            : (MODE == 1) ? (uint32_t)(kw & 1) * par16 + (uint32_t)(kh * 9 + (kw >> 1))
                                                                          : 0u;
                                        const uint32_t a16 = a16s + toff, b16 = b_cb16 + (uint32_t)tap * b_stage16;
#pragma unroll
                                        for (int ks = 0; ks < KSTEPS; ++ks) {
                                            const uint32_t alo = a16 + a_ks[ks], blo = b16 + b_ks[ks];
                                            const uint32_t accf = (tap == 0 && ks == 0) ? first : 1u;
                                            if (STACKED) {
                                                umma_f16_parts(td, alo, a_hi, blo, b_hi, idesc2, accf);            // hi*[hi|lo]
                                                if (!skip_lo) umma_f16_parts(td, alo + a_part16, a_hi, blo, b_hi, idesc, 1u);   // lo*hi
                                            } else {
                                                umma_f16_parts(td, alo, a_hi, blo, b_hi, idesc, accf);
                                                if (SPLIT) {
                                                    if (!skip_lo) umma_f16_parts(td, alo + a_part16, a_hi, blo, b_hi, idesc, 1u);
                                                    umma_f16_parts(td, alo, a_hi, blo + b_part16, b_hi, idesc, 1u);
                                                }
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        __syncwarp();
                    } else {
                    // (Measured in round 2: moving the per-tap weight-stage waits into ONE elected region per channel block, like the
                    // resident-weight path above, is ~8 % SLOWER on the streamed tensor-bound layers -- conv5_1 0.61 vs 0.56 ms --: the
                    // elected lane's spin loops run on the divergent path.)
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
                        uint32_t b16;
                        {
                            mbar_wait(bar_b_full + 8u * sb_slot, sb_phase);
                            tc_fence_after();
                            b16 = b_base16 + (uint32_t)sb_slot * b_stage16;
                        }
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const uint32_t toff = (MODE == 0) ? (uint32_t)(kh * 10 + kw)
                                            : (MODE == 1) ? (uint32_t)(kw & 1) * par16 + (uint32_t)(kh * 9 + (kw >> 1))
                                                          : 0u;
                        if (elect_one()) {
#pragma unroll
                            for (int sub = 0; sub < 2; ++sub) {
                                if (sub >= sub_lo && sub < sub_hi) {
                                    const uint32_t td = sub ? td1 : td0;
                                    const uint32_t a16 = (sub ? a1_16 : a0_16) + toff;
#pragma unroll
                                    for (int ks = 0; ks < KSTEPS; ++ks) {
                                        const uint32_t alo = a16 + a_ks[ks], blo = b16 + b_ks[ks];
                                        const uint32_t accf = (tap == 0 && ks == 0) ? first : 1u;
                                        if (STACKED) {
                                            umma_f16_parts(td, alo, a_hi, blo, b_hi, idesc2, accf);            // hi*[hi|lo]
                                            if (!skip_lo) umma_f16_parts(td, alo + a_part16, a_hi, blo, b_hi, idesc, 1u);   // lo*hi
                                        } else {
                                            umma_f16_parts(td, alo, a_hi, blo, b_hi, idesc, accf);
                                            if (SPLIT) {
                                                if (!skip_lo) umma_f16_parts(td, alo + a_part16, a_hi, blo, b_hi, idesc, 1u);
                                                umma_f16_parts(td, alo, a_hi, blo + b_part16, b_hi, idesc, 1u);
                                            }
                                        }
                                    }
                                }
                            }
                            umma_commit(bar_b_empty + 8u * sb_slot);
                        }
                        __syncwarp();
                        if (++sb_slot == g.SB) { sb_slot = 0; sb_phase ^= 1u; }
                    }
                    }
                    if (elect_one()) {
                        if (sub_lo == 0) umma_commit(bar_a_empty + 8u * slot0);
                        if (sub_hi == 2) umma_commit(bar_a_empty + 8u * slot1);
                    }
                    __syncwarp();
                    rs += g.msub;
                    if (rs >= g.SAr) { rs -= g.SAr; rs_par ^= 1u; }
                }
                if (elect_one()) umma_commit(smem_u32(&ctl->acc_full[buf]));
                __syncwarp();
                if (mw == 0) TRACE(2, iacc, 3);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) tmem_dealloc(tmem_d, (uint32_t)g.tmem_cols);
}

int g_num_sms = 0;

int build_geom(const disco_conv_desc* d, ConvGeom* g) {
    DISCO_REQUIRE(d->taps == 9 || d->taps == 1, "conv: taps must be 9 or 1 (got %d)", d->taps);
    DISCO_REQUIRE(d->stride == 1 || (d->stride == 2 && d->taps == 9), "conv: stride %d unsupported", d->stride);
    DISCO_REQUIRE(d->c_blk == 16 || d->c_blk == 32 || d->c_blk == 64, "conv: c_blk must be 16/32/64");
    DISCO_REQUIRE(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "conv: bad block_n %d", d->block_n);
    DISCO_REQUIRE(d->src_c[0] > 0 && d->src_c[0] % d->c_blk == 0 && d->src_c[1] % d->c_blk == 0,
                  "conv: source channels (%d,%d) must be multiples of c_blk %d", d->src_c[0], d->src_c[1], d->c_blk);
    DISCO_REQUIRE(d->precision == DISCO_PREC_FP16 || d->precision == DISCO_PREC_BF16X3, "conv: bad precision");
    DISCO_REQUIRE(!d->wpack_stacked || (d->precision == DISCO_PREC_BF16X3 && d->block_n <= 128),
                  "conv: stacked weight images need bf16x3 and block_n <= 128");
    DISCO_REQUIRE(d->taps == 9 || (d->src_up[0] == 0 && d->src_up[1] == 0), "conv: 1x1 cannot upsample");
    DISCO_REQUIRE(d->n > 0 && d->h_in > 0 && d->w_in > 0, "conv: empty input");
    if (d->subpix)
        DISCO_REQUIRE(d->taps == 9 && d->stride == 1 && d->src_up[0] == 1 && d->src_up[1] == 0 && d->src_c[1] > 0 && d->c_blk == 16 &&
                          d->precision == DISCO_PREC_BF16X3 && d->chain_c_out == 0 && (d->sub_py | d->sub_px) >= 0 && d->sub_py <= 1 &&
                          d->sub_px <= 1 && d->h_in % 2 == 0 && d->w_in % 2 == 0 && get_encode_tiled() != nullptr,
                      "conv: sub-pixel class needs a 3x3 stride-1 conv over (upsampled, plain) sources, c_blk 16, bf16x3 and TMA");
    const bool fused = d->subpix == 2;   // all four classes per item (MODE 4)
    if (fused)
        DISCO_REQUIRE(d->wpack_stacked && d->block_n <= 64 && d->c_out <= d->block_n && d->out_mode == DISCO_OUT_ACT,
                      "conv: the fused sub-pixel form needs one stacked N tile (c_out <= 64) and an activation output");
    DISCO_REQUIRE(((d->c_out + d->block_n - 1) / d->block_n) * d->block_n * 4 <= kBiasBytes, "conv: c_out %d too large", d->c_out);
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
    else {
        const int oc = d->chain_c_out > 0 ? d->chain_c_out : d->c_out;
        DISCO_REQUIRE(d->out_split % 4 == 0 && (oc - d->out_split) % 4 == 0 && d->out_split <= oc,
                      "conv: fp32 output split must be a multiple of 4");
    }
    if (g_num_sms == 0) {
        int dev = 0;
        DISCO_CHECK_CUDA(cudaGetDevice(&dev));
        DISCO_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }

    g->d = *d;
    {
        const char* tr = getenv("DISCO_CONV_TRACE");
        g->trace = tr ? (long long*)strtoull(tr, nullptr, 0) : nullptr;
        const char* dbg = getenv("DISCO_CONV_DBG");
        g->dbg = dbg ? atoi(dbg) : 0;
        const char* ds = getenv("DISCO_CONV_DIRECT_STORE");
        g->direct = (ds && ds[0] == '1') ? 1 : 0;
    }
    g->ncb0 = d->src_c[0] / d->c_blk;
    g->ncb = g->ncb0 + d->src_c[1] / d->c_blk;
    g->chunks = d->c_blk / 8;
    g->chunk_shift = (g->chunks == 2) ? 1 : (g->chunks == 4) ? 2 : 3;
    g->nparts = (d->precision == DISCO_PREC_BF16X3) ? 2 : 1;
    g->total_pix = (long long)d->n * d->h_out * d->w_out;
    g->plane0 = 0; g->a_part0 = 0;
    if (d->taps == 1) {
        g->PIX = 128; g->parplane = 0; g->sbo_a = 128;
    } else if (fused) {
        g->PIX = 34 * 18; g->parplane = 34 * 9 * 16; g->sbo_a = 2 * 9 * 16;   // 34 x 18 window as two column-parity planes
    } else if (d->subpix) {
        g->PIX = 33 * 17; g->parplane = 33 * 9 * 16; g->sbo_a = 2 * 9 * 16;   // source-1 stages: stride-2 geometry
    } else if (d->stride == 1) {
        g->PIX = 180; g->parplane = 0; g->sbo_a = 160;
    } else {
        g->PIX = 33 * 17; g->parplane = 33 * 9 * 16; g->sbo_a = 2 * 9 * 16;
    }
    // TMA for the A operand: every source that is not zero-stuffed (src_up == 2 is the transposed-conv trick of the
    // training data gradient and keeps the cp.async gather).  DISCO_CONV_NO_TMA=1 forces the cp.async gather (A/B tests).
    g->use_tma = 0; g->scratch_warp = 0; g->scratch_total = 0; g->scr_bufs = 1;
    {
        static int no_tma = -1;
        if (no_tma < 0) { const char* e = getenv("DISCO_CONV_NO_TMA"); no_tma = (e && e[0] == '1') ? 1 : 0; }
        const bool any_plain = (d->src_up[0] != 2) || (d->src_c[1] && d->src_up[1] != 2);
        const bool up_in_s2 = d->stride == 2 && (d->src_up[0] == 1 || (d->src_c[1] && d->src_up[1] == 1));
        if (!no_tma && any_plain && !up_in_s2 && g->total_pix < (1ll << 31) && get_encode_tiled()) g->use_tma = 1;
    }
    int plane = (d->taps == 9 && (d->stride == 2 || d->subpix)) ? 2 * g->parplane : g->PIX * 16;
    if (d->subpix) {
        g->use_tma = 1;
        g->parplane = (g->parplane + 127) / 128 * 128;
        plane = 2 * g->parplane;
        g->plane0 = (18 * 10 * 16 + 127) / 128 * 128;
        g->a_part0 = g->chunks * g->plane0;
    } else if (g->use_tma) {
        // TMA destinations are 128-byte aligned: round the per-chunk planes up (the async proxy has no bank conflicts to dodge)
        if (d->taps == 9 && d->stride == 2) { g->parplane = (g->parplane + 127) / 128 * 128; plane = 2 * g->parplane; }
        else plane = (plane + 127) / 128 * 128;
        if (d->taps == 9 && (d->src_up[0] == 1 || (d->src_c[1] && d->src_up[1] == 1))) {
            // Prefetching the NEXT upsampled stage's box into a second scratch half is implemented (parity-green) but OFF: measured
            // on conv8_1 it is slower (1.37 vs 1.06 ms) because the extra 16 KB of scratch shrink the A ring from 6 to 4 stages,
            // which costs more than the hidden L2 latency buys.  DISCO_CONV_SCR_PREFETCH=1 enables it for 16-channel stages.
            static int pf = -1;
            if (pf < 0) { const char* e = getenv("DISCO_CONV_SCR_PREFETCH"); pf = (e && e[0] == '1') ? 1 : 0; }
            g->scr_bufs = (pf && g->chunks <= 2) ? 2 : 1;
            g->scratch_warp = g->scr_bufs * g->nparts * g->chunks * kScrBox;
            g->scratch_total = kProdWarps * g->scratch_warp;
        }
    } else {
        // pad so that the `chunks` 16-byte writes of one pixel land in distinct bank groups
        const int want = (128 / (g->chunks > 8 ? 8 : g->chunks)) % 128;
        while ((plane % 128) != want) plane += 16;
    }
    g->plane = plane;
    g->a_part_bytes = g->chunks * plane;
    g->a_stage_bytes = ((g->nparts * g->a_part_bytes + 127) / 128) * 128;
    g->b_part_bytes = d->c_blk * d->block_n * 2;
    g->b_stage_bytes = fused ? 9 * 64 * d->block_n : g->nparts * g->b_part_bytes;   // fused: slots of nine weight blocks
    g->n_tiles = (d->c_out + d->block_n - 1) / d->block_n;

    const int stage_region = d->chain_c_out > 0 ? 0 : kStageBytes;
    const int budget = 224 * 1024 - kCtlBytes - kBiasBytes - stage_region - g->scratch_total;
    g->w_iters = fused ? g->ncb0 * 2 + (g->ncb - g->ncb0)
                       : d->subpix ? g->ncb0 * 4 + (g->ncb - g->ncb0) * 9 : g->ncb * d->taps;
    g->w_bytes = g->w_iters * g->b_stage_bytes;
    // MSUB = 2 (256-pixel items) when the N tile leaves room for double-buffered accumulators and the
    // image is wide enough; it halves the weight stream per MAC.
    const int h_grid = d->subpix ? d->h_out / 2 : d->h_out, w_grid = d->subpix ? d->w_out / 2 : d->w_out;   // tiled pixel grid
    const bool wide = (d->taps == 1) ? (g->total_pix >= 256) : (w_grid >= 16);
    const int acc_cols = (d->wpack_stacked ? 2 : 1) * d->block_n;   // TMEM columns one accumulator writes
    g->msub = (wide && 4 * acc_cols <= 512 && !fused) ? 2 : 1;
    {
        // Wave quantisation: with few, large items (N >= 128 tiles at 32^2 / 16^2 resolution) the last round of a launch leaves
        // most SMs idle (e.g. 320 items of two sub-tiles on 148 SMs = 2.16 rounds -> 3).  One sub-tile per item lowers the number of
        // sub-tile rounds but doubles the weight stream per MAC (42 instead of 21 B/clk per SM for N >= 128: over the L2 cap when
        // all SMs stream).  Tuning knob DISCO_CONV_MSUB = 1 (force) | 3 (when it saves >= 10 % of the rounds); default: off.
        static int force = -1;
        if (force < 0) { const char* e = getenv("DISCO_CONV_MSUB"); force = e ? atoi(e) : 0; }
        if (g->msub == 2 && d->block_n >= 128) {
            const long long th = (d->taps == 1) ? 1 : (h_grid + 15) / 16;
            const long long sub_tiles = (d->taps == 1) ? (g->total_pix + 127) / 128 : (long long)d->n * th * ((w_grid + 7) / 8);
            const long long items2 = ((d->taps == 1) ? (g->total_pix + 255) / 256 : (long long)d->n * th * ((w_grid + 15) / 16)) * g->n_tiles;
            const long long items1 = sub_tiles * g->n_tiles;
            const int sms = g_num_sms > 0 ? g_num_sms : 148;
            const long long rounds2 = 2 * ((items2 + sms - 1) / sms), rounds1 = (items1 + sms - 1) / sms;
            if (force == 1 || (force == 3 && rounds1 * 10 <= rounds2 * 9)) g->msub = 1;   // 3: only when >= 10 % fewer sub-tile rounds
            // default: split only launches that cannot even fill the SMs once (small batches, e.g. one agent per GPU): the
            // doubled weight stream is harmless when most of the L2 bandwidth is idle, and twice as many SMs get work
            if (force == 0 && items2 < sms) g->msub = 1;
        }
    }
    g->stationary = (g->n_tiles == 1 && g->w_bytes + 3 * g->a_stage_bytes <= budget && !fused) ? 1 : 0;
    {
        // tuning knob: weight sets larger than DISCO_CONV_NOSTAT_MAXW bytes are streamed even if they would fit (a big resident
        // set leaves only a shallow A-stage ring: conv8_1 = 110 KB)
        static long maxw = -1;
        if (maxw < 0) { const char* e = getenv("DISCO_CONV_NOSTAT_MAXW"); maxw = e ? atol(e) : 0; }
        if (maxw > 0 && g->w_bytes > maxw && d->chain_c_out == 0) g->stationary = 0;
    }
    // Stationary layers: items of one sub-tile with issuers splitting ITEMS (private A rings) -- unless one item
    // needs more channel blocks than a private ring can hold (conv8_1: 3 blocks, 2 slots per ring): then keep
    // MSUB = 2 and let the issuers split the SUB-TILES of every item, which pipelines at channel-block granularity.
    bool stationary_by_sub = false;
    if (g->stationary && d->stride == 2 && (budget - g->w_bytes) / g->a_stage_bytes < 4 && wide && 4 * acc_cols <= 512)
        g->stationary = 0;   // conv1_1: the big stride-2 patches need the room more than the weights do (measured)
    if (g->stationary) {
        const int sa_fit = (budget - g->w_bytes) / g->a_stage_bytes;
        stationary_by_sub = (d->chain_c_out == 0 && g->msub == 2 && acc_cols <= 128 && sa_fit >= 4 && sa_fit / 2 < g->ncb);
        if (!stationary_by_sub) g->msub = 1;
    }
    const int nsub = fused ? 4 : g->msub;   // accumulators per TMEM buffer
    g->nacc = (2 * nsub * acc_cols <= 512) ? 2 : 1;
    g->nmma = 1;
    if (g->stationary && g->nacc == 2) {
        g->nmma = kMmaWarps;
        if (kMaxAcc * g->msub * ((acc_cols + 31) / 32 * 32) <= 512) g->nacc = kMaxAcc;   // two buffers per issuer
    }
    g->acc_stride = (acc_cols + 31) / 32 * 32;
    if (g->nacc * nsub * g->acc_stride > 512) g->acc_stride = acc_cols;
    int cols = 32;
    while (cols < g->nacc * nsub * g->acc_stride) cols *= 2;
    g->tmem_cols = cols;
    int ctas_per_sm = 1;
    g->chain = d->chain_c_out > 0 ? 1 : 0;
    g->chain_bn = 0; g->a2_bytes = 0; g->w2_bytes = 0; g->acc2_col = 0; g->acc2_stride = 0;
    int chain_smem = 0;
    if (g->chain) {
        DISCO_REQUIRE(g->n_tiles == 1 && d->wpack_stacked && d->precision == DISCO_PREC_BF16X3 && d->block_n <= 64 &&
                          d->out_mode == DISCO_OUT_F32 && d->chain_wpack && d->chain_bias && d->relu,
                      "conv: chained 1x1 needs one stacked bf16x3 N tile (<= 64 channels), ReLU and fp32 output");
        g->chain_bn = (d->chain_c_out + 15) / 16 * 16;
        DISCO_REQUIRE(g->chain_bn <= 128 && g->chain_bn <= 4 * (kBiasBytes / 4 - kBias2Off) , "conv: chain too wide");
        g->a2_bytes = 2 * (d->block_n / 8) * 2048;
        g->w2_bytes = (d->block_n / 8) * 2 * g->chain_bn * 16;
        chain_smem = 2 * g->a2_bytes + g->w2_bytes;
        g->stationary = (g->w_bytes + 3 * g->a_stage_bytes + chain_smem <= budget) ? 1 : 0;
        DISCO_REQUIRE(g->stationary, "conv: chained layer does not fit shared memory");
        g->msub = 1;
        g->nacc = 2;
        g->nmma = kMmaWarps;
        g->acc2_stride = (2 * g->chain_bn + 31) / 32 * 32;
        g->acc2_col = g->nacc * g->acc_stride;
        DISCO_REQUIRE(g->acc2_col + 2 * g->acc2_stride <= 512, "conv: chain accumulators exceed TMEM");
        int cols2 = 32;
        while (cols2 < g->acc2_col + 2 * g->acc2_stride) cols2 *= 2;
        g->tmem_cols = cols2;
    }
    if (g->stationary) {
        // small resident weight sets: aim for two CTAs per SM (two MMA issuers, 2x gather streams)
        int sa;
        sa = (budget - g->w_bytes - chain_smem) / g->a_stage_bytes;   // one CTA per SM: the MMA issue port is per SM anyway
        if (sa > kMaxStages) sa = kMaxStages;
        g->SA = sa;
        g->SB = 1;
        g->smem_bytes = kCtlBytes + kBiasBytes + stage_region + g->scratch_total + g->SA * g->a_stage_bytes + g->w_bytes + chain_smem;
    } else {
        int sa = 2 * g->msub;  // current + next channel block
        if (sa < 4 && 4 * g->a_stage_bytes <= budget / 3) sa = 4;
        if (fused) {
            // 39 KB stages carrying 16 (source 0) / 36 (source 1) MMA pairs each: three slots = two stages of lookahead; the rest
            // of shared memory is the weight ring (8 / 16 KB slots).  Tuning knob: DISCO_CONV_FUSED_SA.
            static int fsa = -1;
            if (fsa < 0) { const char* e = getenv("DISCO_CONV_FUSED_SA"); fsa = e ? atoi(e) : 0; }
            sa = (fsa >= 2 && fsa <= kMaxStages) ? fsa : 3;
        }
        int sb = (budget - sa * g->a_stage_bytes) / g->b_stage_bytes;
        while (sb < 2 && sa > g->msub) {
            --sa;
            sb = (budget - sa * g->a_stage_bytes) / g->b_stage_bytes;
        }
        if (sb > kMaxStages) sb = kMaxStages;
        DISCO_REQUIRE(sb >= 1 && sa >= g->msub, "conv: tile does not fit shared memory (a_stage %d, b_stage %d)",
                      g->a_stage_bytes, g->b_stage_bytes);
        g->SA = sa;
        g->SB = sb;
        g->smem_bytes = kCtlBytes + kBiasBytes + stage_region + g->scratch_total + g->SA * g->a_stage_bytes + g->SB * g->b_stage_bytes;
    }
    DISCO_REQUIRE(g->SA >= g->msub && g->SA >= 1, "conv: not enough A stages");
    g->by_sub = 0;
    g->nrings = 1;
    if (g->nmma == 2 && g->SA / 2 >= g->msub && !stationary_by_sub) g->nrings = 2;   // issuers split items, private rings
    else g->nmma = 1;
    if ((!g->stationary || stationary_by_sub) && g->msub == 2 && acc_cols <= 128 && g->SA >= 4) {
        // streamed N <= 64 layers are issue-bound too: two issuers, one per sub-tile, sharing every B stage
        g->by_sub = 1;
        g->nmma = 2;
        g->SA &= ~1;   // even ring: slot parity == sub-tile, i.e. one consumer per slot
    }
    if (fused) {
        g->nmma = 2; g->nrings = 1; g->by_sub = 0;
        DISCO_REQUIRE(g->acc_stride == 2 * d->block_n, "conv: fused sub-pixel accumulators must be contiguous (stride %d, block_n %d)", g->acc_stride, d->block_n);
    }   // two issuers, one per output-row parity, sharing every stage
    g->SAr = g->SA / g->nrings;
    g->nprod = kProdWarps / g->nrings;
    if (g->nprod > g->SAr) g->nprod = g->SAr;
    DISCO_REQUIRE(g->SAr >= g->msub && g->nprod >= 1, "conv: A ring too small");
    g->tiles_h = (h_grid + 15) / 16;
    g->tiles_w = (w_grid + 8 * g->msub - 1) / (8 * g->msub);
    const long long m_tiles = (d->taps == 9) ? (long long)d->n * g->tiles_h * g->tiles_w
                                             : (g->total_pix + 128 * g->msub - 1) / (128 * g->msub);
    DISCO_REQUIRE(m_tiles > 0 && m_tiles * g->n_tiles < (1ll << 30), "conv: bad tile count");
    g->m_tiles = (int)m_tiles;
    g->items = g->m_tiles * g->n_tiles;
    g->fd_m_tiles = make_fastdiv(g->m_tiles);
    g->fd_per_img = make_fastdiv(g->tiles_h * g->tiles_w);
    g->fd_tiles_w = make_fastdiv(g->tiles_w);
    g->grid = g->items < ctas_per_sm * g_num_sms ? g->items : ctas_per_sm * g_num_sms;
    DISCO_REQUIRE((g->plane >> 4) < 16384 && g->smem_bytes <= 227 * 1024, "conv: descriptor / smem range");
    if (d->subpix) {
        for (int part = 0; part < g->nparts; ++part) {
            const uint16_t* b0 = reinterpret_cast<const uint16_t*>(d->src[0]) + (part ? d->src_lo_off[0] : 0);
            const uint16_t* b1 = reinterpret_cast<const uint16_t*>(d->src[1]) + (part ? d->src_lo_off[1] : 0);
            DISCO_REQUIRE(((reinterpret_cast<uintptr_t>(b0) | reinterpret_cast<uintptr_t>(b1)) & 15) == 0, "conv: sources not 16-byte aligned");
            const bool ok = make_tmap4(&g->tmap[part], b0, d->src_c[0], d->w_in / 2, d->h_in / 2, d->n, 10, 18, 1) &&
                            make_tmap4(&g->tmap[2 + part], b1, d->src_c[1], d->w_in, d->h_in, d->n, 18, fused ? 34 : 33, 2);
            DISCO_REQUIRE(ok, "conv: cuTensorMapEncodeTiled failed (sub-pixel class)");
        }
    } else if (g->use_tma) {
        for (int s = 0; s < 2; ++s) {
            if (!d->src_c[s] || d->src_up[s] == 2) continue;
            const int up = d->src_up[s] ? 1 : 0;
            const int Hs = d->h_in >> up, Ws = d->w_in >> up;
            for (int part = 0; part < g->nparts; ++part) {
                const uint16_t* base = reinterpret_cast<const uint16_t*>(d->src[s]) + (part ? d->src_lo_off[s] : 0);
                DISCO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "conv: source %d is not 16-byte aligned", s);
                bool ok;
                if (d->taps == 1) ok = make_tmap2(&g->tmap[s * 2 + part], base, d->src_c[s], g->total_pix);
                else if (up) ok = make_tmap4(&g->tmap[s * 2 + part], base, d->src_c[s], Ws, Hs, d->n, 6, 10, 1);
                else if (d->stride == 1) ok = make_tmap4(&g->tmap[s * 2 + part], base, d->src_c[s], Ws, Hs, d->n, 10, 18, 1);
                else ok = make_tmap4(&g->tmap[s * 2 + part], base, d->src_c[s], Ws, Hs, d->n, 18, 33, 2);
                DISCO_REQUIRE(ok, "conv: cuTensorMapEncodeTiled failed (source %d: C %d, %dx%d, n %d)", s, d->src_c[s], Hs, Ws, d->n);
            }
        }
    }
    return DISCO_OK;
}

template <int MODE, int KSTEPS, int PASSES>
int launch_inst(const ConvGeom& g, cudaStream_t stream) {
    static bool attr_set[64] = {false};
    int dev = 0;
    DISCO_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        DISCO_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<MODE, KSTEPS, PASSES>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    conv_tc_kernel<MODE, KSTEPS, PASSES><<<g.grid, kThreads, g.smem_bytes, stream>>>(g);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}

template <int MODE>
int launch_mode(const ConvGeom& g, cudaStream_t stream) {
    const int passes = (g.nparts == 2) ? (g.d.wpack_stacked ? 2 : 3) : 1;
#define DISCO_LAUNCH_K(KS)                                                      \
    (passes == 1 ? launch_inst<MODE, KS, 1>(g, stream)                          \
                 : passes == 2 ? launch_inst<MODE, KS, 2>(g, stream) : launch_inst<MODE, KS, 3>(g, stream))
    switch (g.d.c_blk) {
        case 16: return DISCO_LAUNCH_K(1);
        case 32: return DISCO_LAUNCH_K(2);
        default: return DISCO_LAUNCH_K(4);
    }
#undef DISCO_LAUNCH_K
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
    cudaStream_t s = (cudaStream_t)stream;
    if (d->taps == 1) return launch_mode<2>(g, s);
    if (d->subpix == 2) return launch_inst<4, 1, 2>(g, s);
    if (d->subpix) return g.d.wpack_stacked ? launch_inst<3, 1, 2>(g, s) : launch_inst<3, 1, 3>(g, s);
    return d->stride == 1 ? launch_mode<0>(g, s) : launch_mode<1>(g, s);
}
