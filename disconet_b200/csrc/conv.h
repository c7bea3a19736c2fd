// Convolution descriptor shared by the tcgen05 kernel (conv_tc.cu), the CUDA-core validator
// (conv_ref.cu) and the C-ABI (capi.cu).  Mirrors `disco_conv_desc` in include/disco_b200.h.
#pragma once
#include <stdint.h>

// Activations are NHWC 16-bit tensors.  precision:
//   DISCO_PREC_FP16   : one fp16 tensor, one tensor-core pass
//   DISCO_PREC_BF16X3 : value = hi + lo (two bf16 tensors, `lo` at element offset lo_off from `hi`),
//                       three tensor-core passes (hi*Whi + lo*Whi + hi*Wlo), fp32 accumulate
#define DISCO_PREC_FP16 0
#define DISCO_PREC_BF16X3 1

#define DISCO_OUT_ACT 0   // 16-bit activation tensor (hi [+ lo]) NHWC, channel stride out_cstride[0]
#define DISCO_OUT_F32 1   // fp32 NHWC; optionally split in two tensors at channel `out_split`

struct disco_conv_desc {
    // ---- input: channel-concatenation of up to two NHWC sources --------------------------------
    const void* src[2];      // hi tensors (16-bit elements)
    long long src_lo_off[2]; // elements from hi to lo tensor (BF16X3 only)
    int src_c[2];            // channels per source (multiple of 16; second may be 0)
    int src_up[2];           // 1: source is (H_in/2 x W_in/2), nearest-upsampled x2 on the fly; 2: same size, ZERO-STUFFED
                             // (value at even (h,w) only) -- the data gradient of a stride-2 conv as a stride-1 conv
    int n, h_in, w_in;       // logical (post-upsample) conv input size
    int h_out, w_out;
    int stride;              // 1 | 2
    int taps;                // 9 (3x3, pad 1) | 1 (1x1, pad 0, stride 1)
    int c_blk;               // channels per K stage: 16 | 32 | 64; divides src_c[0] and src_c[1]
    // ---- weights ---------------------------------------------------------------------------------
    int c_out;               // real output channels
    int block_n;             // N tile (multiple of 16, <= 256); weights are padded to n_tiles*block_n rows
    const void* wpack;       // [n_tile][c_block][tap][part][c_blk/8][block_n][8] 16-bit (part: hi, lo)
    int wpack_stacked;       // 1: [n_tile][c_block][tap][c_blk/8][part][block_n][8] -- hi and lo rows of a chunk are
                             // adjacent, so A_hi*[W_hi;W_lo] is ONE MMA of N = 2*block_n (2 MMAs per k-step, not 3)
    const float* wref;       // validator only: [c_out][tap][c_in] fp32 (BN already folded)
    const float* bias;       // [n_tiles*block_n] fp32 (BN folded; zero padded)
    int relu;
    int precision;           // DISCO_PREC_*
    // ---- output ----------------------------------------------------------------------------------
    int out_mode;            // DISCO_OUT_*
    void* out[2];            // OUT_ACT: out[0] = hi; OUT_F32: out[0] (channels < out_split), out[1] (rest)
    long long out_lo_off;    // OUT_ACT + BF16X3: elements from hi to lo
    int out_split;           // OUT_F32: first channel of out[1]; == c_out when out[1] unused
    // ---- optional chained 1x1 conv on the (ReLU'd) result, computed in the same kernel -------------
    // out = [relu](W2 * relu(conv(x) + bias) + bias2): the intermediate never leaves the SM (heads:
    // DetModelBase.py:295-296,319-330).  With a chain, out/out_split/out_mode describe the CHAIN output
    // (OUT_F32) and c_out the intermediate width (<= 64, one N tile, stacked bf16x3 weights).
    const void* chain_wpack; // [c_out/8][part][chain_block_n][8] bf16 (stacked layout, K = c_out)
    const float* chain_bias; // [chain_block_n]
    int chain_c_out;         // 0 = no chain
    int chain_relu;
    /* optional device flag (int): when it reads 0 every lo element of the sources is zero (exact 0/1 occupancy input written by
     * disco_bev_pack / disco_bev_scatter_batched), so the kernel skips the lo-plane loads and the A_lo*W_hi pass; NULL = use lo */
    const int* src_lo_nonzero;
    /* output-parity ("sub-pixel") class of conv(cat(nearest_up2(src[0]), src[1])): subpix = 1 computes ONLY the output pixels
     * (2a + sub_py, 2b + sub_px); wpack then holds, per 16-channel block, the 4 pre-summed taps of that class for source 0
     * (ascending tap index inside the 3x3 window: rows sub_py..sub_py+1, columns sub_px..sub_px+1) and all 9 taps for source 1
     * (see disconet_b200/plan.py::pack_conv_subpix).  Needs taps 9, stride 1, src_up = {1, 0}, c_blk 16, bf16x3.  Four launches
     * (one per class) replace one launch with src_up[0] = 1 and execute 4/9 of its source-0 MMAs.
     * subpix = 2 is the FUSED form (C_out <= 64, one stacked N tile, activation output): one launch whose work items are 16 x 8
     * tiles of LOW-RES positions and keep the four class accumulators side by side in TMEM, so every staged operand (the low-res
     * patch of source 0, the 34 x 18 window of source 1) is shared by the four classes; wpack is then the slot stream of
     * disconet_b200/plan.py::pack_conv_subpix_fused and sub_py / sub_px are ignored. */
    int subpix, sub_py, sub_px;
};

int disco_conv_tc_launch(const disco_conv_desc* d, void* stream);
int disco_conv_ref_launch(const disco_conv_desc* d, void* stream);
int disco_conv_tc_smem_bytes(const disco_conv_desc* d);
