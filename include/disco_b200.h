/* libdisco_b200 -- C ABI of the B200-native DiscoNet hot path (sm_100a).
 *
 * The reference (ai4ce/DiscoNet -> coperception) has no FFI on this path: its boundary is the Python
 * class coperception.models.det.DiscoNet (DiscoNet.py:21-129) calling stock torch ops.  This header is
 * the ABI underneath the drop-in Python class (disconet_b200/det.py); each entry point names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions: every function returns 0 or a negative DISCO_E* code, never throws or aborts; text via
 * disco_last_error().  Pointers are raw CUDA device pointers unless marked "host".  The caller owns all
 * buffers.  `stream` is a cudaStream_t; no call synchronises.  Activations are NHWC 16-bit tensors
 * ("activation buffers"): precision DISCO_PREC_FP16 = one fp16 tensor; DISCO_PREC_BF16X3 = value is
 * hi + lo, two bf16 tensors, `lo` located `*_lo_off` ELEMENTS after `hi`.
 */
#ifndef DISCO_B200_H
#define DISCO_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DISCO_OK 0
#define DISCO_EINVAL (-1)  /* bad argument / unsupported shape                     */
#define DISCO_ECUDA (-2)   /* CUDA runtime error (launch, memset, attribute)      */
#define DISCO_EARCH (-3)   /* device is not compute capability 10.x               */
#define DISCO_EKERNEL (-4) /* reserved                                            */

#define DISCO_PREC_FP16 0
#define DISCO_PREC_BF16X3 1
#define DISCO_OUT_ACT 0 /* 16-bit activation buffer                                */
#define DISCO_OUT_F32 1 /* fp32 NHWC, optionally split into two tensors            */

int disco_version(void);
int disco_last_error(char* buf /* host */, size_t len);
/* Fails with DISCO_EARCH unless the current device is sm_100 (there is no other code path). */
int disco_device_check(void);

/* One conv + folded BatchNorm(eval) + optional ReLU layer on the tcgen05 tensor cores.
 * Replaces F.conv2d / Conv3D(1x1x1) + bn + relu of Backbone.encode/decode (Backbone.py:102-136,173-237),
 * the nearest-x2 F.interpolate + torch.cat feeding conv5_1..conv8_1 (:176,195,214,233; src_up / two
 * sources), the heads (DetModelBase.py:283-351) and PWF conv1_1 (DiscoNet.py:148). */
typedef struct disco_conv_desc {
    const void* src[2];      /* NHWC 16-bit sources, concatenated along channels (src[1] may be NULL) */
    long long src_lo_off[2]; /* BF16X3: elements from hi to lo                                         */
    int src_c[2];            /* channels per source, multiples of 16                                   */
    int src_up[2];           /* 1: source is (h_in/2, w_in/2) and is nearest-upsampled x2 on the fly   */
    int n, h_in, w_in;       /* logical conv input size                                                */
    int h_out, w_out;        /* (h_in-1)/stride+1, (w_in-1)/stride+1                                   */
    int stride;              /* 1 | 2                                                                  */
    int taps;                /* 9: 3x3 pad 1; 1: 1x1                                                   */
    int c_blk;               /* channels per K stage (16|32|64), divides src_c[*]                      */
    int c_out;               /* real output channels                                                   */
    int block_n;             /* N tile, multiple of 16, <= 256                                         */
    const void* wpack;       /* [n_tile][c_block][tap][part][c_blk/8][block_n][8] 16-bit               */
    int wpack_stacked;       /* 1: [..][tap][c_blk/8][part][block_n][8] (hi|lo rows adjacent; bf16x3, N<=128) */
    const float* wref;       /* disco_conv_reference only: [c_out][tap][c_in] fp32                     */
    const float* bias;       /* [n_tiles*block_n] fp32                                                 */
    int relu;
    int precision;           /* DISCO_PREC_*                                                           */
    int out_mode;            /* DISCO_OUT_*                                                            */
    void* out[2];            /* OUT_ACT: out[0]=hi; OUT_F32: out[0] gets channels < out_split, out[1] the rest */
    long long out_lo_off;
    int out_split;
    /* optional chained 1x1 conv on the ReLU'd result inside the same kernel (heads): when chain_c_out > 0,
     * out/out_split/out_mode describe the chain output (OUT_F32) and c_out (<= 64) the intermediate width */
    const void* chain_wpack; /* [c_out/8][part][chain_block_n][8] bf16, K = c_out                         */
    const float* chain_bias; /* [chain_block_n]                                                          */
    int chain_c_out;         /* 0 = no chain                                                             */
    int chain_relu;
} disco_conv_desc;

int disco_conv_forward(const disco_conv_desc* d /* host */, void* stream);
/* CUDA-core direct convolution with the same descriptor (validation only). */
int disco_conv_reference(const disco_conv_desc* d /* host */, void* stream);
/* Dynamic shared memory the tensor-core kernel needs for this descriptor (or a negative error). */
int disco_conv_smem_bytes(const disco_conv_desc* d /* host */);

/* Dense BEV fp32 [n_pix, z] (the DataLoader's padded_voxel_points, V2XSimDet.py:293-302, in the layout
 * DiscoNet.forward receives it, DiscoNet.py:42) -> 16-channel NHWC activation buffer (z <= 16). */
int disco_bev_pack(const float* bev, long long n_pix, int z, void* out_hi, long long out_lo_off, int precision,
                   void* stream);

/* NHWC activation buffer -> fp32 NCHW (layout of the tensors DiscoNet.forward returns when kd_flag == 1). */
int disco_act_unpack_nchw(const void* act_hi, long long lo_off, int precision, int n, int h, int w, int c,
                          float* out_nchw, void* stream);

/* voxelize_occupy (utils/data_util.py:625-717): points [n_points, point_stride] fp32 (x,y,z first) ->
 * occupancy bitmap (ceil(X*Y*Z/32) words, key = (x*Y + y)*Z + z), lexicographically sorted unique voxel
 * indices [*n_voxels, 3] int32 (buffer sized for min(n_points, X*Y*Z) rows; may be NULL) and an optional
 * dense float grid [X,Y,Z].  extents = {xmin,xmax,ymin,ymax,zmin,zmax} and voxel_size are host float64;
 * the floor-divide runs in float64 like numpy's.  dims (host int[3]) = grid size. */
int disco_voxelize_occupy(const float* points, int n_points, int point_stride, const double* extents /* host */,
                          const double* voxel_size /* host */, const int* dims /* host */, unsigned int* bitmap,
                          int* voxel_indices, int* n_voxels, float* dense, void* stream);

/* Dataset scatter (datasets/V2XSimDet.py:293-302): voxel indices [n,3] int32 -> dense BEV
 * bev[y, X-1-x, z] = 1 (== np.rot90(vox, 3)), as fp32 [Y,X,Z] and/or a 16-bit NHWC activation with
 * act_c channels per cell.  Either output may be NULL. */
int disco_bev_scatter(const int* voxel_indices, int n_voxels, const int* dims /* host */, float* bev_f32,
                      void* act_hi, int act_c, int precision, void* stream);

/* DiscoGraph fusion block: per-ego affine warp of every neighbour map (DetModelBase.py:139-209),
 * PixelWeightedFusionSoftmax tail (DiscoNet.py:150-153), agent-axis softmax and weighted sum
 * (DiscoNet.py:83-111) in one launch. */
typedef struct disco_fusion_desc {
    const void* feat_hi;   /* [A*B, h, w, C] activation buffer, agent-major rows a*B + b             */
    long long feat_lo_off;
    int precision;
    const float* en;       /* [A*B, h, w, 2*hid] fp32: PWF conv1_1+bn1_1 ego half (with bias) | nb half */
    int hid;               /* 128                                                                    */
    const float* w2; const float* b2; /* [32,hid],[32]  conv1_2+bn1_2 folded                        */
    const float* w3; const float* b3; /* [8,32],[8]     conv1_3+bn1_3 folded                        */
    const float* w4; const float* b4; /* [1,8],[1]      conv1_4                                     */
    const double* trans;   /* [B, A, A, 4, 4] float64 trans_matrices (device)                        */
    const int* num_agent;  /* [B] int32 (device)                                                     */
    int B, A, h, w, C;     /* C = 256 | 512                                                          */
    int only_v2i;
    float trans_scale;     /* 4/128 (DetModelBase.py:163)                                            */
    void* out_hi;          /* fused features, same layout as feat                                    */
    long long out_lo_off;
    float* weights;        /* optional [B, A(ego), A(neighbour id), h, w] softmax weights (unflipped) */
    int row_begin, row_end; /* ego rows n = a*B+b computed by this call; output row = n - row_begin      */
    const int* outage;     /* optional [B, A] int32 (device): 1 = outage, ego keeps its own features    */
} disco_fusion_desc;

int disco_fusion_forward(const disco_fusion_desc* d /* host */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISCO_B200_H */
