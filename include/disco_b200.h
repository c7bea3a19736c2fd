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
    int src_up[2];           /* 1: source is (h_in/2, w_in/2) and is nearest-upsampled x2 on the fly;
                              * 2: zero-stuffed x2 instead (transposed stride-2 conv = data gradient)   */
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
} disco_conv_desc;

int disco_conv_forward(const disco_conv_desc* d /* host */, void* stream);
/* CUDA-core direct convolution with the same descriptor (validation only). */
int disco_conv_reference(const disco_conv_desc* d /* host */, void* stream);
/* Dynamic shared memory the tensor-core kernel needs for this descriptor (or a negative error). */
int disco_conv_smem_bytes(const disco_conv_desc* d /* host */);

/* Dense BEV fp32 [n_pix, z] (the DataLoader's padded_voxel_points, V2XSimDet.py:293-302, in the layout
 * DiscoNet.forward receives it, DiscoNet.py:42) -> 16-channel NHWC activation buffer (z <= 16).
 * lo_nonzero (optional device int): set to 0, then to 1 if any element needs a non-zero lo part (i.e. is not exactly
 * representable in bf16; 0/1 occupancy is) -- feeds disco_conv_desc.src_lo_nonzero of the first conv. */
int disco_bev_pack(const float* bev, long long n_pix, int z, void* out_hi, long long out_lo_off, int precision,
                   int* lo_nonzero, void* stream);

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

/* voxelize_occupy for n_sweeps point clouds in three launches: points [n_sweeps, p_max, point_stride] fp32, n_points [n_sweeps]
 * (device) -> voxel_indices [n_sweeps, m_max, 3] int32 (sorted unique per sweep, rows >= n_voxels[s] = -1) and n_voxels
 * [n_sweeps] -- exactly the (indices, counts) pair disco_bev_scatter_batched consumes.  bitmap: n_sweeps * ceil(X*Y*Z/32) words,
 * block_count: n_sweeps * ceil(words/1024) ints (scratch).  One extents / voxel_size for the whole batch (host float64). */
int disco_voxelize_occupy_batched(const float* points, const int* n_points, int n_sweeps, int p_max, int point_stride,
                                  const double* extents /* host */, const double* voxel_size /* host */, const int* dims /* host */,
                                  unsigned int* bitmap, int* block_count, int* voxel_indices, int m_max, int* n_voxels, void* stream);

/* Dataset scatter (datasets/V2XSimDet.py:293-302): voxel indices [n,3] int32 -> dense BEV
 * bev[y, X-1-x, z] = 1 (== np.rot90(vox, 3)), as fp32 [Y,X,Z] and/or a 16-bit NHWC activation with
 * act_c channels per cell.  Either output may be NULL. */
int disco_bev_scatter(const int* voxel_indices, int n_voxels, const int* dims /* host */, float* bev_f32,
                      void* act_hi, int act_c, int precision, void* stream);

/* The same scatter for a whole batch of agents, straight into the encoder's input activation (skips the dense fp32 BEV of
 * V2XSimDet.py:293-302 and its 3.4 MB/agent host->device copy): voxel_indices [n, m_max, 3] int32 (device), counts [n]
 * (device; rows >= counts[a] ignored), act = activation buffer [parts, n, Y, X, act_c] (zeroed by the call). */
int disco_bev_scatter_batched(const int* voxel_indices, const int* counts, int n, int m_max, const int* dims /* host */,
                              void* act_hi, long long act_lo_off, int act_c, int precision, int* lo_nonzero /* optional: set to 0 */,
                              void* stream);

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
    const float* wpre;     /* optional [B, A, A, h, w]: precomputed PWF output maps (training mode); en/w2..b4 unused */
} disco_fusion_desc;

int disco_fusion_forward(const disco_fusion_desc* d /* host */, void* stream);

/* The per-anchor half of apply_nms_det (utils/detection_util.py:256-373) for n_agents agents: foreground probability
 * softmax(cls)[1] of cls [n_agents*anchors_per_agent, 2], threshold (postprocess.py:85), bev_box_decode_torch (:376-400)
 * of loc [.., 6] against anchors (agent stride anchor_agent_stride floats; 0 = shared), rotated corners (obj_util.py:271-359).
 * Survivors are compacted per agent: count[agent] (may exceed max_cand: then only max_cand were stored), corners
 * [n_agents, max_cand, 4, 2], scores / index [n_agents, max_cand] (index = anchor number inside the agent), unordered. */
int disco_det_candidates(const float* loc, const float* cls, const float* anchors, long long anchors_per_agent,
                         long long anchor_agent_stride, int n_agents, float thresh, int max_cand, int* count, float* corners,
                         float* scores, int* index, void* stream);

/* Rotated-box NMS: `non_max_suppression` of utils/postprocess.py:72-115 (called with threshold 0.01 by apply_nms_det,
 * detection_util.py:349-351, and late_fusion, :962-964) for n_sets independent candidate sets in one go.
 * corners [n_sets, cap, 4, 2] (float32, or float64 when corners_f64), scores [n_sets, cap], ids [n_sets, cap] or NULL
 * (tie-break key: the anchor number; NULL = the position), count [n_sets] (device; NULL = every set holds `cap` boxes).
 * Keeps scores > score_thresh, orders by descending score (ties: larger id first = `scores.argsort()[::-1]` of a
 * stable argsort), then greedily drops every box whose polygon IoU (float64 convex clip) with a picked one is > iou_thresh.
 * keep [n_sets, kmax] receives the picked POSITIONS (into corners/scores) in pick order, n_keep [n_sets] their number,
 * n_valid [n_sets] (optional) the number of boxes above score_thresh (> kmax means the set was truncated to the kmax best).
 * cap <= 8192, ids < 2^19.  workspace: disco_nms_workspace_bytes(n_sets, kmax) bytes of device memory. */
long long disco_nms_workspace_bytes(int n_sets, int kmax);
int disco_nms_rotated(const void* corners, int corners_f64, const float* scores, const int* ids, const int* count, int n_sets,
                      int cap, int kmax, float score_thresh, double iou_thresh, void* workspace, long long workspace_bytes,
                      int* keep, int* n_keep, int* n_valid, void* stream);

/* `FaFModule.corner_loss` (utils/CoDetModule.py:80-105), value and gradient in one launch: for every entry e of
 * pred / target [n_entries, 6] with mask[e] != 0 (reg_loss_mask, one byte per entry; t_len consecutive entries share an
 * anchor of anchors [n_entries / t_len, 6]): decode both boxes (bev_box_decode_torch), four rotated corners each
 * (center_to_corner_box2d_torch), sum of the corner distances.  *loss_sum (float64, zeroed by the call) receives the
 * UNSCALED sum; grad (optional, [n_entries, 6]) receives inv_n * d(sum)/d(pred), zeros where the mask is 0. */
int disco_corner_loss(const float* pred, const float* target, const float* anchors, const unsigned char* mask, long long n_entries,
                      int t_len, float inv_n, double* loss_sum, float* grad, void* stream);

/* ---- BEV segmentation U-Net (models/seg/SegModelBase.py) around disco_conv_forward ------------------------
 * nn.MaxPool2d(2) of the Down blocks (:113-123): activation buffer [n,h,w,c] -> [n,h/2,w/2,c]. */
int disco_maxpool2(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision, int n, int h,
                   int w, int c, void* stream);
/* nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) of the Up blocks (:126-142): [n,h,w,c] -> [n,2h,2w,c];
 * the torch.cat([skip, up]) that follows is the two-source gather of disco_conv_forward. */
int disco_upsample_bilinear2x(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision, int n,
                              int h, int w, int c, void* stream);
/* autograd backward of the two: x = the pool's INPUT activation [n,h,w,c], g fp32 [n,h/2,w/2,c] -> gx fp32 [n,h,w,c] (first
 * maximum of each 2x2 block, torch's convention);  g_up fp32 [n,2h,2w,c] -> gs fp32 [n,h,w,c] (transpose of the bilinear
 * interpolation).  (h, w) are the sizes of the pool input / of the upsample source. */
int disco_maxpool2_backward(const void* x_hi, long long x_lo_off, int precision, const float* g, float* gx, int n, int h, int w, int c,
                            void* stream);
int disco_upsample_bilinear2x_backward(const float* g_up, float* gs, int n, int h, int w, int c, void* stream);
/* fp32 NHWC [n,h,w,c_src] -> fp32 NCHW [n,c,h,w] (first c channels): layout of the logits OutConv returns (:145-151). */
int disco_nhwc_to_nchw(const float* src, int n, int h, int w, int c_src, int c, float* dst, void* stream);

/* ======================================================================================================
 * Training mode: replaces, for model.train(), the batch-statistics F.batch_norm calls of
 * Backbone.encode/decode + heads + PixelWeightedFusionSoftmax and torch.autograd's backward of the whole path
 * (utils/CoDetModule.py:289-291 `loss.backward()`).  Forward convs run through disco_conv_forward with the raw
 * (unfolded) weights and fp32 output; data gradients are disco_conv_forward calls with transposed/flipped
 * weights (stride-2 layers: src_up = 2, zero-stuffed source).
 * ====================================================================================================== */

/* A channel slice of an fp32 NHWC gradient tensor; pool = 1: tensor is (2h x 2w) and each 2x2 block is summed
 * (backward of the nearest x2 upsample feeding conv5_1..conv8_1, Backbone.py:176,195,214,233). */
typedef struct disco_grad_src {
    const float* ptr;
    int c_total;
    int c_off;
    int pool;
} disco_grad_src;

/* conv -> BatchNorm2d/3d(train) -> ReLU around the conv kernel (nn.BatchNorm semantics: biased batch variance to
 * normalise, momentum update of running_mean / unbiased running_var, num_batches_tracked += 1). */
typedef struct disco_bn_desc {
    const float* z;            /* [n*h*w, c] conv output incl. bias, fp32 NHWC                         */
    int n, h, w, c;            /* c in {8,16,32,64,128,256,512}                                       */
    const float* gamma;
    const float* beta;
    float* running_mean;       /* updated in place; may be NULL                                       */
    float* running_var;
    long long* num_batches_tracked;
    float momentum, eps;
    double* sums;              /* workspace [2*c]                                                     */
    float* stats;              /* [2*c] mean | rstd: written by forward, read by backward             */
    void* out_hi;              /* forward: y = relu(bn(z)) activation buffer                          */
    long long out_lo_off;
    int relu;
    disco_grad_src g[3];       /* backward: gradient sources wrt y (summed)                           */
    int n_g;
    void* dz_hi;               /* backward: gradient wrt z, activation buffer                         */
    long long dz_lo_off;
    float* dgamma;             /* [c] assigned                                                        */
    float* dbeta;              /* [c] assigned                                                        */
} disco_bn_desc;

int disco_bn_train_forward(const disco_bn_desc* d /* host */, void* stream);
int disco_bn_train_backward(const disco_bn_desc* d /* host */, void* stream);

/* Raw conv weights (fp32 OIHW) -> the bf16x3 operand image of disco_conv_forward (`wpack`), on the device.
 * transpose = 0: forward weights; 1: data-gradient weights of input channels [c0, c0 + n_real) -- B[n = ci][k = co]
 * with the taps flipped.  Replaces the host-side packing when the weights change every step (training). */
typedef struct disco_pack_desc {
    const float* w;          /* [co_src][ci_src][taps]                                                  */
    int co_src, ci_src, taps;
    int transpose;
    int c0;
    int n_real;              /* real rows of B (forward: co_src; transpose: slice width)                */
    int k_pad;               /* padded K channels, multiple of c_blk                                    */
    int block_n, c_blk, n_tiles, stacked;
    void* wpack;
    const float* bias_src;   /* optional [n_real]                                                        */
    float* bias;             /* optional [n_tiles*block_n] zero-padded copy                              */
} disco_pack_desc;

int disco_pack_weights(const disco_pack_desc* d /* host */, void* stream);

/* fp32 NHWC gradient tensors a [n_pix, ca] | b [n_pix, cb] (channel-concatenated; b may be NULL with cb = 0)
 * -> activation buffer [n_pix, ca + cb]. */
int disco_grad_pack(const float* a, int ca, const float* b, int cb, long long n_pix, void* out_hi, long long out_lo_off,
                    void* stream);
/* out[c] = sum over pixels of src[pix, c] (bias gradient of a conv without BatchNorm); sums: workspace [c] double */
int disco_channel_sum(const float* src, long long n_pix, int c, double* sums, float* out, void* stream);
/* fp32 NCHW -> NHWC (gradients of the returned KD feature maps arrive NCHW) */
int disco_nchw_to_nhwc(const float* src, int n, int c, int h, int w, float* dst, void* stream);
int disco_add_f32(float* dst, const float* a, const float* b, long long n, void* stream);

/* One KD term of FaFModule.get_kd_loss (utils/CoDetModule.py:334-382): nn.KLDivLoss(mean over elements) between
 * log_softmax(student) and softmax(teacher) over the channel axis of two NCHW fp32 maps [n, c, hw].  The SUM over elements
 * is added to *loss_sum (device double, caller zeroes it and divides by n*c*hw); grad (optional, NCHW) receives
 * (softmax(student) - softmax(teacher)) * grad_scale, i.e. d(mean KL)/d student for grad_scale = 1/(n*c*hw). */
int disco_kd_kl(const float* student, const float* teacher, int n, int c, long long hw, double* loss_sum, float* grad,
                float grad_scale, void* stream);

/* SoftmaxFocalClassificationLoss._compute_loss (utils/loss.py:322-394) per anchor, k <= 8 classes, logits / one-hot
 * targets [n_anchor, k] fp32.  grad_out == NULL: out = loss [n_anchor, k].  Otherwise out = gradient wrt the logits of
 * sum(grad_out * loss); grad_out_stride = 1 for a dense grad_out, 0 for a broadcast scalar (backward of torch.sum). */
int disco_focal_loss(const float* logits, const float* target, int k, long long n_anchor, float gamma, float alpha, int use_alpha,
                     const float* grad_out, long long grad_out_stride, float* out, void* stream);

/* Weight gradient dW[co][ci][kh][kw] = sum_pixels dz[p][co] * x[p (+) tap][ci] of a conv layer on the tensor cores
 * (MN-major tcgen05 operands, split-K over pixel tiles). */
typedef struct disco_wgrad_desc {
    const void* src[2];        /* forward input of the conv: as in disco_conv_desc                    */
    long long src_lo_off[2];
    int src_c[2];
    int src_up[2];
    int n, h_in, w_in, h_out, w_out;
    int stride, taps;
    const void* dz_hi;         /* gradient wrt the conv output, activation buffer [n,h_out,w_out,c_out] */
    long long dz_lo_off;
    int c_out;                 /* multiple of 16; <= 128 or a multiple of 128                          */
    float* partial;            /* workspace [splits][c_out][taps][c_in] fp32                           */
    int splits;                /* what `partial` was sized for (>= disco_conv_wgrad_splits)            */
    float* dw;                 /* [c_out][c_in_real][taps] fp32 (OIHW), assigned                       */
    int c_in_real;
    int passes;                /* 3 (bf16x3) | 1                                                       */
} disco_wgrad_desc;

int disco_conv_wgrad(const disco_wgrad_desc* d /* host */, void* stream);
int disco_conv_wgrad_reference(const disco_wgrad_desc* d /* host */, void* stream); /* CUDA-core validator */
int disco_conv_wgrad_splits(const disco_wgrad_desc* d /* host */);

/* PixelWeightedFusionSoftmax in train() mode (DiscoNet.py:86-95,148-155: one call per (scene, ego, neighbour) with
 * that call's batch statistics and a sequential running-statistics update) and the backward of the DiscoGraph
 * fusion block (DiscoNet.py:83-111 softmax / weighted sum, DetModelBase.py:139-169 affine warp). */
typedef struct disco_pwf_train_desc {
    const void* feat_hi;
    long long feat_lo_off;
    const float* en;           /* [A*B,h,w,2*hid] fp32: conv1_1 ego half (+bias) | neighbour half, raw weights */
    int hid;
    const float* g1; const float* be1;
    const float* w2; const float* b2; const float* g2; const float* be2;
    const float* w3; const float* b3; const float* g3; const float* be3;
    const float* w4; const float* b4;
    float eps, momentum;
    float* rm1; float* rv1; float* rm2; float* rv2; float* rm3; float* rv3;
    long long* nbt1; long long* nbt2; long long* nbt3;
    const double* trans;
    const int* num_agent;
    const int* outage;
    int B, A, h, w, C;
    int only_v2i;
    float trans_scale;
    double* psum;              /* [B*A*A][2][168] per-pair sum | sum of squares of the pre-BN activations */
    float* wlogit;             /* [B,A,A,h,w] PWF output maps (post-ReLU), 0 for unused pairs          */
    const float* dfused;       /* backward: [A*B,h,w,C] fp32 gradient wrt the fused map                */
    float* dwlogit;            /* [B,A,A,h,w]                                                         */
    float* dfeat;              /* [A*B,h,w,C] fp32, accumulated (caller zeroes)                        */
    float* den;                /* [A*B,h,w,2*hid] fp32, accumulated (caller zeroes)                    */
    float* gsum;               /* [B*A*A][2][168] backward workspace (sum g | sum g*xhat per pair)      */
    float* dparams;            /* [4697] accumulated: dg1 dbe1 dw2 dg2 dbe2 dw3 dg3 dbe3 dw4 db4       */
} disco_pwf_train_desc;

int disco_pwf_train_forward(const disco_pwf_train_desc* d /* host */, void* stream);
int disco_fusion_combine_backward(const disco_pwf_train_desc* d /* host */, void* stream);
int disco_pwf_train_backward(const disco_pwf_train_desc* d /* host */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISCO_B200_H */
