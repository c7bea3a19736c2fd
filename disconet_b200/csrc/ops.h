// Non-GEMM kernels of the hot path: BEV pack / voxelize / scatter / DiscoGraph fusion.
#pragma once
#include <stdint.h>

// DiscoGraph fusion block descriptor (mirrors include/disco_b200.h).
struct disco_fusion_desc {
    // collaboration-layer features, agent-major rows n = a*B + b, NHWC
    const void* feat_hi;       // 16-bit [A*B, h, w, C]
    long long feat_lo_off;     // BF16X3: elements from hi to lo
    int precision;             // DISCO_PREC_*
    // PWF first layer applied per agent by the conv kernel (fp32 [A*B, h, w, 2*hid]):
    //   channels [0,hid)     = s*(W_ego  x) + s*(b - mean) + beta      (BN folded)
    //   channels [hid,2hid)  = s*(W_nb   x)
    const float* en;
    int hid;                   // 128
    // PWF tail (BN folded), fp32 row-major
    const float* w2; const float* b2;   // [h2, hid], [h2]     (h2 = 32)
    const float* w3; const float* b3;   // [h3, h2],  [h3]     (h3 = 8)
    const float* w4; const float* b4;   // [1, h3],   [1]
    // scene description
    const double* trans;       // [B, A, A, 4, 4] float64 (trans_matrices)
    const int* num_agent;      // [B]
    int B, A, h, w, C;
    int only_v2i;
    float trans_scale;         // 4/128 (DetModelBase.py:163)
    // outputs
    void* out_hi;              // fused features, same layout as feat
    long long out_lo_off;
    float* weights;            // optional [B, A(ego), A(neighbour id), h, w] softmax weights, unflipped frame
    // ego rows (n = a*B + b) computed by this call: [row_begin, row_end); output row = n - row_begin.
    // (0, A*B) = everything; a rank of an agent-sharded run passes its own slice.
    int row_begin, row_end;
    const int* outage;         // optional [B, A] int32: 1 = communication outage for that ego (DetModelBase.py:129-137,
                               // DiscoNet.py:68-69): the ego keeps its own features
    const float* wpre;         // optional [B, A(ego), A(neighbour id), h, w]: precomputed PWF output maps (training mode,
                               // per-pair batch statistics -- fusion_train.cu); when set, en / w2..b4 are unused
};

int disco_fusion_launch(const disco_fusion_desc* d, void* stream);
int disco_bev_pack_launch(const float* bev, long long n_pix, int z, void* out_hi, long long out_lo_off, int precision,
                          int* lo_nonzero, void* stream);
int disco_act_unpack_nchw_launch(const void* act_hi, long long lo_off, int precision, int n, int h, int w, int c,
                                 float* out_nchw, void* stream);
int disco_voxelize_launch(const float* points, int n_points, int point_stride, const double* extents,
                          const double* voxel_size, const int* dims, unsigned int* bitmap, int* voxel_indices,
                          int* n_voxels, float* dense, void* stream);
int disco_voxelize_batched_launch(const float* points, const int* n_points, int n_sweeps, int p_max, int point_stride,
                                  const double* extents, const double* voxel_size, const int* dims, unsigned int* bitmap,
                                  int* block_count, int* voxel_indices, int m_max, int* n_voxels, void* stream);
int disco_bev_scatter_launch(const int* voxel_indices, int n_voxels, const int* dims, float* bev_f32, void* act_hi,
                             int act_c, int precision, void* stream);

int disco_bev_scatter_batched_launch(const int* voxel_indices, const int* counts, int n, int m_max, const int* dims, void* act_hi,
                                     long long act_lo_off, int act_c, int precision, int* lo_nonzero, void* stream);

// BEV-segmentation U-Net data movement (seg.cu)
int disco_maxpool2_launch(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision, int n,
                          int h, int w, int c, void* stream);
int disco_upsample_bilinear2x_launch(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision,
                                     int n, int h, int w, int c, void* stream);
int disco_nhwc_to_nchw_launch(const float* src, int n, int h, int w, int c_src, int c, float* dst, void* stream);

// Detection candidates (misc.cu): score > thresh anchors of every agent -> corners / scores / anchor index, compacted
int disco_det_candidates_launch(const float* loc, const float* cls, const float* anchors, long long anchors_per_agent,
                                long long anchor_agent_stride, int n_agents, float thresh, int max_cand, int* count,
                                float* corners, float* scores, int* index, void* stream);
int disco_maxpool2_backward_launch(const void* x_hi, long long x_lo_off, int precision, const float* g, float* gx, int n, int h, int w,
                                   int c, void* stream);
int disco_upsample_bilinear2x_backward_launch(const float* g_up, float* gs, int n, int h, int w, int c, void* stream);

// Detection post-processing + regression loss (post.cu)
#include <stddef.h>
size_t disco_nms_workspace_bytes_impl(int n, int kmax);
int disco_nms_rotated_launch(const void* corners, int corners_f64, const float* scores, const int* ids, const int* count, int n,
                             int cap, int kmax, float score_thresh, double iou_thresh, void* workspace, size_t workspace_bytes,
                             int* keep, int* n_keep, int* n_valid, void* stream);
int disco_corner_loss_launch(const float* pred, const float* target, const float* anchors, const unsigned char* mask,
                             long long n_entries, int t_len, float inv_n, double* loss_sum, float* grad, void* stream);
