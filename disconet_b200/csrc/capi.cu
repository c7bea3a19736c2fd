// C-ABI of libdisco_b200.so (declared in include/disco_b200.h).  Plain pointers and sizes only; every
// entry point returns 0 or a negative DISCO_E* code and never throws / aborts; text via
// disco_last_error().  All launches go to the caller's stream, no internal synchronisation.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"
#include "conv.h"
#include "ops.h"
#include "train.h"

static thread_local char g_err[512] = "";

void disco_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

// ABI version: bumped whenever a descriptor struct or an entry-point signature changes; disconet_b200/_lib.py refuses to drive a
// library built from other sources (raw-pointer descriptors read with the wrong layout would corrupt device memory silently).
int disco_version(void) { return 200; }

int disco_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return DISCO_EINVAL;
    strncpy(buf, g_err, len - 1);
    buf[len - 1] = 0;
    return DISCO_OK;
}

// Fails unless the current device is compute capability 10.x (the only target this library is built for).
int disco_device_check(void) {
    int dev = 0;
    DISCO_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    DISCO_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
    if (p.major != 10) {
        disco_set_error("device %d is sm_%d%d; libdisco_b200 contains sm_100a code only", dev, p.major, p.minor);
        return DISCO_EARCH;
    }
    return DISCO_OK;
}

int disco_conv_forward(const disco_conv_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_conv_tc_launch(d, stream);
}
int disco_conv_reference(const disco_conv_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_conv_ref_launch(d, stream);
}
int disco_conv_smem_bytes(const disco_conv_desc* d) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_conv_tc_smem_bytes(d);
}

int disco_bev_pack(const float* bev, long long n_pix, int z, void* out_hi, long long out_lo_off, int precision,
                   int* lo_nonzero, void* stream) {
    return disco_bev_pack_launch(bev, n_pix, z, out_hi, out_lo_off, precision, lo_nonzero, stream);
}

int disco_act_unpack_nchw(const void* act_hi, long long lo_off, int precision, int n, int h, int w, int c,
                          float* out_nchw, void* stream) {
    return disco_act_unpack_nchw_launch(act_hi, lo_off, precision, n, h, w, c, out_nchw, stream);
}

int disco_voxelize_occupy(const float* points, int n_points, int point_stride, const double* extents,
                          const double* voxel_size, const int* dims, unsigned int* bitmap, int* voxel_indices,
                          int* n_voxels, float* dense, void* stream) {
    return disco_voxelize_launch(points, n_points, point_stride, extents, voxel_size, dims, bitmap, voxel_indices,
                                 n_voxels, dense, stream);
}

int disco_voxelize_occupy_batched(const float* points, const int* n_points, int n_sweeps, int p_max, int point_stride,
                                  const double* extents, const double* voxel_size, const int* dims, unsigned int* bitmap,
                                  int* block_count, int* voxel_indices, int m_max, int* n_voxels, void* stream) {
    return disco_voxelize_batched_launch(points, n_points, n_sweeps, p_max, point_stride, extents, voxel_size, dims, bitmap, block_count,
                                         voxel_indices, m_max, n_voxels, stream);
}

int disco_bev_scatter(const int* voxel_indices, int n_voxels, const int* dims, float* bev_f32, void* act_hi,
                      int act_c, int precision, void* stream) {
    return disco_bev_scatter_launch(voxel_indices, n_voxels, dims, bev_f32, act_hi, act_c, precision, stream);
}

int disco_bev_scatter_batched(const int* voxel_indices, const int* counts, int n, int m_max, const int* dims, void* act_hi,
                              long long act_lo_off, int act_c, int precision, int* lo_nonzero, void* stream) {
    return disco_bev_scatter_batched_launch(voxel_indices, counts, n, m_max, dims, act_hi, act_lo_off, act_c, precision, lo_nonzero, stream);
}

int disco_fusion_forward(const disco_fusion_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_fusion_launch(d, stream);
}

int disco_det_candidates(const float* loc, const float* cls, const float* anchors, long long anchors_per_agent,
                         long long anchor_agent_stride, int n_agents, float thresh, int max_cand, int* count, float* corners,
                         float* scores, int* index, void* stream) {
    return disco_det_candidates_launch(loc, cls, anchors, anchors_per_agent, anchor_agent_stride, n_agents, thresh, max_cand, count,
                                       corners, scores, index, stream);
}

// ---- detection post-processing (row f3) and the regression loss of the training step (row f4) -------------------
long long disco_nms_workspace_bytes(int n_sets, int kmax) {
    if (n_sets <= 0 || kmax <= 0) { disco_set_error("nms_workspace_bytes: bad sizes"); return DISCO_EINVAL; }
    return (long long)disco_nms_workspace_bytes_impl(n_sets, kmax);
}
int disco_nms_rotated(const void* corners, int corners_f64, const float* scores, const int* ids, const int* count, int n_sets,
                      int cap, int kmax, float score_thresh, double iou_thresh, void* workspace, long long workspace_bytes,
                      int* keep, int* n_keep, int* n_valid, void* stream) {
    return disco_nms_rotated_launch(corners, corners_f64, scores, ids, count, n_sets, cap, kmax, score_thresh, iou_thresh, workspace,
                                    (size_t)workspace_bytes, keep, n_keep, n_valid, stream);
}
int disco_corner_loss(const float* pred, const float* target, const float* anchors, const unsigned char* mask, long long n_entries,
                      int t_len, float inv_n, double* loss_sum, float* grad, void* stream) {
    return disco_corner_loss_launch(pred, target, anchors, mask, n_entries, t_len, inv_n, loss_sum, grad, stream);
}

// ---- BEV segmentation U-Net ---------------------------------------------------------------
int disco_maxpool2(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision, int n, int h,
                   int w, int c, void* stream) {
    return disco_maxpool2_launch(src_hi, src_lo_off, dst_hi, dst_lo_off, precision, n, h, w, c, stream);
}
int disco_upsample_bilinear2x(const void* src_hi, long long src_lo_off, void* dst_hi, long long dst_lo_off, int precision, int n,
                              int h, int w, int c, void* stream) {
    return disco_upsample_bilinear2x_launch(src_hi, src_lo_off, dst_hi, dst_lo_off, precision, n, h, w, c, stream);
}
int disco_maxpool2_backward(const void* x_hi, long long x_lo_off, int precision, const float* g, float* gx, int n, int h, int w, int c,
                            void* stream) {
    return disco_maxpool2_backward_launch(x_hi, x_lo_off, precision, g, gx, n, h, w, c, stream);
}
int disco_upsample_bilinear2x_backward(const float* g_up, float* gs, int n, int h, int w, int c, void* stream) {
    return disco_upsample_bilinear2x_backward_launch(g_up, gs, n, h, w, c, stream);
}
int disco_nhwc_to_nchw(const float* src, int n, int h, int w, int c_src, int c, float* dst, void* stream) {
    return disco_nhwc_to_nchw_launch(src, n, h, w, c_src, c, dst, stream);
}

// ---- training mode ----------------------------------------------------------------------
int disco_bn_train_forward(const disco_bn_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_bn_train_forward_launch(d, stream);
}
int disco_bn_train_backward(const disco_bn_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_bn_train_backward_launch(d, stream);
}
int disco_kd_kl(const float* student, const float* teacher, int n, int c, long long hw, double* loss_sum, float* grad,
                float grad_scale, void* stream) {
    return disco_kd_kl_launch(student, teacher, n, c, hw, loss_sum, grad, grad_scale, stream);
}
int disco_focal_loss(const float* logits, const float* target, int k, long long n_anchor, float gamma, float alpha, int use_alpha,
                     const float* grad_out, long long grad_out_stride, float* out, void* stream) {
    return disco_focal_loss_launch(logits, target, k, n_anchor, gamma, alpha, use_alpha, grad_out, grad_out_stride, out, stream);
}
int disco_pack_weights(const disco_pack_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_pack_weights_launch(d, stream);
}
int disco_grad_pack(const float* a, int ca, const float* b, int cb, long long n_pix, void* out_hi, long long out_lo_off,
                    void* stream) {
    return disco_grad_pack_launch(a, ca, b, cb, n_pix, out_hi, out_lo_off, stream);
}
int disco_channel_sum(const float* src, long long n_pix, int c, double* sums, float* out, void* stream) {
    return disco_channel_sum_launch(src, n_pix, c, sums, out, stream);
}
int disco_nchw_to_nhwc(const float* src, int n, int c, int h, int w, float* dst, void* stream) {
    return disco_nchw_to_nhwc_launch(src, n, c, h, w, dst, stream);
}
int disco_add_f32(float* dst, const float* a, const float* b, long long n, void* stream) {
    return disco_add_f32_launch(dst, a, b, n, stream);
}
int disco_conv_wgrad(const disco_wgrad_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_wgrad_tc_launch(d, stream);
}
int disco_conv_wgrad_reference(const disco_wgrad_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_wgrad_ref_launch(d, stream);
}
int disco_conv_wgrad_splits(const disco_wgrad_desc* d) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_wgrad_splits(d);
}
int disco_pwf_train_forward(const disco_pwf_train_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_pwf_train_forward_launch(d, stream);
}
int disco_fusion_combine_backward(const disco_pwf_train_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_fusion_combine_backward_launch(d, stream);
}
int disco_pwf_train_backward(const disco_pwf_train_desc* d, void* stream) {
    if (!d) { disco_set_error("null descriptor"); return DISCO_EINVAL; }
    return disco_pwf_train_backward_launch(d, stream);
}

}  // extern "C"
