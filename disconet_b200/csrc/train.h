// Training-mode kernels of the DiscoNet hot path: batch-statistics BatchNorm forward /
// backward, gradient re-packing, weight-gradient GEMM, PixelWeightedFusion + DiscoGraph fusion backward.
// Mirrors the `disco_bn_desc`, `disco_wgrad_desc`, `disco_pwf_train_desc` structs of include/disco_b200.h.
#pragma once
#include <stdint.h>

// A slice of an fp32 NHWC gradient tensor: channels [c_off, c_off + C) of a tensor with c_total channels.
// pool = 1: the tensor lives at (2h x 2w) and the gradient is the sum of each 2x2 block (backward of the
// nearest x2 upsample that feeds conv5_1..conv8_1, Backbone.py:176,195,214,233).
struct disco_grad_src {
    const float* ptr;
    int c_total;
    int c_off;
    int pool;
};

// conv -> BatchNorm(train) -> ReLU, forward and backward around the conv kernel.
//   forward : z (conv output incl. bias, fp32 NHWC) -> batch mean / biased variance per channel ->
//             y = relu((z - mean) * rstd * gamma + beta) as an activation buffer; running statistics updated
//             like nn.BatchNorm (momentum, unbiased variance, num_batches_tracked += 1).
//   backward: g = (sum of gradient sources wrt y) * [y > 0];  dgamma = sum g*xhat;  dbeta = sum g;
//             dz = gamma * rstd * (g - mean(g) - xhat * mean(g*xhat))  written as an activation buffer
//             (the A operand of the data-gradient conv and of the weight-gradient GEMM).
struct disco_bn_desc {
    const float* z;                 // [n*h*w, c] fp32
    int n, h, w, c;                 // c % 8 == 0, c <= 512
    const float* gamma;
    const float* beta;
    float* running_mean;            // may be null (no update)
    float* running_var;
    long long* num_batches_tracked;
    float momentum, eps;
    double* sums;                   // workspace [2*c]
    float* stats;                   // [2*c]: mean | rstd   (written by forward, read by backward)
    void* out_hi;                   // forward output activation buffer
    long long out_lo_off;
    int relu;
    disco_grad_src g[3];            // backward: gradient sources wrt y
    int n_g;
    void* dz_hi;                    // backward output: gradient wrt z (activation buffer)
    long long dz_lo_off;
    float* dgamma;                  // [c]
    float* dbeta;                   // [c]
};

int disco_bn_train_forward_launch(const disco_bn_desc* d, void* stream);
int disco_bn_train_backward_launch(const disco_bn_desc* d, void* stream);

// fp32 NHWC gradient tensors (channel-concatenated, up to 2) -> activation buffer (hi/lo bf16)
int disco_grad_pack_launch(const float* a, int ca, const float* b, int cb, long long n_pix, void* out_hi,
                           long long out_lo_off, void* stream);
// out[c] = sum over pixels of src[pix, c]   (bias gradients of convs that are not followed by a BatchNorm)
int disco_channel_sum_launch(const float* src, long long n_pix, int c, double* sums, float* out, void* stream);
// fp32 NCHW -> fp32 NHWC (external gradients of the KD feature maps arrive in the layout they were returned in)
int disco_nchw_to_nhwc_launch(const float* src, int n, int c, int h, int w, float* dst, void* stream);
// dst[i] += src[i]  /  dst[i] = a[i] + b[i]
int disco_add_f32_launch(float* dst, const float* a, const float* b, long long n, void* stream);

// Raw conv weights (fp32 OIHW) -> the bf16x3 UMMA operand image disco_conv_forward consumes (same layout as
// disconet_b200/plan.py::pack_conv).  The weights change with every optimizer step, so a training step re-packs
// every layer: one launch per layer instead of a dozen host-side tensor ops.
//   transpose = 0: forward weights        B[n = co][k = ci][tap]
//   transpose = 1: data-gradient weights  B[n = ci - c0][k = co][tap'] = W[co][ci][taps-1-tap']  (flipped taps)
struct disco_pack_desc {
    const float* w;          // [co_src][ci_src][taps]
    int co_src, ci_src, taps;
    int transpose;
    int c0;                  // transpose: first input channel of the slice
    int n_real;              // rows of B that are real (forward: co_src; transpose: slice width)
    int k_pad;               // padded K channels (multiple of c_blk)
    int block_n, c_blk, n_tiles, stacked;
    void* wpack;             // [n_tile][k_pad/c_blk][tap][part][c_blk/8][block_n][8] (stacked: [..][c_blk/8][part][..])
    const float* bias_src;   // optional [n_real]
    float* bias;             // optional [n_tiles*block_n]: bias_src zero-padded (zeros when bias_src is null)
};
int disco_pack_weights_launch(const disco_pack_desc* d, void* stream);

// KD loss term: sum over elements of softmax(t) * (log softmax(t) - log_softmax(s)) over the channel axis of two
// NCHW fp32 maps is ADDED to *loss_sum; grad (optional, NCHW) = (softmax(s) - softmax(t)) * grad_scale.
int disco_kd_kl_launch(const float* student, const float* teacher, int n, int c, long long hw, double* loss_sum, float* grad,
                       float grad_scale, void* stream);

// Softmax focal classification loss per anchor (k <= 8 classes).  grad_out == null: out = loss [n_anchor, k];
// else out = d(sum grad_out * loss)/d logits [n_anchor, k] (grad_out element stride grad_out_stride: 1, or 0 for a
// broadcast scalar -- the backward of torch.sum).
int disco_focal_loss_launch(const float* logits, const float* target, int k, long long n_anchor, float gamma, float alpha,
                            int use_alpha, const float* grad_out, long long grad_out_stride, float* out, void* stream);

// Weight gradient of a 3x3 / 1x1 conv:  dW[co][tap][ci] = sum_pixels dz[p][co] * x[p (+) tap][ci]
// on the tensor cores (MN-major operands: both dz and x are pixel-major NHWC, the contraction runs over pixels).
struct disco_wgrad_desc {
    // input of the forward conv: same source description as disco_conv_desc
    const void* src[2];
    long long src_lo_off[2];
    int src_c[2];
    int src_up[2];
    int n, h_in, w_in, h_out, w_out;
    int stride, taps;
    // gradient wrt the conv output, activation buffer [n, h_out, w_out, c_out]
    const void* dz_hi;
    long long dz_lo_off;
    int c_out;               // multiple of 16
    // output
    float* partial;          // workspace [splits][c_out][taps][c_in] fp32
    int splits;              // split-K factor the caller sized `partial` for (>= 1)
    float* dw;               // [c_out][c_in_real][taps] fp32 (PyTorch OIHW order), assigned
    int c_in_real;           // real input channels (<= src_c[0] + src_c[1]; only the first conv pads 13 -> 16)
    int passes;              // 3: hi*hi + lo*hi + hi*lo (default) | 1: hi*hi only
};

int disco_wgrad_tc_launch(const disco_wgrad_desc* d, void* stream);
int disco_wgrad_ref_launch(const disco_wgrad_desc* d, void* stream);   // CUDA-core validator (tests only)
int disco_wgrad_splits(const disco_wgrad_desc* d);                      // recommended split-K factor

// PixelWeightedFusionSoftmax in training mode (per-pair batch statistics) + DiscoGraph fusion backward.
struct disco_pwf_train_desc {
    // collaboration-layer features (activation buffer, agent-major rows a*B + b) and the conv1_1 halves `en`
    const void* feat_hi;
    long long feat_lo_off;
    const float* en;            // [A*B, h, w, 2*hid] fp32: ego half (with conv bias) | neighbour half, NOT normalised
    int hid;                    // 128
    // raw (unfolded) parameters
    const float* g1; const float* be1;                       // bn1_1 gamma, beta [128]
    const float* w2; const float* b2; const float* g2; const float* be2;   // conv1_2 [32,128],[32]; bn1_2
    const float* w3; const float* b3; const float* g3; const float* be3;   // conv1_3 [8,32],[8]; bn1_3
    const float* w4; const float* b4;                         // conv1_4 [1,8],[1]
    float eps, momentum;
    // running statistics, updated sequentially in the reference's call order (b, ego i, k)
    float* rm1; float* rv1; float* rm2; float* rv2; float* rm3; float* rv3;
    long long* nbt1; long long* nbt2; long long* nbt3;
    // scene
    const double* trans;        // [B, A, A, 4, 4]
    const int* num_agent;       // [B]
    const int* outage;          // optional [B, A]
    int B, A, h, w, C;
    int only_v2i;
    float trans_scale;
    // forward products
    double* psum;               // [B*A*A][2][168]: per pair (b, ego i, neighbour id j) sum | sum of squares of the three
                                // pre-BatchNorm activations (written by the forward, read by the backward)
    float* wlogit;              // [B, A, A, h, w] post-ReLU PWF output w_k (unflipped frame), 0 for unused pairs
    // backward
    const float* dfused;        // [A*B, h, w, C] fp32 gradient wrt the fused map
    float* dwlogit;             // [B, A, A, h, w] gradient wrt wlogit (written by the combine backward)
    float* dfeat;               // [A*B, h, w, C] fp32 gradient wrt the features (accumulated with atomics; zeroed by caller)
    float* den;                 // [A*B, h, w, 2*hid] fp32 gradient wrt `en` (accumulated; zeroed by caller)
    float* gsum;                // [B*A*A][2][168] backward workspace: per pair sum g | sum g*xhat
    float* dparams;             // [4697] accumulated: dg1[128] dbe1[128] dw2[4096] dg2[32] dbe2[32] dw3[256] dg3[8] dbe3[8] dw4[8] db4[1]
};

int disco_pwf_train_forward_launch(const disco_pwf_train_desc* d, void* stream);
int disco_fusion_combine_backward_launch(const disco_pwf_train_desc* d, void* stream);
int disco_pwf_train_backward_launch(const disco_pwf_train_desc* d, void* stream);
