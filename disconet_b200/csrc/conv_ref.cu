// CUDA-core validator for conv_tc.cu: same descriptor, same activations, *unpacked* fp32 weights,
// straightforward direct convolution with fp32 FMA.  Used by the GPU tests to cross-check the tensor
// core kernel at sizes where a CPU oracle would take minutes.  Never on the product path.
#include "common.cuh"
#include "conv.h"

namespace {

__device__ __forceinline__ float load_act(const uint16_t* p, long long lo_off, int precision) {
    if (precision == DISCO_PREC_BF16X3) return bf16_bits_to_f32(p[0]) + bf16_bits_to_f32(p[lo_off]);
    return f16_bits_to_f32(p[0]);
}

__global__ void conv_ref_kernel(const disco_conv_desc d, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int co = (int)(idx % d.c_out);
    const long long pixel = idx / d.c_out;
    const int ow = (int)(pixel % d.w_out);
    const int oh = (int)((pixel / d.w_out) % d.h_out);
    const int img = (int)(pixel / ((long long)d.w_out * d.h_out));
    const int c_in = d.src_c[0] + d.src_c[1];
    const int k = (d.taps == 9) ? 3 : 1;
    const int pad = (d.taps == 9) ? 1 : 0;
    float acc = 0.f;
    for (int kh = 0; kh < k; ++kh) {
        for (int kw = 0; kw < k; ++kw) {
            const int hi = oh * d.stride - pad + kh, wi = ow * d.stride - pad + kw;
            if (hi < 0 || hi >= d.h_in || wi < 0 || wi >= d.w_in) continue;
            const float* w = d.wref + ((long long)co * d.taps + (kh * k + kw)) * c_in;
            int cbase = 0;
            for (int s = 0; s < 2; ++s) {
                const int Cs = d.src_c[s];
                if (!Cs) continue;
                const int up = d.src_up[s] ? 1 : 0;
                if (d.src_up[s] == 2 && ((hi | wi) & 1)) continue;   // zero-stuffed source: odd rows/cols are zero
                const int Hs = d.h_in >> up, Ws = d.w_in >> up;
                const uint16_t* p = reinterpret_cast<const uint16_t*>(d.src[s]) +
                                    (((long long)img * Hs + (hi >> up)) * Ws + (wi >> up)) * Cs;
                for (int c = 0; c < Cs; ++c) acc = fmaf(load_act(p + c, d.src_lo_off[s], d.precision), w[cbase + c], acc);
                cbase += Cs;
            }
        }
    }
    acc += d.bias[co];
    if (d.relu) acc = fmaxf(acc, 0.f);
    if (d.out_mode == DISCO_OUT_ACT) {
        uint16_t* o = reinterpret_cast<uint16_t*>(d.out[0]) + pixel * d.c_out + co;
        if (d.precision == DISCO_PREC_BF16X3) {
            uint16_t h, l;
            split_bf16(acc, h, l);
            o[0] = h;
            o[d.out_lo_off] = l;
        } else {
            o[0] = f32_to_f16_bits(acc);
        }
    } else {
        if (co < d.out_split) reinterpret_cast<float*>(d.out[0])[pixel * d.out_split + co] = acc;
        else reinterpret_cast<float*>(d.out[1])[pixel * (d.c_out - d.out_split) + (co - d.out_split)] = acc;
    }
}

}  // namespace

int disco_conv_ref_launch(const disco_conv_desc* d, void* stream) {
    DISCO_REQUIRE(d->wref != nullptr, "conv_ref: wref (unpacked fp32 weights) is required");
    const long long total = (long long)d->n * d->h_out * d->w_out * d->c_out;
    const int threads = 128;
    const long long blocks = (total + threads - 1) / threads;
    DISCO_REQUIRE(blocks > 0 && blocks < (1ll << 31), "conv_ref: bad size");
    conv_ref_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(*d, total);
    DISCO_CHECK_CUDA(cudaGetLastError());
    return DISCO_OK;
}
