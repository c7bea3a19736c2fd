"""Thin launch wrappers: torch tensors (device memory + stream plumbing) -> C-ABI descriptors."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import OUT_ACT, OUT_F32, PREC_BF16X3, PREC_FP16, ConvDesc, FusionDesc, check, load
from .plan import ConvPlan


def act_dtype(precision: int):
    return torch.bfloat16 if precision == PREC_BF16X3 else torch.float16


def act_parts(precision: int) -> int:
    return 2 if precision == PREC_BF16X3 else 1


def alloc_act(n, h, w, c, precision, device) -> torch.Tensor:
    """Activation buffer [parts, n, h, w, c]: part 0 = hi (or the single fp16 tensor), part 1 = lo."""
    return torch.empty((act_parts(precision), n, h, w, c), dtype=act_dtype(precision), device=device)


def _lo_off(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] == 2 else 0


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise ValueError("disconet_b200 kernels take CUDA tensors only (no CPU fallback path)")


def conv_forward(plan: ConvPlan, srcs: Sequence[torch.Tensor], ups: Sequence[int], out, *, n, h_in, w_in,
                 out_split: Optional[int] = None, reference: bool = False):
    """Run one conv layer.  srcs: activation buffers [parts,n,hs,ws,c]; out: activation buffer (OUT_ACT)
    or a tuple of fp32 NHWC tensors (OUT_F32)."""
    call = ConvCall(plan, srcs, ups, out, n=n, h_in=h_in, w_in=w_in, out_split=out_split)
    call.launch(_stream_ptr(srcs[0].device), reference=reference)


def bev_pack(bev: torch.Tensor, out: torch.Tensor, precision: int, lo_nonzero: Optional[torch.Tensor] = None):
    """bev fp32 [..., Z] contiguous -> out activation buffer [parts, n, h, w, 16].  `lo_nonzero` (int32 [1], optional)
    receives 1 if any value needed a lo part (0/1 occupancy never does: the first conv then skips the lo plane)."""
    _require_cuda(bev, out)
    assert bev.dtype == torch.float32 and bev.is_contiguous()
    z = bev.shape[-1]
    n_pix = bev.numel() // z
    assert out.shape[-1] == 16 and out[0].numel() == n_pix * 16
    check(load().disco_bev_pack(bev.data_ptr(), n_pix, z, out.data_ptr(), _lo_off(out), precision,
                                lo_nonzero.data_ptr() if lo_nonzero is not None else None, _stream_ptr(bev.device)), "bev_pack")


def act_to_nchw_f32(act: torch.Tensor, precision: int) -> torch.Tensor:
    """activation buffer [parts,n,h,w,c] -> fp32 [n,c,h,w] (contiguous)."""
    _require_cuda(act)
    _, n, h, w, c = act.shape
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=act.device)
    check(load().disco_act_unpack_nchw(act.data_ptr(), _lo_off(act), precision, n, h, w, c, out.data_ptr(),
                                       _stream_ptr(act.device)), "act_unpack")
    return out


class ConvCall:
    """A prebuilt conv launch (descriptor filled once; only output pointers may be re-pointed)."""

    def __init__(self, plan: ConvPlan, srcs, ups, out, *, n, h_in, w_in, out_split=None, lo_nonzero: Optional[torch.Tensor] = None):
        self.plan = plan
        self.keep = (plan, list(srcs), out, lo_nonzero)   # keep tensors alive as long as the descriptor exists
        d = ConvDesc()
        _require_cuda(*srcs)
        for i in range(2):
            if i < len(srcs):
                s = srcs[i]
                d.src[i] = s.data_ptr()
                d.src_lo_off[i] = _lo_off(s)
                d.src_c[i] = s.shape[-1]
                d.src_up[i] = int(ups[i])
        if sum(s.shape[-1] for s in srcs) != plan.c_in:
            raise ValueError(f"conv[{plan.name}]: sources {[tuple(s.shape) for s in srcs]} != c_in {plan.c_in}")
        d.n, d.h_in, d.w_in = n, h_in, w_in
        d.stride, d.taps, d.c_blk = plan.stride, plan.taps, plan.c_blk
        d.h_out = (h_in - 1) // plan.stride + 1
        d.w_out = (w_in - 1) // plan.stride + 1
        d.c_out, d.block_n = plan.c_out, plan.block_n
        d.wpack = plan.wpack.data_ptr()
        d.wpack_stacked = int(plan.stacked)
        d.wref = plan.wref.data_ptr() if plan.wref is not None else None
        d.bias = plan.bias.data_ptr()
        d.relu = int(plan.relu)
        d.precision = plan.precision
        fpp = plan.flops_per_pixel
        if plan.chain is not None:
            d.chain_wpack = plan.chain["wpack"].data_ptr()
            d.chain_bias = plan.chain["bias"].data_ptr()
            d.chain_c_out = plan.chain["c_out"]
            d.chain_relu = int(plan.chain["relu"])
            fpp += plan.chain["flops_per_pixel"]
        d.src_lo_nonzero = lo_nonzero.data_ptr() if lo_nonzero is not None else None
        if plan.subpix is not None:       # one output-parity class of an upsample-concat conv (a quarter of the output pixels)
            if list(ups) != [1, 0] or len(srcs) != 2:
                raise ValueError(f"conv[{plan.name}]: a sub-pixel class plan needs (upsampled, plain) sources")
            d.subpix, d.sub_py, d.sub_px = 1, int(plan.subpix[0]), int(plan.subpix[1])
        if plan.fused_subpix:             # the four classes of every low-res tile in one work item
            if list(ups) != [1, 0] or len(srcs) != 2:
                raise ValueError(f"conv[{plan.name}]: a fused sub-pixel plan needs (upsampled, plain) sources")
            d.subpix = 2
        self.desc = d
        self.flops = fpp * n * d.h_out * d.w_out // (4 if plan.subpix is not None else 1)
        # multiply-accumulates the tensor cores execute per split pass (sub-pixel forms run 4 instead of 9 taps on the upsampled source)
        self.mma_flops = self.flops
        if plan.fused_subpix or plan.subpix is not None:
            c0, c1 = srcs[0].shape[-1], srcs[1].shape[-1]
            self.mma_flops = int(self.flops * (4.0 / 9.0 * c0 + c1) / (c0 + c1))
        self.set_output(out, out_split)
        self._fn = load().disco_conv_forward
        self._ref = load().disco_conv_reference

    def set_output(self, out, out_split=None):
        d = self.desc
        if isinstance(out, torch.Tensor) and out.dtype in (torch.bfloat16, torch.float16):
            _require_cuda(out)
            if out.shape[-1] != self.plan.c_out:
                raise ValueError(f"conv[{self.plan.name}]: output channels {out.shape[-1]} != {self.plan.c_out}")
            d.out_mode = OUT_ACT
            d.out[0] = out.data_ptr()
            d.out[1] = None
            d.out_lo_off = _lo_off(out)
            d.out_split = self.plan.c_out
        else:
            outs = out if isinstance(out, (tuple, list)) else (out,)
            _require_cuda(*outs)
            d.out_mode = OUT_F32
            d.out[0] = outs[0].data_ptr()
            d.out[1] = outs[1].data_ptr() if len(outs) > 1 else None
            d.out_lo_off = 0
            oc = self.plan.chain["c_out"] if self.plan.chain is not None else self.plan.c_out
            d.out_split = oc if out_split is None else out_split

    def launch(self, stream_ptr: int, reference: bool = False):
        check((self._ref if reference else self._fn)(C.byref(self.desc), stream_ptr), f"conv[{self.plan.name}]")


def fusion_forward(desc: FusionDesc, stream_ptr: int):
    check(load().disco_fusion_forward(C.byref(desc), stream_ptr), "fusion")
