"""Test helpers: activation-buffer conversions and error metrics."""
import torch

from disconet_b200._lib import PREC_BF16X3
from disconet_b200.ops import act_dtype, act_parts


def to_act(x_nchw: torch.Tensor, precision: int, c_pad: int = None) -> torch.Tensor:
    """fp32 NCHW -> activation buffer [parts, n, h, w, c] (channels zero-padded to c_pad)."""
    x = x_nchw.permute(0, 2, 3, 1).contiguous()
    if c_pad is not None and c_pad > x.shape[-1]:
        x = torch.nn.functional.pad(x, (0, c_pad - x.shape[-1]))
    if precision == PREC_BF16X3:
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        return torch.stack((hi, lo), 0).contiguous()
    return x.to(torch.float16).unsqueeze(0).contiguous()


def act_value(act: torch.Tensor) -> torch.Tensor:
    """activation buffer -> fp32 NHWC value (hi + lo)."""
    v = act[0].float()
    if act.shape[0] == 2:
        v = v + act[1].float()
    return v


def rel_max(a: torch.Tensor, ref: torch.Tensor) -> float:
    """max|a-ref| / max|ref|  -- the parity metric of SURVEY.md §8(d)."""
    return ((a.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30)).item()


def rel_l2(a: torch.Tensor, ref: torch.Tensor) -> float:
    return ((a.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30)).item()
