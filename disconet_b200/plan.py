"""Host-side weight preparation for the tcgen05 conv kernel: BN folding and UMMA operand packing.

Pure tensor bookkeeping (runs on whatever device the parameters live on); no compute of the hot path
happens here.  Layout contract with csrc/conv_tc.cu:

    wpack[n_tile][c_block][tap][part][c_blk/8][block_n][8]   16-bit, part = (hi, lo) for bf16x3

i.e. for every (channel block, filter tap) one contiguous shared-memory image of the B operand in the
UMMA no-swizzle K-major core-matrix layout, streamed by a single bulk copy.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from ._lib import PREC_BF16X3, PREC_FP16

BN_EPS = 1e-5


def fold_bn(weight, bias, bn_w, bn_b, bn_mean, bn_var, eps: float = BN_EPS):
    """conv -> BatchNorm(eval) == conv with w*s, (b-mean)*s+beta,  s = gamma/sqrt(var+eps)."""
    w = weight.detach().float()
    s = bn_w.detach().float() / torch.sqrt(bn_var.detach().float() + eps)
    wf = w * s.view(-1, *([1] * (w.dim() - 1)))
    b0 = bias.detach().float() if bias is not None else torch.zeros_like(s)
    bf = (b0 - bn_mean.detach().float()) * s + bn_b.detach().float()
    return wf, bf


def split_bf16(t: torch.Tensor):
    """fp32 -> (hi, lo) bf16 with hi = bf16(t), lo = bf16(t - hi)."""
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    return hi, lo


def choose_c_blk(src_channels, precision: int, stride: int) -> int:
    cap = 64
    if precision == PREC_BF16X3 or stride == 2:
        cap = 32
    if precision == PREC_BF16X3 and stride == 2:
        cap = 16   # the 33x17 stride-2 patch is 4.4x a tile: 16-channel stages (38 KB) keep >= 3-4 of them in flight
    for cb in (64, 32, 16):
        if cb <= cap and all(c % cb == 0 for c in src_channels if c):
            return cb
    raise ValueError(f"source channels {src_channels} must be multiples of 16")


@dataclass
class ConvPlan:
    """Packed weights + static geometry of one conv layer."""
    taps: int
    stride: int
    c_in: int                 # padded input channels (sum of sources)
    c_out: int
    c_blk: int
    block_n: int
    relu: bool
    precision: int
    wpack: torch.Tensor       # int16 storage
    bias: torch.Tensor        # fp32 [n_tiles*block_n]
    wref: Optional[torch.Tensor] = None   # fp32 [c_out, taps, c_in] (validator only)
    stacked: bool = False     # weight image layout [..][chunk][part][n][8] (see conv.h wpack_stacked)
    chain: Optional[dict] = None   # chained 1x1: {wpack, bias, c_out, relu} (see conv.h chain_*)
    name: str = ""
    flops_per_pixel: int = field(default=0)
    subpix: Optional[tuple] = None   # (py, px): this plan computes ONE output-parity class of an upsample-concat conv (conv.h subpix)
    fused_subpix: bool = False       # all four classes per work item (conv.h subpix == 2, pack_conv_subpix_fused)


def pack_conv(wf: torch.Tensor, bf: torch.Tensor, *, src_channels, stride: int = 1, relu: bool = True,
              precision: int = PREC_BF16X3, block_n: Optional[int] = None, keep_ref: bool = False,
              name: str = "", c_blk: Optional[int] = None) -> ConvPlan:
    """wf [c_out, c_in_real, k, k] fp32 (BN folded), bf [c_out] -> ConvPlan.

    `src_channels` are the (padded) channel counts of the concatenated NHWC sources in order; real
    input channels are laid out at the start of each source (only the first conv pads 13 -> 16).
    """
    c_out, c_in_real, k, _ = wf.shape
    taps = k * k
    assert taps in (1, 9)
    c_in = int(sum(src_channels))
    assert c_in >= c_in_real
    if c_blk is None:
        c_blk = choose_c_blk(src_channels, precision, stride)
    c_out_pad16 = (c_out + 15) // 16 * 16
    if block_n is None:
        # N=256 keeps the MMA's shared-memory operand reads (A 4 KB + B N*32 B per N/2 cycles) under the
        # 128 B/clk SMEM port; smaller layers use <=128 so that two 128-pixel sub-tiles share one B stream
        block_n = 256 if c_out_pad16 >= 256 else min(c_out_pad16, 128)
    n_tiles = (c_out + block_n - 1) // block_n
    n_rows = n_tiles * block_n
    dev = wf.device
    w = torch.zeros(n_rows, c_in, taps, dtype=torch.float32, device=dev)
    w[:c_out, :c_in_real] = wf.reshape(c_out, c_in_real, taps)
    bias = torch.zeros(n_rows, dtype=torch.float32, device=dev)
    bias[:c_out] = bf
    ncb = c_in // c_blk
    # [n_tile, n, cb, chunk, 8, tap] -> [n_tile, cb, tap, chunk, n, 8]
    w6 = w.view(n_tiles, block_n, ncb, c_blk // 8, 8, taps).permute(0, 2, 5, 3, 1, 4).contiguous()
    # N <= 64 layers are MMA-issue-bound (~47 cycles per tcgen05.mma whatever N): stacking W_hi and W_lo
    # as one 2N-row operand turns the three split passes into two instructions per k-step
    stacked = precision == PREC_BF16X3 and block_n <= 64
    if precision == PREC_BF16X3:
        hi, lo = split_bf16(w6)
        parts = torch.stack((hi, lo), dim=4 if stacked else 3)   # [.., tap, chunk, part, n, 8] | [.., tap, part, chunk, n, 8]
        wpack = parts.contiguous().view(torch.int16)
    else:
        wpack = w6.to(torch.float16).contiguous().view(torch.int16)
    wref = w[:c_out].permute(0, 2, 1).contiguous() if keep_ref else None   # [c_out, taps, c_in]
    return ConvPlan(taps=taps, stride=stride, c_in=c_in, c_out=c_out, c_blk=c_blk, block_n=block_n, relu=relu,
                    precision=precision, wpack=wpack.reshape(-1), bias=bias, wref=wref, name=name, stacked=stacked,
                    flops_per_pixel=2 * taps * c_in_real * c_out)


def pack_chain(w2: torch.Tensor, b2: torch.Tensor, k2: int, relu: bool = False) -> dict:
    """1x1 conv chained onto a <=64-channel layer inside the same kernel: w2 [n2, k2] fp32 -> stacked bf16x3
    image [k2/8][part][n2 padded to 16][8] (whole K resident, one bulk copy)."""
    n2 = w2.shape[0]
    bn = (n2 + 15) // 16 * 16
    w = torch.zeros(bn, k2, dtype=torch.float32, device=w2.device)
    w[:n2, :w2.shape[1]] = w2
    w3 = w.view(bn, k2 // 8, 8).permute(1, 0, 2).contiguous()       # [chunk, n, 8]
    hi, lo = split_bf16(w3)
    img = torch.stack((hi, lo), dim=1).contiguous().view(torch.int16).reshape(-1)   # [chunk, part, n, 8]
    bias = torch.zeros(bn, dtype=torch.float32, device=w2.device)
    bias[:n2] = b2
    return {"wpack": img, "bias": bias, "c_out": n2, "relu": bool(relu), "flops_per_pixel": 2 * k2 * n2}


# ---- output-parity ("sub-pixel") decomposition of conv(cat(nearest_up2(a), b)) ------------------------------------------
# For the output pixels of one parity class (oy % 2, ox % 2) = (py, px) the nine taps of the UPSAMPLED source touch only
# 2 x 2 distinct low-resolution pixels, so their weights can be pre-summed: 4 taps instead of 9 on those channels
# (Backbone.py:176,195,214,233 feed 2/3 of the input channels of conv5_1..conv8_1 through F.interpolate(scale_factor=2)).
# Tap kh of the 3x3 kernel reads up-res row oy + kh - 1 = low-res row a + {-1, 0, +1}[kh'] with (oy = 2a + py):
#   py = 0:  kh 0 -> kh' 0 ;  kh 1, 2 -> kh' 1          py = 1:  kh 0, 1 -> kh' 1 ;  kh 2 -> kh' 2      (same for columns)
_SUBPIX_MAP = {0: (0, 1, 1), 1: (1, 1, 2)}


def subpix_active_taps(py: int, px: int):
    """Tap positions (kh' * 3 + kw') of the low-res 3x3 window that class (py, px) uses, ascending."""
    rows, cols = sorted(set(_SUBPIX_MAP[py])), sorted(set(_SUBPIX_MAP[px]))
    return [r * 3 + c for r in rows for c in cols]


def pack_conv_subpix(wf: torch.Tensor, bf: torch.Tensor, *, src_channels, py: int, px: int, relu: bool = True,
                     precision: int = PREC_BF16X3, name: str = "") -> ConvPlan:
    """Weights of output-parity class (py, px) of a 3x3 conv over cat(nearest_up2(src0), src1): source-0 taps pre-summed into
    the 2 x 2 low-res taps the class touches (packed as 4 blocks per channel block), source-1 taps unchanged (9 blocks).
    Block order = [n_tile][channel block][active tap]; 16-channel K stages (the stride-2 parity planes of source 1 are large)."""
    assert precision == PREC_BF16X3 and wf.shape[2:] == (3, 3) and len(src_channels) == 2
    c_out, c_in_real = wf.shape[0], wf.shape[1]
    c0, c1 = int(src_channels[0]), int(src_channels[1])
    assert c0 + c1 == c_in_real and c0 % 16 == 0 and c1 % 16 == 0
    c_blk = 16
    c_out_pad16 = (c_out + 15) // 16 * 16
    block_n = 256 if c_out_pad16 >= 256 else min(c_out_pad16, 128)
    n_tiles = (c_out + block_n - 1) // block_n
    n_rows = n_tiles * block_n
    dev = wf.device
    w = torch.zeros(n_rows, c_in_real, 3, 3, dtype=torch.float32, device=dev)
    w[:c_out] = wf.detach().float()
    w0 = torch.zeros(n_rows, c0, 3, 3, dtype=torch.float32, device=dev)
    for kh in range(3):
        for kw in range(3):
            w0[:, :, _SUBPIX_MAP[py][kh], _SUBPIX_MAP[px][kw]] += w[:, :c0, kh, kw]
    w = torch.cat((w0, w[:, c0:]), 1).reshape(n_rows, c_in_real, 9)
    bias = torch.zeros(n_rows, dtype=torch.float32, device=dev)
    bias[:c_out] = bf
    ncb = c_in_real // c_blk
    w6 = w.view(n_tiles, block_n, ncb, c_blk // 8, 8, 9).permute(0, 2, 5, 3, 1, 4).contiguous()   # [n_tile, cb, tap, chunk, n, 8]
    stacked = block_n <= 64
    hi, lo = split_bf16(w6)
    parts = torch.stack((hi, lo), dim=4 if stacked else 3).contiguous()      # [.., tap, chunk, part, n, 8] | [.., tap, part, chunk, n, 8]
    act = subpix_active_taps(py, px)
    blocks = []
    for t in range(n_tiles):
        for cb in range(ncb):
            taps = act if cb < c0 // c_blk else range(9)
            blocks += [parts[t, cb, tap].reshape(-1) for tap in taps]
    wpack = torch.cat(blocks).view(torch.int16)
    return ConvPlan(taps=9, stride=1, c_in=c_in_real, c_out=c_out, c_blk=c_blk, block_n=block_n, relu=relu, precision=precision,
                    wpack=wpack.reshape(-1), bias=bias, wref=None, name=name + f"[py{py}px{px}]", stacked=stacked,
                    flops_per_pixel=2 * 9 * c_in_real * c_out, subpix=(py, px))


def pack_conv_subpix_fused(wf: torch.Tensor, bf: torch.Tensor, *, src_channels, relu: bool = True,
                           precision: int = PREC_BF16X3, name: str = "") -> ConvPlan:
    """Weights of the FUSED sub-pixel form (conv.h subpix == 2; C_out <= 64): one work item computes the four output-parity
    classes of a 16 x 8 low-res tile.  The image is a stream of slots of nine blocks (block = [chunk 2][hi rows | lo rows][8 ch],
    64 * block_n bytes), in the order the MMA issuers consume them:
      source-0 channel block:  two slots, one per low-res tap row ty: [py][chunk][(px, tx) = (0,0) (0,1) (1,0) (1,1)] + one zero unit
                               (class (py, px), tap (ty, tx) multiplies the low-res pixel at window row py + ty, column px + tx)
      source-1 channel block:  one slot [kh][chunk][kw = 2, 1, 0] (shared by the four classes)
    A unit is the [W_hi | W_lo] rows (2 * block_n x 8 channels) of one tap; inside a group the units are chunk-major, so two
    neighbouring units are ONE UMMA B operand of 4 * block_n rows: the kernel's px-merged MMAs (conv_tc.cu MODE 4)."""
    assert precision == PREC_BF16X3 and wf.shape[2:] == (3, 3) and len(src_channels) == 2
    c_out, c_in_real = wf.shape[0], wf.shape[1]
    c0, c1 = int(src_channels[0]), int(src_channels[1])
    assert c0 + c1 == c_in_real and c0 % 16 == 0 and c1 % 16 == 0 and c0 > 0 and c1 > 0
    block_n = (c_out + 15) // 16 * 16
    assert block_n <= 64, "fused sub-pixel form: C_out <= 64 (four stacked class accumulators must fit TMEM)"
    dev = wf.device
    w = torch.zeros(block_n, c_in_real, 3, 3, dtype=torch.float32, device=dev)
    w[:c_out] = wf.detach().float()
    bias = torch.zeros(block_n, dtype=torch.float32, device=dev)
    bias[:c_out] = bf

    def blocks_of(wt):   # [block_n, C, taps...] -> stacked bf16 blocks [C/16, taps..., chunk, part, n, 8]
        C = wt.shape[1]
        t = wt.reshape(block_n, C // 16, 2, 8, -1).permute(1, 4, 2, 0, 3).contiguous()    # [cb, tap, chunk, n, 8]
        hi, lo = split_bf16(t)
        return torch.stack((hi, lo), dim=3).contiguous()                                   # [cb, tap, chunk, part, n, 8]

    cls_w = []
    for py in (0, 1):
        for px in (0, 1):
            w0 = torch.zeros(block_n, c0, 3, 3, dtype=torch.float32, device=dev)
            for kh in range(3):
                for kw in range(3):
                    w0[:, :, _SUBPIX_MAP[py][kh], _SUBPIX_MAP[px][kw]] += w[:, :c0, kh, kw]
            cls_w.append(blocks_of(w0[:, :, py:py + 2, px:px + 2].contiguous()))            # [cb, 4, ...]
    s0 = torch.stack(cls_w, dim=1)                                                          # [cb0, class, 4 taps (ty, tx), chunk, part, n, 8]
    cb0 = s0.shape[0]
    s0 = s0.view(cb0, 2, 2, 2, 2, 2, 2, block_n, 8)                                         # [cb0, py, px, ty, tx, chunk, part, n, 8]
    s0 = s0.permute(0, 3, 1, 5, 2, 4, 6, 7, 8).contiguous()                                 # [cb0, ty, py, chunk, px, tx, part, n, 8]
    s0 = s0.view(cb0, 2, -1)
    s0 = torch.cat((s0, torch.zeros(cb0, 2, 2 * 2 * block_n * 8, dtype=s0.dtype, device=dev)), dim=2)   # + one pad unit per slot
    s1 = blocks_of(w[:, c0:].contiguous())                                                  # [cb1, 9 taps, chunk, part, n, 8]
    cb1 = s1.shape[0]
    s1 = s1.view(cb1, 3, 3, 2, 2, block_n, 8).flip(2).permute(0, 1, 3, 2, 4, 5, 6).contiguous()   # [cb1, kh, chunk, kw = 2,1,0, part, n, 8]
    wpack = torch.cat((s0.reshape(-1), s1.reshape(-1))).view(torch.int16)
    return ConvPlan(taps=9, stride=1, c_in=c_in_real, c_out=c_out, c_blk=16, block_n=block_n, relu=relu, precision=precision,
                    wpack=wpack.reshape(-1), bias=bias, wref=None, name=name + "[subpix x4]", stacked=True,
                    flops_per_pixel=2 * 9 * c_in_real * c_out, fused_subpix=True)
