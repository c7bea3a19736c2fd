"""Host-side driver of the DiscoNet hot path: folds/packs parameters into conv plans, owns the NHWC
activation workspace and issues the kernel sequence through the C-ABI on torch's current stream.

Kernel sequence per forward (N = agents x scenes images, eval mode, BN folded):
  bev_pack -> 12 encoder convs -> [compress pair] -> PWF 1x1 (fp32) -> fusion -> 8 decoder convs
  (nearest-x2 upsample + skip concat fused into the gather) -> fused heads (3x3 32->64, 1x1 64->48 split)
i.e. 26 launches instead of the reference's ~700 ATen kernels per 5-agent scene.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import torch

from . import ops
from ._lib import PREC_BF16X3, PREC_FP16, FusionDesc
from .plan import ConvPlan, fold_bn, pack_chain, pack_conv, pack_conv_subpix, pack_conv_subpix_fused

# Output-parity ("sub-pixel") decomposition of the four upsample-concat convs (conv5_1 .. conv8_1): 4 launches, one per parity
# class, with 4 instead of 9 taps on the upsampled channels (-37 % MMAs on those layers).  Parity-green (tests/test_conv_gpu.py)
# but OFF by default: measured on B200 (round 2, B = 16, profiles/r02_layers_subpix_B16.csv) the per-class launches are slower
# than the single launch -- conv5_1 0.85 vs 0.56 ms, conv6_1 0.79 vs 0.52, conv7_1 1.05 vs 0.69, conv8_1 2.50 vs 1.08 -- because
# every class re-stages the whole full-resolution skip window (stride-2 parity planes: 4.8x the pixels it uses) and the small
# per-class grids quantise badly over 148 SMs.  The profitable form keeps the four class accumulators of one tile in TMEM and
# shares one staged window.  DISCO_B200_SUBPIX=1 enables it.
USE_SUBPIX = os.environ.get("DISCO_B200_SUBPIX", "0") == "1"

# Fused form (conv.h subpix == 2, conv_tc.cu MODE 4): ONE launch whose work items keep the four class accumulators of a low-res tile
# in TMEM and share every staged operand -- for the C_out <= 64 layers (conv7_1, conv8_1), where TMEM holds four stacked
# accumulators.  ON by default (measured round 2, B = 16: conv8_1 1.13 -> 0.69 ms, conv7_1 0.71 -> 0.53 ms);
# DISCO_B200_FUSED_SUBPIX=0 disables it, a comma list selects layers.
_fused_env = os.environ.get("DISCO_B200_FUSED_SUBPIX", "c7_1,c8_1")
FUSED_SUBPIX_LAYERS = set() if _fused_env in ("0", "none", "") else set(x for x in _fused_env.split(",") if x)

PRECISIONS = {"bf16x3": PREC_BF16X3, "fp16": PREC_FP16}

Getter = Callable[[str], torch.Tensor]


def _conv_bn(get: Getter, conv: str, bn: str):
    return fold_bn(get(conv + ".weight"), get(conv + ".bias"), get(bn + ".weight"), get(bn + ".bias"),
                   get(bn + ".running_mean"), get(bn + ".running_var"))


def _pad_in(w: torch.Tensor, c_in_pad: int) -> torch.Tensor:
    if w.shape[1] == c_in_pad:
        return w
    out = torch.zeros(w.shape[0], c_in_pad, *w.shape[2:], dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w
    return out


def build_encoder_plans(get: Getter, p: str, precision: int, compress: bool = False) -> Dict[str, ConvPlan]:
    """Backbone.encode (Backbone.py:89-143).  `p` = parameter prefix ('u_encoder.' / 'stpn.')."""
    P: Dict[str, ConvPlan] = {}

    def add(name, conv, bn, srcs, stride=1):
        w, b = _conv_bn(get, p + conv, p + bn)
        if w.dim() == 5:  # Conv3d 1x1x1 == pointwise conv at seq 1
            w = w.view(w.shape[0], w.shape[1], 1, 1)
        P[name] = pack_conv(_pad_in(w, sum(srcs)), b, src_channels=srcs, stride=stride, relu=True,
                            precision=precision, name=p + conv)

    add("pre1", "conv_pre_1", "bn_pre_1", [16])
    add("pre2", "conv_pre_2", "bn_pre_2", [32])
    add("c1_1", "conv1_1", "bn1_1", [32], 2)
    add("c1_2", "conv1_2", "bn1_2", [64])
    add("c3d_1", "conv3d_1.conv3d", "conv3d_1.bn3d", [64])
    add("c2_1", "conv2_1", "bn2_1", [64], 2)
    add("c2_2", "conv2_2", "bn2_2", [128])
    add("c3d_2", "conv3d_2.conv3d", "conv3d_2.bn3d", [128])
    add("c3_1", "conv3_1", "bn3_1", [128], 2)
    add("c3_2", "conv3_2", "bn3_2", [256])
    add("c4_1", "conv4_1", "bn4_1", [256], 2)
    add("c4_2", "conv4_2", "bn4_2", [512])
    if compress:
        # x_3 = relu(bn(1x1 256->cc)), relu(bn(1x1 cc->256))  (Backbone.py:139-141); cc padded to 16 lanes
        w, b = _conv_bn(get, p + "com_compresser", p + "bn_compress")
        cc = w.shape[0]
        cc_pad = (cc + 15) // 16 * 16
        wp = torch.zeros(cc_pad, 256, 1, 1, device=w.device)
        wp[:cc] = w
        bp = torch.zeros(cc_pad, device=w.device)
        bp[:cc] = b
        P["compress"] = pack_conv(wp, bp, src_channels=[256], relu=True, precision=precision, name=p + "com_compresser")
        w, b = _conv_bn(get, p + "com_decompresser", p + "bn_decompress")
        P["decompress"] = pack_conv(_pad_in(w, cc_pad), b, src_channels=[cc_pad], relu=True, precision=precision,
                                    name=p + "com_decompresser")
    return P


def build_decoder_plans(get: Getter, p: str, precision: int) -> Dict[str, ConvPlan]:
    """Backbone.decode (Backbone.py:145-242); concat order = (upsampled, skip) (:176,195,214,233)."""
    P: Dict[str, ConvPlan] = {}

    subpix = set(x for x in os.environ.get("DISCO_B200_SUBPIX_LAYERS", "c5_1,c6_1,c7_1,c8_1").split(",") if x)

    def add(name, conv, bn, srcs, c_blk=None):
        c_blk = int(os.environ.get("DISCO_CBLK_" + name.upper(), "0")) or c_blk     # tuning hook (K-stage width of one layer)
        w, b = _conv_bn(get, p + conv, p + bn)
        P[name] = pack_conv(w, b, src_channels=srcs, relu=True, precision=precision, name=p + conv, c_blk=c_blk)
        if name in FUSED_SUBPIX_LAYERS and len(srcs) == 2 and precision == PREC_BF16X3 and w.shape[0] <= 64:
            P[name + "/fused"] = pack_conv_subpix_fused(w, b, src_channels=srcs, relu=True, precision=precision, name=p + conv)
        elif USE_SUBPIX and len(srcs) == 2 and precision == PREC_BF16X3 and name in subpix:
            P[name + "/sub"] = [pack_conv_subpix(w, b, src_channels=srcs, py=py, px=px, relu=True, precision=precision, name=p + conv)
                                for py in (0, 1) for px in (0, 1)]

    add("c5_1", "conv5_1", "bn5_1", [512, 256]); add("c5_2", "conv5_2", "bn5_2", [256])
    add("c6_1", "conv6_1", "bn6_1", [256, 128]); add("c6_2", "conv6_2", "bn6_2", [128])
    add("c7_1", "conv7_1", "bn7_1", [128, 64]);  add("c7_2", "conv7_2", "bn7_2", [64])
    # conv8_1: 16-channel K stages -- its resident weights (110 KB) + the upsample scratch leave room for only two 32-channel
    # stages, i.e. no prefetch depth (measured round 2, B = 16: 1.32 ms -> 1.08 ms)
    add("c8_1", "conv8_1", "bn8_1", [64, 32], c_blk=16 if precision == PREC_BF16X3 else None);   add("c8_2", "conv8_2", "bn8_2", [32])
    return P


def build_head_plans(get: Getter, precision: int) -> Dict[str, ConvPlan]:
    """cls + reg heads share x_8 (DetModelBase.py:283-351): one 3x3 32->64 conv (cls1 | reg1) and one
    block-diagonal 1x1 64->(12+36) conv writing the two NHWC fp32 result tensors directly."""
    wc, bc = _conv_bn(get, "classification.conv1", "classification.bn1")
    wr, br = _conv_bn(get, "regression.box_prediction.0", "regression.box_prediction.1")
    ch = wc.shape[0]
    chained = precision == PREC_BF16X3
    # (chained: 16-channel K stages -> more, smaller A stages next to the resident weights + A2/W2 buffers)
    h3 = pack_conv(torch.cat((wc, wr), 0), torch.cat((bc, br), 0), src_channels=[wc.shape[1]], relu=True,
                   precision=precision, name="heads.3x3", c_blk=16 if chained else None)
    w2c, b2c = get("classification.conv2.weight").detach().float(), get("classification.conv2.bias").detach().float()
    w2r = get("regression.box_prediction.3.weight").detach().float()
    b2r = get("regression.box_prediction.3.bias").detach().float()
    nc, nr = w2c.shape[0], w2r.shape[0]
    w = torch.zeros(nc + nr, 2 * ch, 1, 1, device=wc.device)
    w[:nc, :ch] = w2c
    w[nc:, ch:] = w2r
    b2 = torch.cat((b2c, b2r), 0)
    if chained and h3.stacked:
        # the 1x1 runs as a chained MMA inside the 3x3 kernel: the 64-channel intermediate never touches HBM
        h3.chain = pack_chain(w.view(nc + nr, 2 * ch), b2, 2 * ch, relu=False)
        h3.name = "heads.3x3+1x1"
        return {"h3": h3, "h1": None, "n_cls": nc, "n_reg": nr}
    h1 = pack_conv(w, b2, src_channels=[2 * ch], relu=False, precision=precision, name="heads.1x1")
    return {"h3": h3, "h1": h1, "n_cls": nc, "n_reg": nr}


def build_pwf_plans(get: Getter, precision: int, p: str = "pixel_weighted_fusion."):
    """PixelWeightedFusionSoftmax (DiscoNet.py:132-155) with BN folded.

    conv1_1 over cat[ego, nb] is split into its ego and neighbour halves and evaluated once per agent
    map as a single 1x1 conv C -> 2*128 (fp32 out): [s*W_e x + s*(b-mean)+beta | s*W_n x].
    """
    w1, b1 = _conv_bn(get, p + "conv1_1", p + "bn1_1")         # [128, 2C, 1, 1]
    C = w1.shape[1] // 2
    w_en = torch.cat((w1[:, :C], w1[:, C:]), 0)                 # [256, C, 1, 1]
    b_en = torch.cat((b1, torch.zeros_like(b1)), 0)
    en = pack_conv(w_en, b_en, src_channels=[C], relu=False, precision=precision, name=p + "conv1_1")
    w2, b2 = _conv_bn(get, p + "conv1_2", p + "bn1_2")
    w3, b3 = _conv_bn(get, p + "conv1_3", p + "bn1_3")
    w4 = get(p + "conv1_4.weight").detach().float()
    b4 = get(p + "conv1_4.bias").detach().float()
    tail = [w2.reshape(32, 128).contiguous(), b2.contiguous(), w3.reshape(8, 32).contiguous(), b3.contiguous(),
            w4.reshape(1, 8).contiguous(), b4.contiguous()]
    return {"en": en, "tail": tail, "C": C}


class Workspace:
    """NHWC activation buffers + prebuilt launches for one (N, H, W) problem size."""

    def __init__(self, n: int, h: int, w: int, precision: int, device, enc: Dict[str, ConvPlan],
                 dec: Dict[str, ConvPlan], heads=None, pwf=None, batch_size: int = 1, agents: int = 1,
                 shard=None, fusion_level: int = 3):
        """`n` = image rows held by this process.  `shard = (row_begin, n_global)` for the agent-sharded
        mode: encoder/decoder/heads run on the local rows, the fusion reads the gathered `x3g` of all
        `n_global` rows and produces only the local ego rows."""
        if h % 16 or w % 16:
            raise ValueError(f"BEV size {h}x{w} must be a multiple of 16 (4 stride-2 stages)")
        self.n, self.h, self.w, self.precision, self.device = n, h, w, precision, device
        A = lambda hh, ww, c: ops.alloc_act(n, hh, ww, c, precision, device)
        h1, w1, h2, w2, h3, w3, h4, w4 = h // 2, w // 2, h // 4, w // 4, h // 8, w // 8, h // 16, w // 16
        b = self.buf = {
            "a0": A(h, w, 16), "t0": A(h, w, 32), "x": A(h, w, 32),
            "t1a": A(h1, w1, 64), "t1b": A(h1, w1, 64), "x1": A(h1, w1, 64),
            "t2a": A(h2, w2, 128), "t2b": A(h2, w2, 128), "x2": A(h2, w2, 128),
            "t3": A(h3, w3, 256), "x3": A(h3, w3, 256),
            "t4": A(h4, w4, 512), "x4": A(h4, w4, 512),
            "t5": A(h3, w3, 256), "x5": A(h3, w3, 256),
            "t6": A(h2, w2, 128), "x6": A(h2, w2, 128),
            "t7": A(h1, w1, 64), "x7": A(h1, w1, 64),
            "t8": A(h, w, 32), "x8": A(h, w, 32),
        }
        mk = lambda plan, srcs, ups, out, hh, ww: ops.ConvCall(plan, srcs, ups, out, n=n, h_in=hh, w_in=ww)
        # 1 if the packed input needs its lo plane (written by bev_pack / bev_scatter_batched; 0/1 occupancy never does)
        self.lo_nonzero = torch.ones(1, dtype=torch.int32, device=device)
        self.enc_calls: List[ops.ConvCall] = [
            ops.ConvCall(enc["pre1"], [b["a0"]], [0], b["t0"], n=n, h_in=h, w_in=w, lo_nonzero=self.lo_nonzero),
            mk(enc["pre2"], [b["t0"]], [0], b["x"], h, w),
            mk(enc["c1_1"], [b["x"]], [0], b["t1a"], h, w),
            mk(enc["c1_2"], [b["t1a"]], [0], b["t1b"], h1, w1),
            mk(enc["c3d_1"], [b["t1b"]], [0], b["x1"], h1, w1),
            mk(enc["c2_1"], [b["x1"]], [0], b["t2a"], h1, w1),
            mk(enc["c2_2"], [b["t2a"]], [0], b["t2b"], h2, w2),
            mk(enc["c3d_2"], [b["t2b"]], [0], b["x2"], h2, w2),
            mk(enc["c3_1"], [b["x2"]], [0], b["t3"], h2, w2),
            mk(enc["c3_2"], [b["t3"]], [0], b["x3"], h3, w3),
            mk(enc["c4_1"], [b["x3"]], [0], b["t4"], h3, w3),
            mk(enc["c4_2"], [b["t4"]], [0], b["x4"], h4, w4),
        ]
        self.x3_key = "x3"
        if "compress" in enc:
            cc_pad = enc["compress"].c_out
            b["x3c"] = A(h3, w3, cc_pad)
            b["x3d"] = A(h3, w3, 256)
            self.enc_calls += [mk(enc["compress"], [b["x3"]], [0], b["x3c"], h3, w3),
                               mk(enc["decompress"], [b["x3c"]], [0], b["x3d"], h3, w3)]
            self.x3_key = "x3d"
        # collaboration-layer fusion (DiscoNet only)
        self.fusion: Optional[FusionDesc] = None
        x3_dec = b[self.x3_key]
        self.shard = shard
        x2_dec = b["x2"]
        self.feat_key, self.fused_key = self.x3_key, "x3f"
        if pwf is not None:
            # collaboration level: 3 -> x_3 (256 ch @ H/8), 2 -> x_2 (128 ch @ H/4)   (DiscoNet.py:23-26)
            if fusion_level not in (2, 3):
                raise NotImplementedError("DiscoNet builds its PixelWeightedFusion for layer 2 or 3 only")
            row_begin, n_glob = (0, n) if shard is None else shard
            hf, wf, cf = (h3, w3, 256) if fusion_level == 3 else (h2, w2, 128)
            self.feat_key = self.x3_key if fusion_level == 3 else "x2"
            self.fused_key = "x3f" if fusion_level == 3 else "x2f"
            self.fuse_hw = (hf, wf)
            feat = b[self.feat_key]
            if shard is not None:
                feat = b["x3g"] = ops.alloc_act(n_glob, hf, wf, cf, precision, device)
            b["en"] = torch.empty((n_glob, hf, wf, 256), dtype=torch.float32, device=device)
            b[self.fused_key] = A(hf, wf, cf)
            self.en_call = ops.ConvCall(pwf["en"], [feat], [0], (b["en"],), n=n_glob, h_in=hf, w_in=wf)
            f = FusionDesc()
            f.feat_hi = feat.data_ptr()
            f.feat_lo_off = ops._lo_off(feat)
            f.precision = precision
            f.en = b["en"].data_ptr()
            f.hid = 128
            t = pwf["tail"]
            f.w2, f.b2, f.w3, f.b3, f.w4, f.b4 = (x.data_ptr() for x in t)
            f.B, f.A, f.h, f.w, f.C = batch_size, agents, hf, wf, cf
            f.trans_scale = 4.0 / 128.0
            f.out_hi = b[self.fused_key].data_ptr()
            f.out_lo_off = ops._lo_off(b[self.fused_key])
            f.row_begin, f.row_end = row_begin, row_begin + n
            self.fusion = f
            self._pwf_keep = pwf
            if fusion_level == 3:
                x3_dec = b["x3f"]
            else:
                x2_dec = b["x2f"]
        def up(name, srcs, out, hh, ww):
            """conv(cat(nearest_up2(srcs[0]), srcs[1])): one launch, or one per output-parity class (engine.USE_SUBPIX)."""
            if name + "/fused" in dec and hh % 2 == 0 and ww % 2 == 0:
                return [mk(dec[name + "/fused"], srcs, [1, 0], out, hh, ww)]
            if name + "/sub" in dec and hh % 2 == 0 and ww % 2 == 0:
                return [mk(pl, srcs, [1, 0], out, hh, ww) for pl in dec[name + "/sub"]]
            return [mk(dec[name], srcs, [1, 0], out, hh, ww)]

        self.dec_calls: List[ops.ConvCall] = (
            up("c5_1", [b["x4"], x3_dec], b["t5"], h3, w3) + [mk(dec["c5_2"], [b["t5"]], [0], b["x5"], h3, w3)] +
            up("c6_1", [b["x5"], x2_dec], b["t6"], h2, w2) + [mk(dec["c6_2"], [b["t6"]], [0], b["x6"], h2, w2)] +
            up("c7_1", [b["x6"], b["x1"]], b["t7"], h1, w1) + [mk(dec["c7_2"], [b["t7"]], [0], b["x7"], h1, w1)] +
            up("c8_1", [b["x7"], b["x"]], b["t8"], h, w) + [mk(dec["c8_2"], [b["t8"]], [0], b["x8"], h, w)])
        self.head_calls: List[ops.ConvCall] = []
        if heads is not None:
            self.n_cls, self.n_reg = heads["n_cls"], heads["n_reg"]
            dummy = (torch.empty(4, device=device), torch.empty(4, device=device))
            if heads["h1"] is None:   # fused 3x3 + chained 1x1
                self.head_calls = [ops.ConvCall(heads["h3"], [b["x8"]], [0], dummy, n=n, h_in=h, w_in=w,
                                                out_split=self.n_cls)]
            else:
                b["hh"] = A(h, w, heads["h3"].c_out)
                self.head_calls = [
                    mk(heads["h3"], [b["x8"]], [0], b["hh"], h, w),
                    ops.ConvCall(heads["h1"], [b["hh"]], [0], dummy, n=n, h_in=h, w_in=w, out_split=self.n_cls),
                ]

    def total_flops(self) -> int:
        calls = self.enc_calls + self.dec_calls + self.head_calls + ([self.en_call] if self.fusion else [])
        return sum(c.flops for c in calls)
