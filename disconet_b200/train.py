"""Training-mode driver of the DiscoNet hot path: batch-statistics BatchNorm forward and the
full backward, as the kernel sequence behind one torch.autograd.Function.

Reference behaviour reproduced (R = coperception/):
  * model.train() forward: every conv -> BatchNorm(batch stats, running-stat update) -> ReLU of Backbone.encode /
    decode (R/models/det/backbone/Backbone.py:89-242), the heads (R/models/det/base/DetModelBase.py:268-351) and the
    per-pair PixelWeightedFusionSoftmax calls (R/models/det/DiscoNet.py:86-95,148-155);
  * loss.backward() (R/utils/CoDetModule.py:289-291): gradients of every live parameter, flowing in from
    result["cls"], result["loc"] and -- with kd_flag -- x_8, x_7, x_6, x_5 and the fused map (CoDetModule.py:340-382).

Kernel sequence of one step (N images):
  forward : bev_pack, per layer [conv_tc (raw weights, fp32 out) -> bn stats/finalize/apply], PWF 1x1 -> pwf_train_fwd
            -> fusion (precomputed maps) , heads 1x1 -> cls / loc
  backward: per layer in reverse [bn reduce/apply -> wgrad (tcgen05, split-K) -> data gradient = conv_tc with the
            transposed + flipped weights (zero-stuffed source for the stride-2 layers)], fusion combine backward,
            pwf_train_bwd, PWF 1x1 data/weight gradient.
Weights are re-packed into the UMMA operand images every step (they change with every optimizer step).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import (PREC_BF16X3, BnDesc, FusionDesc, PackDesc, PwfTrainDesc, WgradDesc, check, load)
from .plan import ConvPlan, pack_conv

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


@dataclass
class Layer:
    name: str
    conv: str                    # parameter prefix of the conv ("u_encoder.conv1_1")
    bn: Optional[str]            # parameter prefix of its BatchNorm (None: no BN / ReLU)
    srcs: List[str]              # activation buffers read (channel-concatenated)
    ups: List[int]               # 1: source is nearest-upsampled x2
    out: str                     # activation buffer written ("" for the fp32 head outputs)
    c_in: List[int]              # padded channels per source
    c_out: int
    stride: int = 1
    taps: int = 9
    level: int = 0               # resolution level of the conv INPUT (0: H, 1: H/2, ...)
    c_in_real: int = 0
    need_dgrad: bool = True

    def __post_init__(self):
        if not self.c_in_real:
            self.c_in_real = sum(self.c_in)


def backbone_layers(pe: str, pd: str, fused_key: Optional[str], fusion_level: int = 3) -> Tuple[List[Layer], List[Layer]]:
    """Backbone.encode / decode layer tables (Backbone.py:89-143,145-242)."""
    E = [
        Layer("pre1", pe + "conv_pre_1", pe + "bn_pre_1", ["a0"], [0], "t0", [16], 32, level=0, c_in_real=13, need_dgrad=False),
        Layer("pre2", pe + "conv_pre_2", pe + "bn_pre_2", ["t0"], [0], "x", [32], 32, level=0),
        Layer("c1_1", pe + "conv1_1", pe + "bn1_1", ["x"], [0], "t1a", [32], 64, stride=2, level=0),
        Layer("c1_2", pe + "conv1_2", pe + "bn1_2", ["t1a"], [0], "t1b", [64], 64, level=1),
        Layer("c3d_1", pe + "conv3d_1.conv3d", pe + "conv3d_1.bn3d", ["t1b"], [0], "x1", [64], 64, taps=1, level=1),
        Layer("c2_1", pe + "conv2_1", pe + "bn2_1", ["x1"], [0], "t2a", [64], 128, stride=2, level=1),
        Layer("c2_2", pe + "conv2_2", pe + "bn2_2", ["t2a"], [0], "t2b", [128], 128, level=2),
        Layer("c3d_2", pe + "conv3d_2.conv3d", pe + "conv3d_2.bn3d", ["t2b"], [0], "x2", [128], 128, taps=1, level=2),
        Layer("c3_1", pe + "conv3_1", pe + "bn3_1", ["x2"], [0], "t3", [128], 256, stride=2, level=2),
        Layer("c3_2", pe + "conv3_2", pe + "bn3_2", ["t3"], [0], "x3", [256], 256, level=3),
        Layer("c4_1", pe + "conv4_1", pe + "bn4_1", ["x3"], [0], "t4", [256], 512, stride=2, level=3),
        Layer("c4_2", pe + "conv4_2", pe + "bn4_2", ["t4"], [0], "x4", [512], 512, level=4),
    ]
    x3d = fused_key if (fused_key and fusion_level == 3) else "x3"
    x2d = fused_key if (fused_key and fusion_level == 2) else "x2"
    D = [
        Layer("c5_1", pd + "conv5_1", pd + "bn5_1", ["x4", x3d], [1, 0], "t5", [512, 256], 256, level=3),
        Layer("c5_2", pd + "conv5_2", pd + "bn5_2", ["t5"], [0], "x5", [256], 256, level=3),
        Layer("c6_1", pd + "conv6_1", pd + "bn6_1", ["x5", x2d], [1, 0], "t6", [256, 128], 128, level=2),
        Layer("c6_2", pd + "conv6_2", pd + "bn6_2", ["t6"], [0], "x6", [128], 128, level=2),
        Layer("c7_1", pd + "conv7_1", pd + "bn7_1", ["x6", "x1"], [1, 0], "t7", [128, 64], 64, level=1),
        Layer("c7_2", pd + "conv7_2", pd + "bn7_2", ["t7"], [0], "x7", [64], 64, level=1),
        Layer("c8_1", pd + "conv8_1", pd + "bn8_1", ["x7", "x"], [1, 0], "t8", [64, 32], 32, level=0),
        Layer("c8_2", pd + "conv8_2", pd + "bn8_2", ["t8"], [0], "x8", [32], 32, level=0),
    ]
    return E, D


def _f32(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _w4(w: torch.Tensor) -> torch.Tensor:
    """conv weight as [co, ci, k, k] (Conv3d 1x1x1 weights are [co, ci, 1, 1, 1])."""
    return w.view(w.shape[0], w.shape[1], 1, 1) if w.dim() == 5 else w


class _LayerState:
    """Per-layer device buffers + prebuilt launches."""

    def __init__(self):
        self.plan: Optional[ConvPlan] = None
        self.fwd: Optional[ops.ConvCall] = None
        self.z: Optional[torch.Tensor] = None
        self.stats: Optional[torch.Tensor] = None
        self.dz: Optional[torch.Tensor] = None
        self.bn: Optional[BnDesc] = None
        self.dplans: List[ConvPlan] = []
        self.dcalls: List[ops.ConvCall] = []
        self.gbufs: List[torch.Tensor] = []
        self.wg: Optional[WgradDesc] = None


class TrainRunner:
    """Training-mode workspace of one model for one (N, H, W, B) problem size."""

    def __init__(self, get, n: int, h: int, w: int, device, enc_prefix: str, dec_prefix: str, *, heads: bool = True,
                 pwf_prefix: Optional[str] = None, batch_size: int = 1, agents: int = 1, fusion_level: int = 3,
                 only_v2i: bool = False):
        if h % 16 or w % 16:
            raise ValueError(f"BEV size {h}x{w} must be a multiple of 16")
        self.get, self.n, self.h, self.w, self.dev = get, n, h, w, device
        self.B, self.A, self.fusion_level, self.only_v2i = batch_size, agents, fusion_level, only_v2i
        self.prec = PREC_BF16X3
        self.lib = load()
        self.pwf_prefix = pwf_prefix
        self.fused_key = None
        if pwf_prefix is not None:
            if fusion_level not in (2, 3):
                raise NotImplementedError("DiscoNet builds its PixelWeightedFusion for layer 2 or 3 only")
            self.fused_key = "x3f" if fusion_level == 3 else "x2f"
        E, D = backbone_layers(enc_prefix, dec_prefix, self.fused_key, fusion_level)
        self.enc, self.dec = E, D
        self.head_layers: List[Layer] = []
        if heads:
            self.head_layers = [
                Layer("h3", "heads.3x3", "heads.bn", ["x8"], [0], "hh", [32], 64, level=0),
                Layer("h1", "heads.1x1", None, ["hh"], [0], "", [64], 48, taps=1, level=0),
            ]
        self.layers = E + D + self.head_layers
        self.res = [(h >> k, w >> k) for k in range(5)]
        self.act: Dict[str, torch.Tensor] = {}
        self.st: Dict[str, _LayerState] = {}
        A_ = lambda hh, ww, c: ops.alloc_act(n, hh, ww, c, self.prec, device)
        self.act["a0"] = A_(h, w, 16)
        self.sums = torch.zeros(1024, dtype=torch.float64, device=device)
        max_partial = 0
        for L in self.layers:
            hi, wi = self.res[L.level]
            ho, wo = (hi - 1) // L.stride + 1, (wi - 1) // L.stride + 1
            S = self.st[L.name] = _LayerState()
            S.z = torch.empty((n, ho, wo, L.c_out), dtype=torch.float32, device=device) if L.bn else None
            if L.out:
                self.act[L.out] = A_(ho, wo, L.c_out)
            S.dz = A_(ho, wo, L.c_out)
            S.hw_in, S.hw_out = (hi, wi), (ho, wo)
            if L.bn:
                S.stats = torch.zeros(2 * L.c_out, dtype=torch.float32, device=device)
        if self.fused_key:
            hf, wf = self.res[fusion_level]
            cf = 256 if fusion_level == 3 else 128
            self.feat_key = "x3" if fusion_level == 3 else "x2"
            self.fuse_hw, self.fuse_c = (hf, wf), cf
            self.act[self.fused_key] = A_(hf, wf, cf)
        self._build_static()
        self.partial = None
        self._max_partial = max_partial

    # ------------------------------------------------------------------------------------------------
    def _conv_params(self, L: Layer):
        """(weight [co, ci_pad, k, k], bias [co]) of a layer from the current parameters."""
        g = self.get
        if L.name == "h3":
            w = torch.cat((_f32(g("classification.conv1.weight")), _f32(g("regression.box_prediction.0.weight"))), 0)
            b = torch.cat((_f32(g("classification.conv1.bias")), _f32(g("regression.box_prediction.0.bias"))), 0)
        elif L.name == "h1":
            wc, wr = _f32(g("classification.conv2.weight")), _f32(g("regression.box_prediction.3.weight"))
            nc, nr, ch = wc.shape[0], wr.shape[0], wc.shape[1]
            w = torch.zeros(nc + nr, 2 * ch, 1, 1, device=wc.device)
            w[:nc, :ch] = wc
            w[nc:, ch:] = wr
            b = torch.cat((_f32(g("classification.conv2.bias")), _f32(g("regression.box_prediction.3.bias"))), 0)
        else:
            w, b = _w4(_f32(g(L.conv + ".weight"))), _f32(g(L.conv + ".bias"))
        c_pad = sum(L.c_in)
        if w.shape[1] != c_pad:
            wp = torch.zeros(w.shape[0], c_pad, w.shape[2], w.shape[3], device=w.device)
            wp[:, :w.shape[1]] = w
            w = wp
        return w, b

    def _bn_params(self, L: Layer):
        g = self.get
        if L.name == "h3":
            names = ("classification.bn1", "regression.box_prediction.1")
            cat = lambda f: torch.cat([_f32(g(nm + "." + f)) for nm in names], 0)
            return cat("weight"), cat("bias"), cat("running_mean"), cat("running_var")
        return (_f32(g(L.bn + ".weight")), _f32(g(L.bn + ".bias")), g(L.bn + ".running_mean"), g(L.bn + ".running_var"))

    def _build_static(self):
        """Allocate packed-weight buffers and prebuild every launch descriptor (pointers stay fixed; the
        contents of the weight buffers are refreshed each step by `_repack`)."""
        n = self.n
        for L in self.layers:
            S = self.st[L.name]
            w, b = self._conv_params(L)
            S.plan = pack_conv(w, b, src_channels=L.c_in, stride=L.stride, relu=False, precision=self.prec, name=L.conv)
            srcs = [self.act[k] for k in L.srcs]
            hi, wi = S.hw_in
            if L.bn:
                out = (S.z,)
            else:
                out = None   # head 1x1: result tensors are supplied per call
            if out is not None:
                S.fwd = ops.ConvCall(S.plan, srcs, L.ups, out, n=n, h_in=hi, w_in=wi)
            # ---- BatchNorm descriptor (pointers to gamma/beta/running stats are refreshed per step) ----
            if L.bn:
                d = BnDesc()
                ho, wo = S.hw_out
                d.z, d.n, d.h, d.w, d.c = S.z.data_ptr(), n, ho, wo, L.c_out
                d.momentum, d.eps = BN_MOMENTUM, BN_EPS
                d.sums, d.stats = self.sums.data_ptr(), S.stats.data_ptr()
                d.out_hi, d.out_lo_off = self.act[L.out].data_ptr(), ops._lo_off(self.act[L.out])
                d.relu = 1
                d.dz_hi, d.dz_lo_off = S.dz.data_ptr(), ops._lo_off(S.dz)
                S.bn = d
            # ---- data gradient: one conv per source with the transposed + flipped weights ----------
            if L.need_dgrad:
                c0 = 0
                for si, (key, cs, up) in enumerate(zip(L.srcs, L.c_in, L.ups)):
                    wt = self._dgrad_weight(w, c0, cs)
                    dp = pack_conv(wt, torch.zeros(cs, device=w.device), src_channels=[L.c_out], stride=1, relu=False,
                                   precision=self.prec, name=L.conv + f".dgrad{si}")
                    gb = torch.empty((n, hi, wi, cs), dtype=torch.float32, device=self.dev)
                    call = ops.ConvCall(dp, [S.dz], [2 if L.stride == 2 else 0], (gb,), n=n, h_in=hi, w_in=wi)
                    S.dplans.append(dp); S.dcalls.append(call); S.gbufs.append(gb)
                    c0 += cs
            # ---- weight gradient -------------------------------------------------------------------------
            wg = WgradDesc()
            for i, s in enumerate(srcs):
                wg.src[i], wg.src_lo_off[i], wg.src_c[i], wg.src_up[i] = s.data_ptr(), ops._lo_off(s), s.shape[-1], int(L.ups[i])
            wg.n, wg.h_in, wg.w_in = n, hi, wi
            wg.h_out, wg.w_out = S.hw_out
            wg.stride, wg.taps = L.stride, L.taps
            wg.dz_hi, wg.dz_lo_off, wg.c_out = S.dz.data_ptr(), ops._lo_off(S.dz), L.c_out
            wg.c_in_real, wg.passes = L.c_in_real, 3
            S.wg = wg
        if self.fused_key:
            self._build_fusion()

    @staticmethod
    def _dgrad_weight(w: torch.Tensor, c0: int, cs: int) -> torch.Tensor:
        """W [co, ci, k, k] -> weights of the data-gradient conv for input channels [c0, c0+cs):
        W'[ci, co, kh, kw] = W[co, c0+ci, k-1-kh, k-1-kw]."""
        return w[:, c0:c0 + cs].flip(2, 3).permute(1, 0, 2, 3).contiguous()

    def _pack(self, plan: ConvPlan, w: torch.Tensor, bias: Optional[torch.Tensor], stream, transpose=False, c0=0,
              n_real=None):
        """Device-side re-pack of `w` [co, ci, k, k] (fp32, contiguous) into plan.wpack (+ plan.bias)."""
        d = PackDesc()
        d.w = w.data_ptr()
        d.co_src, d.ci_src, d.taps = w.shape[0], w.shape[1], plan.taps
        d.transpose, d.c0 = int(transpose), c0
        d.n_real = plan.c_out if n_real is None else n_real
        d.k_pad, d.block_n, d.c_blk = plan.c_in, plan.block_n, plan.c_blk
        d.n_tiles = (plan.c_out + plan.block_n - 1) // plan.block_n
        d.stacked = int(plan.stacked)
        d.wpack = plan.wpack.data_ptr()
        d.bias_src = bias.data_ptr() if bias is not None else None
        d.bias = plan.bias.data_ptr() if bias is not None else None
        check(self.lib.disco_pack_weights(C.byref(d), stream), f"pack_weights[{plan.name}]")

    def _raw_params(self, L: Layer):
        """(weight [co, ci_real, k, k] contiguous fp32, bias) WITHOUT channel padding (the pack kernel pads)."""
        if L.name in ("h3", "h1"):
            return self._conv_params(L)
        g = self.get
        return _w4(_f32(g(L.conv + ".weight"))), _f32(g(L.conv + ".bias"))

    def _repack(self, stream=None):
        """Refresh the packed operand images from the current parameter values (same pointers): one pack_weights
        launch per forward / data-gradient weight image."""
        if stream is None:
            stream = torch.cuda.current_stream(self.dev).cuda_stream
        keep = self._pack_keep = []
        for L in self.layers:
            S = self.st[L.name]
            w, b = self._raw_params(L)
            keep += [w, b]
            self._pack(S.plan, w, b, stream)
            c0 = 0
            if L.need_dgrad:
                for si, cs in enumerate(L.c_in):
                    self._pack(S.dplans[si], w, None, stream, transpose=True, c0=c0, n_real=cs)
                    c0 += cs
        if self.fused_key:
            w, b = self._pwf_en_params()
            keep += [w, b]
            self._pack(self.en_plan, w, b, stream)
            self._pack(self.en_dplan, w, None, stream, transpose=True, c0=0, n_real=self.fuse_c)

    # ------------------------------------------------------------------------------------------------
    # fusion block
    def _pwf_en_params(self):
        g, p = self.get, self.pwf_prefix
        w1 = _f32(g(p + "conv1_1.weight"))           # [128, 2C, 1, 1]
        Cc = w1.shape[1] // 2
        w = torch.cat((w1[:, :Cc], w1[:, Cc:]), 0)   # [256, C, 1, 1]: ego half | neighbour half
        b1 = _f32(g(p + "conv1_1.bias"))
        return w, torch.cat((b1, torch.zeros_like(b1)), 0)

    def _build_fusion(self):
        n, dev, B, A = self.n, self.dev, self.B, self.A
        hf, wf = self.fuse_hw
        cf = self.fuse_c
        feat = self.act[self.feat_key]
        w, b = self._pwf_en_params()
        self.en_plan = pack_conv(w, b, src_channels=[cf], relu=False, precision=self.prec, name="pwf.conv1_1")
        self.en = torch.empty((n, hf, wf, 256), dtype=torch.float32, device=dev)
        self.en_call = ops.ConvCall(self.en_plan, [feat], [0], (self.en,), n=n, h_in=hf, w_in=wf)
        self.den = torch.zeros((n, hf, wf, 256), dtype=torch.float32, device=dev)
        self.den_act = ops.alloc_act(n, hf, wf, 256, self.prec, dev)
        self.en_dplan = pack_conv(self._dgrad_weight(w, 0, cf), torch.zeros(cf, device=dev), src_channels=[256], relu=False,
                                  precision=self.prec, name="pwf.conv1_1.dgrad")
        self.en_gbuf = torch.empty((n, hf, wf, cf), dtype=torch.float32, device=dev)
        self.en_dcall = ops.ConvCall(self.en_dplan, [self.den_act], [0], (self.en_gbuf,), n=n, h_in=hf, w_in=wf)
        wg = WgradDesc()
        wg.src[0], wg.src_lo_off[0], wg.src_c[0], wg.src_up[0] = feat.data_ptr(), ops._lo_off(feat), cf, 0
        wg.n, wg.h_in, wg.w_in, wg.h_out, wg.w_out = n, hf, wf, hf, wf
        wg.stride, wg.taps = 1, 1
        wg.dz_hi, wg.dz_lo_off, wg.c_out = self.den_act.data_ptr(), ops._lo_off(self.den_act), 256
        wg.c_in_real, wg.passes = cf, 3
        self.en_wg = wg
        self.trans = torch.zeros((B, A, A, 4, 4), dtype=torch.float64, device=dev)
        self.na = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.outage = torch.zeros((B, A), dtype=torch.int32, device=dev)
        self.psum = torch.zeros((B * A * A, 2, 168), dtype=torch.float64, device=dev)
        self.gsum = torch.zeros((B * A * A, 2, 168), dtype=torch.float32, device=dev)
        self.wlogit = torch.zeros((B, A, A, hf, wf), dtype=torch.float32, device=dev)
        self.dwlogit = torch.zeros((B, A, A, hf, wf), dtype=torch.float32, device=dev)
        self.weights = torch.zeros((B, A, A, hf, wf), dtype=torch.float32, device=dev)
        self.dfeat = torch.zeros((n, hf, wf, cf), dtype=torch.float32, device=dev)
        self.dfused = torch.zeros((n, hf, wf, cf), dtype=torch.float32, device=dev)
        self.dparams = torch.zeros(4697, dtype=torch.float32, device=dev)
        f = FusionDesc()
        f.feat_hi, f.feat_lo_off, f.precision = feat.data_ptr(), ops._lo_off(feat), self.prec
        f.hid = 128
        f.trans, f.num_agent, f.outage = self.trans.data_ptr(), self.na.data_ptr(), self.outage.data_ptr()
        f.B, f.A, f.h, f.w, f.C = B, A, hf, wf, cf
        f.only_v2i = int(bool(self.only_v2i))
        f.trans_scale = 4.0 / 128.0
        fused = self.act[self.fused_key]
        f.out_hi, f.out_lo_off = fused.data_ptr(), ops._lo_off(fused)
        f.weights = self.weights.data_ptr()
        f.row_begin, f.row_end = 0, n
        f.wpre = self.wlogit.data_ptr()
        self.fusion = f
        p = PwfTrainDesc()
        p.feat_hi, p.feat_lo_off, p.en, p.hid = feat.data_ptr(), ops._lo_off(feat), self.en.data_ptr(), 128
        p.eps, p.momentum = BN_EPS, BN_MOMENTUM
        p.trans, p.num_agent, p.outage = self.trans.data_ptr(), self.na.data_ptr(), self.outage.data_ptr()
        p.B, p.A, p.h, p.w, p.C = B, A, hf, wf, cf
        p.only_v2i, p.trans_scale = int(bool(self.only_v2i)), 4.0 / 128.0
        p.psum, p.wlogit, p.gsum = self.psum.data_ptr(), self.wlogit.data_ptr(), self.gsum.data_ptr()
        p.dfused, p.dwlogit = self.dfused.data_ptr(), self.dwlogit.data_ptr()
        p.dfeat, p.den, p.dparams = self.dfeat.data_ptr(), self.den.data_ptr(), self.dparams.data_ptr()
        self.pwf = p

    def _refresh_pwf_params(self):
        g, pp, p = self.get, self.pwf_prefix, self.pwf
        keep = self._pwf_keep = {}

        def ptr(name, reshape=None):
            t = _f32(g(pp + name))
            if reshape:
                t = t.reshape(reshape).contiguous()
            keep[name] = t
            return t.data_ptr()

        p.g1, p.be1 = ptr("bn1_1.weight"), ptr("bn1_1.bias")
        p.w2, p.b2, p.g2, p.be2 = ptr("conv1_2.weight", (32, 128)), ptr("conv1_2.bias"), ptr("bn1_2.weight"), ptr("bn1_2.bias")
        p.w3, p.b3, p.g3, p.be3 = ptr("conv1_3.weight", (8, 32)), ptr("conv1_3.bias"), ptr("bn1_3.weight"), ptr("bn1_3.bias")
        p.w4, p.b4 = ptr("conv1_4.weight", (1, 8)), ptr("conv1_4.bias")
        for k, nm in (("1", "bn1_1"), ("2", "bn1_2"), ("3", "bn1_3")):
            setattr(p, "rm" + k, g(pp + nm + ".running_mean").data_ptr())
            setattr(p, "rv" + k, g(pp + nm + ".running_var").data_ptr())
            setattr(p, "nbt" + k, g(pp + nm + ".num_batches_tracked").data_ptr())

    # ------------------------------------------------------------------------------------------------
    def forward(self, bevs: torch.Tensor, trans=None, num_agent=None, outage_host=None):
        """Runs the training-mode forward; returns dict of result tensors (cls, loc fp32 NHWC) ."""
        dev = self.dev
        stream = torch.cuda.current_stream(dev).cuda_stream
        self._repack()
        bev = bevs.detach()
        if bev.dtype != torch.float32:
            bev = bev.float()
        ops.bev_pack(bev.contiguous(), self.act["a0"], self.prec)
        self._keep = []
        if self.fused_key:
            self.trans.copy_(trans.detach(), non_blocking=True)
            self.na.copy_(num_agent.detach()[:, 0], non_blocking=True)
            if outage_host is not None:
                self.outage.copy_(outage_host, non_blocking=True)
            else:
                self.outage.zero_()
            self._refresh_pwf_params()
        for L in self.enc:
            self._layer_fwd(L, stream)
        if self.fused_key and self.fusion_level in (2, 3):
            self._fusion_fwd(stream)
        for L in self.dec:
            self._layer_fwd(L, stream)
        out = {}
        if self.head_layers:
            self._layer_fwd(self.head_layers[0], stream)
            L = self.head_layers[1]
            S = self.st["h1"]
            n, h, w = self.n, self.h, self.w
            cls = torch.empty((n, h, w, 12), dtype=torch.float32, device=dev)
            loc = torch.empty((n, h, w, 36), dtype=torch.float32, device=dev)
            call = ops.ConvCall(S.plan, [self.act["hh"]], [0], (cls, loc), n=n, h_in=h, w_in=w, out_split=12)
            call.launch(stream)
            out["cls"], out["loc"] = cls, loc
        return out

    def _layer_fwd(self, L: Layer, stream):
        S = self.st[L.name]
        S.fwd.launch(stream)
        gam, bet, rm, rv = self._bn_params(L)
        self._keep += [gam, bet]
        d = S.bn
        d.gamma, d.beta = gam.data_ptr(), bet.data_ptr()
        if L.name == "h3":
            # two BatchNorm modules side by side: run on concatenated copies of the running stats, write back
            self._keep += [rm, rv]
            d.running_mean, d.running_var, d.num_batches_tracked = rm.data_ptr(), rv.data_ptr(), None
            check(self.lib.disco_bn_train_forward(C.byref(d), stream), "bn_fwd[h3]")
            g = self.get
            for nm, sl in (("classification.bn1", slice(0, 32)), ("regression.box_prediction.1", slice(32, 64))):
                g(nm + ".running_mean").copy_(rm[sl]); g(nm + ".running_var").copy_(rv[sl])
                g(nm + ".num_batches_tracked").add_(1)
        else:
            d.running_mean, d.running_var = rm.data_ptr(), rv.data_ptr()
            d.num_batches_tracked = self.get(L.bn + ".num_batches_tracked").data_ptr()
            check(self.lib.disco_bn_train_forward(C.byref(d), stream), f"bn_fwd[{L.name}]")

    def _fusion_fwd(self, stream):
        self.en_call.launch(stream)
        check(self.lib.disco_pwf_train_forward(C.byref(self.pwf), stream), "pwf_train_forward")
        ops.fusion_forward(self.fusion, stream)

    def kd_map(self, key: str) -> torch.Tensor:
        return ops.act_to_nchw_f32(self.act[key], self.prec)

    # ------------------------------------------------------------------------------------------------
    def backward(self, grads: Dict[str, Optional[torch.Tensor]]) -> Dict[str, torch.Tensor]:
        """grads: {"cls": [n,h,w,12] fp32 NHWC, "loc": [n,h,w,36], "x8"/"x7"/"x6"/"x5"/fused_key: NCHW fp32 or None}.
        Returns {parameter name: gradient}."""
        dev, lib = self.dev, self.lib
        stream = torch.cuda.current_stream(dev).cuda_stream
        out: Dict[str, torch.Tensor] = {}
        gsrc: Dict[str, list] = {}
        keep = []

        def add_src(key, t, c_total, c_off, pool):
            gsrc.setdefault(key, []).append((t, c_total, c_off, pool))

        # external gradients of the KD maps arrive NCHW
        for key, g in grads.items():
            if key in ("cls", "loc") or g is None:
                continue
            if key not in self.act:
                raise KeyError(f"no activation buffer {key!r} to receive a gradient")
            g = g.detach().float().contiguous()
            nn_, c, hh, ww = g.shape
            t = torch.empty((nn_, hh, ww, c), dtype=torch.float32, device=dev)
            check(lib.disco_nchw_to_nhwc(g.data_ptr(), nn_, c, hh, ww, t.data_ptr(), stream), "nchw_to_nhwc")
            keep += [g, t]
            add_src(key, t, c, 0, 0)

        def run_wgrad(wg: WgradDesc, shape) -> torch.Tensor:
            dw = torch.empty(shape, dtype=torch.float32, device=dev)
            wg.dw = dw.data_ptr()
            wg.partial, wg.splits = None, 0
            need = lib.disco_conv_wgrad_splits(C.byref(wg))
            check(need, "wgrad_splits")
            c_in = wg.src_c[0] + wg.src_c[1]
            numel = need * wg.c_out * wg.taps * c_in
            if self.partial is None or self.partial.numel() < numel:
                self.partial = torch.empty(numel, dtype=torch.float32, device=dev)
            wg.partial, wg.splits = self.partial.data_ptr(), need
            check(lib.disco_conv_wgrad(C.byref(wg), stream), "wgrad")
            return dw

        def layer_bwd(L: Layer):
            S = self.st[L.name]
            n = self.n
            if L.bn:
                d = S.bn
                srcs = gsrc.get(L.out, [])
                if not srcs:
                    raise RuntimeError(f"no gradient reaches {L.out}")
                d.n_g = len(srcs)
                for i, (t, ct, co, pool) in enumerate(srcs):
                    d.g[i].ptr, d.g[i].c_total, d.g[i].c_off, d.g[i].pool = t.data_ptr(), ct, co, pool
                gam, bet, _, _ = self._bn_params(L)
                dgam = torch.empty(L.c_out, dtype=torch.float32, device=dev)
                dbet = torch.empty(L.c_out, dtype=torch.float32, device=dev)
                keep.extend([gam, bet])
                d.gamma, d.beta, d.dgamma, d.dbeta = gam.data_ptr(), bet.data_ptr(), dgam.data_ptr(), dbet.data_ptr()
                check(lib.disco_bn_train_backward(C.byref(d), stream), f"bn_bwd[{L.name}]")
            dw = run_wgrad(S.wg, (L.c_out, L.c_in_real, L.taps))
            for call, gb, key, cs, up in zip(S.dcalls, S.gbufs, L.srcs, L.c_in, L.ups):
                call.launch(stream)
                add_src(key, gb, cs, 0, 1 if up else 0)
            k = 3 if L.taps == 9 else 1
            if L.name == "h3":
                out["classification.conv1.weight"] = dw[:32].reshape(32, 32, 3, 3)
                out["regression.box_prediction.0.weight"] = dw[32:].reshape(32, 32, 3, 3)
                out["classification.bn1.weight"], out["regression.box_prediction.1.weight"] = dgam[:32], dgam[32:]
                out["classification.bn1.bias"], out["regression.box_prediction.1.bias"] = dbet[:32], dbet[32:]
                # a conv bias in front of a BatchNorm has an exactly zero gradient
                out["classification.conv1.bias"] = torch.zeros(32, device=dev)
                out["regression.box_prediction.0.bias"] = torch.zeros(32, device=dev)
            elif L.name == "h1":
                out["classification.conv2.weight"] = dw[:12, :32].reshape(12, 32, 1, 1)
                out["regression.box_prediction.3.weight"] = dw[12:, 32:].reshape(36, 32, 1, 1)
            else:
                pshape = self.get(L.conv + ".weight").shape
                out[L.conv + ".weight"] = dw.reshape(pshape)
                out[L.conv + ".bias"] = torch.zeros(L.c_out, device=dev)
                out[L.bn + ".weight"], out[L.bn + ".bias"] = dgam, dbet

        # ---- heads ------------------------------------------------------------------------------------
        if self.head_layers:
            gc, gl = grads.get("cls"), grads.get("loc")
            n, h, w = self.n, self.h, self.w
            gc = torch.zeros((n, h, w, 12), device=dev) if gc is None else gc.detach().float().contiguous()
            gl = torch.zeros((n, h, w, 36), device=dev) if gl is None else gl.detach().float().contiguous()
            keep += [gc, gl]
            S = self.st["h1"]
            npix = n * h * w
            check(lib.disco_grad_pack(gc.data_ptr(), 12, gl.data_ptr(), 36, npix, S.dz.data_ptr(), ops._lo_off(S.dz), stream),
                  "grad_pack[heads]")
            bc = torch.empty(12, dtype=torch.float32, device=dev)
            bl = torch.empty(36, dtype=torch.float32, device=dev)
            check(lib.disco_channel_sum(gc.data_ptr(), npix, 12, self.sums.data_ptr(), bc.data_ptr(), stream), "channel_sum")
            check(lib.disco_channel_sum(gl.data_ptr(), npix, 36, self.sums.data_ptr(), bl.data_ptr(), stream), "channel_sum")
            out["classification.conv2.bias"], out["regression.box_prediction.3.bias"] = bc, bl
            layer_bwd(self.head_layers[1])
            layer_bwd(self.head_layers[0])
        for L in reversed(self.dec):
            layer_bwd(L)
        if self.fused_key:
            self._fusion_bwd(gsrc, add_src, run_wgrad, out, stream, keep)
        for L in reversed(self.enc):
            layer_bwd(L)
        self._bwd_keep = keep
        return out

    def _fusion_bwd(self, gsrc, add_src, run_wgrad, out, stream, keep):
        lib, p, dev = self.lib, self.pwf, self.dev
        srcs = gsrc.get(self.fused_key, [])
        if not srcs:
            raise RuntimeError("no gradient reaches the fused map")
        # gradient wrt the fused map = decoder data gradient (+ KD gradient)
        n_el = self.dfused.numel()
        (t0, ct0, co0, pl0) = srcs[0]
        assert ct0 == self.fuse_c and co0 == 0 and pl0 == 0
        if len(srcs) == 1:
            p.dfused = t0.data_ptr()
        else:
            (t1, ct1, co1, pl1) = srcs[1]
            assert ct1 == self.fuse_c and co1 == 0 and pl1 == 0 and len(srcs) == 2
            check(lib.disco_add_f32(self.dfused.data_ptr(), t0.data_ptr(), t1.data_ptr(), n_el, stream), "add_f32")
            p.dfused = self.dfused.data_ptr()
        self.dfeat.zero_(); self.den.zero_(); self.dparams.zero_()
        check(lib.disco_fusion_combine_backward(C.byref(p), stream), "fusion_combine_backward")
        check(lib.disco_pwf_train_backward(C.byref(p), stream), "pwf_train_backward")
        hf, wf = self.fuse_hw
        npix = self.n * hf * wf
        check(lib.disco_grad_pack(self.den.data_ptr(), 256, None, 0, npix, self.den_act.data_ptr(), ops._lo_off(self.den_act),
                                  stream), "grad_pack[den]")
        dw = run_wgrad(self.en_wg, (256, self.fuse_c, 1))
        self.en_dcall.launch(stream)
        add_src(self.feat_key, self.en_gbuf, self.fuse_c, 0, 0)
        add_src(self.feat_key, self.dfeat, self.fuse_c, 0, 0)
        pp = self.pwf_prefix
        dp = self.dparams
        out[pp + "conv1_1.weight"] = torch.cat((dw[:128], dw[128:]), 1).reshape(128, 2 * self.fuse_c, 1, 1)
        out[pp + "conv1_1.bias"] = torch.zeros(128, device=dev)
        out[pp + "bn1_1.weight"], out[pp + "bn1_1.bias"] = dp[0:128].clone(), dp[128:256].clone()
        out[pp + "conv1_2.weight"] = dp[256:4352].clone().reshape(32, 128, 1, 1)
        out[pp + "conv1_2.bias"] = torch.zeros(32, device=dev)
        out[pp + "bn1_2.weight"], out[pp + "bn1_2.bias"] = dp[4352:4384].clone(), dp[4384:4416].clone()
        out[pp + "conv1_3.weight"] = dp[4416:4672].clone().reshape(8, 32, 1, 1)
        out[pp + "conv1_3.bias"] = torch.zeros(8, device=dev)
        out[pp + "bn1_3.weight"], out[pp + "bn1_3.bias"] = dp[4672:4680].clone(), dp[4680:4688].clone()
        out[pp + "conv1_4.weight"] = dp[4688:4696].clone().reshape(1, 8, 1, 1)
        out[pp + "conv1_4.bias"] = dp[4696:4697].clone()
