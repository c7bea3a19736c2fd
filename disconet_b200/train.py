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
import os
from dataclasses import dataclass
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
    kind: str = "conv"           # "conv" | "pool" (MaxPool2d(2): srcs[0] -> out) | "up" (bilinear x2, align_corners: srcs[0] -> out)
    n_real: int = 0              # real output channels when c_out is padded (seg OutConv: 8 of 16)

    def __post_init__(self):
        if not self.c_in_real:
            self.c_in_real = sum(self.c_in)


def backbone_layers(pe: str, pd: str, fused_key: Optional[str], fusion_level: int = 3,
                    compress_cc: int = 0) -> Tuple[List[Layer], List[Layer]]:
    """Backbone.encode / decode layer tables (Backbone.py:89-143,145-242).  compress_cc > 0: the 1x1 bottleneck pair
    on x_3 (Backbone.py:73-87,139-141; applied AFTER conv4 has read the uncompressed x_3)."""
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
    x3_key = "x3"
    if compress_cc:
        E += [
            Layer("compress", pe + "com_compresser", pe + "bn_compress", ["x3"], [0], "x3c", [256], compress_cc, taps=1, level=3),
            Layer("decompress", pe + "com_decompresser", pe + "bn_decompress", ["x3c"], [0], "x3d", [compress_cc], 256, taps=1, level=3),
        ]
        x3_key = "x3d"
    x3d = fused_key if (fused_key and fusion_level == 3) else x3_key
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


USE_GRAPH = os.environ.get("DISCO_B200_TRAIN_GRAPH", "1") != "0"


class TrainRunner:
    """Training-mode workspace of one model for one (N, H, W, B) problem size.

    Every buffer, descriptor and gradient-source wiring is static, so a step is a flat list of prebuilt C-ABI calls
    (`fwd_ops`, `bwd_ops`: ~150 + ~200 launches) and, after two eager steps, two CUDA-graph replays.  The descriptors
    hold raw pointers into the parameters; `_signature()` detects re-allocated parameters and rebuilds.
    """

    def __init__(self, get, n: int, h: int, w: int, device, enc_prefix: str, dec_prefix: str, *, heads: bool = True,
                 pwf_prefix: Optional[str] = None, batch_size: int = 1, agents: int = 1, fusion_level: int = 3,
                 only_v2i: bool = False, kd_keys: Sequence[str] = (), compress_level: int = 0,
                 nodes: Optional[Tuple[List[Layer], List[Layer]]] = None, fusion_spec: Optional[tuple] = None):
        """`nodes` = (encoder, decoder) layer lists replacing the detection backbone tables (the segmentation U-Net
        passes its own, incl. pool / up nodes and the `outc` output conv); `fusion_spec` = (feature key, fused key,
        resolution level, channels) of the collaboration map when it is not x_3 / x_2 of the detection backbone."""
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
        cc = 256 // (2 ** compress_level) if compress_level > 0 else 0
        if cc and cc % 16:
            raise NotImplementedError("training mode supports compress_level <= 4 (bottleneck width a multiple of 16)")
        self.x3_key = "x3d" if cc else "x3"
        self.fusion_spec = fusion_spec
        if fusion_spec is not None:
            self.fused_key = fusion_spec[1]
        if nodes is not None:
            E, D = nodes
        else:
            E, D = backbone_layers(enc_prefix, dec_prefix, self.fused_key, fusion_level, compress_cc=cc)
        self.enc, self.dec = E, D
        self.head_layers: List[Layer] = []
        if heads:
            self.head_layers = [
                Layer("h3", "heads.3x3", "heads.bn", ["x8"], [0], "hh", [32], 64, level=0),
                Layer("h1", "heads.1x1", None, ["hh"], [0], "", [64], 48, taps=1, level=0),
            ]
        self.layers = E + D + self.head_layers
        self.kd_keys = tuple(kd_keys)            # activation buffers that may receive an external (NCHW) gradient
        self.res = [(h >> k, w >> k) for k in range(5)]
        self.act: Dict[str, torch.Tensor] = {}
        self.st: Dict[str, _LayerState] = {}
        A_ = lambda hh, ww, c: ops.alloc_act(n, hh, ww, c, self.prec, device)
        self.act["a0"] = A_(h, w, 16)
        self.bev_in = torch.zeros((n, 1, h, w, 13), dtype=torch.float32, device=device)
        self.sums = torch.zeros(1024, dtype=torch.float64, device=device)
        for L in self.layers:
            hi, wi = self.res[L.level]
            ho, wo = (hi - 1) // L.stride + 1, (wi - 1) // L.stride + 1
            S = self.st[L.name] = _LayerState()
            if L.kind != "conv":
                ho, wo = (hi // 2, wi // 2) if L.kind == "pool" else (hi * 2, wi * 2)
                self.act[L.out] = A_(ho, wo, L.c_out)
                S.hw_in, S.hw_out = (hi, wi), (ho, wo)
                continue
            S.z = torch.empty((n, ho, wo, L.c_out), dtype=torch.float32, device=device) if L.bn else None
            if L.out:
                self.act[L.out] = A_(ho, wo, L.c_out)
            S.dz = A_(ho, wo, L.c_out)
            S.hw_in, S.hw_out = (hi, wi), (ho, wo)
            if L.bn:
                S.stats = torch.zeros(2 * L.c_out, dtype=torch.float32, device=device)
        if self.fused_key:
            if fusion_spec is not None:
                self.feat_key, _, lvl, cf = fusion_spec
                hf, wf = self.res[lvl]
            else:
                hf, wf = self.res[fusion_level]
                cf = 256 if fusion_level == 3 else 128
                self.feat_key = self.x3_key if fusion_level == 3 else "x2"
            self.fuse_hw, self.fuse_c = (hf, wf), cf
            self.act[self.fused_key] = A_(hf, wf, cf)
        self.has_outc = any(L.name == "outc" for L in self.layers)
        if self.has_outc:
            self.logits = torch.empty((n, h, w, 16), dtype=torch.float32, device=device)
            self.glogits = torch.zeros((n, h, w, 8), dtype=torch.float32, device=device)
            self.gzero8 = torch.zeros((n, h, w, 8), dtype=torch.float32, device=device)
        if heads:
            self.cls = torch.empty((n, h, w, 12), dtype=torch.float32, device=device)
            self.loc = torch.empty((n, h, w, 36), dtype=torch.float32, device=device)
            self.gcls, self.gloc = torch.zeros_like(self.cls), torch.zeros_like(self.loc)
        # external (KD) gradients, converted NCHW -> NHWC into static buffers
        self.ext: Dict[str, torch.Tensor] = {}
        self.ext_dirty: Dict[str, bool] = {}
        for k in self.kd_keys:
            a = self.act[k]
            self.ext[k] = torch.zeros(tuple(a.shape[1:]), dtype=torch.float32, device=device)
            self.ext_dirty[k] = False
        self._sig = None
        self._graphs = {"fwd": None, "bwd": None}
        self._steps = 0
        self.generation = 0          # bumped by every forward: the saved activations belong to the LAST forward only
        self._build()

    # ------------------------------------------------------------------------------------------------
    # parameters
    def _p(self, name: str) -> torch.Tensor:
        t = self.get(name)
        if t.dtype not in (torch.float32, torch.int64) or not t.is_contiguous() or not t.is_cuda:
            raise ValueError(f"parameter {name} must be a contiguous CUDA fp32 tensor for the training kernels")
        return t.detach()

    def _signature(self):
        return tuple(self.get(k).data_ptr() for k in self._ptr_names)

    def _special_refresh(self):
        """Composite weights of the fused heads / PWF conv1_1 halves -> static buffers (plain tensor bookkeeping)."""
        p = self._p
        sp = self.sp
        if self.head_layers:
            torch.cat((p("classification.conv1.weight"), p("regression.box_prediction.0.weight")), 0, out=sp["h3.w"])
            torch.cat((p("classification.conv1.bias"), p("regression.box_prediction.0.bias")), 0, out=sp["h3.b"])
            wc, wr = p("classification.conv2.weight"), p("regression.box_prediction.3.weight")
            sp["h1.w"][:12, :32] = wc
            sp["h1.w"][12:, 32:] = wr
            torch.cat((p("classification.conv2.bias"), p("regression.box_prediction.3.bias")), 0, out=sp["h1.b"])
            for f, key in (("weight", "h3.gamma"), ("bias", "h3.beta"), ("running_mean", "h3.rm"), ("running_var", "h3.rv")):
                torch.cat((p("classification.bn1." + f), p("regression.box_prediction.1." + f)), 0, out=sp[key])
        if self.fused_key:
            pp, Cc = self.pwf_prefix, self.fuse_c
            w1 = p(pp + "conv1_1.weight")                # [128, 2C, 1, 1]
            sp["en.w"][:128] = w1[:, :Cc]
            sp["en.w"][128:] = w1[:, Cc:]
            sp["en.b"][:128] = p(pp + "conv1_1.bias")

    def _h3_writeback(self):
        sp, g = self.sp, self.get
        for nm, sl in (("classification.bn1", slice(0, 32)), ("regression.box_prediction.1", slice(32, 64))):
            g(nm + ".running_mean").copy_(sp["h3.rm"][sl])
            g(nm + ".running_var").copy_(sp["h3.rv"][sl])
            g(nm + ".num_batches_tracked").add_(1)

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _dgrad_weight(w: torch.Tensor, c0: int, cs: int) -> torch.Tensor:
        """W [co, ci, k, k] -> weights of the data-gradient conv for input channels [c0, c0+cs):
        W'[ci, co, kh, kw] = W[co, c0+ci, k-1-kh, k-1-kw]."""
        return w[:, c0:c0 + cs].flip(2, 3).permute(1, 0, 2, 3).contiguous()

    def _pack(self, plan: ConvPlan, w: torch.Tensor, bias: Optional[torch.Tensor], stream, transpose=False, c0=0,
              n_real=None):
        """Device-side re-pack of `w` [co, ci, k, k] (fp32, contiguous) into plan.wpack (+ plan.bias), immediately."""
        d = self._pack_desc(plan, w, bias, transpose, c0, n_real)
        check(self.lib.disco_pack_weights(C.byref(d), stream), f"pack_weights[{plan.name}]")

    @staticmethod
    def _pack_desc(plan: ConvPlan, w: torch.Tensor, bias, transpose=False, c0=0, n_real=None) -> PackDesc:
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
        return d

    # ------------------------------------------------------------------------------------------------
    def _galloc(self, name: str, shape) -> int:
        """Reserve a slice of the flat gradient buffer; returns its offset."""
        numel = 1
        for s_ in shape:
            numel *= s_
        off = self._g_size
        self._g_index[name] = (off, numel, tuple(shape))
        self._g_size += (numel + 3) // 4 * 4          # 16-byte aligned slices
        return off

    def _gview(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, numel, shape = self._g_index[name]
        return buf[off:off + numel].view(shape)

    def _build(self):
        """(Re)build packed-weight buffers, descriptors and the flat op lists for the current parameter storage."""
        n, dev, lib = self.n, self.dev, self.lib
        p = self._p
        self._keep = []
        self._g_index: Dict[str, tuple] = {}
        self._g_size = 0
        self._ptr_names: List[str] = []
        sp = self.sp = {}
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
        if self.head_layers:
            sp.update({"h3.w": z(64, 32, 3, 3), "h3.b": z(64), "h1.w": z(48, 64, 1, 1), "h1.b": z(48),
                       "h3.gamma": z(64), "h3.beta": z(64), "h3.rm": z(64), "h3.rv": z(64)})
        if self.fused_key:
            sp.update({"en.w": z(256, self.fuse_c, 1, 1), "en.b": z(256)})
        self._special_refresh()

        def raw(L: Layer):
            if L.name in ("h3", "h1"):
                return sp[L.name + ".w"], sp[L.name + ".b"]
            self._ptr_names += [L.conv + ".weight", L.conv + ".bias"]
            return _w4(p(L.conv + ".weight")), p(L.conv + ".bias")

        pre: list = [("py", self._special_refresh)]
        fwd_layer: Dict[str, list] = {}
        bwd_layer: Dict[str, list] = {}
        wg_list: List[WgradDesc] = []
        # gradient sources per activation buffer: external KD gradients first
        gsrc: Dict[str, list] = {k: [(self.ext[k], self.ext[k].shape[-1], 0, 0)] for k in self.kd_keys}

        # ---- forward descriptors + packs --------------------------------------------------------------------
        for L in self.layers:
            S = self.st[L.name]
            if L.kind != "conv":
                src, dst = self.act[L.srcs[0]], self.act[L.out]
                hi, wi = S.hw_in
                fn = lib.disco_maxpool2 if L.kind == "pool" else lib.disco_upsample_bilinear2x
                fwd_layer[L.name] = [("call", fn, (src.data_ptr(), ops._lo_off(src), dst.data_ptr(), ops._lo_off(dst), self.prec, n, hi, wi,
                                                   L.c_out), f"{L.kind}[{L.name}]")]
                continue
            w, b = raw(L)
            c_pad = sum(L.c_in)
            wz = torch.zeros(L.c_out, c_pad, w.shape[2], w.shape[3], device=dev)
            S.plan = pack_conv(wz, torch.zeros(L.c_out, device=dev), src_channels=L.c_in, stride=L.stride, relu=False,
                               precision=self.prec, name=L.conv)
            pd = self._pack_desc(S.plan, w, b, n_real=w.shape[0])
            pre.append((lib.disco_pack_weights, pd, f"pack[{L.name}]"))
            srcs = [self.act[k] for k in L.srcs]
            hi, wi = S.hw_in
            ho, wo = S.hw_out
            ops_f = []
            if L.bn:
                S.fwd = ops.ConvCall(S.plan, srcs, L.ups, (S.z,), n=n, h_in=hi, w_in=wi)
            elif L.name == "outc":
                S.fwd = ops.ConvCall(S.plan, srcs, L.ups, (self.logits,), n=n, h_in=hi, w_in=wi)
            else:
                S.fwd = ops.ConvCall(S.plan, srcs, L.ups, (self.cls, self.loc), n=n, h_in=hi, w_in=wi, out_split=12)
            ops_f.append((lib.disco_conv_forward, S.fwd.desc, f"conv[{L.name}]"))
            if L.bn:
                d = BnDesc()
                d.z, d.n, d.h, d.w, d.c = S.z.data_ptr(), n, ho, wo, L.c_out
                d.momentum, d.eps = BN_MOMENTUM, BN_EPS
                d.sums, d.stats = self.sums.data_ptr(), S.stats.data_ptr()
                d.out_hi, d.out_lo_off = self.act[L.out].data_ptr(), ops._lo_off(self.act[L.out])
                d.relu = 1
                d.dz_hi, d.dz_lo_off = S.dz.data_ptr(), ops._lo_off(S.dz)
                if L.name == "h3":
                    d.gamma, d.beta = sp["h3.gamma"].data_ptr(), sp["h3.beta"].data_ptr()
                    d.running_mean, d.running_var, d.num_batches_tracked = sp["h3.rm"].data_ptr(), sp["h3.rv"].data_ptr(), None
                else:
                    names = [L.bn + f for f in (".weight", ".bias", ".running_mean", ".running_var", ".num_batches_tracked")]
                    self._ptr_names += names
                    d.gamma, d.beta = p(names[0]).data_ptr(), p(names[1]).data_ptr()
                    d.running_mean, d.running_var = p(names[2]).data_ptr(), p(names[3]).data_ptr()
                    d.num_batches_tracked = p(names[4]).data_ptr()
                S.bn = d
                ops_f.append((lib.disco_bn_train_forward, d, f"bn_fwd[{L.name}]"))
                if L.name == "h3":
                    ops_f.append(("py", self._h3_writeback))
            fwd_layer[L.name] = ops_f

        # ---- backward descriptors, in backward order so that the gradient-source lists are complete ------------
        def wgrad_desc(srcs, ups, hw_in, hw_out, stride, taps, dz, c_out, c_in_real, gname):
            wg = WgradDesc()
            for i, s_ in enumerate(srcs):
                wg.src[i], wg.src_lo_off[i], wg.src_c[i], wg.src_up[i] = s_.data_ptr(), ops._lo_off(s_), s_.shape[-1], int(ups[i])
            wg.n, wg.h_in, wg.w_in = n, hw_in[0], hw_in[1]
            wg.h_out, wg.w_out = hw_out
            wg.stride, wg.taps = stride, taps
            wg.dz_hi, wg.dz_lo_off, wg.c_out = dz.data_ptr(), ops._lo_off(dz), c_out
            wg.c_in_real, wg.passes = c_in_real, 3
            wg._goff = self._galloc(gname, (c_out, c_in_real, taps))
            wg_list.append(wg)
            return wg

        def layer_bwd_ops(L: Layer):
            S = self.st[L.name]
            out_ops = []
            if L.kind != "conv":
                srcs = gsrc.get(L.out, [])
                if len(srcs) != 1 or srcs[0][1] != L.c_out or srcs[0][2] != 0 or srcs[0][3] != 0:
                    raise RuntimeError(f"{L.name}: a {L.kind} node needs exactly one dense gradient source for {L.out}")
                g_out = srcs[0][0]
                hi, wi = S.hw_in
                gx = torch.empty((n, hi, wi, L.c_out), dtype=torch.float32, device=dev)
                S.gbufs.append(gx)
                if L.kind == "pool":
                    x = self.act[L.srcs[0]]
                    out_ops.append(("call", lib.disco_maxpool2_backward, (x.data_ptr(), ops._lo_off(x), self.prec, g_out.data_ptr(),
                                                                          gx.data_ptr(), n, hi, wi, L.c_out), f"pool_bwd[{L.name}]"))
                else:
                    out_ops.append(("call", lib.disco_upsample_bilinear2x_backward, (g_out.data_ptr(), gx.data_ptr(), n, hi, wi, L.c_out),
                                    f"up_bwd[{L.name}]"))
                gsrc.setdefault(L.srcs[0], []).append((gx, L.c_out, 0, 0))
                return out_ops
            if L.bn:
                d = S.bn
                srcs = gsrc.get(L.out, [])
                if not srcs:
                    raise RuntimeError(f"no gradient reaches {L.out} (kd_keys={self.kd_keys})")
                if len(srcs) > 3:
                    raise RuntimeError(f"{L.out}: more than 3 gradient sources")
                d.n_g = len(srcs)
                for i, (t, ct, co, pool) in enumerate(srcs):
                    d.g[i].ptr, d.g[i].c_total, d.g[i].c_off, d.g[i].pool = t.data_ptr(), ct, co, pool
                d._goff = (self._galloc("_dgamma." + L.name, (L.c_out,)), self._galloc("_dbeta." + L.name, (L.c_out,)))
                out_ops.append((lib.disco_bn_train_backward, d, f"bn_bwd[{L.name}]"))
            S.wg = wgrad_desc([self.act[k] for k in L.srcs], L.ups, S.hw_in, S.hw_out, L.stride, L.taps, S.dz, L.c_out,
                              L.c_in_real, "_dw." + L.name)
            out_ops.append((lib.disco_conv_wgrad, S.wg, f"wgrad[{L.name}]"))
            if L.need_dgrad:
                w, _ = raw_cache[L.name]
                hi, wi = S.hw_in
                c0 = 0
                for si, (key, cs, up) in enumerate(zip(L.srcs, L.c_in, L.ups)):
                    dp = pack_conv(torch.zeros(cs, L.c_out, w.shape[2], w.shape[3], device=dev), torch.zeros(cs, device=dev),
                                   src_channels=[L.c_out], stride=1, relu=False, precision=self.prec, name=L.conv + f".dgrad{si}")
                    pre.append((lib.disco_pack_weights, self._pack_desc(dp, w, None, transpose=True, c0=c0, n_real=cs),
                                f"pack[{L.name}.dgrad{si}]"))
                    gb = torch.empty((n, hi, wi, cs), dtype=torch.float32, device=dev)
                    call = ops.ConvCall(dp, [S.dz], [2 if L.stride == 2 else 0], (gb,), n=n, h_in=hi, w_in=wi)
                    S.dplans.append(dp); S.dcalls.append(call); S.gbufs.append(gb)
                    out_ops.append((lib.disco_conv_forward, call.desc, f"dgrad[{L.name}.{si}]"))
                    gsrc.setdefault(key, []).append((gb, cs, 0, 1 if up else 0))
                    c0 += cs
            return out_ops

        raw_cache = {}
        for L in self.layers:
            if L.kind != "conv":
                continue
            if L.name in ("h3", "h1"):
                raw_cache[L.name] = (sp[L.name + ".w"], sp[L.name + ".b"])
            else:
                raw_cache[L.name] = (_w4(p(L.conv + ".weight")), p(L.conv + ".bias"))

        bwd: list = []
        if self.head_layers:
            S1 = self.st["h1"]
            npix = n * self.h * self.w
            ob_c, ob_l = self._galloc("classification.conv2.bias", (12,)), self._galloc("regression.box_prediction.3.bias", (36,))
            self._head_bias_off = (ob_c, ob_l)
            bwd.append(("grad_pack_heads", npix))
            bwd += layer_bwd_ops(self.head_layers[1])
            bwd += layer_bwd_ops(self.head_layers[0])
        for L in reversed(self.dec):
            if L.name == "outc":
                npix = n * self.h * self.w
                self._outc_bias_off = self._galloc("outc.conv.bias", (8,))
                bwd.append(("grad_pack_outc", npix))
            bwd += layer_bwd_ops(L)
        if self.fused_key:
            bwd += self._build_fusion(pre, gsrc, wgrad_desc)
        for L in reversed(self.enc):
            bwd += layer_bwd_ops(L)

        # ---- flat gradient buffer, wgrad partial workspace, final pointer fix-ups --------------------------------
        self._galloc("_zero", (512,))
        self.G = torch.zeros(self._g_size, dtype=torch.float32, device=dev)
        gbase = self.G.data_ptr()
        need = 0
        for wg in wg_list:
            wg.dw = gbase + 4 * wg._goff
            wg.partial, wg.splits = None, 0
            sp_ = lib.disco_conv_wgrad_splits(C.byref(wg))
            check(sp_, "wgrad_splits")
            wg.splits = sp_
            need = max(need, sp_ * wg.c_out * wg.taps * (wg.src_c[0] + wg.src_c[1]))
        self.partial = torch.empty(need, dtype=torch.float32, device=dev)
        for wg in wg_list:
            wg.partial = self.partial.data_ptr()
        for L in self.layers:
            d = self.st[L.name].bn
            if d is not None and hasattr(d, "_goff"):
                d.dgamma, d.dbeta = gbase + 4 * d._goff[0], gbase + 4 * d._goff[1]
        if self.fused_key:
            self.pwf.dparams = gbase + 4 * self._g_index["_pwf.dparams"][0]

        # ---- op lists ---------------------------------------------------------------------------------------------
        fwd: list = list(pre)
        fwd.append(("bev_pack",))
        for L in self.enc:
            fwd += fwd_layer[L.name]
        if self.fused_key:
            fwd += self.ops_fusion_fwd
        for L in self.dec:
            fwd += fwd_layer[L.name]
        for L in self.head_layers:
            fwd += fwd_layer[L.name]
        self.ops_pre, self.fwd_ops, self.bwd_ops = pre, fwd, bwd
        self._ptr_names = sorted(set(self._ptr_names))
        self._sig = self._signature()
        self._graphs = {"fwd": None, "bwd": None}
        self._steps = 0

    # ------------------------------------------------------------------------------------------------
    # fusion block
    def _build_fusion(self, pre, gsrc, wgrad_desc):
        n, dev, B, A, lib = self.n, self.dev, self.B, self.A, self.lib
        p, pp, sp = self._p, self.pwf_prefix, self.sp
        hf, wf = self.fuse_hw
        cf = self.fuse_c
        feat = self.act[self.feat_key]
        self.en_plan = pack_conv(torch.zeros(256, cf, 1, 1, device=dev), torch.zeros(256, device=dev), src_channels=[cf], relu=False,
                                 precision=self.prec, name="pwf.conv1_1")
        pre.append((lib.disco_pack_weights, self._pack_desc(self.en_plan, sp["en.w"], sp["en.b"]), "pack[en]"))
        self.en = torch.empty((n, hf, wf, 256), dtype=torch.float32, device=dev)
        self.en_call = ops.ConvCall(self.en_plan, [feat], [0], (self.en,), n=n, h_in=hf, w_in=wf)
        self.den = torch.zeros((n, hf, wf, 256), dtype=torch.float32, device=dev)
        self.den_act = ops.alloc_act(n, hf, wf, 256, self.prec, dev)
        self.en_dplan = pack_conv(torch.zeros(cf, 256, 1, 1, device=dev), torch.zeros(cf, device=dev), src_channels=[256], relu=False,
                                  precision=self.prec, name="pwf.conv1_1.dgrad")
        pre.append((lib.disco_pack_weights, self._pack_desc(self.en_dplan, sp["en.w"], None, transpose=True, c0=0, n_real=cf),
                    "pack[en.dgrad]"))
        self.en_gbuf = torch.empty((n, hf, wf, cf), dtype=torch.float32, device=dev)
        self.en_dcall = ops.ConvCall(self.en_dplan, [self.den_act], [0], (self.en_gbuf,), n=n, h_in=hf, w_in=wf)
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
        q = PwfTrainDesc()
        q.feat_hi, q.feat_lo_off, q.en, q.hid = feat.data_ptr(), ops._lo_off(feat), self.en.data_ptr(), 128
        q.eps, q.momentum = BN_EPS, BN_MOMENTUM
        q.trans, q.num_agent, q.outage = self.trans.data_ptr(), self.na.data_ptr(), self.outage.data_ptr()
        q.B, q.A, q.h, q.w, q.C = B, A, hf, wf, cf
        q.only_v2i, q.trans_scale = int(bool(self.only_v2i)), 4.0 / 128.0
        q.psum, q.wlogit, q.gsum = self.psum.data_ptr(), self.wlogit.data_ptr(), self.gsum.data_ptr()
        q.dwlogit, q.dfeat, q.den = self.dwlogit.data_ptr(), self.dfeat.data_ptr(), self.den.data_ptr()
        names = {"g1": "bn1_1.weight", "be1": "bn1_1.bias", "w2": "conv1_2.weight", "b2": "conv1_2.bias", "g2": "bn1_2.weight",
                 "be2": "bn1_2.bias", "w3": "conv1_3.weight", "b3": "conv1_3.bias", "g3": "bn1_3.weight", "be3": "bn1_3.bias",
                 "w4": "conv1_4.weight", "b4": "conv1_4.bias", "rm1": "bn1_1.running_mean", "rv1": "bn1_1.running_var",
                 "rm2": "bn1_2.running_mean", "rv2": "bn1_2.running_var", "rm3": "bn1_3.running_mean", "rv3": "bn1_3.running_var",
                 "nbt1": "bn1_1.num_batches_tracked", "nbt2": "bn1_2.num_batches_tracked", "nbt3": "bn1_3.num_batches_tracked"}
        for field_, nm in names.items():
            setattr(q, field_, p(pp + nm).data_ptr())
            self._ptr_names.append(pp + nm)
        self._ptr_names += [pp + "conv1_1.weight", pp + "conv1_1.bias"]
        self.pwf = q
        self._galloc("_pwf.dparams", (4697,))
        self.ops_fusion_fwd = [(lib.disco_conv_forward, self.en_call.desc, "conv[en]"),
                               (lib.disco_pwf_train_forward, q, "pwf_train_forward"),
                               (lib.disco_fusion_forward, f, "fusion")]
        # ---- backward ----
        srcs = gsrc.get(self.fused_key, [])
        if not srcs or len(srcs) > 3:
            raise RuntimeError("the fused map needs one to three gradient sources")
        ops_b = []
        for (t, ct, co, pl) in srcs:
            assert ct == cf and co == 0 and pl == 0
        if len(srcs) == 1:
            q.dfused = srcs[0][0].data_ptr()
        else:
            ops_b.append(("add_f32", self.dfused, srcs[0][0], srcs[1][0]))
            if len(srcs) == 3:
                ops_b.append(("add_f32", self.dfused, self.dfused, srcs[2][0]))
            q.dfused = self.dfused.data_ptr()
        ops_b.append(("py", self._fusion_zero))
        ops_b.append((lib.disco_fusion_combine_backward, q, "fusion_combine_backward"))
        ops_b.append((lib.disco_pwf_train_backward, q, "pwf_train_backward"))
        ops_b.append(("grad_pack_den", n * hf * wf))
        self.en_wg = wgrad_desc([feat], [0], (hf, wf), (hf, wf), 1, 1, self.den_act, 256, cf, "_dw.en")
        ops_b.append((lib.disco_conv_wgrad, self.en_wg, "wgrad[en]"))
        ops_b.append((lib.disco_conv_forward, self.en_dcall.desc, "dgrad[en]"))
        gsrc.setdefault(self.feat_key, []).append((self.en_gbuf, cf, 0, 0))
        gsrc[self.feat_key].append((self.dfeat, cf, 0, 0))
        self.ops_bwd_fusion = ops_b
        return ops_b

    def _fusion_zero(self):
        self.dfeat.zero_(); self.den.zero_()
        self._gview(self.G, "_pwf.dparams").zero_()

    # ------------------------------------------------------------------------------------------------
    def run_ops(self, op_list, stream):
        lib = self.lib
        for op in op_list:
            f = op[0]
            if f == "py":
                op[1]()
            elif f == "bev_pack":
                ops.bev_pack(self.bev_in, self.act["a0"], self.prec)
            elif f == "add_f32":
                check(lib.disco_add_f32(op[1].data_ptr(), op[2].data_ptr(), op[3].data_ptr(), op[1].numel(), stream), "add_f32")
            elif f == "grad_pack_heads":
                S = self.st["h1"]
                gb = self.G.data_ptr()
                check(lib.disco_grad_pack(self.gcls.data_ptr(), 12, self.gloc.data_ptr(), 36, op[1], S.dz.data_ptr(),
                                          ops._lo_off(S.dz), stream), "grad_pack[heads]")
                check(lib.disco_channel_sum(self.gcls.data_ptr(), op[1], 12, self.sums.data_ptr(), gb + 4 * self._head_bias_off[0],
                                            stream), "channel_sum")
                check(lib.disco_channel_sum(self.gloc.data_ptr(), op[1], 36, self.sums.data_ptr(), gb + 4 * self._head_bias_off[1],
                                            stream), "channel_sum")
            elif f == "call":
                check(op[1](*op[2], stream), op[3])
            elif f == "grad_pack_outc":
                S = self.st["outc"]
                check(lib.disco_grad_pack(self.glogits.data_ptr(), 8, self.gzero8.data_ptr(), 8, op[1], S.dz.data_ptr(),
                                          ops._lo_off(S.dz), stream), "grad_pack[outc]")
                check(lib.disco_channel_sum(self.glogits.data_ptr(), op[1], 8, self.sums.data_ptr(),
                                            self.G.data_ptr() + 4 * self._outc_bias_off, stream), "channel_sum")
            elif f == "grad_pack_den":
                check(lib.disco_grad_pack(self.den.data_ptr(), 256, None, 0, op[1], self.den_act.data_ptr(),
                                          ops._lo_off(self.den_act), stream), "grad_pack[den]")
            else:
                check(f(C.byref(op[1]), stream), op[2])

    def _run(self, which: str, op_list, stream):
        """Eager for the first two steps (lazy kernel attributes, allocator warm-up), then one CUDA-graph replay."""
        g = self._graphs[which]
        if USE_GRAPH and g is None and self._steps >= 2 and not torch.cuda.is_current_stream_capturing():
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    self.run_ops(op_list, torch.cuda.current_stream(self.dev).cuda_stream)
                g = self._graphs[which] = graph
            except Exception as e:      # capture unsupported in this context: stay eager (the reason is kept for diagnosis)
                g = self._graphs[which] = False
                self.graph_error = f"{which}: {type(e).__name__}: {e}"[:400]
                torch.cuda.synchronize()
        if g:
            g.replay()
        else:
            self.run_ops(op_list, stream)

    def forward(self, bevs: torch.Tensor, trans=None, num_agent=None, outage_host=None):
        """Runs the training-mode forward; returns {"cls", "loc"} (fp32 NHWC, fresh tensors) when the model has heads."""
        dev = self.dev
        if self._signature() != self._sig:
            self._build()                      # parameters were re-allocated (e.g. .to(), new load)
        stream = torch.cuda.current_stream(dev).cuda_stream
        self.generation += 1
        self.bev_in.copy_(bevs.detach().reshape(self.bev_in.shape), non_blocking=True)
        if self.fused_key:
            self.trans.copy_(trans.detach(), non_blocking=True)
            self.na.copy_(num_agent.detach()[:, 0], non_blocking=True)
            if outage_host is not None:
                self.outage.copy_(outage_host, non_blocking=True)
                self._outage_dirty = True
            elif getattr(self, "_outage_dirty", False):
                self.outage.zero_()
                self._outage_dirty = False
        self._run("fwd", self.fwd_ops, stream)
        out = {}
        if self.head_layers:
            out["cls"], out["loc"] = self.cls.clone(), self.loc.clone()
        if self.has_outc:
            lg = torch.empty((self.n, 8, self.h, self.w), dtype=torch.float32, device=dev)
            check(self.lib.disco_nhwc_to_nchw(self.logits.data_ptr(), self.n, self.h, self.w, 16, 8, lg.data_ptr(), stream), "nhwc_to_nchw")
            out["logits"] = lg
        return out

    def kd_map(self, key: str) -> torch.Tensor:
        return ops.act_to_nchw_f32(self.act[key], self.prec)

    # ------------------------------------------------------------------------------------------------
    def backward(self, grads: Dict[str, Optional[torch.Tensor]]) -> Dict[str, torch.Tensor]:
        """grads: {"cls": [n,h,w,12] fp32 NHWC, "loc": [n,h,w,36], <kd key>: NCHW fp32 or None}.
        Returns {parameter name: gradient}."""
        dev, lib = self.dev, self.lib
        stream = torch.cuda.current_stream(dev).cuda_stream
        keep = []
        for key, g in grads.items():
            if key in ("cls", "loc", "logits"):
                continue
            if key not in self.ext:
                if g is None:
                    continue
                raise KeyError(f"no static gradient buffer for {key!r} (kd_keys={self.kd_keys})")
            if g is None:
                if self.ext_dirty[key]:
                    self.ext[key].zero_()
                    self.ext_dirty[key] = False
                continue
            g = g.detach().float().contiguous()
            nn_, c, hh, ww = g.shape
            check(lib.disco_nchw_to_nhwc(g.data_ptr(), nn_, c, hh, ww, self.ext[key].data_ptr(), stream), "nchw_to_nhwc")
            self.ext_dirty[key] = True
            keep.append(g)
        if self.has_outc:
            g = grads.get("logits")
            if g is None:
                self.glogits.zero_()
            else:
                g = g.detach().float().contiguous()
                check(lib.disco_nchw_to_nhwc(g.data_ptr(), self.n, 8, self.h, self.w, self.glogits.data_ptr(), stream), "nchw_to_nhwc")
                keep.append(g)
        if self.head_layers:
            for name, buf in (("cls", self.gcls), ("loc", self.gloc)):
                g = grads.get(name)
                if g is None:
                    buf.zero_()
                else:
                    buf.copy_(g.detach().reshape(buf.shape), non_blocking=True)
        self._run("bwd", self.bwd_ops, stream)
        self._steps += 1
        self._bwd_keep = keep
        G = self.G.clone()
        if getattr(self, "grad_group", None) is not None:
            from . import parallel
            self.last_local_grad = G.clone() if getattr(self, "keep_local_grad", False) else None
            parallel.allreduce_mean_(G, self.grad_group)     # scene-sharded data parallelism: ONE flat all-reduce per step
        return self._collect(G)

    def _collect(self, G: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Views of the (cloned) flat gradient buffer under the reference's parameter names."""
        out: Dict[str, torch.Tensor] = {}
        v = lambda name: self._gview(G, name)

        class _Zeros:   # DISJOINT zero slices, one per bias (AccumulateGrad may steal the tensor: no two .grad may alias)
            def __init__(self_):
                self_.buf, self_.pos = torch.zeros(16384, dtype=G.dtype, device=G.device), 0

            def __getitem__(self_, sl):
                a = self_.pos
                self_.pos += sl.stop
                assert self_.pos <= self_.buf.numel()
                return self_.buf[a:self_.pos]
        zero = _Zeros()
        for L in self.enc + self.dec:
            if L.kind != "conv":
                continue
            if L.name == "outc":
                out["outc.conv.weight"] = v("_dw.outc")[:8].reshape(8, 64, 1, 1)
                out["outc.conv.bias"] = v("outc.conv.bias")
                continue
            out[L.conv + ".weight"] = v("_dw." + L.name).view(self.get(L.conv + ".weight").shape)
            out[L.conv + ".bias"] = zero[:L.c_out]     # conv bias in front of a BatchNorm: exactly zero gradient
            out[L.bn + ".weight"], out[L.bn + ".bias"] = v("_dgamma." + L.name), v("_dbeta." + L.name)
        if self.head_layers:
            dw3, dw1 = v("_dw.h3"), v("_dw.h1").view(48, 64)
            out["classification.conv1.weight"] = dw3[:32].view(32, 32, 3, 3)
            out["regression.box_prediction.0.weight"] = dw3[32:].view(32, 32, 3, 3)
            dg, db = v("_dgamma.h3"), v("_dbeta.h3")
            out["classification.bn1.weight"], out["regression.box_prediction.1.weight"] = dg[:32], dg[32:]
            out["classification.bn1.bias"], out["regression.box_prediction.1.bias"] = db[:32], db[32:]
            out["classification.conv1.bias"], out["regression.box_prediction.0.bias"] = zero[:32], zero[:32]
            out["classification.conv2.weight"] = dw1[:12, :32].reshape(12, 32, 1, 1)
            out["regression.box_prediction.3.weight"] = dw1[12:, 32:].reshape(36, 32, 1, 1)
            out["classification.conv2.bias"], out["regression.box_prediction.3.bias"] = v("classification.conv2.bias"), v("regression.box_prediction.3.bias")
        if self.fused_key:
            pp, dp, dw = self.pwf_prefix, v("_pwf.dparams"), v("_dw.en").view(256, self.fuse_c)
            out[pp + "conv1_1.weight"] = torch.cat((dw[:128], dw[128:]), 1).reshape(128, 2 * self.fuse_c, 1, 1)
            out[pp + "conv1_1.bias"] = zero[:128]
            out[pp + "bn1_1.weight"], out[pp + "bn1_1.bias"] = dp[0:128], dp[128:256]
            out[pp + "conv1_2.weight"] = dp[256:4352].view(32, 128, 1, 1)
            out[pp + "conv1_2.bias"] = zero[:32]
            out[pp + "bn1_2.weight"], out[pp + "bn1_2.bias"] = dp[4352:4384], dp[4384:4416]
            out[pp + "conv1_3.weight"] = dp[4416:4672].view(8, 32, 1, 1)
            out[pp + "conv1_3.bias"] = zero[:8]
            out[pp + "bn1_3.weight"], out[pp + "bn1_3.bias"] = dp[4672:4680], dp[4680:4688]
            out[pp + "conv1_4.weight"] = dp[4688:4696].view(1, 8, 1, 1)
            out[pp + "conv1_4.bias"] = dp[4696:4697]
        return out
