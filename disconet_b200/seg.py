"""Host-side mirror of `coperception.models.seg.DiscoNet` (BEV segmentation, SURVEY §8 row f1, BASELINE config 5).

Same constructor / forward signature / return structure / state_dict names as the reference
(models/seg/DiscoNet.py:8-20, FusionBase.py:9-84, SegModelBase.py:6-151): a U-Net (DoubleConv / Down / Up /
OutConv) whose 512-channel H/8 feature map goes through the same DiscoGraph fusion block as the detection model.
Eval-mode forward on the sm_100a kernels of libdisco_b200: the 18 3x3 convs + OutConv + PWF conv1_1 on the tcgen05
conv kernel (BatchNorm folded, concat [skip, up] as a two-source gather), MaxPool2d(2) / bilinear x2 upsample /
layout changes as streaming kernels, the fusion block as the one fused kernel (C = 512).  No CPU / torch fallback.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import engine, ops, train as train_mod
from ._lib import FusionDesc, check, load
from .modules import ParamHolder, PixelWeightedFusionParams
from .plan import fold_bn, pack_conv


class DoubleConvParams(ParamHolder):
    """SegModelBase.py:93-110: `double_conv.{0,1,3,4}`."""

    def __init__(self, cin, cout, mid=None):
        super().__init__()
        mid = mid or cout
        self.double_conv = nn.Sequential(
            nn.Conv2d(cin, mid, kernel_size=3, padding=1), nn.BatchNorm2d(mid), nn.ReLU(inplace=True),
            nn.Conv2d(mid, cout, kernel_size=3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class DownParams(ParamHolder):
    """SegModelBase.py:113-123: `maxpool_conv.1.double_conv.*`."""

    def __init__(self, cin, cout):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConvParams(cin, cout))


class UpParams(ParamHolder):
    """SegModelBase.py:126-142 (bilinear=True): `conv.double_conv.*`; `up` has no parameters."""

    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = DoubleConvParams(cin, cout, cin // 2)


class OutConvParams(ParamHolder):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size=1)


class SegDiscoNet(nn.Module):
    """BEV-segmentation DiscoNet (reference: models/seg/DiscoNet.py:8-126), B200-native eval forward."""

    def __init__(self, n_channels, n_classes, num_agent, kd_flag=True, compress_level=0, only_v2i=False,
                 precision: Optional[str] = None):
        super().__init__()
        if not 0 <= compress_level <= 9:
            raise ValueError("compress_level must be in [0, 9] (SegModelBase.py:30-31)")
        self.n_channels, self.n_classes, self.bilinear = n_channels, n_classes, True
        self.num_agent, self.only_v2i, self.kd_flag, self.compress_level = num_agent, only_v2i, kd_flag, compress_level
        # registration order = reference (SegModelBase.__init__, then DiscoNet.__init__)
        self.inc = DoubleConvParams(n_channels, 64)
        self.down1 = DownParams(64, 128)
        self.down2 = DownParams(128, 256)
        self.down3 = DownParams(256, 512)
        self.down4 = DownParams(512, 512)
        self.up1 = UpParams(1024, 256)
        self.up2 = UpParams(512, 128)
        self.up3 = UpParams(256, 64)
        self.up4 = UpParams(128, 64)
        self.outc = OutConvParams(64, n_classes)
        if compress_level > 0:     # communication bottleneck on the shared 512-channel map (SegModelBase.py:29-44, FusionBase.py:31-33)
            cc = 512 // (2 ** compress_level)
            self.com_compresser = nn.Conv2d(512, cc, kernel_size=1, stride=1)
            self.bn_compress = nn.BatchNorm2d(cc)
            self.com_decompresser = nn.Conv2d(cc, 512, kernel_size=1, stride=1)
            self.bn_decompress = nn.BatchNorm2d(512)
        self.pixel_weighted_fusion = PixelWeightedFusionParams(512)
        self.neighbor_feat_list = None
        self.tg_agent = None
        self.current_num_agent = None
        from .det import DEFAULT_PRECISION
        self.precision_name = precision or DEFAULT_PRECISION
        if self.precision_name not in engine.PRECISIONS:
            raise ValueError(f"precision must be one of {list(engine.PRECISIONS)}")
        self._plans = None
        self._plans_key = None
        self._ws: Dict[tuple, "_SegWorkspace"] = {}
        self._runners: Dict[tuple, train_mod.TrainRunner] = {}

    @property
    def precision(self) -> int:
        return engine.PRECISIONS[self.precision_name]

    def _getter(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        return sd.__getitem__

    def plans(self):
        ts = list(self.parameters()) + list(self.buffers())
        key = (self.precision_name, ts[0].device, tuple(t._version for t in ts), tuple(t.data_ptr() for t in ts[:4]),
               getattr(self, "_train_epoch", 0))   # training kernels write BN running stats through raw pointers
        if self._plans is None or key != self._plans_key:
            with torch.no_grad():
                self._plans = build_seg_plans(self._getter(), self.precision, self.n_channels, self.n_classes)
            self._plans_key = key
            self._ws.clear()
        return self._plans

    def forward(self, x, trans_matrices, num_agent_tensor):
        """x [A*B, n_channels, H, W] float (agent-major), trans_matrices [B, A, A, 4, 4], num_agent_tensor [B, A].
        Returns (logits, x9, x8, x7, x6, x5, feat_mat) if kd_flag else logits  (FusionBase.py:24-84)."""
        load()
        if self.training and self.precision_name != "bf16x3":
            raise NotImplementedError("training mode runs in the default bf16x3 precision only")
        if not x.is_cuda:
            raise ValueError("disconet_b200 runs on CUDA tensors only (no CPU fallback); got a CPU input")
        if x.dim() != 4 or x.shape[1] != self.n_channels:
            raise ValueError(f"x must be [N, {self.n_channels}, H, W] (got {tuple(x.shape)})")
        N, _, H, W = x.shape
        A = self.num_agent
        if N % A:
            raise ValueError(f"{N} rows are not a multiple of num_agent = {A}")
        B = N // A
        if tuple(trans_matrices.shape) != (B, A, A, 4, 4):
            raise ValueError(f"trans_matrices must be [{B},{A},{A},4,4] (got {tuple(trans_matrices.shape)})")
        dev = x.device
        if self.training:
            return self._forward_train(x, trans_matrices, num_agent_tensor, B)
        P = self.plans()
        key = (N, H, W, str(dev))
        ws = self._ws.get(key)
        if ws is None:
            ws = self._ws[key] = _SegWorkspace(N, H, W, A, B, self.precision, dev, P, self.n_channels, self.n_classes)
        stream = torch.cuda.current_stream(dev).cuda_stream
        ws.trans.copy_(trans_matrices.detach(), non_blocking=True)
        ws.na.copy_(num_agent_tensor.detach()[:, 0], non_blocking=True)
        ws.fusion.only_v2i = int(bool(self.only_v2i))
        xin = x.detach()
        if xin.dtype != torch.float32:
            xin = xin.float()
        ws.run(xin.contiguous(), stream)
        logits = ws.logits_nchw()
        if self.kd_flag:
            nchw = lambda k: ops.act_to_nchw_f32(ws.buf[k], self.precision)
            return logits, nchw("x9"), nchw("x8"), nchw("x7"), nchw("x6"), nchw("x5"), nchw("feat")
        return logits

    def _forward_train(self, x, trans_matrices, num_agent_tensor, B):
        """model.train() forward + autograd (SegModule.step, utils/SegModule.py:45-120, calls it like this)."""
        from .det import _TrainFn, runner_param_names
        if self.n_channels != 13 or self.n_classes != 8:
            raise NotImplementedError("segmentation training is built for the reference configuration (13 channels, 8 classes)")
        if self.compress_level > 0:
            raise NotImplementedError("segmentation training with compress_level > 0 is not built (eval mode is)")
        N, _, H, W = x.shape
        dev = x.device
        kd_keys = list(SEG_KD_KEYS) if self.kd_flag else []
        key = (N, H, W, str(dev), bool(self.only_v2i), tuple(kd_keys))
        runner = self._runners.get(key)
        if runner is None:
            runner = train_mod.TrainRunner(self._getter(), N, H, W, dev, "", "", heads=False, pwf_prefix="pixel_weighted_fusion.",
                                           batch_size=B, agents=self.num_agent, only_v2i=bool(self.only_v2i), kd_keys=kd_keys,
                                           nodes=seg_layers(), fusion_spec=("x4", "feat", 3, 512))
            self._runners[key] = runner
        runner.get = self._getter()
        live = runner_param_names(runner)
        named = dict(self.named_parameters())
        names = tuple(k for k in named if k in live)
        bev = x.detach().permute(0, 2, 3, 1).unsqueeze(1)          # [N, 1, H, W, 13], the layout the pack kernel reads
        self._train_epoch = getattr(self, "_train_epoch", 0) + 1
        outs = list(_TrainFn.apply(runner, bev, trans_matrices, num_agent_tensor, None, tuple(kd_keys), names,
                                   *[named[k] for k in names]))
        return (outs[0], *outs[1:]) if self.kd_flag else outs[0]


SEG_KD_KEYS = ("x9", "x8", "x7", "x6", "x5", "feat")


def seg_layers():
    """U-Net layer tables for the training driver (SegModelBase.py:17-26,93-151; FusionBase.py:24-84)."""
    L = train_mod.Layer

    def dc(name, prefix, srcs, mid, out, cin, cmid, cout, level, c_in_real=0):
        return [L(name + "a", f"{prefix}.0", f"{prefix}.1", srcs, [0] * len(srcs), mid, cin, cmid, level=level, c_in_real=c_in_real,
                  need_dgrad=srcs != ["a0"]),          # the network input needs no gradient
                L(name + "b", f"{prefix}.3", f"{prefix}.4", [mid], [0], out, [cmid], cout, level=level)]

    pool = lambda name, src, out, c, level: L(name, "", None, [src], [0], out, [c], c, level=level, kind="pool")
    up = lambda name, src, out, c, level: L(name, "", None, [src], [0], out, [c], c, level=level, kind="up")
    E = (dc("inc", "inc.double_conv", ["a0"], "x1a", "x1", [16], 64, 64, 0, c_in_real=13) + [pool("pool1", "x1", "p1", 64, 0)] +
         dc("d1", "down1.maxpool_conv.1.double_conv", ["p1"], "x2a", "x2", [64], 128, 128, 1) + [pool("pool2", "x2", "p2", 128, 1)] +
         dc("d2", "down2.maxpool_conv.1.double_conv", ["p2"], "x3a", "x3", [128], 256, 256, 2) + [pool("pool3", "x3", "p3", 256, 2)] +
         dc("d3", "down3.maxpool_conv.1.double_conv", ["p3"], "x4a", "x4", [256], 512, 512, 3))
    D = ([pool("pool4", "feat", "p4", 512, 3)] + dc("d4", "down4.maxpool_conv.1.double_conv", ["p4"], "x5a", "x5", [512], 512, 512, 4) +
         [up("up5", "x5", "u5", 512, 4)] + dc("u1", "up1.conv.double_conv", ["feat", "u5"], "x6a", "x6", [512, 512], 512, 256, 3) +
         [up("up6", "x6", "u6", 256, 3)] + dc("u2", "up2.conv.double_conv", ["x3", "u6"], "x7a", "x7", [256, 256], 256, 128, 2) +
         [up("up7", "x7", "u7", 128, 2)] + dc("u3", "up3.conv.double_conv", ["x2", "u7"], "x8a", "x8", [128, 128], 128, 64, 1) +
         [up("up8", "x8", "u8", 64, 1)] + dc("u4", "up4.conv.double_conv", ["x1", "u8"], "x9a", "x9", [64, 64], 64, 64, 0) +
         [L("outc", "outc.conv", None, ["x9"], [0], "", [64], 16, taps=1, level=0, n_real=8)])
    return E, D


def build_seg_plans(get, precision: int, n_channels: int, n_classes: int):
    def cbr(prefix, idx_conv, idx_bn, srcs):
        w, b = fold_bn(get(f"{prefix}.{idx_conv}.weight"), get(f"{prefix}.{idx_conv}.bias"), get(f"{prefix}.{idx_bn}.weight"),
                       get(f"{prefix}.{idx_bn}.bias"), get(f"{prefix}.{idx_bn}.running_mean"), get(f"{prefix}.{idx_bn}.running_var"))
        c_pad = sum(srcs)
        if w.shape[1] != c_pad:
            wp = torch.zeros(w.shape[0], c_pad, 3, 3, device=w.device)
            wp[:, :w.shape[1]] = w
            w = wp
        return pack_conv(w, b, src_channels=srcs, relu=True, precision=precision, name=f"{prefix}.{idx_conv}")

    cin_pad = (n_channels + 15) // 16 * 16
    P = {
        "inc": (cbr("inc.double_conv", 0, 1, [cin_pad]), cbr("inc.double_conv", 3, 4, [64])),
        "down1": (cbr("down1.maxpool_conv.1.double_conv", 0, 1, [64]), cbr("down1.maxpool_conv.1.double_conv", 3, 4, [128])),
        "down2": (cbr("down2.maxpool_conv.1.double_conv", 0, 1, [128]), cbr("down2.maxpool_conv.1.double_conv", 3, 4, [256])),
        "down3": (cbr("down3.maxpool_conv.1.double_conv", 0, 1, [256]), cbr("down3.maxpool_conv.1.double_conv", 3, 4, [512])),
        "down4": (cbr("down4.maxpool_conv.1.double_conv", 0, 1, [512]), cbr("down4.maxpool_conv.1.double_conv", 3, 4, [512])),
        # torch.cat([x2 (skip), x1 (upsampled)], dim=1): skip channels first (SegModelBase.py:141)
        "up1": (cbr("up1.conv.double_conv", 0, 1, [512, 512]), cbr("up1.conv.double_conv", 3, 4, [512])),
        "up2": (cbr("up2.conv.double_conv", 0, 1, [256, 256]), cbr("up2.conv.double_conv", 3, 4, [256])),
        "up3": (cbr("up3.conv.double_conv", 0, 1, [128, 128]), cbr("up3.conv.double_conv", 3, 4, [128])),
        "up4": (cbr("up4.conv.double_conv", 0, 1, [64, 64]), cbr("up4.conv.double_conv", 3, 4, [64])),
    }
    wo, bo = get("outc.conv.weight").detach().float(), get("outc.conv.bias").detach().float()
    nc4 = (n_classes + 3) // 4 * 4
    wpad = torch.zeros(nc4, 64, 1, 1, device=wo.device)
    wpad[:n_classes] = wo
    bpad = torch.zeros(nc4, device=wo.device)
    bpad[:n_classes] = bo
    P["outc"] = pack_conv(wpad, bpad, src_channels=[64], relu=False, precision=precision, name="outc.conv")
    try:
        get("com_compresser.weight")
        has_comp = True
    except KeyError:
        has_comp = False
    if has_comp:   # relu(bn(1x1 512 -> cc)), relu(bn(1x1 cc -> 512)); cc padded to the 16-channel lane width
        w, b = fold_bn(get("com_compresser.weight"), get("com_compresser.bias"), get("bn_compress.weight"), get("bn_compress.bias"),
                       get("bn_compress.running_mean"), get("bn_compress.running_var"))
        cc = w.shape[0]
        cc_pad = (cc + 15) // 16 * 16
        wp = torch.zeros(cc_pad, 512, 1, 1, device=w.device); wp[:cc] = w
        bp = torch.zeros(cc_pad, device=w.device); bp[:cc] = b
        P["compress"] = pack_conv(wp, bp, src_channels=[512], relu=True, precision=precision, name="com_compresser")
        w, b = fold_bn(get("com_decompresser.weight"), get("com_decompresser.bias"), get("bn_decompress.weight"), get("bn_decompress.bias"),
                       get("bn_decompress.running_mean"), get("bn_decompress.running_var"))
        wp = torch.zeros(512, cc_pad, 1, 1, device=w.device); wp[:, :cc] = w
        P["decompress"] = pack_conv(wp, b, src_channels=[cc_pad], relu=True, precision=precision, name="com_decompresser")
    P["pwf"] = engine.build_pwf_plans(get, precision)
    P["cin_pad"], P["nc4"] = cin_pad, nc4
    return P


class _SegWorkspace:
    """NHWC activation buffers + prebuilt launches of the U-Net for one (N, H, W)."""

    def __init__(self, n, h, w, agents, batch, precision, device, P, n_channels, n_classes):
        if h % 16 or w % 16:
            raise ValueError(f"BEV size {h}x{w} must be a multiple of 16 (4 pooling stages)")
        self.n, self.h, self.w, self.precision, self.dev = n, h, w, precision, device
        self.n_channels, self.n_classes = n_channels, n_classes
        self.lib = load()
        A_ = lambda lvl, c: ops.alloc_act(n, h >> lvl, w >> lvl, c, precision, device)
        b = self.buf = {
            "a0": A_(0, 16), "x1a": A_(0, 64), "x1": A_(0, 64), "p1": A_(1, 64),
            "x2a": A_(1, 128), "x2": A_(1, 128), "p2": A_(2, 128),
            "x3a": A_(2, 256), "x3": A_(2, 256), "p3": A_(3, 256),
            "x4a": A_(3, 512), "x4": A_(3, 512), "feat": A_(3, 512), "p4": A_(4, 512),
            "x5a": A_(4, 512), "x5": A_(4, 512), "u5": A_(3, 512),
            "x6a": A_(3, 512), "x6": A_(3, 256), "u6": A_(2, 256),
            "x7a": A_(2, 256), "x7": A_(2, 128), "u7": A_(1, 128),
            "x8a": A_(1, 128), "x8": A_(1, 64), "u8": A_(0, 64),
            "x9a": A_(0, 64), "x9": A_(0, 64),
        }
        feat_src = "x4"
        if "compress" in P:
            b["x4c"] = A_(3, P["compress"].c_out)
            b["x4d"] = A_(3, 512)
            feat_src = "x4d"
        if P["cin_pad"] != 16:
            raise NotImplementedError("segmentation input with more than 16 channels")
        self.x_nhwc = torch.empty((n, h, w, n_channels), dtype=torch.float32, device=device)
        self.logits_nhwc = torch.empty((n, h, w, P["nc4"]), dtype=torch.float32, device=device)
        self.nc4 = P["nc4"]

        def conv(plan, srcs, out, lvl):
            return ("conv", ops.ConvCall(plan, [b[s] for s in srcs], [0] * len(srcs), b[out] if isinstance(out, str) else out,
                                         n=n, h_in=h >> lvl, w_in=w >> lvl))

        def block(name, srcs, mid, out, lvl):
            return [conv(P[name][0], srcs, mid, lvl), conv(P[name][1], [mid], out, lvl)]

        pool = lambda s, d, lvl: ("pool", s, d, lvl)
        up = lambda s, d, lvl: ("up", s, d, lvl)        # lvl = level of the SOURCE
        hf, wf = h >> 3, w >> 3
        self.en = torch.empty((n, hf, wf, 256), dtype=torch.float32, device=device)
        self.trans = torch.zeros((batch, agents, agents, 4, 4), dtype=torch.float64, device=device)
        self.na = torch.zeros((batch,), dtype=torch.int32, device=device)
        pwf = P["pwf"]
        f = FusionDesc()
        f.feat_hi, f.feat_lo_off, f.precision = b[feat_src].data_ptr(), ops._lo_off(b[feat_src]), precision
        f.en, f.hid = self.en.data_ptr(), 128
        t = pwf["tail"]
        f.w2, f.b2, f.w3, f.b3, f.w4, f.b4 = (x.data_ptr() for x in t)
        f.trans, f.num_agent = self.trans.data_ptr(), self.na.data_ptr()
        f.B, f.A, f.h, f.w, f.C = batch, agents, hf, wf, 512
        f.trans_scale = 4.0 / 128.0
        f.out_hi, f.out_lo_off = b["feat"].data_ptr(), ops._lo_off(b["feat"])
        f.row_begin, f.row_end = 0, n
        self.fusion, self._keep = f, pwf
        self.steps = (
            block("inc", ["a0"], "x1a", "x1", 0) + [pool("x1", "p1", 0)] +
            block("down1", ["p1"], "x2a", "x2", 1) + [pool("x2", "p2", 1)] +
            block("down2", ["p2"], "x3a", "x3", 2) + [pool("x3", "p3", 2)] +
            block("down3", ["p3"], "x4a", "x4", 3) +
            ([conv(P["compress"], ["x4"], "x4c", 3), conv(P["decompress"], ["x4c"], "x4d", 3)] if "compress" in P else []) +
            [("conv", ops.ConvCall(pwf["en"], [b[feat_src]], [0], (self.en,), n=n, h_in=hf, w_in=wf)), ("fusion",),
             pool("feat", "p4", 3)] +
            block("down4", ["p4"], "x5a", "x5", 4) + [up("x5", "u5", 4)] +
            block("up1", ["feat", "u5"], "x6a", "x6", 3) + [up("x6", "u6", 3)] +
            block("up2", ["x3", "u6"], "x7a", "x7", 2) + [up("x7", "u7", 2)] +
            block("up3", ["x2", "u7"], "x8a", "x8", 1) + [up("x8", "u8", 1)] +
            block("up4", ["x1", "u8"], "x9a", "x9", 0) +
            [("conv", ops.ConvCall(P["outc"], [b["x9"]], [0], (self.logits_nhwc,), n=n, h_in=h, w_in=w))]
        )
        self.flops = sum(s[1].flops for s in self.steps if s[0] == "conv")

    def run(self, x_nchw: torch.Tensor, stream):
        n, h, w, p, lib, b = self.n, self.h, self.w, self.precision, self.lib, self.buf
        check(lib.disco_nchw_to_nhwc(x_nchw.data_ptr(), n, self.n_channels, h, w, self.x_nhwc.data_ptr(), stream), "nchw_to_nhwc")
        ops.bev_pack(self.x_nhwc, b["a0"], p)
        for s in self.steps:
            if s[0] == "conv":
                s[1].launch(stream)
            elif s[0] == "fusion":
                ops.fusion_forward(self.fusion, stream)
            else:
                _, src, dst, lvl = s
                S, D = b[src], b[dst]
                fn = lib.disco_maxpool2 if s[0] == "pool" else lib.disco_upsample_bilinear2x
                check(fn(S.data_ptr(), ops._lo_off(S), D.data_ptr(), ops._lo_off(D), p, n, h >> lvl, w >> lvl, S.shape[-1], stream),
                      s[0])

    def logits_nchw(self) -> torch.Tensor:
        n, h, w = self.n, self.h, self.w
        out = torch.empty((n, self.n_classes, h, w), dtype=torch.float32, device=self.dev)
        check(self.lib.disco_nhwc_to_nchw(self.logits_nhwc.data_ptr(), n, h, w, self.nc4, self.n_classes, out.data_ptr(),
                                          torch.cuda.current_stream(self.dev).cuda_stream), "nhwc_to_nchw")
        return out
