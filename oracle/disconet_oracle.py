"""CPU oracle for the DiscoNet collaborative-perception forward path (TEST INFRASTRUCTURE).

A functional, state_dict-driven restatement in plain PyTorch fp32 (CPU) of what the reference
computes on the hot path.  It is *not* the product and is never imported by ``disconet_b200``.

Pinned against the live reference (imported from /root/reference in the build container by
``oracle/ref_import.py``) by ``oracle/make_golden.py`` -> ``tests/golden/*.npz``; the reference's own
tests hold no golden vectors for this path, so those fixtures are the pin.

Reference lines followed (R = /root/reference/coperception/coperception):
  encode          R/models/det/backbone/Backbone.py:89-143   (Conv3D 1x1x1: :280-300)
  decode          R/models/det/backbone/Backbone.py:145-242
  regroup / flip  R/models/det/base/DetModelBase.py:53-127
  affine warp     R/models/det/base/DetModelBase.py:139-169
  neighbour list  R/models/det/base/DetModelBase.py:171-209
  fusion loop     R/models/det/DiscoNet.py:28-129
  PWF MLP         R/models/det/DiscoNet.py:132-155
  heads           R/models/det/base/DetModelBase.py:226-351
  teacher / FaF   R/models/det/backbone/Backbone.py:245-257, R/models/det/FaFNet.py:28-39
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm default eps


class training:
    """Context manager: inside it `_bn` uses batch statistics (nn.BatchNorm train mode: biased variance to
    normalise, momentum-0.1 update of running_mean / unbiased running_var, num_batches_tracked += 1 per call --
    torch/nn/modules/batchnorm.py semantics the reference relies on) and records the updated buffers in
    `self.buffers` (the input state dict is left untouched).  Autograd flows through every tensor of `sd`
    that requires grad, so the oracle's backward is torch.autograd over this restatement."""
    active = None

    def __init__(self, sd):
        self.buffers = {k: v.clone() for k, v in sd.items()
                        if k.endswith(("running_mean", "running_var", "num_batches_tracked"))}

    def __enter__(self):
        training.active = self
        return self

    def __exit__(self, *exc):
        training.active = None


def _bn(x, sd, name):
    """Batch norm (BatchNorm2d/3d, eps 1e-5): running statistics in eval mode, batch statistics inside a
    `training` context."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    g = sd[name + ".weight"].view(shape)
    b = sd[name + ".bias"].view(shape)
    ctx = training.active
    if ctx is None:
        mean = sd[name + ".running_mean"].view(shape)
        var = sd[name + ".running_var"].view(shape)
        return (x - mean) / torch.sqrt(var + EPS) * g + b
    dims = [0] + list(range(2, x.dim()))
    m = x.numel() // x.shape[1]
    mean = x.mean(dims)
    var = x.var(dims, unbiased=False)
    with torch.no_grad():
        buf = ctx.buffers
        buf[name + ".running_mean"] = 0.9 * buf[name + ".running_mean"] + 0.1 * mean
        buf[name + ".running_var"] = 0.9 * buf[name + ".running_var"] + 0.1 * var * (m / max(m - 1, 1))
        buf[name + ".num_batches_tracked"] = buf[name + ".num_batches_tracked"] + 1
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + EPS) * g + b


def _cbr(x, sd, conv, bn, stride=1):
    w = sd[conv + ".weight"]
    pad = w.shape[-1] // 2
    y = F.conv2d(x, w, sd[conv + ".bias"], stride=stride, padding=pad)
    return F.relu(_bn(y, sd, bn))


def _pointwise3d(x, sd, name):
    """Conv3D(k=1x1x1)+BN3d+ReLU on a seq-1 tensor == per-pixel linear map (Backbone.py:280-300)."""
    w = sd[name + ".conv3d.weight"]
    y = F.conv2d(x, w.view(w.shape[0], w.shape[1], 1, 1), sd[name + ".conv3d.bias"])
    return F.relu(_bn(y, sd, name + ".bn3d"))


def encode(sd, pfx, bev_nchw):
    """bev_nchw [N,13,H,W] float -> [x, x_1, x_2, x_3, x_4]  (Backbone.py:89-143, seq == 1)."""
    p = pfx
    if bev_nchw.dtype != torch.float64:   # (float64 inputs: the tests' high-precision "truth" run of this oracle)
        bev_nchw = bev_nchw.float()       # Backbone.py:101
    x = _cbr(bev_nchw, sd, p + "conv_pre_1", p + "bn_pre_1")
    x = _cbr(x, sd, p + "conv_pre_2", p + "bn_pre_2")
    x1 = _cbr(x, sd, p + "conv1_1", p + "bn1_1", stride=2)
    x1 = _cbr(x1, sd, p + "conv1_2", p + "bn1_2")
    x1 = _pointwise3d(x1, sd, p + "conv3d_1")
    x2 = _cbr(x1, sd, p + "conv2_1", p + "bn2_1", stride=2)
    x2 = _cbr(x2, sd, p + "conv2_2", p + "bn2_2")
    x2 = _pointwise3d(x2, sd, p + "conv3d_2")
    x3 = _cbr(x2, sd, p + "conv3_1", p + "bn3_1", stride=2)
    x3 = _cbr(x3, sd, p + "conv3_2", p + "bn3_2")
    x4 = _cbr(x3, sd, p + "conv4_1", p + "bn4_1", stride=2)
    x4 = _cbr(x4, sd, p + "conv4_2", p + "bn4_2")
    if (p + "com_compresser.weight") in sd:
        x3 = _cbr(x3, sd, p + "com_compresser", p + "bn_compress")
        x3 = _cbr(x3, sd, p + "com_decompresser", p + "bn_decompress")
    return [x, x1, x2, x3, x4]


def _up_cat(lo, skip):
    up = F.interpolate(lo, scale_factor=(2, 2))  # default mode: nearest
    return torch.cat((up, skip), dim=1)


def decode(sd, pfx, x, x1, x2, x3, x4):
    """-> (x_8, x_7, x_6, x_5)  (Backbone.py:145-242; seq==1 so the max-pool over seq is identity)."""
    p = pfx
    x5 = _cbr(_up_cat(x4, x3), sd, p + "conv5_1", p + "bn5_1")
    x5 = _cbr(x5, sd, p + "conv5_2", p + "bn5_2")
    x6 = _cbr(_up_cat(x5, x2), sd, p + "conv6_1", p + "bn6_1")
    x6 = _cbr(x6, sd, p + "conv6_2", p + "bn6_2")
    x7 = _cbr(_up_cat(x6, x1), sd, p + "conv7_1", p + "bn7_1")
    x7 = _cbr(x7, sd, p + "conv7_2", p + "bn7_2")
    x8 = _cbr(_up_cat(x7, x), sd, p + "conv8_1", p + "bn8_1")
    x8 = _cbr(x8, sd, p + "conv8_2", p + "bn8_2")
    return x8, x7, x6, x5


def pwf(sd, cat_feat, pfx="pixel_weighted_fusion."):
    """PixelWeightedFusionSoftmax on [1,2C,h,w] -> [1,1,h,w]  (DiscoNet.py:148-155)."""
    y = _cbr(cat_feat, sd, pfx + "conv1_1", pfx + "bn1_1")
    y = _cbr(y, sd, pfx + "conv1_2", pfx + "bn1_2")
    y = _cbr(y, sd, pfx + "conv1_3", pfx + "bn1_3")
    return F.relu(F.conv2d(y, sd[pfx + "conv1_4.weight"], sd[pfx + "conv1_4.bias"]))


def warp_to_ego(nb_feat, tfm_ji):
    """Neighbour map [C,h,w] (already H-flipped) -> ego frame (DetModelBase.py:158-168).

    tfm_ji = trans_matrices[b, j, i] (4x4).  theta = [R_2x2 | -t_xy * 4/128]; bilinear, zeros padding,
    align_corners=False (the torch>=1.3 default the reference relies on).
    """
    m = torch.hstack((tfm_ji[:2, :2], -tfm_ji[:2, 3:4])).float().unsqueeze(0)   # fp32 cast as in the reference
    m = (m * torch.tensor([[[1, 1, 4 / 128], [1, 1, 4 / 128]]])).to(nb_feat.dtype)
    c, h, w = nb_feat.shape
    grid = F.affine_grid(m, size=[1, c, h, w], align_corners=False)
    return F.grid_sample(nb_feat.unsqueeze(0), grid, mode="bilinear", padding_mode="zeros",
                         align_corners=False)[0]


def fuse(sd, x3, trans_matrices, num_agent_tensor, batch_size, agent_num, only_v2i=False, outage=None):
    """DiscoGraph fusion of the collaboration layer.

    x3 [A*B,C,h,w] agent-major (row a*B+b) -> (fused [A*B,C,h,w], weights[b][i] = list of [h,w]).
    Follows DetModelBase.py:80-127 (flip, regroup), DiscoNet.py:59-113 (ego loop, exp/sum softmax,
    weighted sum), DetModelBase.py:53-69 (regroup back, flip back).  p_com_outage == 0.
    """
    B, A = batch_size, agent_num
    feat = torch.flip(x3, (2,))
    C, h, w = feat.shape[1:]
    com = torch.stack([feat[B * a: B * (a + 1)] for a in range(A)], dim=1)  # [B,A,C,h,w]
    out = com.clone()
    all_w = []
    for b in range(B):
        n_ag = int(num_agent_tensor[b, 0])
        per_b = []
        for i in range(n_ag):
            ego = com[b, i]
            nbs = [ego]
            if outage is not None and bool(outage[b][i]):   # DiscoNet.py:68-69: ego keeps its own features
                per_b.append([torch.ones(h, w, dtype=feat.dtype)])
                continue
            for j in range(n_ag):
                if j == i:
                    continue
                if only_v2i and i != 0 and j != 0:
                    continue
                nbs.append(warp_to_ego(com[b, j], trans_matrices[b, j, i]))
            e = [torch.exp(pwf(sd, torch.cat([ego, nb], 0).unsqueeze(0))[0, 0]) for nb in nbs]
            tot = sum(e)
            ws = [ek / tot for ek in e]
            out[b, i] = sum(wk.unsqueeze(0) * nb for wk, nb in zip(ws, nbs))
            per_b.append(ws)
        all_w.append(per_b)
    fused = torch.cat([out[:, a] for a in range(A)], 0)
    return torch.flip(fused, (2,)), all_w


def heads(sd, x8, category_num=2, anchors=6, box_code=6):
    """-> cls [N, H*W*anchors, category_num], loc [N,H,W,anchors,1,box_code] (DetModelBase.py:226-265)."""
    c = _cbr(x8, sd, "classification.conv1", "classification.bn1")
    c = F.conv2d(c, sd["classification.conv2.weight"], sd["classification.conv2.bias"])
    r = _cbr(x8, sd, "regression.box_prediction.0", "regression.box_prediction.1")
    r = F.conv2d(r, sd["regression.box_prediction.3.weight"], sd["regression.box_prediction.3.bias"])
    n, _, H, W = x8.shape
    cls = c.permute(0, 2, 3, 1).contiguous().view(n, -1, category_num)
    loc = r.permute(0, 2, 3, 1).contiguous().view(n, H, W, anchors, 1, box_code)
    return cls, loc


def disconet_forward_graph(sd, bevs, trans_matrices, num_agent_tensor, batch_size, agent_num=5,
                     layer=3, only_v2i=False, return_all=False, outage=None):
    """DiscoNet.forward (DiscoNet.py:28-129), eval mode unless called inside `training(sd)`; builds an autograd
    graph when tensors of `sd` require grad.  bevs [A*B,1,H,W,13]."""
    if layer not in (2, 3):
        raise NotImplementedError("the reference builds a PixelWeightedFusion for layer 2 or 3 only")
    bev = bevs.permute(0, 1, 4, 2, 3)
    bev = bev.reshape(-1, bev.shape[2], bev.shape[3], bev.shape[4])
    x, x1, x2, x3, x4 = encode(sd, "u_encoder.", bev)
    enc = [x, x1, x2, x3, x4]
    fused, weights = fuse(sd, enc[layer], trans_matrices, num_agent_tensor, batch_size, agent_num, only_v2i, outage)
    enc[layer] = fused                         # DetModelBase.py:222 (get_decoded_layers)
    x8, x7, x6, x5 = decode(sd, "decoder.", *enc)
    cls, loc = heads(sd, x8)
    out = {"cls": cls, "loc": loc}
    if return_all:
        out.update(x=x, x_1=x1, x_2=x2, x_3=x3, x_4=x4, fused=fused, x_5=x5, x_6=x6, x_7=x7, x_8=x8,
                   weights=weights)
    return out


@torch.no_grad()
def disconet_forward(*args, **kwargs):
    """Eval-mode forward without autograd (the parity tests' default entry)."""
    return disconet_forward_graph(*args, **kwargs)


def fafnet_forward_graph(sd, bevs, pfx="stpn."):
    """FaFNet (no fusion; BASELINE config 1): FaFNet.py:28-39 + STPN_KD (Backbone.py:245-257)."""
    bev = bevs.permute(0, 1, 4, 2, 3)
    bev = bev.reshape(-1, bev.shape[2], bev.shape[3], bev.shape[4])
    x, x1, x2, x3, x4 = encode(sd, pfx, bev)
    x8, x7, x6, x5 = decode(sd, pfx, x, x1, x2, x3, x4)
    cls, loc = heads(sd, x8)
    return {"cls": cls, "loc": loc, "x_8": x8, "x_7": x7, "x_6": x6, "x_5": x5, "x_3": x3, "x_4": x4}


@torch.no_grad()
def fafnet_forward(sd, bevs, pfx="stpn."):
    return fafnet_forward_graph(sd, bevs, pfx)


def probe_loss(outputs: dict, seed: int = 0):
    """Scalar test loss  sum_k <out_k, R_k>  with numpy-seeded fixed cotangents R_k = |N(0,1)| / sqrt(numel_k).

    Every output named in `outputs` receives a dense cotangent, so one backward exercises all gradient paths of
    the hot path (heads, decoder, fusion, encoder).  The cotangents are kept NON-NEGATIVE on purpose: with
    random-sign cotangents every parameter gradient is a sum of ~1e5 cancelling terms, which amplifies the fp32
    reduction noise of the reference's own BatchNorm kernels (its outputs move by 3e-4 and such gradients by
    2-8% between 1 and 8 CPU threads) far beyond any meaningful parity tolerance.
    Returns (loss, {name: R_k})."""
    import numpy as np
    rng = np.random.default_rng(seed)
    loss = 0.0
    cot = {}
    for k in sorted(outputs):
        t = outputs[k]
        r = torch.from_numpy(np.abs(rng.standard_normal(tuple(t.shape))).astype(np.float32)) / float(np.sqrt(t.numel()))
        cot[k] = r
        loss = loss + (t * r.to(t.device)).sum()
    return loss, cot


from disconet_b200.synth import synth_state_dict, synth_poses, synth_bev  # noqa: E402,F401  (shared seeded generators)
