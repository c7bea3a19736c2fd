"""CPU oracle for the BEV-segmentation DiscoNet forward (BASELINE config 5) -- TEST INFRASTRUCTURE.

Functional fp32 PyTorch restatement, pinned against the live reference by oracle/make_golden.py
(tests/golden/seg_*.npz).  Reference lines followed (R = /root/reference/coperception/coperception/models/seg):
  U-Net blocks      R/SegModelBase.py:90-151  (DoubleConv :93-110, Down :113-123 MaxPool2d(2), Up :126-142 bilinear x2
                    align_corners=True + pad + cat([skip, up]), OutConv :145-151)
  encoder / decoder R/FusionBase.py:24-84     (inc, down1..3 -> fusion on x4 [512 ch @ H/8] -> down4, up1..4, outc)
  regroup / flips   R/SegModelBase.py:44-85   (flip H, agent-major <-> [B, A, ...])
  affine warp       R/SegModelBase.py:59-75   (same theta = [R | -t * 4/128] as the detection model)
  fusion            R/DiscoNet.py:77-101      (PWF per neighbour, exp / sum softmax, weighted sum)
  PWF MLP           R/DiscoNet.py:104-126     (2*512 -> 128 -> 32 -> 8 -> 1)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle.disconet_oracle import _bn, pwf, warp_to_ego


def _double_conv(x, sd, p):
    """conv3x3 + BN + ReLU twice; p = '<block>.double_conv.' (SegModelBase.py:93-110)."""
    for c, b in (("0", "1"), ("3", "4")):
        x = F.relu(_bn(F.conv2d(x, sd[p + c + ".weight"], sd[p + c + ".bias"], padding=1), sd, p + b))
    return x


def _down(x, sd, name):
    return _double_conv(F.max_pool2d(x, 2), sd, name + ".maxpool_conv.1.double_conv.")


def _up(x1, x2, sd, name):
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(torch.cat([x2, x1], 1), sd, name + ".conv.double_conv.")


def fuse(sd, x4, trans_matrices, num_agent_tensor, batch_size, agent_num, only_v2i=False):
    """FusionBase.forward:38-70 + DiscoNet.fusion:77-101.  x4 [A*B, C, h, w] agent-major."""
    B, A = batch_size, agent_num
    feat = torch.flip(x4, (2,))
    com = torch.stack([feat[B * a: B * (a + 1)] for a in range(A)], dim=1)
    out = com.clone()
    for b in range(B):
        n_ag = int(num_agent_tensor[b, 0])
        for i in range(n_ag):
            ego = com[b, i]
            nbs = [ego]
            for j in range(n_ag):
                if j == i or (only_v2i and i != 0 and j != 0):
                    continue
                nbs.append(warp_to_ego(com[b, j], trans_matrices[b, j, i]))
            e = [torch.exp(pwf(sd, torch.cat([ego, nb], 0).unsqueeze(0))[0, 0]) for nb in nbs]
            tot = sum(e)
            out[b, i] = sum((ek / tot).unsqueeze(0) * nb for ek, nb in zip(e, nbs))
    fused = torch.cat([out[:, a] for a in range(A)], 0)
    return torch.flip(fused, (2,))


def seg_disconet_forward_graph(sd, x, trans_matrices, num_agent_tensor, agent_num=5, only_v2i=False, return_all=False):
    """seg DiscoNet forward (eval mode unless inside `disconet_oracle.training(sd)`; autograd flows through `sd`).
    x [A*B, 13, H, W] float (agent-major) -> logits [A*B, n_classes, H, W]."""
    if x.dtype != torch.float64:
        x = x.float()
    x1 = _double_conv(x, sd, "inc.double_conv.")
    x2 = _down(x1, sd, "down1")
    x3 = _down(x2, sd, "down2")
    x4 = _down(x3, sd, "down3")
    if "com_compresser.weight" in sd:      # compress_level > 0: 1x1 bottleneck on the shared map (FusionBase.py:31-33)
        x4 = torch.relu(_bn(F.conv2d(x4, sd["com_compresser.weight"], sd["com_compresser.bias"]), sd, "bn_compress"))
        x4 = torch.relu(_bn(F.conv2d(x4, sd["com_decompresser.weight"], sd["com_decompresser.bias"]), sd, "bn_decompress"))
    B = x.shape[0] // agent_num
    feat = fuse(sd, x4, trans_matrices, num_agent_tensor, B, agent_num, only_v2i)
    x5 = _down(feat, sd, "down4")
    x6 = _up(x5, feat, sd, "up1")
    x7 = _up(x6, x3, sd, "up2")
    x8 = _up(x7, x2, sd, "up3")
    x9 = _up(x8, x1, sd, "up4")
    logits = F.conv2d(x9, sd["outc.conv.weight"], sd["outc.conv.bias"])
    if return_all:
        return {"logits": logits, "x9": x9, "x8": x8, "x7": x7, "x6": x6, "x5": x5, "feat": feat, "x4": x4}
    return {"logits": logits}


@torch.no_grad()
def seg_disconet_forward(*args, **kwargs):
    """Eval-mode forward without autograd."""
    return seg_disconet_forward_graph(*args, **kwargs)
