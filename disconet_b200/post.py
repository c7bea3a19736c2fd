"""Device half of the detection post-processing.

`apply_nms_det` (coperception/utils/detection_util.py:256-373) softmaxes the class logits, decodes every anchor's box,
copies ALL scores and boxes to the host, builds corners in numpy and only then keeps `scores > 0.7`
(`non_max_suppression`, utils/postprocess.py:72-115) for the sequential shapely polygon NMS.  `det_candidates` does the
per-anchor part in one kernel and returns only the survivors, highest score first -- the exact input of that NMS loop.
"""
from __future__ import annotations

from typing import List

import torch

from ._lib import check, load


def det_candidates(loc: torch.Tensor, cls: torch.Tensor, anchors: torch.Tensor, score_thresh: float = 0.7,
                   max_candidates: int = 8192) -> List[dict]:
    """loc [N, H, W, A, 1, 6], cls [N, H*W*A, 2] (the model's `result`), anchors [N, H, W, A, 6] or [H, W, A, 6].

    Returns one dict per agent: corners [K, 4, 2] fp32, score [K] fp32, index [K] int64 (anchor number, as
    `selected_idx` of the reference indexes them), sorted by descending score."""
    if not (loc.is_cuda and cls.is_cuda and anchors.is_cuda):
        raise ValueError("disconet_b200.post runs on CUDA tensors only (no CPU fallback)")
    n = cls.shape[0]
    per = cls.shape[1]
    if cls.shape[-1] != 2 or loc.numel() != n * per * 6:
        raise ValueError(f"expected binary cls [N, anchors, 2] and loc with 6 codes per anchor (got {tuple(cls.shape)}, {tuple(loc.shape)})")
    loc_c, cls_c, anc = loc.detach().float().contiguous(), cls.detach().float().contiguous(), anchors.detach().float().contiguous()
    if anc.numel() == per * 6:
        stride = 0
    elif anc.numel() == n * per * 6:
        stride = per * 6
    else:
        raise ValueError(f"anchors {tuple(anchors.shape)} do not match {per} anchors per agent")
    dev = cls.device
    count = torch.zeros(n, dtype=torch.int32, device=dev)
    corners = torch.empty((n, max_candidates, 4, 2), dtype=torch.float32, device=dev)
    scores = torch.empty((n, max_candidates), dtype=torch.float32, device=dev)
    index = torch.empty((n, max_candidates), dtype=torch.int32, device=dev)
    check(load().disco_det_candidates(loc_c.data_ptr(), cls_c.data_ptr(), anc.data_ptr(), per, stride, n, float(score_thresh),
                                      max_candidates, count.data_ptr(), corners.data_ptr(), scores.data_ptr(), index.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream), "det_candidates")
    counts = count.tolist()
    out = []
    for a, k in enumerate(counts):
        if k > max_candidates:
            raise RuntimeError(f"agent {a}: {k} anchors above the score threshold exceed max_candidates={max_candidates}")
        sc, order = torch.sort(scores[a, :k], descending=True, stable=True)
        out.append({"corners": corners[a, :k][order], "score": sc, "index": index[a, :k][order].long()})
    return out
