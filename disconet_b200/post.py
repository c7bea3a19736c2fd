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
        # descending score; ties: larger anchor number first (`scores.argsort()[::-1]`, utils/postprocess.py:86)
        key = (scores[a, :k].view(torch.int32).long() << 32) | index[a, :k].long()
        order = torch.argsort(key, descending=True)
        out.append({"corners": corners[a, :k][order], "score": scores[a, :k][order], "index": index[a, :k][order].long()})
    return out


# ---------------------------------------------------------------------------------------------------------------
# Rotated-box NMS on the device and the reference-shaped wrappers around it
# ---------------------------------------------------------------------------------------------------------------
import numpy as np

SCORE_THRESH = 0.7      # utils/postprocess.py:85
NMS_IOU_THRESH = 0.01   # detection_util.py:349-351 (apply_nms_det), :962-964 (late_fusion)


def nms_rotated_batched(corners: torch.Tensor, scores: torch.Tensor, ids=None, count=None, *, threshold: float = NMS_IOU_THRESH,
                        score_thresh: float = SCORE_THRESH, max_boxes: int = 4096):
    """`non_max_suppression` (utils/postprocess.py:72-115) for n independent candidate sets in one launch sequence.

    corners [n, cap, 4, 2] float32 | float64, scores [n, cap] float32, ids [n, cap] int32 (tie-break key, optional),
    count [n] int32 (device; None = all `cap` boxes).  Returns device tensors (keep [n, kmax] int32 = picked positions in
    pick order, n_keep [n] int32, n_valid [n] int32 = boxes above `score_thresh`); no host synchronisation."""
    if not (corners.is_cuda and scores.is_cuda):
        raise ValueError("disconet_b200.post runs on CUDA tensors only (no CPU fallback)")
    if corners.dim() != 4 or corners.shape[2:] != (4, 2) or tuple(scores.shape) != tuple(corners.shape[:2]):
        raise ValueError(f"corners [n, cap, 4, 2] / scores [n, cap] expected (got {tuple(corners.shape)}, {tuple(scores.shape)})")
    if corners.dtype not in (torch.float32, torch.float64):
        raise ValueError("corners must be float32 or float64")
    n, cap = scores.shape
    if cap > 8192:
        raise ValueError("at most 8192 candidate boxes per set")
    kmax = min(int(max_boxes), cap)
    dev = corners.device
    corners, scores = corners.contiguous(), scores.float().contiguous()
    ids = ids.to(torch.int32).contiguous() if ids is not None else None
    count = count.to(torch.int32).contiguous() if count is not None else None
    lib = load()
    ws = torch.empty((int(lib.disco_nms_workspace_bytes(n, kmax)),), dtype=torch.uint8, device=dev)
    keep = torch.empty((n, kmax), dtype=torch.int32, device=dev)
    n_keep = torch.empty((n,), dtype=torch.int32, device=dev)
    n_valid = torch.empty((n,), dtype=torch.int32, device=dev)
    check(lib.disco_nms_rotated(corners.data_ptr(), int(corners.dtype == torch.float64), scores.data_ptr(),
                                ids.data_ptr() if ids is not None else None, count.data_ptr() if count is not None else None,
                                n, cap, kmax, float(score_thresh), float(threshold), ws.data_ptr(), ws.numel(), keep.data_ptr(),
                                n_keep.data_ptr(), n_valid.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "nms_rotated")
    return keep, n_keep, n_valid


def non_max_suppression(boxes, scores, threshold, device=None):
    """Drop-in for coperception.utils.postprocess.non_max_suppression: boxes [K, 4, 2], scores [K] (numpy or torch) ->
    int32 numpy array of the picked positions.  The polygon IoU loop runs on the GPU (float64 clip)."""
    dev = device or (boxes.device if isinstance(boxes, torch.Tensor) and boxes.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    b = torch.as_tensor(np.asarray(boxes) if not isinstance(boxes, torch.Tensor) else boxes)
    s = torch.as_tensor(np.asarray(scores) if not isinstance(scores, torch.Tensor) else scores)
    assert b.shape[0] > 0
    if b.dtype not in (torch.float32, torch.float64):
        b = b.float()
    k = b.shape[0]
    pick = []
    # sets larger than the kernel's 8192-box capacity: only the boxes above the score threshold take part anyway
    if k > 8192:
        sel = torch.nonzero(s > SCORE_THRESH).flatten()
        if sel.numel() > 8192:
            raise RuntimeError(f"{sel.numel()} boxes above the score threshold exceed the 8192-box NMS capacity")
        b, s, base = b[sel], s[sel], sel
    else:
        base = None
    if b.shape[0] == 0:
        return np.array(pick, dtype=np.int32)
    ids = (base if base is not None else torch.arange(b.shape[0])).to(torch.int32)
    keep, n_keep, _ = nms_rotated_batched(b.reshape(1, -1, 4, 2).to(dev), s.reshape(1, -1).float().to(dev), ids.reshape(1, -1).to(dev),
                                          threshold=threshold, max_boxes=b.shape[0])
    nk = int(n_keep.item())
    out = keep[0, :nk].cpu()
    if base is not None:
        out = base.cpu()[out.long()]
    print("selected: ", nk)                       # (the reference function prints this, postprocess.py:114)
    return out.numpy().astype(np.int32)


def detect(loc: torch.Tensor, cls: torch.Tensor, anchors: torch.Tensor, *, score_thresh: float = SCORE_THRESH,
           threshold: float = NMS_IOU_THRESH, max_candidates: int = 4096, device_only: bool = False):
    """Everything `apply_nms_det` does after the network, for ALL agents at once and without leaving the device:
    softmax score / threshold / box decode / corners (det_candidates_kernel) -> sort -> polygon-IoU NMS.

    loc [N, H, W, A, 1, 6], cls [N, H*W*A, 2], anchors [N, H, W, A, 6] | [H, W, A, 6].  Returns per agent
    {"pred": [K', 1, 4, 2] float64, "score": [K'] float32, "selected_idx": [K'] int32} as numpy (one small D2H copy), or --
    with `device_only` -- the packed device tensors (corners, scores, index, keep, n_keep, n_valid) with no host sync."""
    if not (loc.is_cuda and cls.is_cuda and anchors.is_cuda):
        raise ValueError("disconet_b200.post runs on CUDA tensors only (no CPU fallback)")
    n, per = cls.shape[0], cls.shape[1]
    if cls.shape[-1] != 2 or loc.numel() != n * per * 6:
        raise ValueError(f"expected binary cls [N, anchors, 2] and loc with 6 codes per anchor (got {tuple(cls.shape)}, {tuple(loc.shape)})")
    if per >= (1 << 19):
        raise ValueError("at most 2^19 anchors per agent")
    loc_c, cls_c, anc = loc.detach().float().contiguous(), cls.detach().float().contiguous(), anchors.detach().float().contiguous()
    if anc.numel() == per * 6:
        stride = 0
    elif anc.numel() == n * per * 6:
        stride = per * 6
    else:
        raise ValueError(f"anchors {tuple(anchors.shape)} do not match {per} anchors per agent")
    dev = cls.device
    cap = int(max_candidates)
    count = torch.empty(n, dtype=torch.int32, device=dev)
    corners = torch.empty((n, cap, 4, 2), dtype=torch.float32, device=dev)
    scores = torch.empty((n, cap), dtype=torch.float32, device=dev)
    index = torch.empty((n, cap), dtype=torch.int32, device=dev)
    check(load().disco_det_candidates(loc_c.data_ptr(), cls_c.data_ptr(), anc.data_ptr(), per, stride, n, float(score_thresh), cap,
                                      count.data_ptr(), corners.data_ptr(), scores.data_ptr(), index.data_ptr(),
                                      torch.cuda.current_stream(dev).cuda_stream), "det_candidates")
    keep, n_keep, n_valid = nms_rotated_batched(corners, scores, index, count, threshold=threshold, score_thresh=score_thresh,
                                                max_boxes=cap)
    if device_only:
        return corners, scores, index, keep, n_keep, n_valid, count
    # one compact device -> host transfer: counts first, then only the kept rows
    cnt = torch.stack((count, n_keep, n_valid)).cpu()
    out = []
    for a in range(n):
        if int(cnt[0, a]) > cap:
            raise RuntimeError(f"agent {a}: {int(cnt[0, a])} anchors above the score threshold exceed max_candidates={cap}")
        k = keep[a, :int(cnt[1, a])].long()
        out.append({"pred": corners[a][k].double().cpu().numpy()[:, None], "score": scores[a][k].cpu().numpy(),
                    "selected_idx": index[a][k].cpu().numpy().astype(np.int32)})
    return out


def apply_nms_det(batch_box_preds, batch_cls_preds, anchors, code_type, config, batch_motion=None):
    """Drop-in for coperception.utils.detection_util.apply_nms_det (:256-373) on the default detection config
    (code_type 'faf', binary head, pred_len 1): same arguments, same `(predictions_dicts, cls_pred_first_nms)` return."""
    # (Config's default pred_type is "motion"; with T = 1 its i == 0 branch is the plain decode, detection_util.py:304-313)
    if code_type[0] != "f" or getattr(config, "motion_state", False):
        raise NotImplementedError("disconet_b200.post.apply_nms_det covers code_type='faf' without the motion-state head")
    assert len(batch_box_preds.shape) == 6, "bbox must have shape [N ,W , H , num_per_loc, T, box_code]"
    if batch_box_preds.shape[4] != 1 or batch_cls_preds.shape[-1] != 2:
        raise NotImplementedError("only_det (T = 1) / binary classification only")
    n = batch_box_preds.shape[0]
    res = detect(batch_box_preds, batch_cls_preds.reshape(n, -1, 2), anchors.reshape((n,) + tuple(batch_box_preds.shape[1:4]) + (6,)))
    predictions_dicts = [[r] for r in res]
    for r in res:
        print("selected: ", len(r["selected_idx"]))   # (printed by the reference's non_max_suppression, postprocess.py:114)
    sel = torch.as_tensor(res[-1]["selected_idx"].astype(np.int64), device=batch_cls_preds.device)
    cls_pred_first_nms = batch_cls_preds[n - 1][sel, :]
    return predictions_dicts, cls_pred_first_nms


def late_fusion(ego_agent, num_agent, result, trans_matrices, box_color_map):
    """Drop-in for coperception.utils.detection_util.late_fusion (:927-973): the neighbours' kept boxes are moved into the
    ego frame on the host exactly as the reference does (float64 numpy, a few hundred corners) and the merged set goes
    through the GPU polygon NMS."""
    box_colors = np.array([box_color_map[ego_agent] for _ in result[ego_agent][0][0][0]["pred"]])
    for j in range(num_agent):
        if j == ego_agent or len(result[ego_agent]) == 0 or len(result[j]) == 0:
            continue
        trans_mat_j2ego = np.asarray(trans_matrices[0, ego_agent, j])
        trans_mat_j2ego = np.delete(trans_mat_j2ego, 2, axis=1)
        trans_mat_j2ego = np.delete(trans_mat_j2ego, 2, axis=0)
        boxes_j = np.array(result[j][0][0][0]["pred"])
        points = boxes_j.reshape(-1, 2).T
        points[0, :] = -points[0, :]
        points = np.dot(trans_mat_j2ego, np.vstack((points, np.ones(points.shape[1]))))[:2, :]
        points[0, :] = -points[0, :]
        points = points.T.reshape(-1, 1, 4, 2)
        ego = result[ego_agent][0][0][0]
        ego["pred"] = np.vstack((ego["pred"], points))
        ego["score"] = np.append(ego["score"], result[j][0][0][0]["score"])
        ego["selected_idx"] = np.append(ego["selected_idx"], result[j][0][0][0]["selected_idx"])
        box_colors = np.append(box_colors, [box_color_map[j] for _ in points])
    if len(result[ego_agent]) > 0:
        ego = result[ego_agent][0][0][0]
        boxes = np.squeeze(ego["pred"])
        pick = non_max_suppression(boxes, ego["score"], NMS_IOU_THRESH)
        ego["pred"] = np.take(ego["pred"], pick, axis=0)
        box_colors = np.take(box_colors, pick, axis=0)
    return box_colors
