"""CPU oracle for the GPU half of the detection post-processing -- TEST INFRASTRUCTURE.

numpy restatement of what `apply_nms_det` computes for every anchor BEFORE the polygon NMS
(R = /root/reference/coperception/coperception/utils):
  scores        softmax over the class axis, foreground column            R/detection_util.py:276-277
  box decode    bev_box_decode_torch                                      R/detection_util.py:376-400
  corners       center_to_corner_box2d -> corners_nd, rotation_2d         R/obj_util.py:271-359 (via detection_util.py:330-332)
  candidates    scores > 0.7, highest score first                         R/postprocess.py:83-88
Pinned against those reference functions by tests/test_oracle_cpu.py::test_post_oracle_matches_reference.
"""
import numpy as np


def decode_boxes(enc: np.ndarray, anchors: np.ndarray) -> np.ndarray:
    """enc / anchors [..., 6] = (x, y, w, h, sin, cos) -> decoded boxes [..., 6]."""
    xa, ya, wa, ha, sina, cosa = [anchors[..., i] for i in range(6)]
    xp, yp, wp, hp, sinp, cosp = [enc[..., i] for i in range(6)]
    h = ha / np.exp(hp)
    w = wa / np.exp(wp)
    x = xa - w * xp
    y = ya - h * yp
    s = sina * cosp + cosa * sinp
    c = cosa * cosp - sina * sinp
    return np.stack([x, y, w, h, s, c], -1)


def box_corners(dec: np.ndarray) -> np.ndarray:
    """decoded boxes [K, 6] -> corners [K, 4, 2]: (x0y1, x1y1, x1y0, x0y0) of the w x h box, rotated with
    x' = x cos + y sin, y' = -x sin + y cos (sin / cos as decoded, not re-normalised), moved to the centre."""
    norm = np.array([[-0.5, 0.5], [0.5, 0.5], [0.5, -0.5], [-0.5, -0.5]], dtype=dec.dtype)
    pts = dec[:, None, 2:4] * norm[None]
    s, c = dec[:, 4, None], dec[:, 5, None]
    rot = np.stack([pts[..., 0] * c + pts[..., 1] * s, -pts[..., 0] * s + pts[..., 1] * c], -1)
    return rot + dec[:, None, :2]


def det_candidates(loc: np.ndarray, cls: np.ndarray, anchors: np.ndarray, thresh: float = 0.7):
    """One agent: loc [H, W, A, 1, 6], cls [H*W*A, 2], anchors [H, W, A, 6] -> (corners [K,4,2], scores [K], index [K])
    of the anchors with foreground probability > thresh, highest first."""
    z = cls.astype(np.float64)
    z = z - z.max(-1, keepdims=True)
    p = np.exp(z)
    score = (p[:, 1] / p.sum(-1)).astype(np.float32)
    idx = np.where(score > thresh)[0]
    order = idx[np.argsort(-score[idx], kind="stable")]
    dec = decode_boxes(loc.reshape(-1, 6)[order].astype(np.float32), anchors.reshape(-1, 6)[order].astype(np.float32))
    return box_corners(dec), score[order], order.astype(np.int32)
