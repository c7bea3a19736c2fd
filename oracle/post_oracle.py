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
    order = nms_order(score, None, thresh)          # descending score, ties: larger anchor number first
    dec = decode_boxes(loc.reshape(-1, 6)[order].astype(np.float32), anchors.reshape(-1, 6)[order].astype(np.float32))
    return box_corners(dec), score[order], order.astype(np.int32)


# ---------------------------------------------------------------------------------------------------------------
# Rotated-box NMS (row f3): restatement of `non_max_suppression` (R/postprocess.py:72-115) with the shapely
# polygon arithmetic (`box.intersection(b).area / box.union(b).area`, :49-50) replaced by an explicit float64
# convex-quad clip.  shapely / GEOS is a third-party dependency that is absent from /root/reference and from this
# image (requirements.txt pins no version); GEOS computes the overlay of two simple polygons in float64, which for
# two convex quads is their Sutherland-Hodgman clip, and `union.area` = area_a + area_b - intersection.area up to
# rounding.  The LOOP (filter > 0.7, argsort()[::-1], greedy removal with `iou > threshold`) is pinned by running the
# unmodified reference function on `StubPolygon` (oracle/make_golden.py); the area arithmetic is our restatement.
# Every product / sum below is an individually rounded float64 operation, in the same order as csrc/post.cu.
# ---------------------------------------------------------------------------------------------------------------
def _area_signed(x, y):
    n = len(x)
    s = 0.0
    for i in range(n):
        j = 0 if i + 1 == n else i + 1
        s = s + (x[i] * y[j] - x[j] * y[i])
    return 0.5 * s


def _cross3(ax, ay, bx, by, px, py):
    return (bx - ax) * (py - ay) - (by - ay) * (px - ax)


def quad_intersection_area(P, Q) -> float:
    """P, Q: [4, 2] float64 convex quads in either orientation -> area of their intersection."""
    px, py = [float(v) for v in P[:, 0]], [float(v) for v in P[:, 1]]
    qx, qy = [float(v) for v in Q[:, 0]], [float(v) for v in Q[:, 1]]
    if _area_signed(px, py) < 0.0:
        px[1], px[3] = px[3], px[1]
        py[1], py[3] = py[3], py[1]
    if _area_signed(qx, qy) < 0.0:
        qx[1], qx[3] = qx[3], qx[1]
        qy[1], qy[3] = qy[3], qy[1]
    sx, sy = px, py
    for e in range(4):
        if not sx:
            break
        ax, ay, bx, by = qx[e], qy[e], qx[(e + 1) & 3], qy[(e + 1) & 3]
        ox, oy = [], []
        n = len(sx)
        for i in range(n):
            j = 0 if i + 1 == n else i + 1
            ds = _cross3(ax, ay, bx, by, sx[i], sy[i])
            de = _cross3(ax, ay, bx, by, sx[j], sy[j])
            in_s, in_e = ds >= 0.0, de >= 0.0
            if in_s and len(ox) < 8:
                ox.append(sx[i]); oy.append(sy[i])
            if in_s != in_e and len(ox) < 8:
                t = ds / (ds - de)
                ox.append(sx[i] + t * (sx[j] - sx[i])); oy.append(sy[i] + t * (sy[j] - sy[i]))
        sx, sy = ox, oy
    if len(sx) < 3:
        return 0.0
    return abs(_area_signed(sx, sy))


def quad_area(P) -> float:
    return abs(_area_signed([float(v) for v in P[:, 0]], [float(v) for v in P[:, 1]]))


def quad_iou(P, Q) -> float:
    inter = quad_intersection_area(P, Q)
    uni = (quad_area(P) + quad_area(Q)) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(inter) / np.float64(uni))


class StubPolygon:
    """Stand-in for shapely.geometry.Polygon restricted to what R/postprocess.py touches: the constructor from a
    vertex list, `.intersection(other).area`, `.union(other).area`, `.area`.  Used ONLY to drive the unmodified
    reference loop when generating goldens."""

    class _Area:
        def __init__(self, a):
            self.area = a

    def __init__(self, pts):
        self.q = np.asarray(pts, dtype=np.float64).reshape(4, 2)

    @property
    def area(self):
        return quad_area(self.q)

    def intersection(self, other):
        return StubPolygon._Area(quad_intersection_area(self.q, other.q))

    def union(self, other):
        return StubPolygon._Area((quad_area(self.q) + quad_area(other.q)) - quad_intersection_area(self.q, other.q))


def nms_order(scores: np.ndarray, ids=None, score_thresh: float = 0.7) -> np.ndarray:
    """Positions with score > thresh, by descending score; ties: larger id first (= argsort()[::-1] of a stable argsort,
    R/postprocess.py:85-90)."""
    scores = np.asarray(scores)
    fil = np.where(scores > score_thresh)[0]
    key = fil if ids is None else np.asarray(ids)[fil]
    order = np.lexsort((key, scores[fil]))[::-1]
    return fil[order]


def non_max_suppression(boxes: np.ndarray, scores: np.ndarray, threshold: float, ids=None) -> np.ndarray:
    """boxes [K, 4, 2], scores [K] -> picked positions (int32) in pick order.  R/postprocess.py:72-115."""
    ixs = nms_order(scores, ids)
    quads = [np.asarray(boxes[i], dtype=np.float64).reshape(4, 2) for i in ixs]
    alive = list(range(len(ixs)))
    pick = []
    while alive:
        i = alive[0]
        pick.append(ixs[i])
        rest = []
        for j in alive[1:]:
            iou = quad_iou(quads[i], quads[j])
            if not (iou > threshold):     # NaN (degenerate boxes) keeps the box, like np.where(iou > threshold)
                rest.append(j)
        alive = rest
    return np.array(pick, dtype=np.int32)


def apply_nms_det_agent(loc: np.ndarray, cls: np.ndarray, anchors: np.ndarray):
    """One agent of `apply_nms_det` (R/detection_util.py:256-373), class 1 of the binary head: returns
    (pred [K', 1, 4, 2] float64, score [K'] float32, selected_idx [K'] int32)."""
    cor, sc, idx = det_candidates(loc, cls, anchors, 0.7)
    pick = non_max_suppression(cor.astype(np.float64), sc, 0.01, ids=idx)
    return cor.astype(np.float64)[pick][:, None], sc[pick], idx[pick].astype(np.int32)


def late_fusion_boxes(boxes_j: np.ndarray, trans_ego_j: np.ndarray) -> np.ndarray:
    """Corner transform of `late_fusion` (R/detection_util.py:936-951): boxes_j [K, 1, 4, 2] in agent j's frame ->
    ego frame with the 3x3 (z dropped) part of trans_matrices[0, ego, j]; x is negated before and after."""
    m = np.delete(np.delete(np.asarray(trans_ego_j, dtype=np.float64), 2, axis=1), 2, axis=0)
    pts = np.array(boxes_j, dtype=np.float64).reshape(-1, 2).T.copy()
    pts[0, :] = -pts[0, :]
    pts = np.dot(m, np.vstack((pts, np.ones(pts.shape[1]))))[:2, :]
    pts[0, :] = -pts[0, :]
    return pts.T.reshape(-1, 1, 4, 2)


# ---------------------------------------------------------------------------------------------------------------
# Corner loss (row f4): `FaFModule.corner_loss` (R/CoDetModule.py:80-105), numpy float64, value + gradient wrt pred
# ---------------------------------------------------------------------------------------------------------------
def corner_loss(anchors: np.ndarray, mask: np.ndarray, targets: np.ndarray, pred: np.ndarray):
    """anchors [N,H,W,A,6], mask [N,H,W,A,T] bool, targets / pred [N,H,W,A,T,6] -> (loss, dloss/dpred)."""
    N = pred.shape[0]
    T = mask.shape[-1]
    anc = np.broadcast_to(anchors[..., None, :], anchors.shape[:-1] + (T, 6))[mask].astype(np.float64)
    p = pred[mask].astype(np.float64)
    t = targets[mask].astype(np.float64)
    nx = np.array([-0.5, 0.5, 0.5, -0.5]); ny = np.array([0.5, 0.5, -0.5, -0.5])

    def corners(enc):
        d = decode_boxes(enc, anc)
        px, py = d[:, 2:3] * nx, d[:, 3:4] * ny
        c, s = d[:, 5:6], d[:, 4:5]
        return d, px, py, px * c + py * s + d[:, 0:1], -px * s + py * c + d[:, 1:2]

    dp, px, py, cx, cy = corners(p)
    _, _, _, tx, ty = corners(t)
    dx, dy = cx - tx, cy - ty
    dist = np.sqrt(dx * dx + dy * dy)
    loss = dist.sum() / N
    with np.errstate(divide="ignore", invalid="ignore"):
        gx = np.where(dist > 0, dx / dist, 0.0); gy = np.where(dist > 0, dy / dist, 0.0)
    c, s, w, h = dp[:, 5:6], dp[:, 4:5], dp[:, 2], dp[:, 3]
    Gx, Gy = gx.sum(1), gy.sum(1)
    Gw = (nx * (gx * c - gy * s)).sum(1); Gh = (ny * (gx * s + gy * c)).sum(1)
    Gc = (gx * px + gy * py).sum(1); Gs = (gx * py - gy * px).sum(1)
    g = np.stack([-w * Gx, -h * Gy, -w * (Gw - p[:, 0] * Gx), -h * (Gh - p[:, 1] * Gy),
                  Gs * anc[:, 5] - Gc * anc[:, 4], Gs * anc[:, 4] + Gc * anc[:, 5]], 1) / N
    grad = np.zeros(pred.shape, dtype=np.float64)
    grad[mask] = g
    return loss, grad


# ---------------------------------------------------------------------------------------------------------------
# Seeded inputs shared by oracle/make_golden.py (which feeds them to the live reference) and the tests
# ---------------------------------------------------------------------------------------------------------------
def synth_rotated_boxes(seed: int, k: int, extent: float = 24.0, ties: bool = False):
    """k random rotated boxes [k, 4, 2] float64 (corner order of center_to_corner_box2d) + scores [k] float32 in (0.5, 1)."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(-extent, extent, (k, 2))
    wh = rng.uniform(1.0, 5.0, (k, 2))
    ang = rng.uniform(-np.pi, np.pi, k)
    dec = np.concatenate([ctr, wh, np.sin(ang)[:, None], np.cos(ang)[:, None]], 1).astype(np.float32)
    boxes = box_corners(dec).astype(np.float64)
    scores = rng.uniform(0.5, 1.0, k).astype(np.float32)
    if ties:   # exact score ties (tie order = larger position first)
        scores[1::3] = scores[0:-1:3][: scores[1::3].size]
    return boxes, scores


def synth_head_outputs(seed: int, n: int, H: int, W: int, A: int = 6, pos_frac: float = 0.01):
    """Seeded (loc [n,H,W,A,1,6], cls [n,H*W*A,2], anchors [H,W,A,6]) float32 shaped like the detection head's result."""
    rng = np.random.default_rng(seed)
    loc = (rng.standard_normal((n, H, W, A, 1, 6)) * 0.3).astype(np.float32)
    cls = rng.standard_normal((n, H * W * A, 2)).astype(np.float32)
    hot = rng.random((n, H * W * A)) < pos_frac
    cls[..., 1] += np.where(hot, 3.0, -1.5).astype(np.float32)
    anchors = np.zeros((H, W, A, 6), dtype=np.float32)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    anchors[..., 0] = (xs[..., None] * (64.0 / W) - 32.0 + 32.0 / W)
    anchors[..., 1] = (ys[..., None] * (64.0 / H) - 32.0 + 32.0 / H)
    size = np.array([[2.0, 4.0], [4.0, 2.0], [2.5, 5.0], [5.0, 2.5], [1.5, 3.0], [3.0, 1.5]], dtype=np.float32)[:A]
    rot = np.array([0.0, 0.0, 0.3, -0.3, 0.8, -0.8], dtype=np.float32)[:A]
    anchors[..., 2:4] = size
    anchors[..., 4] = np.sin(rot)
    anchors[..., 5] = np.cos(rot)
    return loc, cls, anchors


def synth_reg_targets(seed: int, n: int, H: int, W: int, A: int = 6, pos_frac: float = 0.02):
    """Seeded regression-loss inputs: (anchors [n,H,W,A,6], mask [n,H,W,A,1] bool, targets, pred [n,H,W,A,1,6]) float32."""
    loc, _, anc = synth_head_outputs(seed, n, H, W, A)
    rng = np.random.default_rng(seed + 1)
    targets = (rng.standard_normal(loc.shape) * 0.1).astype(np.float32)
    mask = rng.random((n, H, W, A, 1)) < pos_frac
    anchors = np.broadcast_to(anc, (n,) + anc.shape).copy()
    return anchors, mask, targets, loc


def synth_kd_maps(seed: int, n: int):
    """Seeded student / teacher KD feature maps [(x7, t7), (x6, t6), (x5, t5), (fused, t3)] float32 at the resolutions
    `get_kd_loss` hard-codes in its reshapes (128^2, 64^2, 32^2, 32^2; CoDetModule.py:344-377)."""
    rng = np.random.default_rng(seed)
    shapes = [(n, 64, 128, 128), (n, 128, 64, 64), (n, 256, 32, 32), (n, 256, 32, 32)]
    return [(rng.standard_normal(s).astype(np.float32), rng.standard_normal(s).astype(np.float32) * 1.5) for s in shapes]


def synth_focal_inputs(seed: int, n: int, anchors: int):
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((n, anchors, 2)) * 2).astype(np.float32)
    pos = rng.random((n, anchors)) < 0.01
    target = np.stack([~pos, pos], -1).astype(np.float32)
    return logits, target


def late_fusion(ego: int, results, trans_matrices: np.ndarray):
    """Oracle of `late_fusion` (R/detection_util.py:927-973) on per-agent NMS outputs: results[k] = (pred [K,1,4,2], score [K])
    or None.  Returns (fused pred of the ego [K',1,4,2] float64, source agent of every kept box [K'])."""
    pred, score = results[ego][0].copy(), results[ego][1].copy()
    src = np.full(len(pred), ego)
    for j, r in enumerate(results):
        if j == ego or r is None:
            continue
        pts = late_fusion_boxes(r[0], trans_matrices[0, ego, j])
        pred = np.vstack((pred, pts))
        score = np.append(score, r[1])
        src = np.append(src, np.full(len(pts), j))
    pick = non_max_suppression(np.squeeze(pred, 1), score, 0.01)
    return pred[pick], src[pick]
