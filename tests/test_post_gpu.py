"""GPU parity of rows f2 / f3 / f4: the CUDA kernels against the reference-pinned goldens (tests/golden/{post,losses}.npz)
and, on more shapes / edge cases, against the oracles that those goldens pin."""
import os

import numpy as np
import pytest
import torch

from oracle import loss_oracle as L
from oracle import make_golden_losses as G
from oracle import post_oracle as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_nms_kernel_matches_reference_picks(cuda_dev):
    from disconet_b200 import post
    g = np.load(os.path.join(GOLD, "post.npz"))
    for name, c in G.NMS_CASES.items():
        boxes, scores = P.synth_rotated_boxes(c["seed"], c["k"], extent=c.get("extent", 24.0))
        pick = post.non_max_suppression(boxes, scores, 0.01, device=cuda_dev)
        assert pick.dtype == np.int32 and np.array_equal(pick, g[name + "_pick"]), name     # bit-exact kept-index list


def test_nms_kernel_matches_oracle_edge_cases(cuda_dev):
    from disconet_b200 import post
    # exact score ties, dense overlap, a single box, nothing above the score threshold, > 64 and > 1024 candidates, fp32 corners
    for seed, k, ext, ties in [(1, 64, 24.0, True), (2, 65, 6.0, True), (3, 1, 24.0, False), (4, 1500, 40.0, False), (5, 700, 10.0, True)]:
        boxes, scores = P.synth_rotated_boxes(seed, k, extent=ext, ties=ties)
        want = P.non_max_suppression(boxes, scores, 0.01)
        got = post.non_max_suppression(boxes, scores, 0.01, device=cuda_dev)
        assert np.array_equal(got, want), (seed, k)
        got32 = post.non_max_suppression(boxes.astype(np.float32), scores, 0.01, device=cuda_dev)
        assert np.array_equal(got32, P.non_max_suppression(boxes.astype(np.float32).astype(np.float64), scores, 0.01))
    boxes, scores = P.synth_rotated_boxes(6, 50)
    assert post.non_max_suppression(boxes, np.full(50, 0.5, np.float32), 0.01, device=cuda_dev).size == 0
    # other IoU thresholds
    boxes, scores = P.synth_rotated_boxes(7, 400, extent=8.0)
    for thr in (0.0, 0.3, 0.7):
        assert np.array_equal(post.non_max_suppression(boxes, scores, thr, device=cuda_dev), P.non_max_suppression(boxes, scores, thr))
    # batched, ragged counts: three sets in one launch sequence
    sets = [P.synth_rotated_boxes(10 + i, k, extent=12.0) for i, k in enumerate((300, 17, 0))]
    cap = 320
    cor = np.zeros((3, cap, 4, 2)); sc = np.zeros((3, cap), np.float32); cnt = np.array([300, 17, 0], np.int32)
    for i, (b, s) in enumerate(sets):
        cor[i, :len(b)], sc[i, :len(s)] = b, s
    keep, n_keep, n_valid = post.nms_rotated_batched(torch.from_numpy(cor).to(cuda_dev), torch.from_numpy(sc).to(cuda_dev),
                                                     count=torch.from_numpy(cnt).to(cuda_dev))
    for i, (b, s) in enumerate(sets):
        want = P.non_max_suppression(b, s, 0.01) if len(b) else np.zeros(0, np.int32)
        assert int(n_keep[i]) == len(want) and np.array_equal(keep[i, :len(want)].cpu().numpy(), want)
        assert int(n_valid[i]) == int((s > 0.7).sum())


def test_detect_and_apply_nms_det_match_reference(cuda_dev):
    """det_candidates -> sort -> NMS for all agents at once, and the apply_nms_det mirror, against the live-reference golden."""
    from disconet_b200 import post
    g = np.load(os.path.join(GOLD, "post.npz"))
    c = G.DET_CASE
    loc, cls, anc = P.synth_head_outputs(c["seed"], c["n"], c["H"], c["W"])
    d = lambda x: torch.from_numpy(x).to(cuda_dev)
    res = post.detect(d(loc), d(cls), d(anc))
    for a in range(c["n"]):
        assert np.array_equal(res[a]["selected_idx"], g[f"det{a}_idx"])          # bit-exact kept anchor list
        assert res[a]["pred"].shape == g[f"det{a}_pred"].shape and res[a]["pred"].dtype == np.float64
        assert np.abs(res[a]["pred"] - g[f"det{a}_pred"]).max() <= 1e-5
        assert np.abs(res[a]["score"] - g[f"det{a}_score"]).max() <= 1e-6

    class Cfg:
        motion_state = False; pred_type = "center"
    for a in range(c["n"]):   # the way predict_all calls it: one agent per call (CoDetModule.py:484-511)
        pd, first = post.apply_nms_det(d(loc[a:a + 1]), d(cls[a:a + 1]), d(anc[None]), "faf", Cfg(), None)
        assert len(pd) == 1 and len(pd[0]) == 1
        assert np.array_equal(pd[0][0]["selected_idx"], g[f"det{a}_idx"])
        assert np.array_equal(first.cpu().numpy(), g[f"det{a}_first"])


def test_late_fusion_matches_reference(cuda_dev):
    from disconet_b200 import post
    g = np.load(os.path.join(GOLD, "post.npz"))
    c = G.LATE_CASE
    loc, cls, anc, T = G.late_fusion_inputs(c)
    d = lambda x: torch.from_numpy(x).to(cuda_dev)
    res = post.detect(d(loc), d(cls), d(anc))
    result = [[[[r]]] for r in res]                           # result[k][0][0][0] as test_codet.py:292-316 indexes it
    colors = post.late_fusion(0, c["n"], result, T, ["red", "green", "blue"])
    assert result[0][0][0][0]["pred"].shape == g["late_pred"].shape
    assert np.abs(result[0][0][0][0]["pred"] - g["late_pred"]).max() <= 1e-5
    assert np.array_equal(np.array(colors), g["late_colors"])


def test_corner_loss_kernel(cuda_dev):
    from disconet_b200.loss import corner_loss
    g = np.load(os.path.join(GOLD, "losses.npz"))
    anchors, mask, targets, pred = P.synth_reg_targets(G.CORNER_SEED, 2, 32, 32)
    d = lambda x: torch.from_numpy(x).to(cuda_dev)
    p = d(pred).requires_grad_(True)
    loss = corner_loss(d(anchors), d(mask), d(targets), p)
    (loss * 3.0).backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - g["corner_loss"][0]) <= 1e-5 * g["corner_loss"][0]               # vs FaFModule.corner_loss
    gr = p.grad.cpu().numpy() / 3.0
    assert np.abs(gr[mask] - g["corner_grad_nz"]).max() <= 1e-4 * np.abs(g["corner_grad_nz"]).max()
    assert np.abs(gr[~mask]).max() == 0.0
    # full-size maps (256 x 256 x 6 anchors, 5 agents) against the float64 oracle; and an empty mask
    anchors, mask, targets, pred = P.synth_reg_targets(5, 5, 256, 256, pos_frac=1e-3)
    want, wgrad = P.corner_loss(anchors, mask, targets, pred)
    p = d(pred).requires_grad_(True)
    loss = corner_loss(d(anchors), d(mask), d(targets), p)
    loss.backward()
    assert abs(loss.item() - want) <= 1e-5 * want
    assert np.abs(p.grad.cpu().numpy() - wgrad).max() <= 1e-4 * np.abs(wgrad).max()
    p = d(pred).requires_grad_(True)
    loss = corner_loss(d(anchors), d(np.zeros_like(mask)), d(targets), p)
    loss.backward()
    assert loss.item() == 0.0 and float(p.grad.abs().max()) == 0.0


def test_kd_loss_matches_reference_golden(cuda_dev):
    """f2 pinned: our get_kd_loss (bound onto FaFModule by the patcher) on the same seeded maps as the live reference's."""
    from disconet_b200 import kd
    g = np.load(os.path.join(GOLD, "losses.npz"))
    maps = P.synth_kd_maps(G.KD_SEED, 2)
    stu = [torch.from_numpy(s).to(cuda_dev).requires_grad_(True) for s, _ in maps]
    tea = [torch.from_numpy(t).to(cuda_dev) for _, t in maps]

    class Self:
        kd_flag = 1
        teacher = staticmethod(lambda bev: (None, tea[0], tea[1], tea[2], tea[3], None))
    loss = kd.get_kd_loss(Self(), 1, {"bev_seq_teacher": None, "kd_weight": 100000}, stu[3], 2, stu[2], stu[1], stu[0])
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - g["kd_loss"][0]) <= 2e-5 * g["kd_loss"][0]
    for t, name in zip(stu, ("x7", "x6", "x5", "fused")):
        sub = t.grad.reshape(-1)[::499].cpu().numpy()
        assert np.abs(sub - g[f"kd_grad_{name}_sub"]).max() <= 2e-5 * g[f"kd_grad_{name}_norm"][1], name


def test_focal_loss_matches_reference_golden(cuda_dev):
    from disconet_b200.loss import SoftmaxFocalClassificationLoss
    g = np.load(os.path.join(GOLD, "losses.npz"))
    logits, target = P.synth_focal_inputs(G.FOCAL_SEED, 2, 6000)
    z = torch.from_numpy(logits).to(cuda_dev).requires_grad_(True)
    out = SoftmaxFocalClassificationLoss()(z, torch.from_numpy(target).to(cuda_dev))
    loss = torch.sum(out) / 2
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - g["focal_loss"][0]) <= 1e-5 * g["focal_loss"][0]
    assert np.abs(out.detach().reshape(-1)[::7].cpu().numpy() - g["focal_out_sub"]).max() <= 1e-5 * np.abs(g["focal_out_sub"]).max()
    assert np.abs(z.grad.cpu().numpy() - g["focal_grad"]).max() <= 1e-5 * np.abs(g["focal_grad"]).max()
