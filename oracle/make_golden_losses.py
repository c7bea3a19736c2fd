"""Generate tests/golden/{losses,post}.npz from the LIVE reference objects (build container only; needs /root/reference).

    python -m oracle.make_golden_losses

Pins rows f2 / f3 / f4 of SURVEY.md §8 to the reference's own code:
  f2  FaFModule.get_kd_loss                    (coperception/utils/CoDetModule.py:312-388)
  f4  SoftmaxFocalClassificationLoss           (coperception/utils/loss.py:322-394)
      FaFModule.corner_loss                    (CoDetModule.py:80-105)
  f3  non_max_suppression / apply_nms_det / late_fusion (utils/postprocess.py:72-115, detection_util.py:256-373,927-973)
      -- run UNMODIFIED, with `shapely.geometry.Polygon` (absent from this image) replaced by oracle.post_oracle.StubPolygon,
      so the reference's loop / ordering / thresholds are pinned while the polygon areas are our float64 restatement.
Inputs are regenerated from seeds by the tests (oracle.post_oracle.synth_*); only outputs are stored.
"""
import os
import sys
import types

import numpy as np
import torch

from oracle import post_oracle as P
from oracle import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

KD_SEED, FOCAL_SEED, CORNER_SEED = 71, 72, 73
# (no exact score ties here: `scores.argsort()[::-1]` uses numpy's unstable default sort, so the reference's order among
# equal scores is implementation-defined -- measured on this host: positions [6, 7], [9, 10] ascending but [4, 3]
# descending inside ONE call.  Our rule for ties (larger position / anchor number first) is tested GPU-vs-oracle only.)
NMS_CASES = {"nms_k300": dict(seed=81, k=300), "nms_k40": dict(seed=82, k=40), "nms_k12": dict(seed=83, k=12),
             "nms_dense": dict(seed=84, k=400, extent=8.0)}
DET_CASE = dict(seed=91, n=2, H=32, W=32)
LATE_CASE = dict(seed=95, n=3, H=32, W=32)


def install_stub_shapely():
    ref_import.install_bypass(mock_heavy=True)
    ref_import.install_stub_shapely()
    for m in ("coperception.utils.postprocess", "coperception.utils.detection_util", "coperception.utils.CoDetModule"):
        sys.modules.pop(m, None)


def late_fusion_inputs(case):
    """Per-agent NMS results (the structure predict_all returns) + poses for the late-fusion case."""
    from oracle import disconet_oracle as O
    loc, cls, anc = P.synth_head_outputs(case["seed"], case["n"], case["H"], case["W"])
    T = O.synth_poses(1, case["n"], seed=case["seed"] + 1).numpy()
    return loc, cls, anc, T


def main():
    install_stub_shapely()
    from coperception.utils import postprocess as R_post
    from coperception.utils import detection_util as R_du
    from coperception.utils.CoDetModule import FaFModule
    from coperception.utils.loss import SoftmaxFocalClassificationLoss
    from coperception.configs.Config import Config
    cfg = Config("train", binary=True, only_det=True)

    rec = {}
    # ---- f2: KD loss through the reference method (teacher = a stub returning the seeded teacher maps) --------------
    maps = P.synth_kd_maps(KD_SEED, 2)
    stu = [torch.from_numpy(s).requires_grad_(True) for s, _ in maps]
    tea = [torch.from_numpy(t) for _, t in maps]
    fm = FaFModule.__new__(FaFModule)
    fm.kd_flag = 1
    fm.teacher = lambda bev: (None, tea[0], tea[1], tea[2], tea[3], None)
    kd = fm.get_kd_loss(1, {"bev_seq_teacher": None, "kd_weight": 100000}, stu[3], 2, stu[2], stu[1], stu[0])
    kd.backward()
    rec["kd_loss"] = np.array([kd.item()])
    for name, t in zip(("x7", "x6", "x5", "fused"), stu):
        g = t.grad.reshape(-1).double()
        rec[f"kd_grad_{name}_sub"] = g[::499].float().numpy()
        rec[f"kd_grad_{name}_norm"] = np.array([g.norm().item(), g.abs().max().item()])

    # ---- f4: focal classification loss ------------------------------------------------------------------------------
    logits, target = P.synth_focal_inputs(FOCAL_SEED, 2, 6000)
    z = torch.from_numpy(logits).requires_grad_(True)
    out = SoftmaxFocalClassificationLoss()(z, torch.from_numpy(target))
    loss = torch.sum(out) / 2
    loss.backward()
    rec["focal_out_sub"] = out.detach().reshape(-1)[::7].numpy()
    rec["focal_loss"] = np.array([loss.item()])
    rec["focal_grad"] = z.grad.numpy()

    # ---- f4: corner loss through the reference method ---------------------------------------------------------------
    anchors, mask, targets, pred = P.synth_reg_targets(CORNER_SEED, 2, 32, 32)
    p = torch.from_numpy(pred).requires_grad_(True)
    cl = FaFModule.corner_loss(fm, torch.from_numpy(anchors), torch.from_numpy(mask), torch.from_numpy(targets), p)
    cl.backward()
    rec["corner_loss"] = np.array([cl.item()])
    rec["corner_grad_nz"] = p.grad.numpy()[mask]
    rec["corner_grad_absmax_unmasked"] = np.array([np.abs(p.grad.numpy()[~mask]).max()])
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **rec)
    print("losses", {k: v.shape for k, v in rec.items()}, "kd", kd.item(), "focal", loss.item(), "corner", cl.item())

    # ---- f3 ---------------------------------------------------------------------------------------------------------
    rec = {}
    for name, c in NMS_CASES.items():
        boxes, scores = P.synth_rotated_boxes(c["seed"], c["k"], extent=c.get("extent", 24.0), ties=c.get("ties", False))
        rec[name + "_pick"] = R_post.non_max_suppression(boxes, scores, threshold=0.01)
    loc, cls, anc = P.synth_head_outputs(DET_CASE["seed"], DET_CASE["n"], DET_CASE["H"], DET_CASE["W"])
    cfg.motion_state = False
    for a in range(DET_CASE["n"]):   # predict_all calls apply_nms_det once per agent (CoDetModule.py:484-511)
        pd, first = R_du.apply_nms_det(torch.from_numpy(loc[a:a + 1]), torch.from_numpy(cls[a:a + 1]), torch.from_numpy(anc[None]),
                                       cfg.code_type, cfg, None)
        r = pd[0][0]
        rec[f"det{a}_pred"], rec[f"det{a}_score"], rec[f"det{a}_idx"] = r["pred"], r["score"], r["selected_idx"]
        rec[f"det{a}_first"] = first.numpy()
    loc, cls, anc, T = late_fusion_inputs(LATE_CASE)
    result = []
    for a in range(LATE_CASE["n"]):
        pd, _ = R_du.apply_nms_det(torch.from_numpy(loc[a:a + 1]), torch.from_numpy(cls[a:a + 1]), torch.from_numpy(anc[None]), cfg.code_type,
                                   cfg, None)
        result.append([[[{k: np.array(v) for k, v in pd[0][0].items()}]]])   # result[k][0][0][0] as test_codet.py:292-316 indexes it
    colors = R_du.late_fusion(0, LATE_CASE["n"], result, T, ["red", "green", "blue"])
    rec["late_pred"] = result[0][0][0][0]["pred"]
    rec["late_colors"] = np.array(colors)
    np.savez_compressed(os.path.join(OUT, "post.npz"), **rec)
    print("post", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
