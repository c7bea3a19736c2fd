"""Swap the reference's model classes for the B200-native ones without touching the reference tree.

    import disconet_b200.patch as p; p.patch_coperception()
    # or run an unmodified reference tool:
    python -m disconet_b200.patch /path/to/coperception/tools/det/test_codet.py --com disco ...

`tools/det/train_codet.py:12` / `test_codet.py:14` do `from coperception.models.det import *`, so the
classes are replaced on the already-imported `coperception.models.det` package before the tool runs.
"""
from __future__ import annotations

import importlib
import runpy
import sys


def patch_coperception(classes=("DiscoNet", "FaFNet", "TeacherNet")) -> None:
    from . import det as ours
    pkg = importlib.import_module("coperception.models.det")
    for name in classes:
        setattr(pkg, name, getattr(ours, name))
        sub = sys.modules.get(f"coperception.models.det.{name}")
        if sub is not None:
            setattr(sub, name, getattr(ours, name))
    # KD loss (FaFModule.get_kd_loss), corner loss (FaFModule.corner_loss) and the detection post-processing of
    # predict_all (apply_nms_det -> polygon NMS) on the fused kernels
    try:
        mod = importlib.import_module("coperception.utils.CoDetModule")
        from . import kd, loss as ours_loss, post
        mod.FaFModule.get_kd_loss = kd.get_kd_loss
        mod.FaFModule.corner_loss = ours_loss._corner_loss_method
        mod.apply_nms_det = post.apply_nms_det          # CoDetModule does `from ...detection_util import *`
        du = importlib.import_module("coperception.utils.detection_util")
        du.apply_nms_det = post.apply_nms_det
        du.late_fusion = post.late_fusion
        du.non_max_suppression = post.non_max_suppression
        pp = importlib.import_module("coperception.utils.postprocess")
        pp.non_max_suppression = post.non_max_suppression
    except Exception:
        pass
    # focal classification loss (train_codet.py:173-176 builds it from coperception.utils.loss)
    try:
        lmod = importlib.import_module("coperception.utils.loss")
        from . import loss as ours_loss
        lmod.SoftmaxFocalClassificationLoss = ours_loss.SoftmaxFocalClassificationLoss
    except Exception:
        pass
    # BEV segmentation (tools/seg/*.py do `from coperception.models.seg import *`)
    try:
        seg_pkg = importlib.import_module("coperception.models.seg")
    except Exception:   # the seg package pulls optional dependencies; the detection tools do not need it
        return
    from .seg import SegDiscoNet
    setattr(seg_pkg, "DiscoNet", SegDiscoNet)
    sub = sys.modules.get("coperception.models.seg.DiscoNet")
    if sub is not None:
        setattr(sub, "DiscoNet", SegDiscoNet)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m disconet_b200.patch <reference tool .py> [tool args...]")
    patch_coperception()
    tool = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(tool, run_name="__main__")
