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


_originals = []   # (object, attribute name, original value) of everything patch_coperception() replaced


def _swap(obj, name, value) -> None:
    _originals.append((obj, name, getattr(obj, name)))
    setattr(obj, name, value)


def unpatch_coperception() -> None:
    """Undo patch_coperception(): put the reference's own classes / functions back (used by the tests that run the stock
    reference and the drop-in side by side in one process)."""
    while _originals:
        obj, name, value = _originals.pop()
        setattr(obj, name, value)


def patch_coperception(classes=("DiscoNet", "FaFNet", "TeacherNet")) -> None:
    from . import det as ours
    pkg = importlib.import_module("coperception.models.det")
    for name in classes:
        _swap(pkg, name, getattr(ours, name))
        sub = sys.modules.get(f"coperception.models.det.{name}")
        if sub is not None:
            _swap(sub, name, getattr(ours, name))
    # KD loss (FaFModule.get_kd_loss), corner loss (FaFModule.corner_loss) and the detection post-processing of
    # predict_all (apply_nms_det -> polygon NMS) on the fused kernels.  Optional: the utils modules pull heavy third-party
    # imports (shapely, matplotlib, nuscenes ...) that a model-only user may not have.
    try:
        mod = importlib.import_module("coperception.utils.CoDetModule")
        du = importlib.import_module("coperception.utils.detection_util")
        pp = importlib.import_module("coperception.utils.postprocess")
    except ImportError:
        mod = None
    if mod is not None:
        from . import kd, loss as ours_loss, post
        _swap(mod.FaFModule, "get_kd_loss", kd.get_kd_loss)
        _swap(mod.FaFModule, "corner_loss", ours_loss._corner_loss_method)
        _swap(mod, "apply_nms_det", post.apply_nms_det)          # CoDetModule does `from ...detection_util import *`
        _swap(du, "apply_nms_det", post.apply_nms_det)
        _swap(du, "late_fusion", post.late_fusion)
        _swap(du, "non_max_suppression", post.non_max_suppression)
        _swap(pp, "non_max_suppression", post.non_max_suppression)
    # focal classification loss (train_codet.py:173-176 builds it from coperception.utils.loss)
    try:
        lmod = importlib.import_module("coperception.utils.loss")
    except ImportError:
        lmod = None
    if lmod is not None:
        from . import loss as ours_loss
        _swap(lmod, "SoftmaxFocalClassificationLoss", ours_loss.SoftmaxFocalClassificationLoss)
    # BEV segmentation (tools/seg/*.py do `from coperception.models.seg import *`)
    try:
        seg_pkg = importlib.import_module("coperception.models.seg")
    except Exception:   # the seg package pulls optional dependencies; the detection tools do not need it
        return
    from .seg import SegDiscoNet
    _swap(seg_pkg, "DiscoNet", SegDiscoNet)
    sub = sys.modules.get("coperception.models.seg.DiscoNet")
    if sub is not None:
        _swap(sub, "DiscoNet", SegDiscoNet)


if __name__ == "__main__":
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m disconet_b200.patch <reference tool .py> [tool args...]")
    patch_coperception()
    tool = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(tool, run_name="__main__")
