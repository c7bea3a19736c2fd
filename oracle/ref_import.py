"""Import the UNMODIFIED reference model classes from /root/reference (build container only).

TEST INFRASTRUCTURE.  `import coperception` fails here (shapely/nuscenes/matplotlib are absent), so
the package __init__ is bypassed exactly as SURVEY.md §8(c) describes.  /root/reference does not
exist on the GPU box: nothing that runs there may call this module.
"""
import os
import sys
import types
from unittest.mock import MagicMock

_LIVE = "/root/reference/coperception/coperception"
_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "coperception")   # oracle/stage_ref.py (GPU box)
REF_PKG = _LIVE if os.path.isdir(_LIVE) else _STAGED


def available() -> bool:
    return os.path.isdir(REF_PKG)


def install_stub_shapely():
    """shapely is absent from this image: give the reference's postprocess.py a Polygon built on the oracle's float64
    quad clip (oracle/post_oracle.py::StubPolygon).  Call before importing coperception.utils.*."""
    from oracle.post_oracle import StubPolygon
    geo = types.ModuleType("shapely.geometry")
    geo.Polygon = StubPolygon
    sh = types.ModuleType("shapely")
    sh.geometry = geo
    sys.modules["shapely"] = sh
    sys.modules["shapely.geometry"] = geo


def install_bypass(mock_heavy=False):
    if not available():
        raise RuntimeError("reference tree not present (neither /root/reference nor the staged oracle/_ref copy)")
    if "coperception" not in sys.modules or not getattr(sys.modules["coperception"], "__path__", None):
        pkg = types.ModuleType("coperception")
        pkg.__path__ = [REF_PKG]
        sys.modules["coperception"] = pkg
    if mock_heavy:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "shapely", "shapely.geometry",
                     "nuscenes", "nuscenes.utils", "nuscenes.utils.data_classes",
                     "nuscenes.utils.geometry_utils", "pyquaternion", "mmcv", "mmcv.utils",
                     "terminaltables", "seaborn"):
            sys.modules.setdefault(name, MagicMock())


def reference_classes():
    install_bypass()
    from coperception.models.det.DiscoNet import DiscoNet
    from coperception.models.det.FaFNet import FaFNet
    from coperception.models.det.TeacherNet import TeacherNet
    from coperception.configs.Config import Config
    return DiscoNet, FaFNet, TeacherNet, Config
