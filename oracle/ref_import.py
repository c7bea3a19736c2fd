"""Import the UNMODIFIED reference model classes from /root/reference (build container only).

TEST INFRASTRUCTURE.  `import coperception` fails here (shapely/nuscenes/matplotlib are absent), so
the package __init__ is bypassed exactly as SURVEY.md §8(c) describes.  /root/reference does not
exist on the GPU box: nothing that runs there may call this module.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REF_PKG = "/root/reference/coperception/coperception"


def available() -> bool:
    return os.path.isdir(REF_PKG)


def install_bypass(mock_heavy=False):
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
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
