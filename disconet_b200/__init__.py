"""disconet_b200 -- B200-native (sm_100a) implementation of DiscoNet's collaborative-perception hot path.

Public surface = the reference's model classes for this path (`DiscoNet`, `FaFNet`, `TeacherNet`; `seg.SegDiscoNet` for
the BEV-segmentation variant) plus the data-format entry points (`voxelize_occupy`, `bev_scatter`).  Everything computes in libdisco_b200.so.
"""
from .det import DiscoNet, FaFNet, TeacherNet, AgentWeightList  # noqa: F401
from .voxel import voxelize_occupy, voxelize_occupy_batched, bev_scatter, bev_scatter_batched  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from . import seg  # noqa: F401  (disconet_b200.seg.SegDiscoNet == coperception.models.seg.DiscoNet)

__all__ = ["DiscoNet", "FaFNet", "TeacherNet", "voxelize_occupy", "voxelize_occupy_batched", "bev_scatter", "bev_scatter_batched",
           "HostPipeline"]
