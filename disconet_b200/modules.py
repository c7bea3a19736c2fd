"""Parameter containers with the reference's state_dict names, shapes and registration order.

The drop-in boundary includes checkpoint compatibility: `load_state_dict` of a reference
checkpoint (321 entries incl. the dead parameters that exist because the reference's LidarEncoder and
LidarDecoder each own a full Backbone -- IntermediateModelBase.py:24-25) and `Adam(model.parameters())`
state reload must line up.  These modules only *hold* parameters (stock nn.Conv2d / nn.BatchNorm
objects so default initialisation also matches); they have no forward -- compute happens in the CUDA
kernels driven by disconet.py.
"""
from __future__ import annotations

import torch.nn as nn


class ParamHolder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: compute runs in libdisco_b200, not in torch modules")


class PointwiseSeq(ParamHolder):
    """Names of the reference's Conv3D block (`conv3d`, `bn3d`; Backbone.py:280-292)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv3d = nn.Conv3d(cin, cout, kernel_size=(1, 1, 1), stride=1, padding=(0, 0, 0))
        self.bn3d = nn.BatchNorm3d(cout)


# (name, c_in, c_out, kernel, stride) in the reference's registration order (Backbone.py:11-47)
_BACKBONE_CONVS = [
    ("conv1_1", 32, 64, 3, 2), ("conv1_2", 64, 64, 3, 1),
    ("conv2_1", 64, 128, 3, 2), ("conv2_2", 128, 128, 3, 1),
    ("conv3_1", 128, 256, 3, 2), ("conv3_2", 256, 256, 3, 1),
    ("conv4_1", 256, 512, 3, 2), ("conv4_2", 512, 512, 3, 1),
    ("conv5_1", 512 + 256, 256, 3, 1), ("conv5_2", 256, 256, 3, 1),
    ("conv6_1", 256 + 128, 128, 3, 1), ("conv6_2", 128, 128, 3, 1),
    ("conv7_1", 128 + 64, 64, 3, 1), ("conv7_2", 64, 64, 3, 1),
    ("conv8_1", 64 + 32, 32, 3, 1), ("conv8_2", 32, 32, 3, 1),
]
_BACKBONE_BNS = [("bn1_1", 64), ("bn1_2", 64), ("bn2_1", 128), ("bn2_2", 128), ("bn3_1", 256), ("bn3_2", 256),
                 ("bn4_1", 512), ("bn4_2", 512), ("bn5_1", 256), ("bn5_2", 256), ("bn6_1", 128), ("bn6_2", 128),
                 ("bn7_1", 64), ("bn7_2", 64), ("bn8_1", 32), ("bn8_2", 32)]


class BackboneParams(ParamHolder):
    """All parameters of the reference `Backbone` (Backbone.py:9-87)."""

    def __init__(self, height_feat_size=13, compress_level=0):
        super().__init__()
        self.conv_pre_1 = nn.Conv2d(height_feat_size, 32, kernel_size=3, stride=1, padding=1)
        self.conv_pre_2 = nn.Conv2d(32, 32, kernel_size=3, stride=1, padding=1)
        self.bn_pre_1 = nn.BatchNorm2d(32)
        self.bn_pre_2 = nn.BatchNorm2d(32)
        self.conv3d_1 = PointwiseSeq(64, 64)
        self.conv3d_2 = PointwiseSeq(128, 128)
        for name, cin, cout, k, s in _BACKBONE_CONVS:
            setattr(self, name, nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=k // 2))
        for name, c in _BACKBONE_BNS:
            setattr(self, name, nn.BatchNorm2d(c))
        self.compress_level = compress_level
        if compress_level > 0:
            assert compress_level <= 8
            cc = 256 // (2 ** compress_level)
            self.com_compresser = nn.Conv2d(256, cc, kernel_size=1, stride=1)
            self.bn_compress = nn.BatchNorm2d(cc)
            self.com_decompresser = nn.Conv2d(cc, 256, kernel_size=1, stride=1)
            self.bn_decompress = nn.BatchNorm2d(256)


class ClassificationHeadParams(ParamHolder):
    """DetModelBase.py:268-298."""

    def __init__(self, channel, category_num, anchors):
        super().__init__()
        self.conv1 = nn.Conv2d(channel, channel, kernel_size=3, stride=1, padding=1)
        self.conv2 = nn.Conv2d(channel, category_num * anchors, kernel_size=1, stride=1, padding=0)
        self.bn1 = nn.BatchNorm2d(channel)


class RegressionHeadParams(ParamHolder):
    """DetModelBase.py:301-351 (binary + only_det branch): box_prediction.{0,1,3}."""

    def __init__(self, channel, out_ch):
        super().__init__()
        self.box_prediction = nn.Sequential(
            nn.Conv2d(channel, channel, kernel_size=3, stride=1, padding=1),
            nn.BatchNorm2d(channel),
            nn.ReLU(),
            nn.Conv2d(channel, out_ch, kernel_size=1, stride=1, padding=0),
        )


class PixelWeightedFusionParams(ParamHolder):
    """DiscoNet.py:132-146."""

    def __init__(self, channel):
        super().__init__()
        self.conv1_1 = nn.Conv2d(channel * 2, 128, kernel_size=1, stride=1, padding=0)
        self.bn1_1 = nn.BatchNorm2d(128)
        self.conv1_2 = nn.Conv2d(128, 32, kernel_size=1, stride=1, padding=0)
        self.bn1_2 = nn.BatchNorm2d(32)
        self.conv1_3 = nn.Conv2d(32, 8, kernel_size=1, stride=1, padding=0)
        self.bn1_3 = nn.BatchNorm2d(8)
        self.conv1_4 = nn.Conv2d(8, 1, kernel_size=1, stride=1, padding=0)
