"""ResNet-FPN local feature CNN (stays PyTorch / cuDNN; BASELINE.json north_star).

Same topology and state_dict key names as the reference backbones (src/loftr/backbone/resnet_fpn.py:
ResNetFPN_8_2 :43-118, ResNetFPN_16_4 :121-199) so released checkpoints load unchanged, but written as
one depth-parametrised module: a stem, `depth` residual stages, and a top-down pathway that stops
`out_levels` stages above the input resolution.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv(cin, cout, k, stride=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)


class ResidualUnit(nn.Module):
    """Two 3x3 conv+BN with a (possibly strided, 1x1-projected) skip; keys conv1/bn1/conv2/bn2/downsample.{0,1}."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = _conv(cin, cout, 3, stride)
        self.conv2 = _conv(cout, cout, 3)
        self.bn1 = nn.BatchNorm2d(cout)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None if stride == 1 else nn.Sequential(_conv(cin, cout, 1, stride), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        skip = x if self.downsample is None else self.downsample(x)
        return F.relu(skip + y)


class ResNetFPN(nn.Module):
    """depth=3 -> outputs at 1/8 (coarse) and 1/2 (fine); depth=4 -> 1/16 and 1/4."""

    def __init__(self, initial_dim, block_dims):
        super().__init__()
        self.depth = len(block_dims)
        dims = list(block_dims)
        self.conv1 = nn.Conv2d(1, initial_dim, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(initial_dim)
        cin = initial_dim
        for lvl, d in enumerate(dims, start=1):
            stride = 1 if lvl == 1 else 2
            setattr(self, f"layer{lvl}", nn.Sequential(ResidualUnit(cin, d, stride), ResidualUnit(d, d, 1)))
            cin = d
        top = self.depth
        setattr(self, f"layer{top}_outconv", _conv(dims[top - 1], dims[top - 1], 1))
        # two top-down merge steps: (top-1) and (top-2)
        for lvl in (top - 1, top - 2):
            up = dims[lvl]          # channels of the level above (index lvl -> level lvl+1)
            here = dims[lvl - 1]
            setattr(self, f"layer{lvl}_outconv", _conv(here, up, 1))
            setattr(self, f"layer{lvl}_outconv2", nn.Sequential(
                _conv(up, up, 3), nn.BatchNorm2d(up), nn.LeakyReLU(), _conv(up, here, 3)))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        feats = []
        y = F.relu(self.bn1(self.conv1(x)))
        for lvl in range(1, self.depth + 1):
            y = getattr(self, f"layer{lvl}")(y)
            feats.append(y)
        top = self.depth
        coarse = getattr(self, f"layer{top}_outconv")(feats[top - 1])
        out = coarse
        for lvl in (top - 1, top - 2):
            up = F.interpolate(out, scale_factor=2.0, mode="bilinear", align_corners=True)
            lat = getattr(self, f"layer{lvl}_outconv")(feats[lvl - 1])
            out = getattr(self, f"layer{lvl}_outconv2")(lat + up)
        return [coarse, out]


def build_backbone(config):
    """Mirror of src/loftr/backbone/__init__.py:4-11."""
    if config["backbone_type"] != "ResNetFPN":
        raise ValueError(f"LOFTR.BACKBONE_TYPE {config['backbone_type']} not supported.")
    res = tuple(config["resolution"])
    dims = list(config["resnetfpn"]["block_dims"])
    if res == (8, 2):
        return ResNetFPN(config["resnetfpn"]["initial_dim"], dims[:3])
    if res == (16, 4):
        return ResNetFPN(config["resnetfpn"]["initial_dim"], dims[:4])
    raise ValueError(f"resolution {res} not supported")
