"""Feature / context encoders that feed the B200 hot path.

These stay in PyTorch (cuDNN) by design: the north-star keeps the feature
extractor on the host framework and only replaces the correlation volume,
lookup and GRU loop.  The module tree reproduces the parameter names of the
reference encoders so that DKT / RAFT-Stereo checkpoints load with
``strict=True``:

* ``BasicEncoder``      <- /root/reference/core/extractor.py:122-197
* ``MultiBasicEncoder`` <- /root/reference/core/extractor.py:199-300
* ``ResidualBlock``     <- /root/reference/core/extractor.py:6-60

The implementation is table driven (one stage spec per line) instead of the
reference's hand-unrolled constructors.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _make_norm(kind: str, ch: int, groups: int | None = None) -> nn.Module:
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "group":
        return nn.GroupNorm(num_groups=groups if groups is not None else ch // 8, num_channels=ch)
    if kind == "none":
        return nn.Sequential()
    raise ValueError(f"unknown norm_fn {kind!r}")


class ResidualBlock(nn.Module):
    """conv3x3-norm-relu x2 with an optional strided 1x1 projection shortcut."""

    def __init__(self, in_planes: int, planes: int, norm_fn: str = "group", stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_planes, planes, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _make_norm(norm_fn, planes)
        self.norm2 = _make_norm(norm_fn, planes)
        self.downsample = None
        if stride != 1 or in_planes != planes:
            # the shortcut norm is visible under two names in the state dict
            # (norm3.* and downsample.1.*) exactly as in the reference
            self.norm3 = _make_norm(norm_fn, planes)
            self.downsample = nn.Sequential(nn.Conv2d(in_planes, planes, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


def _init_encoder(module: nn.Module) -> None:
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
            if m.weight is not None:
                nn.init.constant_(m.weight, 1)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class _Trunk(nn.Module):
    """Shared stem: 7x7 conv + three residual stages (64, 96, 128 channels)."""

    # (attribute name, channels, "stride is 2 when downsample > threshold")
    _STAGES = (("layer1", 64, None), ("layer2", 96, 1), ("layer3", 128, 0))

    def __init__(self, norm_fn: str, downsample: int):
        super().__init__()
        self.norm_fn = norm_fn
        self.downsample = downsample
        self.norm1 = _make_norm(norm_fn, 64, groups=8)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=1 + (downsample > 2), padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        in_planes = 64
        for name, dim, thr in self._STAGES:
            stride = 1 if thr is None else 1 + (downsample > thr)
            setattr(self, name, self._stage(in_planes, dim, stride))
            in_planes = dim

    def _stage(self, in_planes: int, dim: int, stride: int) -> nn.Sequential:
        return nn.Sequential(ResidualBlock(in_planes, dim, self.norm_fn, stride),
                             ResidualBlock(dim, dim, self.norm_fn, 1))

    def trunk(self, x):
        x = self.relu1(self.norm1(self.conv1(x)))
        return self.layer3(self.layer2(self.layer1(x)))


class BasicEncoder(_Trunk):
    """fnet: instance-normalised matching features, ``output_dim`` channels at 1/2^downsample."""

    def __init__(self, output_dim: int = 128, norm_fn: str = "batch", dropout: float = 0.0, downsample: int = 3):
        super().__init__(norm_fn, downsample)
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        _init_encoder(self)

    def forward(self, x, dual_inp: bool = False):
        pair = isinstance(x, (tuple, list))
        if pair:
            n = x[0].shape[0]
            x = torch.cat(list(x), dim=0)
        x = self.conv2(self.trunk(x))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return x.split(n, dim=0) if pair else x


class MultiBasicEncoder(_Trunk):
    """cnet: per-scale (hidden, context) heads at 1/4, 1/8, 1/16 for n_downsample=2."""

    def __init__(self, output_dim=((128, 128, 128),), norm_fn: str = "batch", dropout: float = 0.0, downsample: int = 3,
                 head_names=("outputs08", "outputs16", "outputs32")):
        """``head_names``: RAFT-Stereo calls the three head lists outputs08/16/32
        (core/extractor.py:234-258), IGEV-Stereo outputs04/08/16 (igev_stereo/extractor.py:239-256)."""
        super().__init__(norm_fn, downsample)
        self.head_names = tuple(head_names)
        self.layer4 = self._stage(128, 128, 2)
        self.layer5 = self._stage(128, 128, 2)

        def heads(idx: int, with_block: bool) -> nn.ModuleList:
            out = []
            for dim in output_dim:
                conv = nn.Conv2d(128, dim[idx], 3, padding=1)
                out.append(nn.Sequential(ResidualBlock(128, 128, norm_fn, 1), conv) if with_block else conv)
            return nn.ModuleList(out)

        setattr(self, self.head_names[0], heads(2, True))
        setattr(self, self.head_names[1], heads(1, True))
        setattr(self, self.head_names[2], heads(0, False))
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        _init_encoder(self)

    def forward(self, x, dual_inp: bool = False, num_layers: int = 3):
        x = self.trunk(x)
        v = None
        if dual_inp:
            v = x
            x = x[: x.shape[0] // 2]
        h0, h1, h2 = (getattr(self, n) for n in self.head_names)
        scales = [[f(x) for f in h0]]
        if num_layers >= 2:
            y = self.layer4(x)
            scales.append([f(y) for f in h1])
        if num_layers >= 3:
            z = self.layer5(y)
            scales.append([f(z) for f in h2])
        if dual_inp:
            scales.append(v)
        return tuple(scales)
