"""Host-side helpers kept from the reference's ``core/utils/utils.py`` surface."""
from __future__ import annotations

import torch
import torch.nn.functional as F


class InputPadder:
    """Pads images so that H and W are divisible by ``divis_by`` (reference core/utils/utils.py:7-26)."""

    def __init__(self, dims, mode: str = "sintel", divis_by: int = 8):
        self.ht, self.wd = dims[-2:]
        pad_ht = (-self.ht) % divis_by
        pad_wd = (-self.wd) % divis_by
        if mode == "sintel":
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, pad_ht // 2, pad_ht - pad_ht // 2]
        else:
            self._pad = [pad_wd // 2, pad_wd - pad_wd // 2, 0, pad_ht]

    def pad(self, *inputs):
        assert all(x.ndim == 4 for x in inputs)
        return [F.pad(x, self._pad, mode="replicate") for x in inputs]

    def unpad(self, x):
        assert x.ndim == 4
        ht, wd = x.shape[-2:]
        return x[..., self._pad[2]:ht - self._pad[3], self._pad[0]:wd - self._pad[1]]


def coords_grid(batch: int, ht: int, wd: int, device=None) -> torch.Tensor:
    """(B,2,H,W) with channel 0 = x, channel 1 = y (reference core/utils/utils.py:77-80)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)
