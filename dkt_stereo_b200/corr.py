"""Correlation-volume operators behind the reference's ``corr_implementation`` plug-in point.

Every class keeps the reference call shape (reference meta_arch/raft_stereo/raft_stereo.py:118-142):
``Block(fmap1, fmap2, num_levels=, radius=)`` then ``block(coords (B,2,h,w)) -> (B, L*(2r+1), h, w)``.

* ``B200CorrBlock1D``      replaces ``CorrBlock1D`` (reference core/corr.py:110-156)
* ``corr_sampler_forward`` / ``corr_sampler_backward`` / ``CorrSampler`` replace the un-vendored ``corr_sampler``
  extension that ``CorrBlockFast1D`` binds (reference core/corr.py:17-29,49): ``forward(volume, coords, radius) ->
  (out,)``, ``backward(volume, coords, grad_output, radius) -> (grad_volume,)``
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

from . import _lib as L
from . import ops


class B200CorrBlock1D:
    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4, radius: int = 4,
                 impl: str = "tc", scale: float | None = None):
        L.require_device(fmap1)
        self.num_levels, self.radius = num_levels, radius
        D = fmap1.shape[1]
        # the reference divides by sqrt(D) (core/corr.py:155); IGEV's variant does not (geometry.py:67-69)
        self.scale = (1.0 / math.sqrt(D)) if scale is None else scale
        self.corr_pyramid: List[torch.Tensor] = ops.corr1d_build(fmap1.float(), fmap2.float(), num_levels,
                                                                 self.scale, impl=impl)

    def __call__(self, coords: torch.Tensor) -> torch.Tensor:
        B, _, H, W = coords.shape
        cx = coords[:, 0].contiguous().float()
        out = torch.empty(B, self.num_levels * (2 * self.radius + 1), H, W, device=coords.device, dtype=torch.float32)
        ops.corr1d_lookup(self.corr_pyramid, cx, self.radius, out, out_layout="nchw")
        return out


def corr_sampler_forward(volume: torch.Tensor, coords: torch.Tensor, radius: int) -> Tuple[torch.Tensor]:
    """Drop-in for ``corr_sampler.forward``: volume (B,H,W1,W2_l) one pyramid level, coords (B,1,H,W1)
    already divided by 2^l  ->  ((B, 2r+1, H, W1),)."""
    L.require_device(volume)
    B, H, W1, _ = volume.shape
    out = torch.empty(B, 2 * radius + 1, H, W1, device=volume.device, dtype=torch.float32)
    ops.corr1d_lookup([volume.contiguous().float()], coords[:, 0].contiguous().float(), radius, out, out_layout="nchw")
    return (out,)


def corr_sampler_backward(volume: torch.Tensor, coords: torch.Tensor, grad_output: torch.Tensor,
                          radius: int) -> Tuple[torch.Tensor]:
    """Drop-in for ``corr_sampler.backward`` (reference core/corr.py:25-29): volume (B,H,W1,W2_l), coords (B,1,H,W1),
    grad_output (B, 2r+1, H, W1)  ->  (grad_volume (B,H,W1,W2_l),).  Only the volume's SHAPE is used."""
    L.require_device(grad_output)
    B, H, W1, W2 = volume.shape
    assert grad_output.shape == (B, 2 * radius + 1, H, W1), grad_output.shape
    g = grad_output.contiguous().float()
    cx = coords[:, 0].contiguous().float()
    grad_volume = torch.empty(B, H, W1, W2, device=g.device, dtype=torch.float32)
    L.check(L.load().dkt_corr1d_lookup_backward(g.data_ptr(), cx.data_ptr(), radius, grad_volume.data_ptr(),
                                                B, H, W1, W2, L.stream_ptr()), "corr1d_lookup_backward")
    return (grad_volume,)


class CorrSampler(torch.autograd.Function):
    """The reference's autograd wrapper (core/corr.py:17-29) on this library's two kernels."""

    @staticmethod
    def forward(ctx, volume, coords, radius):
        ctx.save_for_backward(volume, coords)
        ctx.radius = radius
        corr, = corr_sampler_forward(volume, coords, radius)
        return corr

    @staticmethod
    def backward(ctx, grad_output):
        volume, coords = ctx.saved_tensors
        grad_volume, = corr_sampler_backward(volume, coords, grad_output.contiguous(), ctx.radius)
        return grad_volume, None, None


class Combined_Geo_Encoding_Volume:
    """Drop-in for reference meta_arch/igev_stereo/geometry.py:6-69: same constructor
    ``(init_fmap1, init_fmap2, geo_volume, num_levels=2, radius=4)`` and call
    ``(disp (B,1,h,w), coords (B,h,w,1)) -> (B, 2*9*(2r+1), h, w)`` fp32.  ``coords`` is accepted for
    signature compatibility; the kernel derives the pixel x coordinate itself (the reference always
    passes ``arange(w)``, igev_stereo.py:195)."""

    def __init__(self, init_fmap1: torch.Tensor, init_fmap2: torch.Tensor, geo_volume: torch.Tensor,
                 num_levels: int = 2, radius: int = 4, impl: str = "tc"):
        L.require_device(init_fmap1)
        if num_levels != 2:
            raise NotImplementedError("the B200 geometry lookup serves the shipped 2-level configuration")
        self.num_levels, self.radius = num_levels, radius
        self.init_corr_pyramid = ops.corr1d_build(init_fmap1.float(), init_fmap2.float(), num_levels, 1.0, impl=impl)
        self.geo_volume_pyramid = list(ops.geo_pool(geo_volume.float()))

    def __call__(self, disp: torch.Tensor, coords: torch.Tensor | None = None) -> torch.Tensor:
        B, _, H, W = disp.shape
        Cg = self.geo_volume_pyramid[0].shape[3]
        out = torch.empty(B, self.num_levels * (Cg + 1) * (2 * self.radius + 1), H, W, device=disp.device, dtype=torch.float32)
        d = disp[:, 0].contiguous().float().clone()
        ops.geo_lookup(self.geo_volume_pyramid, self.init_corr_pyramid, d, self.radius, out, out_layout="nchw")
        return out
