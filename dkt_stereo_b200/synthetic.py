"""Deterministic synthetic inputs and weights for benchmarks, smoke runs and parity tests.

There are no datasets or checkpoints offline, so the measured configuration is the one
SURVEY.md section 8(d) prescribes: uniform-noise image pairs (generator seed 1234) and
random-init weights.  ``synthetic_state_dict`` draws every tensor from a generator seeded by
the parameter *name*, so the same weights can be loaded into the reference model (when the
golden vectors are made, ``oracle/make_golden.py``) and into this engine on a box where the
reference does not exist -- independent of module construction order.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Mapping, Tuple

import torch


def synthetic_pair(batch: int, height: int, width: int, seed: int = 1234,
                   mode: str = "noise", max_disp: float = 24.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Two (B,3,H,W) fp32 images in [0,255].

    mode="noise": independent uniform noise (worst case for matching).
    mode="shift": the right image is the left one shifted by a smooth disparity ramp in
    [0,max_disp] px, which gives the matcher a true solution (a contractive regime).
    """
    g = torch.Generator().manual_seed(seed)
    im1 = torch.rand(batch, 3, height, width, generator=g) * 255
    if mode == "noise":
        im2 = torch.rand(batch, 3, height, width, generator=g) * 255
        return im1, im2
    if mode != "shift":
        raise ValueError(mode)
    # low-pass the noise a little so sub-pixel shifts are meaningful
    k = torch.ones(3, 1, 5, 5) / 25.0
    im1 = torch.nn.functional.conv2d(torch.nn.functional.pad(im1, (2, 2, 2, 2), mode="reflect"), k, groups=3)
    im1 = (im1 - im1.amin()) / (im1.amax() - im1.amin()) * 255
    ys = torch.linspace(0, 1, height).view(1, 1, height, 1)
    xs = torch.arange(width, dtype=torch.float32).view(1, 1, 1, width)
    disp = max_disp * (0.25 + 0.75 * ys).expand(batch, 1, height, width)
    src = xs + disp                      # right(x) = left(x + d)
    gx = 2 * src / (width - 1) - 1
    gy = (2 * torch.arange(height, dtype=torch.float32) / (height - 1) - 1).view(1, 1, height, 1).expand_as(gx)
    grid = torch.stack([gx[:, 0], gy[:, 0]], dim=-1)
    im2 = torch.nn.functional.grid_sample(im1, grid, mode="bilinear", padding_mode="border", align_corners=True)
    return im1.contiguous(), im2.contiguous()


def _key_generator(name: str, seed: int) -> torch.Generator:
    # the encoders register the shortcut norm twice (norm3.* and downsample.1.*): one tensor, two names
    name = name.replace(".downsample.1.", ".norm3.")
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)


def synthetic_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """name -> fp32 tensor, drawn per key.

    Distributions follow what the reference's constructors produce: encoder convs
    (cnet./fnet./feature.) Kaiming-normal fan_out (core/extractor.py:155-162), every other conv /
    deconv PyTorch's default U(+-1/sqrt(fan_in)); norm scales near 1, running stats near (0,1).
    """
    out: Dict[str, torch.Tensor] = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        g = _key_generator(name, seed)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = torch.zeros(shape, dtype=torch.long)
        elif leaf == "running_mean":
            out[name] = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            out[name] = 1.0 + 0.1 * torch.rand(shape, generator=g)
        elif len(shape) >= 3:                                   # conv / deconv kernel
            rf = math.prod(shape[2:])
            if name.startswith(("cnet.", "fnet.", "module.cnet.", "module.fnet.")):
                out[name] = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] * rf))
            else:
                bound = 1.0 / math.sqrt(shape[1] * rf)
                out[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif leaf == "weight":                                   # norm scale
            out[name] = 1.0 + 0.05 * torch.randn(shape, generator=g)
        else:                                                    # bias
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return out


def shapes_of(state_dict: Mapping[str, torch.Tensor]) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(v.shape) for k, v in state_dict.items()}
