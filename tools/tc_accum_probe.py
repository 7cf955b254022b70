#!/usr/bin/env python
"""Is the tensor-core path's residual error an ACCUMULATION effect?  Operands that are exactly representable in the
16-bit format (lo planes = 0) leave only the fp32 accumulation inside tcgen05.mma as an error source; the result is
compared with an fp64 convolution.  Reported: rms relative error and the mean SIGNED error along the result's own
direction (a round-toward-zero accumulator shows up as a negative bias that grows with the number of MMAs).

    python tools/tc_accum_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dkt_stereo_b200 import ops, _lib as L  # noqa: E402

dev = torch.device("cuda:0")
dt = L.split_dtype()


def run(Cin, N, H, W, k, impl, two_mma, positive):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, Cin, H, W, generator=g)
    wt = torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    if positive:                       # all products positive: the accumulator only grows (worst case for truncation)
        x, wt = x.abs(), wt.abs()
    x = x.to(dt).float()
    wt = wt.to(dt).float()
    ref = torch.nn.functional.conv2d(x.double().to(dev), wt.double().to(dev), None, padding=k // 2)
    xn = x.permute(0, 2, 3, 1).contiguous().to(dev)
    out = torch.zeros(1, H, W, N, device=dev)
    W_ = ops.pack_conv(wt.to(dev), None, tc=(impl == "tc"))
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, None, None, 0, N))
    if impl == "tc":
        hi = xn.to(dt).contiguous()
        lo = None if two_mma else torch.zeros_like(hi)
        src = L.tensor_slice(None, hi, lo)
    else:
        src = L.tensor_slice(xn, None, None)
    ops.conv2d([src], W_, e, 1, H, W, impl)
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).double()
    err = got - ref
    scale = ref.abs().mean()
    rms = float((err ** 2).mean().sqrt() / scale)
    signed = float((err * torch.sign(ref)).mean() / scale)
    return rms, signed


for positive in (False, True):
    print(f"--- {'all-positive operands' if positive else 'zero-mean operands'} (exact 16-bit operands, lo = 0) ---")
    for Cin, N, k in ((64, 64, 3), (384, 256, 3), (384, 128, 1)):
        for impl, two in (("simt", False), ("tc", True), ("tc", False)):
            rms, signed = run(Cin, N, 40, 48, k, impl, two, positive)
            name = "fp32 CUDA cores" if impl == "simt" else ("tcgen05 2 MMA/K16" if two else "tcgen05 3 MMA/K16 (lo = 0)")
            print(f"conv{k}x{k} {Cin:3d}->{N:3d}  {name:28s} rms rel err {rms:.3e}   mean signed err along result {signed:+.3e}")
