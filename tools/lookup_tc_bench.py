#!/usr/bin/env python
"""Stand-alone timing of the fused lookups at cfg2 / cfg3 (B8, 136x240): exact-fp32 CUDA-core kernels (corr.cu) vs the
tensor-core kernels (lookup_tc.cu).  Run on the GPU box: python tools/lookup_tc_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dkt_stereo_b200 import ops, _lib as L  # noqa: E402

dev = torch.device("cuda:0")
B, h, w = 8, 136, 240
P = B * h * w
g = torch.Generator(device=dev).manual_seed(0)


def t(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


pyr = [torch.randn(B, h, w, w >> l, device=dev, generator=g) for l in range(4)]
cx = (torch.arange(w, device=dev).float().view(1, 1, w) - torch.rand(B, h, w, device=dev, generator=g) * 40).contiguous()
wt = torch.randn(64, 36, 1, 1, device=dev, generator=g) / 6
bias = torch.randn(64, device=dev, generator=g)
hi = torch.zeros(B, h, w, 64, device=dev, dtype=L.split_dtype())
lo = torch.zeros_like(hi)
Wf = ops.pack_conv(wt, bias, cin_pad=64, tc=False)
w_img, b_dev = ops.pack_lookup_tc(wt, bias)
plain = torch.zeros(B, h, w, 36, device=dev)
alg = lambda planes: P * (4 * 10 * 4 + 4 + planes * 64 * 2)      # noqa: E731
print("RAFT plain lookup (36 ch fp32)            : %6.1f us" % t(lambda: ops.corr1d_lookup(pyr, cx, 4, plain, "nhwc")))
for name, out, planes in (("hi+lo out", L.tensor_slice(None, hi, lo, 0, 64), 2), ("hi out   ", L.tensor_slice(None, hi, None, 0, 64), 1)):
    us = t(lambda: ops.corr1d_lookup_enc(pyr, cx, 4, Wf, out))
    print(f"RAFT fused fp32 FMAs, {name}           : {us:6.1f} us  ({alg(planes) / us / 1e3:6.0f} GB/s algorithmic)")
    for tp in (2, 1):
        us = t(lambda: ops.corr1d_lookup_enc_tc(pyr, cx, 4, w_img, b_dev, out, tp))
        print(f"RAFT fused tcgen05 taps x{tp}, {name}       : {us:6.1f} us  ({alg(planes) / us / 1e3:6.0f} GB/s algorithmic)")

geo = [torch.randn(B, h, w, 8, 48 >> l, device=dev, generator=g) for l in range(2)]
geo_dc = [x.permute(0, 1, 2, 4, 3).contiguous() for x in geo]
init = [torch.randn(B, h, w, w >> l, device=dev, generator=g) for l in range(2)]
disp = (torch.rand(B, h, w, device=dev, generator=g) * 48).contiguous()
wg = torch.randn(64, 162, 1, 1, device=dev, generator=g) / 12
Wg = ops.pack_conv(wg, bias, cin_pad=192, tc=False)
wg_img, _ = ops.pack_lookup_tc(wg, bias)
algg = lambda planes: P * (2 * 9 * 10 * 4 + 4 + planes * 64 * 2)  # noqa: E731
out1 = L.tensor_slice(None, hi, None, 0, 64)
us = t(lambda: ops.geo_lookup_enc(geo, init, disp, 4, Wg, out1))
print(f"IGEV fused fp32 FMAs, hi out              : {us:6.1f} us  ({algg(1) / us / 1e3:6.0f} GB/s algorithmic)")
for tp in (1, 2):
    us = t(lambda: ops.geo_lookup_enc_tc(geo_dc, init, disp, 4, wg_img, b_dev, out1, tp))
    print(f"IGEV fused tcgen05 taps x{tp}, hi out         : {us:6.1f} us  ({algg(1) / us / 1e3:6.0f} GB/s algorithmic)")
