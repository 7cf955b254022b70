#!/usr/bin/env python
"""Stand-alone timing of the IGEV combined lookup at cfg3 (B8, 136x240, 8x48 geometry volume): plain vs fused with
convc1.  Run on the GPU box (optionally under ncu): python tools/geo_lookup_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dkt_stereo_b200 import ops, _lib as L

dev = torch.device("cuda:0")
B, h, w, Cg, D = 8, 136, 240, 8, 48
g = torch.Generator(device=dev).manual_seed(0)
geo = (torch.randn(B, h, w, Cg, D, device=dev, generator=g), torch.randn(B, h, w, Cg, D // 2, device=dev, generator=g))
init = (torch.randn(B, h, w, w, device=dev, generator=g), torch.randn(B, h, w, w // 2, device=dev, generator=g))
disp = torch.rand(B, h, w, device=dev, generator=g) * 48
wt = torch.randn(64, 162, 1, 1, device=dev, generator=g) / 12
bias = torch.randn(64, device=dev, generator=g)
W = ops.pack_conv(wt, bias, cin_pad=192, tc=True)
out_hi = torch.zeros(B, h, w, 64, device=dev, dtype=L.split_dtype())
out_lo = torch.zeros_like(out_hi)
enc_out = L.tensor_slice(None, out_hi, out_lo, 0, 64)
plain = torch.zeros(B, h, w, 192, device=dev)
p_hi = torch.zeros(B, h, w, 192, device=dev, dtype=L.split_dtype())
p_lo = torch.zeros_like(p_hi)

def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

print("fused lookup + convc1 : %.1f us" % t(lambda: ops.geo_lookup_enc(geo, init, disp, 4, W, enc_out)))
print("plain lookup (f32+hi/lo, 192 ch): %.1f us" % t(lambda: ops.geo_lookup(geo, init, disp, 4, plain, "nhwc", out_hi=p_hi, out_lo=p_lo)))
print("plain lookup (f32 only): %.1f us" % t(lambda: ops.geo_lookup(geo, init, disp, 4, plain, "nhwc")))
