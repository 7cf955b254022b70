"""Stand-alone timing of the RAFT-Stereo lookup at cfg2 (B8, 136x240, 4 levels): plain vs fused with convc1.
Run on the GPU box: python tools/lookup_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dkt_stereo_b200 import ops, _lib as L
dev = torch.device("cuda:0")
B, h, w = 8, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
pyr = [torch.randn(B, h, w, w >> l, device=dev, generator=g) for l in range(4)]
cx = (torch.arange(w, device=dev).float().view(1, 1, w) - torch.rand(B, h, w, device=dev, generator=g) * 40).contiguous()
wt = torch.randn(64, 36, 1, 1, device=dev, generator=g) / 6
bias = torch.randn(64, device=dev, generator=g)
hi = torch.zeros(B, h, w, 64, device=dev, dtype=L.split_dtype()); lo = torch.zeros_like(hi)
out = L.tensor_slice(None, hi, lo, 0, 64)
def t(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
Wf = ops.pack_conv(wt, bias, cin_pad=64, tc=False)
print("fused lookup + convc1: %.1f us" % t(lambda: ops.corr1d_lookup_enc(pyr, cx, 4, Wf, out)))
plain = torch.zeros(B, h, w, 36, device=dev)
print("plain lookup (36 ch fp32): %.1f us" % t(lambda: ops.corr1d_lookup(pyr, cx, 4, plain, "nhwc")))
