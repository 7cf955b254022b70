#!/usr/bin/env python
"""Timing of one stride-1 3x3x3 hourglass layer at cfg3 sizes: exact-fp32 SIMT kernel vs the tensor-core form
(layout pass in, 2-D conv over depth planes, layout pass out).  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dkt_stereo_b200 import ops, _lib as L
from dkt_stereo_b200.igev_modules import Hourglass, ConvNormAct

dev = torch.device("cuda:0")
torch.manual_seed(0)
hg = Hourglass(8).eval().to(dev)


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


for name, m, shape in (("16->16 @ 24x68x120", hg.conv1[1], (8, 16, 24, 68, 120)), ("32->32 @ 12x34x60", hg.conv2[1], (8, 32, 12, 34, 60)),
                       ("48->48 @ 6x17x30", hg.conv3[1], (8, 48, 6, 17, 30))):
    v = torch.randn(*shape, device=dev)
    att = torch.randn(shape[0], shape[1], shape[3], shape[4], device=dev)
    bn = m.bn
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach()
    shift = (bn.bias - bn.running_mean * scale).detach()
    with torch.no_grad():
        ref = ops.conv3d_k3(v, m.conv.weight, scale, shift, 0.01, att, 1)
        out = hg._k3_tc(m, v, att)
        err = float((out - ref).abs().max()) / (float(ref.abs().max()) + 1e-9)
        us_simt = t(lambda: ops.conv3d_k3(v, m.conv.weight, scale, shift, 0.01, att, 1))
        us_tc = t(lambda: hg._k3_tc(m, v, att))
        # pieces
        B, CI, D, H, W = shape
        xh, xl, of = hg._tc_cache[("b", B, CI, CI, D, H, W, str(v.device))]
        lib = L.load()
        us_in = t(lambda: lib.dkt_ncdhw_to_ndhwc_pad(v.data_ptr(), xh.data_ptr(), xl.data_ptr(), B, CI, D, H, W, L.stream_ptr()))
        o = torch.empty_like(v)
        us_out = t(lambda: lib.dkt_ndhwc_pad_to_ncdhw(of.data_ptr(), att.data_ptr(), o.data_ptr(), B, CI, D, H, W, L.stream_ptr()))
    print(f"{name}: SIMT {us_simt:7.1f} us   TC total {us_tc:7.1f} us (layout in {us_in:6.1f}, out {us_out:6.1f}, conv ~{us_tc - us_in - us_out:7.1f})   rel max err {err:.2e}")
