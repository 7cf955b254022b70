#!/usr/bin/env python
"""Probe (GPU box): IGEV's PyTorch feature side (MobileNetV2 pyramid, stems, match convs) in NCHW vs channels_last,
fp32 (no TF32).  python tools/igev_feature_layout_probe.py"""
import os, sys, json
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import IGEV_CFG
from dkt_stereo_b200.igev_stereo import IGEVStereo
from dkt_stereo_b200.raft_stereo import _fp32_math
from dkt_stereo_b200.synthetic import synthetic_pair

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = IGEVStereo(Namespace(mixed_precision=False, corr_implementation="b200", **IGEV_CFG)).eval().to(dev)
im1, im2 = (t.to(dev) for t in synthetic_pair(8, 544, 960, seed=1234))


def side(i1, i2):
    fl, fr = m.feature(i1), m.feature(i2)
    s2 = m.stem_2(i1); s4 = m.stem_4(s2); s4y = m.stem_4(m.stem_2(i2))
    a = torch.cat((fl[0], s4), 1); b = torch.cat((fr[0], s4y), 1)
    return m.desc(m.conv(a)), m.desc(m.conv(b)), fl


def timed(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


res = {}
with torch.no_grad(), _fp32_math(True):
    i1 = (2 * (im1 / 255.0) - 1.0).contiguous(); i2 = (2 * (im2 / 255.0) - 1.0).contiguous()
    t, ref = timed(lambda: side(i1, i2)); res["nchw_ms"] = round(t, 2)
    torch.backends.cudnn.benchmark = True
    t, _ = timed(lambda: side(i1, i2)); res["nchw_cudnn_benchmark_ms"] = round(t, 2)
    torch.backends.cudnn.benchmark = False
    for mod in (m.feature, m.stem_2, m.stem_4, m.conv, m.desc):
        mod.to(memory_format=torch.channels_last)
    c1, c2 = i1.contiguous(memory_format=torch.channels_last), i2.contiguous(memory_format=torch.channels_last)
    t, out = timed(lambda: side(c1, c2)); res["channels_last_ms"] = round(t, 2)
    torch.backends.cudnn.benchmark = True
    t, out2 = timed(lambda: side(c1, c2)); res["channels_last_cudnn_benchmark_ms"] = round(t, 2)
    res["max_abs_diff_match_left"] = float((out[0] - ref[0]).abs().max())
print(json.dumps(res))
