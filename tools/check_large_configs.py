"""Functional check at the shapes of BASELINE configs[3] / configs[4] (RAFT-Stereo 736x1280 batch 8, IGEV-Stereo
1024x1536): runs, finite output, time and peak memory.  Run on the GPU box: python tools/check_large_configs.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from argparse import Namespace
from bench import IGEV_CFG, RAFT_CFG
from dkt_stereo_b200.igev_stereo import IGEVStereo
from dkt_stereo_b200.raft_stereo import RAFTStereo
from dkt_stereo_b200.synthetic import synthetic_pair
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = IGEVStereo(Namespace(mixed_precision=False, corr_implementation="b200", **IGEV_CFG)).eval().to(dev)
im1, im2 = (t.to(dev) for t in synthetic_pair(2, 1024, 1536, seed=3, mode="shift"))
for _ in range(3):
    _, up = m(im1, im2, iters=22, test_mode=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); _, up = m(im1, im2, iters=22, test_mode=True); b.record(); torch.cuda.synchronize()
print("IGEV 1024x1536 B2 22 iters:", tuple(up.shape), "finite", bool(torch.isfinite(up).all()), "%.1f ms" % a.elapsed_time(b), "mean disp %.2f" % float(-up.mean()))
# the same model with the volume stage / hourglass in PyTorch fp32 (DKT_NATIVE_VOLUME=0): parity of the native stage at this shape
os.environ["DKT_NATIVE_VOLUME"] = "0"
m0 = IGEVStereo(Namespace(mixed_precision=False, corr_implementation="b200", **IGEV_CFG)).eval().to(dev)
os.environ["DKT_NATIVE_VOLUME"] = "1"
m0.load_state_dict(m.state_dict(), strict=True)
_, up0 = m0(im1, im2, iters=22, test_mode=True)
torch.cuda.synchronize()
print("IGEV 1024x1536 native vs PyTorch volume stage: mean-abs %.3e px, max-abs %.3e px (gate 1e-3 mean)" %
      (float((up - up0).abs().mean()), float((up - up0).abs().max())))
del m0, up0
# cfg5 batch (4 pairs per GPU)
im1, im2 = (t.to(dev) for t in synthetic_pair(4, 1024, 1536, seed=4, mode="shift"))
for _ in range(2):
    _, up = m(im1, im2, iters=22, test_mode=True)
torch.cuda.synchronize()
a.record(); _, up = m(im1, im2, iters=22, test_mode=True); b.record(); torch.cuda.synchronize()
print("IGEV 1024x1536 B4 22 iters (cfg5 per-GPU batch):", tuple(up.shape), "finite", bool(torch.isfinite(up).all()), "%.1f ms" % a.elapsed_time(b))
del m
torch.cuda.empty_cache()
r = RAFTStereo(Namespace(mixed_precision=False, **dict(RAFT_CFG, corr_implementation="b200"))).eval().to(dev)
im1, im2 = (t.to(dev) for t in synthetic_pair(8, 736, 1280, seed=3, mode="shift"))
for _ in range(3):
    _, up = r(im1, im2, iters=32, test_mode=True)
torch.cuda.synchronize()
a.record(); _, up = r(im1, im2, iters=32, test_mode=True); b.record(); torch.cuda.synchronize()
print("RAFT 736x1280 B8 32 iters:", tuple(up.shape), "finite", bool(torch.isfinite(up).all()), "%.1f ms" % a.elapsed_time(b), "mean disp %.2f" % float(-up.mean()))
print("max mem GB %.1f" % (torch.cuda.max_memory_allocated() / 1e9))
