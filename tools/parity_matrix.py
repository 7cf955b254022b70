#!/usr/bin/env python
"""A/B of the engine's precision policies on one GPU box: end-to-end error against the reference goldens
(tests/golden/raft_fwd_cfg2*.npz, igev_fwd_cfg3.npz: the headline workload, two weight / image samples) and the step
time of the headline benchmark, one subprocess per policy (the knobs are read at engine construction / library load).

    python tools/parity_matrix.py [--no-bench] > profiles/r2_parity_matrix.txt
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dkt_stereo_b200", "lib")

POLICIES = [
    ("f16: GRUs 2 MMA, motion enc 2 MMA (default)", {}),
    ("f16: fine GRU 2 MMA, coarse GRUs 1 MMA, motion enc 2 MMA", {"DKT_COARSE_GRU_TERMS": "1"}),
    ("f16 default, single accumulator (DKT_ACC_SPLIT=0)", {"DKT_ACC_SPLIT": "0"}),
    ("f16: coarse GRUs 1 MMA, single accumulator", {"DKT_COARSE_GRU_TERMS": "1", "DKT_ACC_SPLIT": "0"}),
    ("f16: GRUs 2 MMA, motion enc 3 MMA", {"DKT_MENC_TERMS": "3"}),
    ("f16: GRUs 3 MMA, motion enc 2 MMA", {"DKT_GRU_TERMS": "3"}),
    ("f16: everything 3 MMA", {"DKT_GRU_TERMS": "3", "DKT_MENC_TERMS": "3"}),
    ("bf16 build: everything 3 MMA (round 1)", {"DKT_STEREO_LIB": os.path.join(LIB, "libdkt_stereo_b200_bf16.so")}),
]

CHILD = r"""
import sys, os, json
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch
from argparse import Namespace
from helpers import load_golden, golden_shapes, golden_seeds, stats, tv_to_timm, RAFT_CFG, IGEV_CFG
from dkt_stereo_b200.synthetic import synthetic_pair, synthetic_state_dict
dev = torch.device("cuda:0")
out = {}
for tag in ("raft_fwd_cfg2", "raft_fwd_cfg2_s2", "raft_fwd_cfg4shape", "igev_fwd_cfg3"):
    g = load_golden(tag)
    B, H, W, iters = [int(v) for v in g["meta"]]
    ws, is_ = golden_seeds(g)
    sd = synthetic_state_dict(golden_shapes(g), seed=ws)
    if tag.startswith("igev"):
        from dkt_stereo_b200.igev_stereo import IGEVStereo
        m = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval()
        m.load_state_dict({tv_to_timm(k): v for k, v in sd.items()}, strict=True)
        key = "disp_up"
    else:
        from dkt_stereo_b200.raft_stereo import RAFTStereo
        m = RAFTStereo(Namespace(mixed_precision=False, **dict(RAFT_CFG, corr_implementation="b200"))).eval()
        m.load_state_dict(sd, strict=True)
        key = "flow_up"
    m = m.to(dev)
    im1, im2 = synthetic_pair(B, H, W, seed=is_, mode=str(g["mode"]))
    _, up = m(im1.to(dev), im2.to(dev), iters=iters, test_mode=True)
    out[tag] = stats(up.cpu(), g[key])
    del m
    torch.cuda.empty_cache()
print("RESULT " + json.dumps(out))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-bench", action="store_true")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--only", type=int, nargs="*", default=None, help="indices into POLICIES")
    a = ap.parse_args()
    for idx, (name, env) in enumerate(POLICIES):
        if a.only is not None and idx not in a.only:
            continue
        e = dict(os.environ, **env)
        if "DKT_STEREO_LIB" in env and not os.path.exists(env["DKT_STEREO_LIB"]):
            print(f"{name}: variant library not built, skipped")
            continue
        r = subprocess.run([sys.executable, "-c", CHILD % dict(root=ROOT)], env=e, capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
        if not line:
            print(f"{name}: FAILED\n{r.stderr[-1500:]}")
            continue
        res = json.loads(line[0][7:])
        print(f"== {name}")
        for tag, (mean, mx) in res.items():
            print(f"   {tag:22s} mean |d - ref| {mean:.3e} px   max {mx:.3e} px   (gate 1e-3 mean)")
        if not a.no_bench:
            b = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(a.steps), "--warmup", "3",
                                "--no-cpu-baseline", "--no-secondary"], env=e, capture_output=True, text=True)
            try:
                d = json.loads([ln for ln in b.stdout.splitlines() if ln.startswith("{")][-1])
                bd = d["breakdown_ms_per_step"]
                top = ", ".join(f"{k} {v:.2f}" for k, v in list(bd.items())[:6])
                print(f"   bench: {d['ms_per_step']:.2f} ms/step = {d['value']:.2f} pairs/s; e2e {d['e2e']['value']:.2f}; "
                      f"zr0 {d['roofline']['ms_per_launch']:.3f} ms; clocks {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
                print(f"   top: {top}")
            except Exception as ex:          # noqa: BLE001
                print(f"   bench failed: {ex}\n{b.stderr[-1500:]}")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
