"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Operand-precision sensitivity study of the RAFT-Stereo forward on CPU.

Question (task 3): the tensor-core path issues 3 MMAs per K step (bf16 hi*hi + lo*hi + hi*lo).
Which layer groups keep the 1e-3 px gate with TWO (or one) MMAs per K step, and with which operand formats?

Method: the CPU oracle's convolutions are re-run with their operands rounded exactly as a given tensor-core variant
would see them (fp32 accumulation either way), one layer group at a time, on the headline workload itself
(544 x 960, 32 iterations, one noise pair, name-seeded weights = tests/golden/raft_fwd_cfg2.npz) and the mean / max
|disparity difference| against the exact-fp32 oracle is tabulated.

    python -m oracle.precision_study [--height 544 --width 960 --iters 32] [--quick]

Variants (x = activation, w = weight; "2" = (hi, lo) pair, products kept):
    f32        exact
    bf16x3     x2 * w2 without lo*lo            3 MMAs (round 1's path)
    h_x1_w2    fp16(x) * (fp16 w_hi + w_lo)     2 MMAs
    h_x2_w1    (fp16 x_hi + x_lo) * fp16(w)     2 MMAs
    b_x1_w2    bf16(x) * (bf16 w_hi + w_lo)     2 MMAs
    b_x2_w1    (bf16 x_hi + x_lo) * bf16(w)     2 MMAs
    h_x1_w1    fp16(x) * fp16(w)                1 MMA
    hb_x1_w2   fp16(x) * (bf16 w_hi + w_lo)     2 MMAs (kind::f16 mixes A = f16, B = bf16)
    h_x2_w2    fp16 pairs, lo*lo dropped        3 MMAs (fp16 subnormals as torch converts them)
The correlation-volume build (K1) is a sixth group, "corr": its two feature maps are the operands.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import hotpath as O  # noqa: E402

GROUPS = ("enc", "menc", "gru08", "gru16_32", "head", "corr")     # + "convc1" (mixes only; exact fp32 otherwise)


FINE_ENC = False      # --fine-enc: the encoders split into stem / layer1 / layer2 / layer3 / output groups per network


def group_of(name: str) -> str:
    if name.endswith("encoder.convc1"):
        return "convc1"
    if FINE_ENC and name.startswith(("cnet.", "fnet.", "context_zqr_convs.")):
        if name.startswith("context_zqr_convs."):
            return "zqr"
        net, rest = name.split(".", 1)
        tag = net[0]
        if rest.startswith("conv1"):
            return tag + ".stem"
        for l in ("layer1", "layer2", "layer3"):
            if rest.startswith(l):
                return tag + "." + l
        return tag + ".rest"          # fnet.conv2; cnet.layer4/5 and the output heads
    if name.startswith(("cnet.", "fnet.", "context_zqr_convs.")):
        return "enc"
    if name.startswith("update_block.encoder."):
        return "menc"
    if name.startswith("update_block.gru08."):
        return "gru08"
    if name.startswith(("update_block.gru16.", "update_block.gru32.")):
        return "gru16_32"
    return "head"


def _split(t, dt):
    hi = t.to(dt).float()
    lo = (t - hi).to(dt).float()
    return hi, lo


def emulated_conv(x, w, b, stride, padding, variant):
    if variant == "f32":
        return F.conv2d(x, w, b, stride=stride, padding=padding)
    kind, xs, ws = variant.split("_") if variant != "bf16x3" else ("b", "x2", "w2")
    dtx = torch.float16 if kind in ("h", "hb") else torch.bfloat16
    dtw = torch.float16 if kind == "h" else torch.bfloat16
    xh, xl = _split(x, dtx)
    wh, wl = _split(w, dtw)
    if variant == "bf16x3" or (xs == "x2" and ws == "w2"):          # 3 MMAs: hi*hi + lo*hi + hi*lo (lo*lo dropped)
        return F.conv2d(xh + xl, wh, b, stride=stride, padding=padding) + F.conv2d(xh, wl, None, stride=stride, padding=padding)
    xe = xh + xl if xs == "x2" else xh
    we = wh + wl if ws == "w2" else wh
    return F.conv2d(xe, we, b, stride=stride, padding=padding)


def run(sd, im1, im2, iters, cfg, assign):
    """assign: group -> variant (missing = f32).  convc1 (1x1 on the 36 lookup taps) and the 7x7 flow stem's 2-channel
    input stay exact fp32 in the product's fused lookup / are negligible; they follow their group here except convc1."""
    orig = O._conv

    def conv(sd_, name, x, stride=1, padding=0):
        v = assign.get(group_of(name), "f32")
        return emulated_conv(x, sd_[name + ".weight"], sd_.get(name + ".bias"), stride, padding, v)

    orig_corr = O.corr1d_all_pairs

    def corr(fmap1, fmap2, scale=True):
        v = assign.get("corr", "f32")
        if v == "f32":
            return orig_corr(fmap1, fmap2, scale)
        # K1 as a 1x1 "conv" of fmap1's pixels with fmap2's pixels as the filters, one image row at a time
        B, D, H, W = fmap1.shape
        rows = []
        for bi in range(B):
            for y in range(H):
                a = fmap1[bi, :, y, :].t().reshape(W, D, 1, 1).permute(2, 1, 0, 3)          # (1, D, W, 1) "image"
                wgt = fmap2[bi, :, y, :].t().reshape(W, D, 1, 1)                              # (W2, D, 1, 1) "filters"
                rows.append(emulated_conv(a, wgt, None, 1, 0, v)[0, :, :, 0].t())            # (W1, W2)
        out = torch.stack(rows).reshape(B, H, W, W)
        return out / torch.sqrt(torch.tensor(D).float()) if scale else out

    O._conv = conv
    O.corr1d_all_pairs = corr
    try:
        return O.raft_forward(sd, im1, im2, iters, cfg)[1]
    finally:
        O._conv = orig
        O.corr1d_all_pairs = orig_corr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=544)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--variants", default="bf16x3,h_x1_w2,hb_x1_w2,h_x2_w1,b_x1_w2,h_x1_w1")
    ap.add_argument("--mixes", default="")
    ap.add_argument("--out", default="")
    ap.add_argument("--fine-enc", action="store_true")
    ap.add_argument("--wseed", type=int, default=0)
    ap.add_argument("--iseed", type=int, default=1234)
    a = ap.parse_args()
    global FINE_ENC
    FINE_ENC = a.fine_enc
    from dkt_stereo_b200.synthetic import synthetic_pair, synthetic_state_dict
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import RAFT_CFG, load_golden, golden_shapes
    torch.set_num_threads(os.cpu_count())
    g = load_golden("raft_fwd_cfg2")
    sd = synthetic_state_dict(golden_shapes(g), seed=a.wseed)
    im1, im2 = synthetic_pair(1, a.height, a.width, seed=a.iseed, mode="noise")
    t0 = time.time()
    ref = run(sd, im1, im2, a.iters, RAFT_CFG, {})
    print(f"# exact fp32 oracle: {time.time() - t0:.1f} s; mean |disp| {float(ref.abs().mean()):.2f} px", flush=True)
    rows = []

    def report(tag, assign):
        t = time.time()
        up = run(sd, im1, im2, a.iters, RAFT_CFG, assign)
        d = (up.double() - ref.double()).abs()
        row = dict(tag=tag, assign=assign, mean=float(d.mean()), max=float(d.max()), p99=float(d.flatten().kthvalue(int(d.numel() * 0.99)).values))
        rows.append(row)
        print(f"{tag:34s} mean {row['mean']:.3e}  p99 {row['p99']:.3e}  max {row['max']:.3e}   ({time.time() - t:.0f} s)", flush=True)

    variants = [v for v in a.variants.split(",") if v]
    for v in variants:
        report(f"all:{v}", {gname: v for gname in GROUPS})
        if a.quick:
            continue
        for gname in GROUPS:
            report(f"{gname}:{v}", {gname: v})
    for mix in [m for m in a.mixes.split(";") if m]:
        assign = dict(kv.split("=") for kv in mix.split(","))
        report("mix " + mix, assign)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
