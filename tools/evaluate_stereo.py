#!/usr/bin/env python
"""Evaluation entry point of the B200 engine; mirrors the reference CLI
(reference tools/evaluate_stereo.py:339-404):

    python tools/evaluate_stereo.py --config configs/raft_stereo/base.json --restore_ckpt X.pth \
        --valid_iters 32 [--mixed_precision] [--logdir DIR]

Same flags, same JSON configs (merged into one Namespace, :349-352), same checkpoint formats (raw
state_dict with or without the DataParallel ``module.`` prefix, or ``{'state_dict': ...}``), same
metrics (EPE, D1 / bad-tau with the per-dataset thresholds and validity masks of :89-90,152-154,
259-261,321-322).  The datasets themselves are the reference's Python readers
(``core.stereo_datasets``, out of scope here): pass ``--reference_root`` (or put the reference on
PYTHONPATH) to run the five validators on real data.  Offline, ``--synthetic HxW`` runs the same
padded batch-1 loop on synthetic pairs and reports throughput; ``--batch`` feeds the engine more
than one pair per call.
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dkt_stereo_b200 import __models__            # noqa: E402
from dkt_stereo_b200.utils import InputPadder     # noqa: E402

DIVIDE_FACTOR = 32        # reference tools/evaluate_stereo.py:37

# name -> (dataset factory kwargs, bad-pixel threshold, uses maxdisp bound, needs non-occlusion mask)
VALIDATORS = {
    "eth3d": dict(cls="ETH3D", kw={}, thr=1.0, bound=False, nocc="eth3d"),
    "middlebury-H": dict(cls="Middlebury", kw={"resolution": "H"}, thr=2.0, bound=True, nocc="middlebury"),
    "kitti-2012": dict(cls="KITTI", kw={"split": "2012", "image_set": "training"}, thr=3.0, bound=True, nocc=None),
    "kitti-2015": dict(cls="KITTI", kw={"split": "2015", "image_set": "training"}, thr=3.0, bound=True, nocc=None),
    "booster-Q": dict(cls="Booster", kw={"resolution": "Q"}, thr=2.0, bound=True, nocc=None),
}


def count_parameters(model) -> int:
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


def load_checkpoint(model: torch.nn.Module, path: str) -> None:
    """reference :366-370 (strict load into a DataParallel wrapper) and ft_dkt.py:136-139."""
    assert path.endswith(".pth") or path.endswith(".ckpt")
    ckpt = torch.load(path, map_location="cpu")
    if isinstance(ckpt, dict) and "state_dict" in ckpt and not any(k.endswith(".weight") for k in ckpt):
        ckpt = ckpt["state_dict"]
    ckpt = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in ckpt.items()}
    model.load_state_dict(ckpt, strict=True)


def _nocc_mask(kind, dataset, idx):
    from PIL import Image
    if kind == "eth3d":
        gt = dataset.disparity_list[idx] if hasattr(dataset, "disparity_list") else None
        path = gt.replace("disp0GT.pfm", "mask0nocc.png") if gt else None
    else:
        img = dataset.image_list[idx][0]
        path = img.replace("im0.png", "mask0nocc.png")
    if path is None or not os.path.exists(path):
        return None
    return np.ascontiguousarray(Image.open(path).convert("L"), dtype=np.float32)


@torch.no_grad()
def validate(model, dataset, name: str, thr: float, bound: bool, nocc, iters: int = 32, maxdisp: int = 192):
    """One pass over a reference dataset object; batch 1, padded to /32 like the reference loops."""
    model.eval()
    epe_list, out_list, elapsed = [], [], []
    for idx in range(len(dataset)):
        sample = dataset[idx]
        image1, image2, flow_gt, valid_gt = sample[-4:]
        image1, image2 = image1[None].cuda(), image2[None].cuda()
        padder = InputPadder(image1.shape, divis_by=DIVIDE_FACTOR)
        image1, image2 = padder.pad(image1, image2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _, flow_pr = model(image1, image2, iters=iters, test_mode=True)
        torch.cuda.synchronize()
        if idx > 2:
            elapsed.append(time.perf_counter() - t0)
        flow_pr = padder.unpad(flow_pr).cpu().squeeze(0)
        assert flow_pr.shape == flow_gt.shape, (flow_pr.shape, flow_gt.shape)
        epe = torch.sum((flow_pr - flow_gt) ** 2, dim=0).sqrt().flatten()
        val = (valid_gt.reshape(-1) >= 0.5) & (flow_gt[0].reshape(-1) < 0)
        if bound:
            val &= flow_gt[0].reshape(-1) > -maxdisp
        if nocc:
            m = _nocc_mask(nocc, dataset, idx)
            if m is not None:
                val &= torch.from_numpy(m.flatten() == 255)
        epe_list.append(epe[val].mean().item())
        out_list.append((epe > thr)[val].float().numpy())
    epe, d1 = float(np.mean(epe_list)), 100 * float(np.mean(np.concatenate(out_list)))
    fps = 1.0 / float(np.mean(elapsed)) if elapsed else float("nan")
    print(f"Validation {name}: EPE {epe:f}, D1 {d1:f}, {fps:.2f}-FPS")
    return {f"{name}-epe": epe, f"{name}-d1": d1}


@torch.no_grad()
def run_synthetic(model, size: str, iters: int, batch: int, reps: int):
    from dkt_stereo_b200.synthetic import synthetic_pair
    H, W = (int(v) for v in size.lower().split("x"))
    im1, im2 = synthetic_pair(batch, H, W, seed=1234)
    im1, im2 = im1.cuda(), im2.cuda()
    padder = InputPadder(im1.shape, divis_by=DIVIDE_FACTOR)
    im1, im2 = padder.pad(im1, im2)
    for _ in range(3):
        model(im1, im2, iters=iters, test_mode=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        _, up = model(im1, im2, iters=iters, test_mode=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    up = padder.unpad(up)
    print(f"synthetic {H}x{W} batch {batch}, {iters} iters: {batch / dt:.2f} pairs/s ({dt * 1e3:.1f} ms/call), "
          f"mean disparity {float(-up.mean()):.3f} px")
    out = {"synthetic-pairs-per-s": batch / dt}
    # the same through the host feeding path: pinned batches uploaded while the previous one computes, pinned read-back
    from dkt_stereo_b200.pipeline import HostPipeline
    h1, h2 = im1.cpu().pin_memory(), im2.cpu().pin_memory()
    pipe = HostPipeline(model, iters=iters)
    pipe.prefetch(h1, h2)
    pipe.step((h1, h2))
    t0 = time.perf_counter()
    for _ in range(reps):
        host_up = pipe.step((h1, h2))
    dt = (time.perf_counter() - t0) / reps
    print(f"  host pipeline (pinned H2D + D2H every call): {batch / dt:.2f} pairs/s ({dt * 1e3:.1f} ms/call), "
          f"mean disparity {float(-padder.unpad(host_up).mean()):.3f} px")
    out["synthetic-pairs-per-s-host-pipeline"] = batch / dt
    return out


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--config", default=None, help="config file to create model")
    p.add_argument("--restore_ckpt", default=None, help="restore checkpoint")
    p.add_argument("--mixed_precision", action="store_true", help="use mixed precision (extractors only)")
    p.add_argument("--valid_iters", type=int, default=32, help="number of flow-field updates during forward pass")
    p.add_argument("--logdir", default=False, help="the directory to save logs")
    # additions of this engine
    p.add_argument("--reference_root", default=os.environ.get("DKT_REFERENCE"), help="reference checkout providing core.stereo_datasets")
    p.add_argument("--datasets", default="eth3d,middlebury-H,kitti-2012,kitti-2015,booster-Q")
    p.add_argument("--synthetic", default=None, metavar="HxW", help="run on synthetic pairs of this size instead of datasets")
    p.add_argument("--batch", type=int, default=1)
    p.add_argument("--reps", type=int, default=5)
    args = p.parse_args(argv)
    with open(args.config) as f:
        cfg = json.load(f)
    args = argparse.Namespace(**vars(args), **cfg)
    print(args)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)-8s [%(filename)s:%(lineno)d] %(message)s")

    model = __models__[args.model](args)
    logging.info("%s: %.2f M parameters", args.model, count_parameters(model) / 1e6)
    if args.restore_ckpt is not None:
        logging.info("Loading checkpoint...")
        load_checkpoint(model, args.restore_ckpt)
        logging.info("Done loading checkpoint")
    model.cuda().eval()

    results = {}
    if args.synthetic:
        results.update(run_synthetic(model, args.synthetic, args.valid_iters, args.batch, args.reps))
    else:
        if args.reference_root and args.reference_root not in sys.path:
            sys.path.insert(0, args.reference_root)
        try:
            import core.stereo_datasets as datasets          # the reference's readers (out of scope here)
        except Exception as e:                                # noqa: BLE001
            raise SystemExit(f"dataset readers unavailable ({e}); pass --reference_root or use --synthetic HxW")
        for name in args.datasets.split(","):
            v = VALIDATORS[name]
            ds = getattr(datasets, v["cls"])({}, **v["kw"])
            results.update(validate(model, ds, name, v["thr"], v["bound"], v["nocc"], iters=args.valid_iters))
    if args.logdir:
        from torch.utils.tensorboard import SummaryWriter
        w = SummaryWriter(args.logdir)
        for k, val in results.items():
            w.add_scalar(f"results/{k}", val, -1)
        w.close()
    return results


if __name__ == "__main__":
    main()
