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
PYTHONPATH) to run the five validators on real data; ``--batch N`` (default 1 like the reference)
groups same-shape samples into forwards of up to N pairs fed through ``HostPipeline`` (pinned upload
of the next batch overlapped with the current forward, dataset decoding in a background thread --
SURVEY 8f rank 3).  Offline, ``--synthetic HxW`` measures the same path on synthetic pairs.
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

# name -> dataset factory, bad-pixel threshold, whether the ground truth is bounded by maxdisp, which non-occlusion mask
# file the validator needs, and how D1 is averaged: the reference pools the outlier flags of ALL pixels for KITTI
# (tools/evaluate_stereo.py:160,162-166) but averages PER-IMAGE outlier rates for ETH3D, Middlebury and Booster
# (:92-101, :263-272, :324-333)
VALIDATORS = {
    "eth3d": dict(cls="ETH3D", kw={}, thr=1.0, bound=False, nocc="eth3d", pool="image"),
    "middlebury-H": dict(cls="Middlebury", kw={"resolution": "H"}, thr=2.0, bound=True, nocc="middlebury", pool="image"),
    "kitti-2012": dict(cls="KITTI", kw={"split": "2012", "image_set": "training"}, thr=3.0, bound=True, nocc=None, pool="pixel"),
    "kitti-2015": dict(cls="KITTI", kw={"split": "2015", "image_set": "training"}, thr=3.0, bound=True, nocc=None, pool="pixel"),
    "booster-Q": dict(cls="Booster", kw={"resolution": "Q"}, thr=2.0, bound=True, nocc=None, pool="image"),
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


def _nocc_mask(kind: str, sample) -> np.ndarray:
    """The non-occlusion mask the reference opens next to the sample's files (:56-57 ETH3D, :236-237 Middlebury).  A
    missing file is an error, as in the reference (PIL raises there) -- silently dropping the mask would change the metric."""
    from PIL import Image
    files = sample[0]
    if kind == "eth3d":
        path = files[2].replace("disp0GT.pfm", "mask0nocc.png")
        return np.ascontiguousarray(Image.open(path))
    path = files[0].replace("im0.png", "mask0nocc.png")
    return np.ascontiguousarray(Image.open(path).convert("L"), dtype=np.float32)


def sample_metrics(flow_pr: torch.Tensor, flow_gt: torch.Tensor, valid_gt: torch.Tensor, thr: float, bound: bool,
                   occ_mask, maxdisp: int = 192):
    """EPE and outlier flags of one image exactly as the reference computes them (:85-90 ETH3D, :151-158 KITTI,
    :255-262 Middlebury, :318-323 Booster).  flow_pr / flow_gt (1,H,W) = -disparity.  -> (image EPE, outlier flags of
    the valid pixels)."""
    assert flow_pr.shape == flow_gt.shape, (flow_pr.shape, flow_gt.shape)
    epe = torch.sum((flow_pr - flow_gt) ** 2, dim=0).sqrt().flatten()
    val = (valid_gt.reshape(-1) >= 0.5) & (flow_gt[0].reshape(-1) < 0)
    if bound:
        val &= flow_gt[0].reshape(-1) > -maxdisp
    if occ_mask is not None:
        val &= torch.from_numpy(np.asarray(occ_mask).flatten() == 255)
    return epe[val].mean().item(), (epe > thr)[val]


class _Batch:
    __slots__ = ("idx", "im1", "im2", "padder", "gts")


def _batch_stream(dataset, batch: int, nocc, depth: int = 2):
    """Background thread: reads samples in dataset order, pads each to /32 (replicate, reference :63-64) and groups
    samples of EQUAL padded shape into pinned (B,3,H,W) batches of up to ``batch`` pairs; yields _Batch objects.
    Decoding and padding of the next batches overlap the GPU work on the current one."""
    import queue
    import threading
    q: "queue.Queue" = queue.Queue(maxsize=depth)
    pin = torch.cuda.is_available()

    def emit(bucket):
        b = _Batch()
        b.idx = [e[0] for e in bucket]
        b.padder = bucket[0][3]
        b.im1 = torch.stack([e[1] for e in bucket])
        b.im2 = torch.stack([e[2] for e in bucket])
        if pin:
            b.im1, b.im2 = b.im1.pin_memory(), b.im2.pin_memory()
        b.gts = [e[4] for e in bucket]
        q.put(b)

    def work():
        try:
            buckets = {}
            for idx in range(len(dataset)):
                sample = dataset[idx]
                image1, image2, flow_gt, valid_gt = sample[-4:]
                occ = _nocc_mask(nocc, sample) if nocc else None
                padder = InputPadder(image1[None].shape, divis_by=DIVIDE_FACTOR)
                p1, p2 = padder.pad(image1[None].float(), image2[None].float())
                key = tuple(p1.shape[-2:]) + tuple(image1.shape[-2:])
                bk = buckets.setdefault(key, [])
                bk.append((idx, p1[0], p2[0], padder, (flow_gt, valid_gt, occ)))
                if len(bk) == batch:
                    emit(buckets.pop(key))
            for bk in buckets.values():
                emit(bk)
            q.put(None)
        except BaseException as e:            # noqa: BLE001 -- re-raised in the consumer
            q.put(e)

    threading.Thread(target=work, daemon=True).start()
    while True:
        item = q.get()
        if item is None:
            return
        if isinstance(item, BaseException):
            raise item
        yield item


@torch.no_grad()
def validate(model, dataset, name: str, thr: float, bound: bool, nocc, iters: int = 32, maxdisp: int = 192,
             pool: str = "image", batch: int = 8, feeder=None):
    """One pass over a reference dataset object (reference validate_* loops, tools/evaluate_stereo.py:46-336), batched
    and pipelined: same-shape samples ride in one forward of up to ``batch`` pairs; the pinned upload of batch i+1
    overlaps the forward of batch i (``HostPipeline``) and the dataset decoding runs in a background thread.  The
    metrics are the reference's, per sample, in dataset order.  ``feeder``: object with prefetch(im1, im2) / step(next)
    and, optionally, step_async(next) / drain() (default: HostPipeline on the model, whose step_async is used)."""
    model.eval()
    if feeder is None:
        from dkt_stereo_b200.pipeline import HostPipeline
        feeder = HostPipeline(model, iters=iters)
    epes, outs = {}, {}
    t0, pairs = time.perf_counter(), 0
    stream = _batch_stream(dataset, batch, nocc)

    def score(b, up):
        up = b.padder.unpad(up)
        for j, idx in enumerate(b.idx):
            flow_gt, valid_gt, occ = b.gts[j]
            image_epe, flags = sample_metrics(up[j].float(), flow_gt, valid_gt, thr, bound, occ, maxdisp)
            epes[idx] = image_epe
            outs[idx] = flags.numpy() if pool == "pixel" else flags.float().mean().item()
        return len(b.idx)

    cur = next(stream, None)
    if cur is not None:
        feeder.prefetch(cur.im1, cur.im2)
    if hasattr(feeder, "step_async"):
        # throughput form: the maps of a batch come back one call late, so the host scores batch i - 1 (and the reader
        # thread decodes batch i + 1) while the GPU computes batch i
        waiting = None
        while cur is not None:
            nxt = next(stream, None)
            up = feeder.step_async((nxt.im1, nxt.im2) if nxt is not None else None)
            if waiting is not None:
                pairs += score(waiting, up)
            waiting, cur = cur, nxt
        if waiting is not None:
            pairs += score(waiting, feeder.drain())
    else:
        while cur is not None:
            nxt = next(stream, None)
            up = feeder.step((nxt.im1, nxt.im2) if nxt is not None else None)  # pinned host (B,1,Hp,Wp), reused next step
            pairs += score(cur, up)
            cur = nxt
    order = sorted(epes)
    epe = float(np.mean([epes[i] for i in order]))
    if pool == "pixel":
        d1 = 100 * float(np.mean(np.concatenate([outs[i] for i in order])))
    else:
        d1 = 100 * float(np.mean([outs[i] for i in order]))
    fps = pairs / (time.perf_counter() - t0)
    print(f"Validation {name}: EPE {epe:f}, D1 {d1:f}, {fps:.2f}-FPS (batch {batch}, pipelined)")
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
            results.update(validate(model, ds, name, v["thr"], v["bound"], v["nocc"], iters=args.valid_iters,
                                    pool=v["pool"], batch=max(1, args.batch)))
    if args.logdir:
        from torch.utils.tensorboard import SummaryWriter
        w = SummaryWriter(args.logdir)
        for k, val in results.items():
            w.add_scalar(f"results/{k}", val, -1)
        w.close()
    return results


if __name__ == "__main__":
    main()
