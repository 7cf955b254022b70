"""CPU: the batched / pipelined validator of tools/evaluate_stereo.py reproduces the reference's per-dataset metric
definitions (reference tools/evaluate_stereo.py:46-336) on a fake dataset with ragged image shapes."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


class FakeDataset:
    """Samples shaped like the reference's datasets: ((files), image1, image2, flow_gt (1,H,W) = -disparity, valid)."""

    def __init__(self, shapes, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for i, (h, w) in enumerate(shapes):
            im1 = torch.rand(3, h, w, generator=g) * 255
            im2 = torch.rand(3, h, w, generator=g) * 255
            gt = -(torch.rand(1, h, w, generator=g) * 250)            # some beyond maxdisp 192
            gt[0, 0, :3] = 0.5                                        # non-negative "flow": invalid in the reference
            valid = (torch.rand(h, w, generator=g) > 0.2).float()
            self.items.append(((f"L{i}", f"R{i}", f"GT{i}"), im1, im2, gt, valid))

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


class FakeFeeder:
    """Stands in for HostPipeline (same prefetch/step protocol): 'disparity' = a fixed function of the padded images."""

    def __init__(self):
        self.cur, self.calls = None, []

    @staticmethod
    def predict(im1, im2):
        return -(im1.mean(1, keepdim=True) * 0.5 + im2[:, :1] * 0.25)

    def prefetch(self, im1, im2):
        self.cur = (im1, im2)

    def step(self, nxt=None):
        im1, im2 = self.cur
        self.calls.append(tuple(im1.shape))
        self.cur = nxt
        return self.predict(im1, im2)


class FakeAsyncFeeder(FakeFeeder):
    """The throughput protocol of HostPipeline: step_async returns the PREVIOUS call's maps, drain() the last ones."""

    def __init__(self):
        super().__init__()
        self.last = None
        self.blocking_calls = 0

    def step(self, nxt=None):
        self.blocking_calls += 1
        return super().step(nxt)

    def step_async(self, nxt=None):
        prev, self.last = self.last, FakeFeeder.step(self, nxt)
        return prev

    def drain(self):
        prev, self.last = self.last, None
        return prev


def _reference_style(ds, thr, bound, pool, maxdisp=192):
    """The reference's batch-1 loop, written out literally."""
    from dkt_stereo_b200.utils import InputPadder
    epe_list, out_list = [], []
    for i in range(len(ds)):
        _, im1, im2, flow_gt, valid_gt = ds[i]
        padder = InputPadder(im1[None].shape, divis_by=32)
        p1, p2 = padder.pad(im1[None], im2[None])
        flow_pr = padder.unpad(FakeFeeder.predict(p1, p2)).squeeze(0)
        epe = torch.sum((flow_pr - flow_gt) ** 2, dim=0).sqrt().flatten()
        val = (valid_gt.reshape(-1) >= 0.5) & (flow_gt[0].reshape(-1) < 0)
        if bound:
            val &= flow_gt[0].reshape(-1) > -maxdisp
        out = epe > thr
        epe_list.append(epe[val].mean().item())
        out_list.append(out[val].numpy() if pool == "pixel" else out[val].float().mean().item())
    d1 = 100 * (np.mean(np.concatenate(out_list)) if pool == "pixel" else np.mean(out_list))
    return float(np.mean(epe_list)), float(d1)


@pytest.mark.parametrize("pool,bound,thr", [("pixel", True, 3.0), ("image", True, 2.0), ("image", False, 1.0)])
@pytest.mark.parametrize("batch", [1, 3])
def test_batched_validator_matches_reference_metric_definitions(pool, bound, thr, batch):
    import evaluate_stereo as E
    shapes = [(37, 50), (40, 64), (37, 50), (37, 50), (33, 70), (40, 64), (37, 50)]
    ds = FakeDataset(shapes)
    feeder = FakeFeeder()
    model = torch.nn.Identity()
    res = E.validate(model, ds, "fake", thr, bound, None, iters=4, pool=pool, batch=batch, feeder=feeder)
    epe, d1 = _reference_style(ds, thr, bound, pool)
    assert res["fake-epe"] == pytest.approx(epe, rel=1e-6)
    assert res["fake-d1"] == pytest.approx(d1, rel=1e-6)
    # pixel pooling and per-image averaging are different numbers on ragged data (the ADVICE finding)
    other = _reference_style(ds, thr, bound, "image" if pool == "pixel" else "pixel")[1]
    assert abs(other - d1) > 1e-9
    # same-shape samples rode together, every sample exactly once
    assert sum(c[0] for c in feeder.calls) == len(shapes)
    if batch == 3:
        assert max(c[0] for c in feeder.calls) == 3 and len(feeder.calls) < len(shapes)
    else:
        assert all(c[0] == 1 for c in feeder.calls)


@pytest.mark.parametrize("batch", [1, 3])
def test_validator_scores_one_batch_late_through_step_async(batch):
    """With a feeder that offers step_async / drain the validator uses them (never the blocking step) and reports the same
    metrics: every batch is scored against the maps that belong to it, the last one through drain()."""
    import evaluate_stereo as E
    shapes = [(37, 50), (40, 64), (37, 50), (37, 50), (33, 70), (40, 64), (37, 50)]
    ds = FakeDataset(shapes)
    feeder = FakeAsyncFeeder()
    res = E.validate(torch.nn.Identity(), ds, "fake", 2.0, True, None, iters=4, pool="image", batch=batch, feeder=feeder)
    epe, d1 = _reference_style(ds, 2.0, True, "image")
    assert res["fake-epe"] == pytest.approx(epe, rel=1e-6)
    assert res["fake-d1"] == pytest.approx(d1, rel=1e-6)
    assert feeder.blocking_calls == 0 and feeder.last is None
    assert sum(c[0] for c in feeder.calls) == len(shapes)


def test_missing_nocc_mask_is_an_error():
    import evaluate_stereo as E
    ds = FakeDataset([(37, 50)])
    with pytest.raises((FileNotFoundError, OSError)):
        E.validate(torch.nn.Identity(), ds, "eth3d", 1.0, False, "eth3d", iters=1, feeder=FakeFeeder())
