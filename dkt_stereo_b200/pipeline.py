"""Host-side feeding of the engine: pinned-memory input upload overlapped with the previous batch's compute and a
pinned read-back of the disparity maps (the evaluator's per-sample ``.cuda()`` / ``.cpu()`` round
trips, reference tools/evaluate_stereo.py:116-134, batched and pipelined).

    pipe = HostPipeline(model, iters=32)
    pipe.prefetch(im1_host, im2_host)                 # pinned (B,3,H,W) fp32 tensors in [0,255]
    for nxt in batches:                               # upload of `nxt` overlaps the compute of the batch before it
        disp_host = pipe.step(nxt)                    # (B,1,H,W) pinned fp32 = -disparity of the PREVIOUS prefetch
    disp_host = pipe.step(None)

PyTorch is used for streams, events and pinned memory only.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class HostPipeline:
    def __init__(self, model, iters: int = 32):
        self.model, self.iters = model, iters
        self.copy_stream: Optional[torch.cuda.Stream] = None
        self.slots = [None, None]          # device input slots (im1, im2)
        self.ready = [None, None]          # upload-finished events
        self.cur = 0                       # slot holding the batch the next step() computes
        self.pending = False
        self.out_host: Optional[torch.Tensor] = None

    def _slot(self, i: int, like: torch.Tensor, dev) -> Tuple[torch.Tensor, torch.Tensor]:
        s = self.slots[i]
        if s is None or s[0].shape != like.shape or s[0].device != dev:
            # the slot is first WRITTEN on the copy stream and READ on the compute stream: allocate it under the copy
            # stream (so the caching allocator cannot hand out a block whose previous compute-stream user is still
            # queued) and tell the allocator about the second stream
            with torch.cuda.stream(self.copy_stream):
                s = self.slots[i] = (torch.empty(like.shape, device=dev, dtype=torch.float32),
                                     torch.empty(like.shape, device=dev, dtype=torch.float32))
            for t in s:
                t.record_stream(torch.cuda.current_stream(dev))
        return s

    def _upload(self, slot: int, batch) -> None:
        im1, im2 = batch
        dev = next(self.model.parameters()).device
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=dev)
        d1, d2 = self._slot(slot, im1, dev)
        # the slot's previous reader (the forward two steps ago) finished: step() ends with a stream sync
        with torch.cuda.stream(self.copy_stream):
            d1.copy_(im1, non_blocking=True)
            d2.copy_(im2, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[slot] = ev

    def prefetch(self, im1_host: torch.Tensor, im2_host: torch.Tensor) -> None:
        """Start uploading the first batch."""
        self._upload(self.cur, (im1_host, im2_host))
        self.pending = True

    def step(self, next_batch=None) -> torch.Tensor:
        """Compute the prefetched batch; meanwhile upload ``next_batch`` (a pair of pinned host tensors) if given.
        Returns the pinned host tensor with this batch's full-resolution result (valid until the next step())."""
        assert self.pending, "call prefetch() first"
        cs = torch.cuda.current_stream()
        cs.wait_event(self.ready[self.cur])
        d1, d2 = self.slots[self.cur]
        _, up = self.model(d1, d2, iters=self.iters, test_mode=True)
        if next_batch is not None:
            self._upload(self.cur ^ 1, next_batch)         # overlaps the forward just enqueued
        if self.out_host is None or self.out_host.shape != up.shape:
            self.out_host = torch.empty(up.shape, dtype=up.dtype, pin_memory=True)
        self.out_host.copy_(up, non_blocking=True)
        cs.synchronize()
        self.cur ^= 1
        self.pending = next_batch is not None
        return self.out_host
