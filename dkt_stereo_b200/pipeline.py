"""Host-side feeding of the engine: pinned-memory input upload overlapped with the previous batch's compute and a
pinned read-back of the disparity maps (the evaluator's per-sample ``.cuda()`` / ``.cpu()`` round
trips, reference tools/evaluate_stereo.py:116-134, batched and pipelined).

    pipe = HostPipeline(model, iters=32)
    pipe.prefetch(im1_host, im2_host)                 # pinned (B,3,H,W) fp32 tensors in [0,255]
    for nxt in batches:                               # upload of `nxt` overlaps the compute of the batch before it
        disp_host = pipe.step(nxt)                    # (B,1,H,W) pinned fp32 = -disparity of the PREVIOUS prefetch
    disp_host = pipe.step(None)

Throughput form (what ``bench.py``'s ``e2e`` runs): ``step_async`` enqueues the batch's forward AND its read-back and returns
the result of the call BEFORE it, so the host never waits for the GPU between batches (graph launch and the read-back of
batch i overlap the forward of batch i + 1); ``drain()`` returns the last one.

    pipe.prefetch(im1_host, im2_host)
    prev = None
    for nxt in batches + [None]:
        out = pipe.step_async(nxt)                    # result of the previous step_async (None on the first call)
    last = pipe.drain()

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
        # step_async state: read-back stream, staged device copies / pinned results / events per slot
        self.d2h_stream: Optional[torch.cuda.Stream] = None
        self.out_dev = [None, None]
        self.out_hosts = [None, None]
        self.fwd_done = [None, None]       # forward that read input slot i (and staged its result) finished
        self.d2h_done = [None, None]
        self._last = None                  # (read-back event, pinned result) of the latest step_async

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
        # the slot's previous reader (the forward two steps ago) finished: step() ends with a stream sync; step_async()
        # leaves that forward's event in fwd_done
        with torch.cuda.stream(self.copy_stream):
            if self.fwd_done[slot] is not None:
                self.copy_stream.wait_event(self.fwd_done[slot])
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

    # ---- throughput form -----------------------------------------------------------------------------------------------
    def step_async(self, next_batch=None) -> Optional[torch.Tensor]:
        """Enqueue the prefetched batch's forward, the staging of its result and its read-back into pinned memory; start
        uploading ``next_batch``; return the PREVIOUS call's result (pinned, valid until the call after the next one) or None.
        Nothing here waits for work enqueued by this call."""
        assert self.pending, "call prefetch() first"
        cs = torch.cuda.current_stream()
        dev = cs.device
        if self.d2h_stream is None:
            self.d2h_stream = torch.cuda.Stream(device=dev)
        k = self.cur
        cs.wait_event(self.ready[k])
        d1, d2 = self.slots[k]
        _, up = self.model(d1, d2, iters=self.iters, test_mode=True)
        if self.out_dev[k] is None or self.out_dev[k].shape != up.shape:
            # both slots at once: pinning host memory is slow and synchronises the device -- it must not happen in the
            # middle of a stream of batches (r4n: the second slot's cudaHostAlloc inside the timed region cost 4 - 20 ms)
            for j in (0, 1):                 # (after a shape change the previous result stays alive through self._last)
                self.out_dev[j] = torch.empty_like(up)
                self.out_dev[j].record_stream(self.d2h_stream)
                self.out_hosts[j] = torch.empty(up.shape, dtype=up.dtype, pin_memory=True)
                self.d2h_done[j] = None
        if self.d2h_done[k] is not None:
            cs.wait_event(self.d2h_done[k])                # the read-back that last used this staging buffer
        self.out_dev[k].copy_(up)                          # the model's output buffer belongs to the next forward
        ev = torch.cuda.Event()
        ev.record(cs)
        self.fwd_done[k] = ev
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(ev)
            self.out_hosts[k].copy_(self.out_dev[k], non_blocking=True)
            e2 = torch.cuda.Event()
            e2.record(self.d2h_stream)
        self.d2h_done[k] = e2
        if next_batch is not None:
            self._upload(k ^ 1, next_batch)                # waits for the forward that last read that slot
        prev, self._last = self._last, (e2, self.out_hosts[k])
        self.cur ^= 1
        self.pending = next_batch is not None
        if prev is None:
            return None
        prev[0].synchronize()
        return prev[1]

    def drain(self) -> Optional[torch.Tensor]:
        """Result of the last step_async(); the current stream also waits for the outstanding copies, so an event recorded
        after drain() covers them."""
        if self._last is None:
            return None
        (ev_last, host), self._last = self._last, None
        cs = torch.cuda.current_stream()
        cs.wait_event(ev_last)
        for ev in self.ready:
            if ev is not None:
                cs.wait_event(ev)
        ev_last.synchronize()
        return host
