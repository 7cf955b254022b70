#!/usr/bin/env python
"""Where the gap between the device-resident step and the end-to-end step (HostPipeline.step_async) comes from: the same timed
loop with the upload and / or the read-back switched off, and with the upload issued from a second pinned buffer.

    python tools/e2e_gap_probe.py        (B200; RAFT-Stereo cfg2)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dkt_stereo_b200.pipeline import HostPipeline  # noqa: E402
from dkt_stereo_b200.synthetic import synthetic_pair  # noqa: E402


class Probe(HostPipeline):
    upload, readback = True, True

    def _upload(self, slot, batch):
        if self.upload or self.slots[slot] is None:
            return super()._upload(slot, batch)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.ready[slot] = ev


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    model, _ = bench.make_model("raft", dev, "tc", False)
    im1_h, im2_h = (t.pin_memory() for t in synthetic_pair(8, 544, 960, seed=1234))
    im1_d, im2_d = im1_h.to(dev), im2_h.to(dev)
    with torch.no_grad():
        for _ in range(4):
            model(im1_d, im2_d, iters=32, test_mode=True)
        torch.cuda.synchronize()
        base = bench.timed_steps(lambda: model(im1_d, im2_d, iters=32, test_mode=True), 5, dev) / 5
        print(f"device-resident                      {base:8.2f} ms/step")
        for up, rb in ((True, True), (False, True), (True, False), (False, False)):
            pipe = Probe(model, iters=32)
            pipe.upload, pipe.readback = True, rb
            pipe.prefetch(im1_h, im2_h)
            pipe.step_async((im1_h, im2_h))
            pipe.step_async((im1_h, im2_h))
            pipe.upload = up
            if not rb:
                orig = pipe.out_hosts
                # no read-back: point the pinned result at a device tensor so that the copy is device-to-device
                pipe.out_hosts = [torch.empty_like(o, device=dev) if o is not None else None for o in orig]
            ms = bench.timed_steps(lambda: pipe.step_async((im1_h, im2_h)), 5, dev, finish=pipe.drain) / 5
            print(f"step_async upload={up!s:5} readback={rb!s:5}   {ms:8.2f} ms/step")


if __name__ == "__main__":
    main()
