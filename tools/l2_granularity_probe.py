#!/usr/bin/env python
"""Does cudaLimitMaxL2FetchGranularity change the gather kernels?  ncu (profiles/ncu_traffic.json) shows the fused
lookup pulling 165 MB from DRAM for 44 MB of taps: every 40-byte tap run costs ~160 bytes, i.e. 128-byte L2 fills.
This probe times the three lookups at 32 / 64 / 128 bytes (and the zr0 conv + K1 build as controls for streaming
kernels).  The limit is device-wide and only a hint; it is NOT changed by the library.

    python tools/l2_granularity_probe.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dkt_stereo_b200 import ops, _lib as L  # noqa: E402

dev = torch.device("cuda:0")
torch.zeros(1, device=dev)
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = ctypes.CDLL(name)
        break
    except OSError:
        pass
if rt is None:
    import glob
    cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
    rt = ctypes.CDLL(cands[0])
LIMIT = 0x05          # cudaLimitMaxL2FetchGranularity


def get_limit():
    v = ctypes.c_size_t(0)
    rc = rt.cudaDeviceGetLimit(ctypes.byref(v), LIMIT)
    return rc, v.value


def t(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


B, h, w = 8, 136, 240
g = torch.Generator(device=dev).manual_seed(0)
pyr = [torch.randn(B, h, w, w >> l, device=dev, generator=g) for l in range(4)]
cx = (torch.arange(w, device=dev).float().view(1, 1, w) - torch.rand(B, h, w, device=dev, generator=g) * 40).contiguous()
wt = torch.randn(64, 36, 1, 1, device=dev, generator=g) / 6
bias = torch.randn(64, device=dev, generator=g)
hi = torch.zeros(B, h, w, 64, device=dev, dtype=L.split_dtype())
out = L.tensor_slice(None, hi, None, 0, 64)
Wf = ops.pack_conv(wt, bias, cin_pad=64, tc=False)
plain = torch.zeros(B, h, w, 36, device=dev)
# IGEV volumes
geo = [torch.randn(B, h, w, 8, 48 >> l, device=dev, generator=g) for l in range(2)]
init = [torch.randn(B, h, w, w >> l, device=dev, generator=g) for l in range(2)]
disp = (torch.rand(B, h, w, device=dev, generator=g) * 48).contiguous()
wg = torch.randn(64, 162, 1, 1, device=dev, generator=g) / 12
Wg = ops.pack_conv(wg, bias, cin_pad=192, tc=False)
src = torch.randn(256 * 1024 * 1024 // 4, device=dev)
dst = torch.empty_like(src)

print("default limit:", get_limit())
for gran in (128, 64, 32, 128):
    rc = rt.cudaDeviceSetLimit(LIMIT, ctypes.c_size_t(gran))
    print(f"--- set {gran} B (rc {rc}); now {get_limit()} ---")
    print("  RAFT plain lookup      : %.1f us" % t(lambda: ops.corr1d_lookup(pyr, cx, 4, plain, "nhwc")))
    print("  RAFT lookup + convc1   : %.1f us" % t(lambda: ops.corr1d_lookup_enc(pyr, cx, 4, Wf, out)))
    print("  IGEV lookup + convc1   : %.1f us" % t(lambda: ops.geo_lookup_enc(geo, init, disp, 4, Wg, out)))
    print("  256 MB device copy     : %.1f us" % t(lambda: dst.copy_(src), 10))
