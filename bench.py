#!/usr/bin/env python
"""Headline benchmark: stereo pairs/sec of RAFT-Stereo forward(test_mode=True) at 544x960, 32 GRU
iterations, batch 8 per GPU (BASELINE.json configs[1]), synthetic inputs, random-init weights.

    python bench.py --gpus 1 --steps 10 --warmup 3                    # this engine
    python bench.py --impl reference --steps 2 --warmup 1             # CPU baseline arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

RAFT_CFG = dict(model="RAFTStereo", loss_func="sequence_loss_raft", backbone_type="default",
                corr_implementation="reg", shared_backbone=False, corr_levels=4, corr_radius=4,
                n_downsample=2, context_norm="batch", slow_fast_gru=False, n_gru_layers=3,
                hidden_dims=[128, 128, 128])

METRIC = "stereo pairs/sec @ 544x960, 32 iters"
IGEV_CFG = dict(model="IGEVStereo", loss_func="sequence_loss_raft", corr_levels=2, corr_radius=4,
                n_downsample=2, context_norm="batch", slow_fast_gru=False, n_gru_layers=3,
                hidden_dims=[128, 128, 128], max_disp=192)


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/extract_traffic.py); None when that capture is missing."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(kernel_key)
    except (OSError, ValueError):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (of measured)")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="B200_PROFILING.md fallback (of fallback)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


def precision_tag(eng) -> str:
    """Arithmetic type of the tensor-core path: operands are 16-bit (hi, lo) splits of fp32 values, fp32 accumulate."""
    from dkt_stereo_b200 import _lib as L
    base = "f16" if L.split_dtype() == torch.float16 else "bf16"
    if eng.gru2 or eng.menc2:
        return f"{base}x3 (encoders, heads, volume) + {base}x2 (" + "+".join(
            n for n, on in (("GRUs", eng.gru2), ("motion encoder", eng.menc2)) if on) + "), fp32 accumulate"
    return f"{base}x3, fp32 accumulate"


def host_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    model = ln.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return dict(cores=os.cpu_count(), cpu=model)


# ---------------------------------------------------------------------------------------------
# CPU baseline -- the only place bench.py executes oracle/: the reference's own RAFTStereo (compiled to bytecode into
# oracle/_ref by oracle/build_ref.py, kind "reference") or, when that is absent, the oracle port (kind "port")
# ---------------------------------------------------------------------------------------------
def cpu_forward_fn(height, width, iters, batch, threads):
    """-> (callable running ONE forward of `batch` pairs on the host cores, kind)."""
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    from dkt_stereo_b200.synthetic import synthetic_pair
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    mine = RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG)).eval()     # weights only
    sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
    im1, im2 = synthetic_pair(batch, height, width, seed=1234)
    try:
        from oracle import build_ref
        import warnings
        warnings.filterwarnings("ignore")
        ref = build_ref.load("raft")(Namespace(mixed_precision=False, **RAFT_CFG)).eval()
        ref.load_state_dict(sd, strict=True)

        def fwd():
            with torch.no_grad():
                return ref(im1, im2, iters=iters, test_mode=True)
        return fwd, "reference"
    except Exception as e:                      # noqa: BLE001 -- bytecode not staged / other CPython: the port still runs
        sys.stderr.write(f"[bench] reference bytecode unavailable ({type(e).__name__}: {e}); timing the oracle port\n")
        from oracle import hotpath as O
        return (lambda: O.raft_forward(sd, im1, im2, iters, RAFT_CFG)), "port"


def cpu_pairs_per_sec(height, width, iters, batch, steps, warmup, threads):
    """-> (pairs/s over the timed steps, mean ms per step, median ms per step, kind)"""
    fwd, kind = cpu_forward_fn(height, width, iters, batch, threads)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fwd()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, 1e3 * total / len(times), 1e3 * statistics.median(times), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    hi = host_info()
    threads = hi["cores"] or 1
    H, W, iters = args.height, args.width, args.iters
    value, ms, _, kind = cpu_pairs_per_sec(H, W, iters, 1, args.steps, args.warmup, threads)
    what = ("the reference's own RAFTStereo.forward(test_mode=True) (oracle/_ref bytecode of the unmodified modules)"
            if kind == "reference" else "the oracle port of the reference path")
    sample = f"1 pair per step (B=1) of the same {H}x{W}, {iters}-iter workload; {args.steps} timed steps; {what}"
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"RAFT-Stereo {H}x{W}, {iters} iters, CPU fp32, B=1 per step", "host": hi},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def make_model(kind: str, dev, kernels: str = "tc", extractor_tf32: bool = False):
    from dkt_stereo_b200 import parallel
    torch.manual_seed(0)
    if kind == "igev":
        from dkt_stereo_b200.igev_stereo import IGEVStereo
        model = IGEVStereo(Namespace(mixed_precision=False, corr_implementation="b200", **IGEV_CFG)).eval().to(dev)
    else:
        from dkt_stereo_b200.raft_stereo import RAFTStereo
        cfg = dict(RAFT_CFG, corr_implementation="b200" if kernels == "tc" else "b200_fp32")
        model = RAFTStereo(Namespace(mixed_precision=False, extractor_tf32=extractor_tf32, **cfg)).eval().to(dev)
    nbytes = parallel.broadcast_weights(model, src=0)          # the one collective of the path
    return model, nbytes


def timed_steps(fn, steps, dev, finish=None):
    """K calls of fn between barrier + synchronize on both sides; CUDA events on the launch stream; max over ranks.
    finish (optional) runs before the closing event: it makes the launch stream wait for work fn left on other streams."""
    from dkt_stereo_b200 import parallel
    parallel.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    if finish is not None:
        finish()
    b.record()
    torch.cuda.synchronize()
    parallel.barrier()
    return parallel.all_reduce_max(a.elapsed_time(b), dev)


def measure_config(kind, H, W, iters, Bg, steps, warmup, rank, world, dev, kernels="tc", extractor_tf32=False):
    """One BASELINE configuration: forward(test_mode=True) on this rank's shard of Bg pairs, (a) inputs resident in HBM,
    (b) end to end through HostPipeline (pinned H2D of both images + D2H of the disparity maps every step)."""
    from dkt_stereo_b200 import _lib as L
    from dkt_stereo_b200.pipeline import HostPipeline
    from dkt_stereo_b200.synthetic import synthetic_pair
    model, bcast = make_model(kind, dev, kernels, extractor_tf32)
    im1_h, im2_h = synthetic_pair(Bg, H, W, seed=1234 + rank)   # every rank works on its own shard (weak scaling)
    im1_h, im2_h = im1_h.pin_memory(), im2_h.pin_memory()
    im1_d, im2_d = im1_h.to(dev), im2_h.to(dev)
    step_device = lambda: model(im1_d, im2_d, iters=iters, test_mode=True)      # noqa: E731
    pipe = HostPipeline(model, iters=iters)
    # the throughput form of the public feeding API: every call uploads a batch from pinned memory, runs forward(), reads the
    # disparity maps back into pinned memory and hands the caller the previous batch's maps
    step_e2e = lambda: pipe.step_async((im1_h, im2_h))                          # noqa: E731
    for _ in range(max(warmup, 3)):             # also builds the CUDA graph (2nd / 3rd call)
        step_device()
    torch.cuda.synchronize()
    # the step runs at the board's power cap: for the first ~second after idle the clocks are still above their sustained
    # value (r4l: 92.4 ms for the first five steps, 95.1 for the next five).  Both timed regions start from the sustained state.
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.5:
        step_device()
        torch.cuda.synchronize()
    res = dict(model=model, pipe=pipe, step_device=step_device, step_e2e=step_e2e, im=(im1_h, im2_h, im1_d, im2_d),
               bcast=bcast)
    return res


def finish_config(res, Bg, H, W, steps, world, dev):
    ms_total = timed_steps(res["step_device"], steps, dev)
    im1_h, im2_h = res["im"][:2]
    res["pipe"].prefetch(im1_h, im2_h)      # the first batch's upload is the only one outside the timed region ...
    res["step_e2e"]()                       # ... and this untimed step consumes it: K timed steps = K uploads + K reads
    res["step_e2e"]()                       # (a second one: every slot of the pipeline exists before the clock starts)
    # drain(): the launch stream waits for the last read-back (and the upload the last call started) before the closing event
    ms_e2e = timed_steps(res["step_e2e"], steps, dev, finish=res["pipe"].drain)
    return dict(value=world * Bg * steps / (ms_total * 1e-3), ms_per_step=ms_total / steps,
                e2e={"value": world * Bg * steps / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e / steps,
                     "h2d_bytes_per_step": 2 * im1_h.numel() * 4, "d2h_bytes_per_step": Bg * H * W * 4,
                     "api": "HostPipeline.step_async: pinned H2D of the batch, forward(), D2H of its maps into pinned memory, "
                            "every step; the read-back of batch i overlaps the forward of batch i + 1, the closing event "
                            "waits for the last read-back"})


def fixed_sample_checksum(model, iters, dev, H=544, W=960):
    """Every rank runs the SAME fixed-seed pair (batch 1) through its own GPU: the CRC of the raw output bytes must be
    identical on all ranks (per-sample results bit-identical between 1-GPU and N-GPU runs) and is printed so
    that the N = 1 and N = 8 records can be compared too."""
    import zlib
    from dkt_stereo_b200 import parallel
    from dkt_stereo_b200.synthetic import synthetic_pair
    im1, im2 = synthetic_pair(1, H, W, seed=4242)
    _, up = model(im1.to(dev), im2.to(dev), iters=iters, test_mode=True)
    crc = zlib.crc32(up.float().cpu().numpy().tobytes())
    allc = parallel.all_gather_int(crc, dev)
    return {"crc32": f"{crc:08x}", "ranks": len(allc), "identical": all(c == allc[0] for c in allc),
            "sample": f"seed 4242, 1 x {H}x{W}, {iters} iters, full-resolution output"}


def igev_lookup_roofline(model, Bg, h, w, peaks):
    """Combined geometry-encoding lookup (a5) alone on the volumes of the last step; algorithmic bytes of SURVEY 8d."""
    from dkt_stereo_b200 import ops
    eng = model.engine
    P = Bg * h * w
    nplanes = 1 if eng.menc2 else 2
    lk_bytes = P * (2 * 9 * 10 * 4 + 4 + nplanes * 64 * 2)       # tap reads + disparity, writes 64 ch 16-bit (fused convc1)
    eng.DELTA["f32"].zero_()                                      # the fused `disp += delta` then leaves the disparities alone
    for _ in range(3):
        model._lookup(eng)                                        # exactly the launch the loop makes
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        model._lookup(eng)
    b.record()
    torch.cuda.synchronize()
    lk_ms = a.elapsed_time(b) / 20
    tr = ncu_traffic("geo_lookup_enc") if (h, w, Bg) == (136, 240, 8) else None
    tma = eng.lookup_tc and eng.lookup_tap_planes == 1 and os.environ.get("DKT_LOOKUP_TMA", "1") != "0"
    return {"kernel": ("geo_lookup_tma_kernel (Combined_Geo_Encoding_Volume lookup: TMA bulk-copy gather, convc1 on tcgen05)" if tma else
                       "lookup_tc_kernel<GEO> (Combined_Geo_Encoding_Volume lookup + convc1 on tcgen05)" if eng.lookup_tc else
                       "geo_lookup_kernel<4, ENC> (Combined_Geo_Encoding_Volume lookup + convc1, fp32 FMAs)"), "bound": "hbm",
            "achieved": lk_bytes / (lk_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": lk_bytes / (lk_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": tr["dram_bytes"] if tr else None,
            "traffic_source": tr["source"] if tr else None, "ms_per_launch": lk_ms,
            "bytes_per_launch": lk_bytes, "peak_source": peaks["source"]}


# the BASELINE.json configurations beside the headline one: (key, model, H, W, iters, pairs per GPU, BASELINE wording)
SECONDARY = [
    ("igev_cfg3", "igev", 544, 960, 32, 8, "configs[2]: IGEV-Stereo 544x960, 32 iters, batch 8 per GPU"),
    ("raft_cfg4", "raft", 736, 1280, 32, 8, "configs[3]: RAFT-Stereo 736x1280, 32 iters, batch 64 over 8 GPUs = 8 per GPU"),
    ("igev_cfg5", "igev", 1024, 1536, 22, 4, "configs[4]: IGEV-Stereo 1024x1536, 22 iters, batch 32 over 8 GPUs = 4 per GPU"),
]


def run_secondary(rank, world, dev, peaks, steps=5):
    """The other BASELINE configurations, measured after the headline region with the same method (device-resident and
    end-to-end pairs/s, max over ranks; fixed batch per GPU)."""
    out = {}
    for key, kind, H, W, iters, Bg, wording in SECONDARY:
        res = measure_config(kind, H, W, iters, Bg, steps, 3, rank, world, dev)
        r = finish_config(res, Bg, H, W, steps, world, dev)
        entry = {"workload": wording, "metric": "stereo pairs/sec", "unit": "pairs/s", "value": r["value"],
                 "ms_per_step": r["ms_per_step"], "e2e": r["e2e"], "steps": steps, "warmup": 3,
                 "global_batch": world * Bg, "n_gpus": world, "dtype": precision_tag(res["model"].engine)}
        if key == "igev_cfg3":
            entry["roofline"] = igev_lookup_roofline(res["model"], Bg, H // 4, W // 4, peaks)
        out[key] = entry
        del res
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    from dkt_stereo_b200 import _lib as L, ops, parallel

    rank, local, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the engine)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    H, W, iters, Bg = args.height, args.width, args.iters, args.batch
    h, w = H // 4, W // 4

    res = measure_config("raft", H, W, iters, Bg, args.steps, args.warmup, rank, world, dev, args.kernels, args.extractor_tf32)
    model, step_device, bcast_bytes = res["model"], res["step_device"], res["bcast"]
    im1_h = res["im"][0]

    if args.ncu_step:
        # profiler window = exactly one step (ncu --profile-from-start off); nothing is timed or printed
        model.use_cuda_graph = False          # same kernels, launched eagerly so ncu sees each one
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    # launches per step, counted on one eager (graph-less) pass
    model.use_cuda_graph = False
    n0 = L.LAUNCHES
    step_device()
    torch.cuda.synchronize()
    launches_per_step = L.LAUNCHES - n0
    # per-kernel breakdown of one step (CUDA events around every launch; eager, so slightly pessimistic)
    with ops.LaunchProfiler() as prof:
        step_device()
    breakdown = prof.summary()
    model.use_cuda_graph = True
    step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    r = finish_config(res, Bg, H, W, args.steps, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_step, value = r["ms_per_step"], r["value"]

    # ---- roofline of the dominant kernel: the gru08 z||r gate conv (3x3, 384 -> 256 at 1/4 res) ----
    eng = model.engine
    dtype_tag = precision_tag(eng) if args.kernels == "tc" else "f32"
    extractor_desc = ("libdkt tcgen05 convs (EncoderEngine), 3-MMA 16-bit (hi, lo) split" if model.encoder is not None
                      else "PyTorch cuDNN fp32" + (" (TF32 allowed)" if args.extractor_tf32 else ""))
    P = Bg * h * w
    flops_zr = 2.0 * P * 256 * 9 * 384
    reps = 20
    evs = []
    for _ in range(3):
        eng._gru(0, 256)
    torch.cuda.synchronize()
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        Sx = eng._slice
        split, simt = eng.impl == "tc", eng.impl == "simt"
        glo = not eng.gru2                    # (hi, lo) activations = 3 MMAs per K step; hi only = 2
        e = ops.make_epilogue(L.EPI_GRU_ZR, out=Sx(eng.RH[0], 0, 128, simt, split, glo), ctx=eng.CTX[0]["f32"], ctx_c0=0,
                              z=Sx(eng.Z[0], 0, 128, True, False), h=Sx(eng.X[0], 0, 128, True, False))
        ops.conv2d([Sx(eng.X[0], 0, 384, simt, split, glo)], eng.weights["zr0"], e, Bg, h, w, eng.impl)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    zr_ms = sum(a.elapsed_time(b) for a, b in evs) / reps
    tflops = flops_zr / (zr_ms * 1e-3) / 1e12
    tr = ncu_traffic("gru08_zr") if (H, W, Bg) == (544, 960, 8) else None
    roofline = {"kernel": "conv_tc_pair_kernel<GRU_ZR> (gru08 z||r gates, 3x3 384->256, tcgen05 cta_group::2)", "bound": "tensor",
                "achieved": tflops, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": tflops / peaks["bf16_burst"],
                "traffic": tr["dram_bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "mma_per_k_step": 2 if eng.gru2 else 3,
                "issued_frac": (2.0 if eng.gru2 else 3.0) * tflops / peaks["bf16_burst"],
                "ms_per_launch": zr_ms, "flops_per_launch": flops_zr,
                "note": "achieved = useful fp32-equivalent FLOPs; the split-precision path issues mma_per_k_step x this many "
                        "MMA flops (issued_frac = tensor-pipe work actually done / peak)",
                "peak_source": peaks["source"] + ", burst (kernel timed alone)"}

    # ---- roofline of the correlation-volume build (K1), HBM bound ----
    # algorithmic bytes: both feature maps once (2*B*D*h*w*4: fp32 in the reference layout; the
    # encoder engine hands them over as bf16 hi+lo = the same 4 bytes per element) + every pyramid level once
    pyr = model._pyr
    k1_bytes = 2.0 * Bg * 256 * h * w * 4 + sum(P * (w >> l) * 4 for l in range(4))
    if model.encoder is not None:
        f = model.encoder.FMAP
        k1_fn = lambda: ops.corr1d_build_split(f.hi[:Bg], f.lo[:Bg], f.hi[Bg:], f.lo[Bg:], 4, 1.0 / 16, pyr)
        k1_launches = "1 (tcgen05 build straight from the encoder's NHWC bf16 hi/lo feature maps)"
    else:
        fm = torch.randn(Bg, 256, h, w, device=dev)
        fm2 = torch.randn(Bg, 256, h, w, device=dev)
        k1_fn = lambda: ops.corr1d_build(fm, fm2, 4, 1.0 / 16, impl=eng.impl, pyr=pyr)
        k1_launches = "3 (2x fp32->bf16 split + tcgen05 build)" if eng.impl == "tc" else "1 (fp32 SIMT)"
    for _ in range(3):
        k1_fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        k1_fn()
    b.record()
    torch.cuda.synchronize()
    k1_ms = a.elapsed_time(b) / 10
    # lookup (K2), timed alone on the volume of the last step: (a) the plain operator (reference
    # CorrBlock1D.__call__: taps to HBM, fp32 NHWC) and (b) the fused lookup + convc1 the loop runs.
    # bytes per iteration (a) = P * [L*(2r+2)*4 + 4 + L*(2r+1)*4]  
    k2_bytes = P * (4 * 10 * 4 + 4 + 4 * 9 * 4)
    k2_enc_bytes = P * (4 * 10 * 4 + 4 + (1 if eng.menc2 else 2) * 64 * 2)   # reads + coord, writes 64 ch 16-bit hi [+ lo]
    cx = eng.coords_x.clone()
    lk_out = torch.empty(Bg, h, w, 36, device=dev)

    def time_it(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    k2_ms = time_it(lambda: ops.corr1d_lookup(pyr, cx, 4, lk_out, "nhwc"))
    eng.DELTA["f32"].zero_()                                      # the fused `coords1 += delta` then leaves the coordinates alone
    k2e_ms = time_it(lambda: model._lookup(eng))                  # exactly the launch the loop makes
    cfg2 = (H, W, Bg) == (544, 960, 8)
    tr_k1 = ncu_traffic("corr1d_build_tc") if cfg2 else None
    tr_k2 = ncu_traffic("corr1d_lookup_enc") if cfg2 else None
    roofline_corr = {
        "build": {"bound": "hbm", "achieved": k1_bytes / (k1_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                  "frac": k1_bytes / (k1_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": k1_ms, "bytes": k1_bytes,
                  "traffic": tr_k1["dram_bytes"] if tr_k1 else None, "launches": k1_launches},
        "lookup": {"bound": "hbm", "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": k2_ms, "bytes": k2_bytes},
        "lookup_enc": {"bound": "hbm", "achieved": k2_enc_bytes / (k2e_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                       "unit": "GB/s", "frac": k2_enc_bytes / (k2e_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": k2e_ms,
                       "bytes": k2_enc_bytes, "traffic": tr_k2["dram_bytes"] if tr_k2 else None,
                       "kernel": "lookup_tc_kernel (tcgen05 convc1)" if eng.lookup_tc else "corr1d_lookup_kernel<ENC> (fp32 FMAs)",
                       "note": "lookup fused with convc1 (1x1, 36->64, ReLU): taps never reach HBM; DRAM traffic is ~2.5x the "
                               "algorithmic bytes because each 40-byte tap run costs a 128-byte line of its own volume row"},
    }

    # every rank: the fixed-sample checksum (all-gather) and the other BASELINE configurations (barriers inside)
    ident = fixed_sample_checksum(model, iters, dev, H, W)
    del res, step_device
    model = eng = pyr = None
    torch.cuda.empty_cache()
    secondary = None if args.no_secondary else run_secondary(rank, world, dev, peaks)
    if rank != 0:
        return
    # CPU baseline on rank 0 at N=1 only: one pair of the same workload through the reference's own modules
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        hi = host_info()
        _, _, med_ms, kind = cpu_pairs_per_sec(H, W, iters, 1, 3, 1, hi["cores"] or 1)
        cpu = {"value": 1e3 / med_ms, "unit": "pairs/s", "cores": hi["cores"], "kind": kind, "cpu": hi["cpu"],
               "sample": f"1 pair (B=1) of the same {H}x{W}, {iters}-iter workload; 1 warm-up + median of 3 runs "
                         f"({med_ms / 1e3:.1f} s per pair)"}

    top = sorted(breakdown.items(), key=lambda kv: -kv[1][1])[:40]
    total_prof = sum(t for _, t in breakdown.values())
    out = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype_tag, "data": "synthetic",
        "config": {"workload": f"RAFT-Stereo {H}x{W}, {iters} iters, batch {Bg} per GPU (BASELINE configs[1])",
                   "global_batch": world * Bg, "parallelism": f"dp{world} (batch shards, no steady-state collective)",
                   "kernels": args.kernels,
                   "extractor": extractor_desc,
                   "cache": "working set per step (470 MB volume + 1.3 GB activations) >> 126 MB L2; no flush needed",
                   "weight_broadcast_bytes": bcast_bytes},
        "e2e": r["e2e"],
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline, "roofline_corr": roofline_corr, "cpu_baseline": cpu, "clocks": clocks,
        "breakdown_ms_per_step": {k: round(v[1], 3) for k, v in top},
        "breakdown_total_ms": round(total_prof, 3),
        "rank_outputs_bit_identical": ident["identical"], "fixed_sample": ident,
        "secondary": secondary,
    }
    emit(out)


def run_b200_igev(args):
    """`--model igev`: BASELINE configs[2] as its own line (the default run reports it under `secondary`): IGEV-Stereo
    forward(test_mode=True); MobileNetV2 pyramid / 2-D stems in PyTorch, everything else on the library's kernels."""
    from dkt_stereo_b200 import _lib as L, parallel
    rank, local, world = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the engine)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    H, W, iters, Bg = args.height, args.width, args.iters, args.batch
    res = measure_config("igev", H, W, iters, Bg, args.steps, args.warmup, rank, world, dev)
    model, step_device = res["model"], res["step_device"]
    im1_d, im2_d = res["im"][2:]
    if args.ncu_step:                        # profiler window = exactly one eager step; nothing is timed or printed
        model.use_cuda_graph = False
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_device()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    model.use_cuda_graph = False             # count the library's launches on one eager pass
    n0 = L.LAUNCHES
    step_device()
    torch.cuda.synchronize()
    launches_per_step = L.LAUNCHES - n0
    model.use_cuda_graph = True
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    r = finish_config(res, Bg, H, W, args.steps, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    # split: feature side + volume stage vs hot path
    with torch.no_grad():
        pre = model.prepare(im1_d, im2_d)
        torch.cuda.synchronize()
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        pre = model.prepare(im1_d, im2_d)
        b.record()
        model.hot_path(*pre, iters)
        c.record()
        torch.cuda.synchronize()
        pre_ms, hot_ms = a.elapsed_time(b), b.elapsed_time(c)
    roof = igev_lookup_roofline(model, Bg, H // 4, W // 4, peaks)
    if rank != 0:
        return
    out = {
        "metric": METRIC + " (IGEV-Stereo)", "value": r["value"], "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision_tag(model.engine), "data": "synthetic",
        "config": {"workload": f"IGEV-Stereo {H}x{W}, {iters} iters, batch {Bg} per GPU (BASELINE configs[2])",
                   "global_batch": world * Bg, "parallelism": f"dp{world} (batch shards, no steady-state collective)",
                   "pre_loop": "MobileNetV2 pyramid + 2-D stems in PyTorch fp32; context encoder, GWC volume, 3-D hourglass, "
                               "classifier on libdkt kernels",
                   "cache": "volumes (init-corr 376 MB + GEV 602 MB) >> 126 MB L2; no flush needed"},
        "e2e": r["e2e"], "gpu_launches": launches_per_step * args.steps,
        "split_ms": {"pre_loop (PyTorch MobileNetV2 / stems + libdkt volume stage + context encoder)": pre_ms, "hot_path_kernels": hot_ms},
        "roofline": roof, "cpu_baseline": None, "clocks": clocks,
    }
    emit(out)


_JSON_OUT = None


def claim_stdout():
    """Keep the process's real stdout for the ONE JSON line and point fd 1 at stderr for everything else: libraries
    write banners to stdout (NCCL prints its version line there when NCCL_DEBUG is set in the environment)."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(out: dict) -> None:
    stream = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    stream.write(json.dumps(out) + "\n")
    stream.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--kernels", default="tc", choices=["tc", "simt"])
    ap.add_argument("--model", default="raft", choices=["raft", "igev"],
                    help="raft = the headline workload (BASELINE configs[1]); igev = secondary line (configs[2])")
    ap.add_argument("--height", type=int, default=544)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8, help="pairs per GPU per step")
    ap.add_argument("--extractor-tf32", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configurations (`secondary` map)")
    ap.add_argument("--ncu-step", action="store_true",
                    help="warm up, then run ONE step inside cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.model == "igev":
        run_b200_igev(args)
    else:
        run_b200(args)
    if args.impl != "reference":
        from dkt_stereo_b200 import parallel
        parallel.shutdown()


if __name__ == "__main__":
    main()
