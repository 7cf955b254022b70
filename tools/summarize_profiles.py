#!/usr/bin/env python
"""Turn the raw ncu outputs of tools/profile_gpu.sh (gpurun_out/) into the small, tracked
summaries under profiles/:

  profiles/<tag>_launches.md   per-kernel count / total device time of ONE bench step
  profiles/<tag>_<name>.csv    key `ncu --set full` metrics per captured launch

Usage: python tools/summarize_profiles.py <tag>      (run in the build container; needs `ncu`)
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    ix = {n: i for i, n in enumerate(rows[start])}
    agg, total, n = collections.OrderedDict(), 0.0, 0
    for r in rows[start + 1:]:
        if len(r) < len(ix) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}[r[ix["Metric Unit"]]]
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])
        name = re.sub(r"^void ", "", name)[:90]
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + v)
        total += v
        n += 1
    out = os.path.join(ROOT, "profiles", f"{tag}_launches.md")
    ours = sum(t for k, (c, t) in agg.items() if k.startswith("dkt::"))
    with open(out, "w") as f:
        f.write(f"# ncu launch list of ONE bench step ({tag})\n\n")
        f.write("`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` around one\n"
                "`bench.py --ncu-step` step (RAFT-Stereo 544x960, 32 iters, batch 8).  Times are serialised and\n"
                "cold-cache: read the SHARES, not the absolutes.\n\n")
        f.write(f"launches: {n}, total device time {total:.1f} ms; this library's kernels (`dkt::`) {ours:.1f} ms "
                f"({100 * ours / total:.1f} %), PyTorch extractor + glue {total - ours:.1f} ms\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if t / total < 0.001:
                continue
            f.write(f"| `{k}` | {c} | {t:.3f} | {100 * t / total:.1f} % |\n")
    print("wrote", out)


def full(tag, name):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        print("missing", rep)
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_{name}.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in cols])
        w.writerow([units[i] for _, i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for _, i in cols])
    print("wrote", out)


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(tag)
    for name in ("conv_tc", "corr"):
        full(tag, name)
