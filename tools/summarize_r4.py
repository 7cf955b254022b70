#!/usr/bin/env python
"""Round-2 (r4*) profile summaries: a launch list with TWO metrics per launch (duration + tensor-pipe activity) -> a per-kernel
table under profiles/, and `ncu --set full` reports -> key metrics per captured launch (same columns as
tools/summarize_profiles.py) + the stall-reason split.

    python tools/summarize_r4.py launches gpurun_out/launches_r4c_raft.csv profiles/r4c_raft_launches.md "title"
    python tools/summarize_r4.py full gpurun_out/prof_r4a_enc.ncu-rep profiles/r4a_enc_convs_full.csv
"""
import collections
import csv
import re
import subprocess
import sys

from summarize_profiles import KEYS

TEN = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"


def launches(src, dst, title):
    lines = [l for l in open(src) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {n: i for i, n in enumerate(hdr)}
    per = collections.OrderedDict()
    for row in r:
        if len(row) <= ix["Metric Value"]:
            continue
        e = per.setdefault(int(row[ix["ID"]]), {"k": re.sub(r"^void ", "", re.sub(r"\(.*", "", row[ix["Kernel Name"]]))[:70]})
        e[row[ix["Metric Name"]]] = float(row[ix["Metric Value"]].replace(",", ""))
    agg = collections.OrderedDict()
    total = 0.0
    for e in per.values():
        t = e.get("gpu__time_duration.sum", 0.0) / 1e6          # ns -> ms
        c, tt, ten = agg.get(e["k"], (0, 0.0, 0.0))
        agg[e["k"]] = (c + 1, tt + t, ten + t * e.get(TEN, 0.0))
        total += t
    ours = sum(t for k, (c, t, _) in agg.items() if k.startswith("dkt::"))
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,{TEN}` around one\n"
                "`bench.py --ncu-step` step.  Times are serialised and cold-cache: read the SHARES, not the absolutes.\n"
                "tensor % = time-weighted tensor-pipe activity of the kernel's launches.\n\n")
        f.write(f"launches: {len(per)}, total device time {total:.1f} ms; this library's kernels (`dkt::`) {ours:.1f} ms "
                f"({100 * ours / total:.1f} %)\n\n| kernel | launches | total ms | share | tensor % |\n|---|---:|---:|---:|---:|\n")
        for k, (c, t, ten) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if t / total < 0.001:
                continue
            f.write(f"| `{k}` | {c} | {t:.3f} | {100 * t / total:.1f} % | {ten / t if t else 0:.0f} |\n")
    print("wrote", dst)


def full(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    stall = [(h.replace("smsp__pcsamp_warps_issue_stalled_", "stall_"), i) for i, h in enumerate(hdr)
             if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in cols] + ["top stall reasons (share of samples)"])
        w.writerow([units[i] for _, i in cols] + [""])
        for r in rows[2:]:
            vals = [(float(r[i].replace(",", "")), n) for n, i in stall if r[i] not in ("", "n/a")]
            tot = sum(v for v, _ in vals) or 1.0
            top = "; ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(vals, reverse=True)[:4])
            w.writerow([r[i] for _, i in cols] + [top])
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
