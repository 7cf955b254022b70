#!/usr/bin/env python
"""Per-launch table from an `ncu --metrics ... --csv --log-file` pass (tools/profile_igev_preloop.sh)."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
idx = {n: i for i, n in enumerate(hdr)}
L = OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    k = (r[idx["ID"]], r[idx["Kernel Name"]].split("(")[0][:44], r[idx["Grid Size"]])
    L.setdefault(k, {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
print(f"{'id':>3} {'kernel':44s} {'grid':>16s} {'us':>8s} {'issue%':>7s} {'fma%':>6s} {'warps%':>7s} {'dram MB':>8s}")
for k, v in L.items():
    print(f"{k[0]:>3} {k[1]:44s} {k[2]:>16s} {v['gpu__time_duration.sum'] / 1e3:8.1f} "
          f"{v.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):7.1f} "
          f"{v.get('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 0):6.1f} "
          f"{v.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):7.1f} "
          f"{(v.get('dram__bytes_read.sum', 0) + v.get('dram__bytes_write.sum', 0)) / 1e6:8.1f}")
