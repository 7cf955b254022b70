#!/usr/bin/env python
"""Pull dram__bytes_read.sum + dram__bytes_write.sum of one captured launch out of an `ncu --set full` report and
record it under profiles/ncu_traffic.json, the file bench.py reads for `roofline.traffic`.

    python tools/extract_traffic.py gpurun_out/prof_iter_r01j.ncu-rep 0 gru08_zr     # launch index inside the report
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, idx, key = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
r = rows[2 + idx]
val = {}
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
    i = h.index(name)
    v = float(r[i].replace(",", ""))
    val[name] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6,
                     "nsecond": 1, "usecond": 1e3, "msecond": 1e6}[units[i]]
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[key] = {"kernel": r[h.index("Kernel Name")], "dram_bytes": val["dram__bytes_read.sum"] + val["dram__bytes_write.sum"],
             "dram_read_bytes": val["dram__bytes_read.sum"], "dram_write_bytes": val["dram__bytes_write.sum"],
             "ncu_duration_us": val["gpu__time_duration.sum"] / 1e3,
             "source": f"ncu --set full, {os.path.basename(rep)} launch {idx}"}
json.dump(data, open(path, "w"), indent=1)
print(data[key])
