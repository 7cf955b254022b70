#!/usr/bin/env python
"""Print the per-launch table of gpurun_out/convs_<tag>.csv (tools/profile_convs.sh)."""
import csv, sys
tag = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = list(csv.reader(open(f'gpurun_out/convs_{tag}.csv')))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[start]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
ix = [h.index(w) for w in want]
print('idx kernel                         us   dramR_MB dramW_MB hit%  xbar_MB xbar_TB/s tensor% sm%')
for n, r in enumerate(rows[start + 2:]):
    if n < lo or n >= hi:
        continue
    v = [float(r[i].replace(',', '')) for i in ix]
    name = r[4].replace('void dkt::', '').replace('(dkt::TcConvParams)', '').replace('(int)', '')
    print(f"{n:3d} {name:30s} {v[0]/1e3:7.1f} {v[1]/1e6:8.1f} {v[2]/1e6:8.1f} {v[3]:5.1f} {v[4]/1e6:8.1f} {v[4]/v[0]/1e3:6.2f} {v[5]:6.1f} {v[6]:5.1f}")
