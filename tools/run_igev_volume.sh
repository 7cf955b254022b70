set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "igev or gwc or conv3d or softargmin or deconv" > gpurun_out/r03e_pytest_igev.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r03e_pytest_igev.log
timeout 300 python tools/igev_preloop_breakdown.py > gpurun_out/r03e_preloop_native.json 2> gpurun_out/r03e_preloop_native.err; cat gpurun_out/r03e_preloop_native.json
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"conv3d_k|deconv3d|gwc_volume|softargmin" -c 20 --csv --log-file gpurun_out/r03e_igev_preloop_launches.csv python tools/igev_preloop_breakdown.py > gpurun_out/r03e_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 400 python bench.py --model igev --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03e_bench_igev.json 2> gpurun_out/r03e_bench_igev.err; python -c "import json;d=json.load(open('gpurun_out/r03e_bench_igev.json'));print('IGEV',d['ms_per_step'],d['value'],d['clocks'])"
