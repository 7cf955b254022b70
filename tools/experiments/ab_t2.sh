set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r03g_pytest_gpu.log
for i in 1 2; do
DKT_CONV_T2=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r03g_bench_t1_$i.json 2> gpurun_out/r03g_bench_t1_$i.err; python -c "import json;d=json.load(open('gpurun_out/r03g_bench_t1_$i.json'));print('T1',d['ms_per_step'],d['clocks']); b=d['breakdown_ms_per_step']; print({k:b[k] for k in list(b)[:12]})"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r03g_bench_t2_$i.json 2> gpurun_out/r03g_bench_t2_$i.err; python -c "import json;d=json.load(open('gpurun_out/r03g_bench_t2_$i.json'));print('T2',d['ms_per_step'],d['clocks']); b=d['breakdown_ms_per_step']; print({k:b[k] for k in list(b)[:12]})"
done
