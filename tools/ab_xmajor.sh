set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r03a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r03a_pytest_gpu.log
for i in 1 2; do
DKT_STEREO_LIB=/root/repo/_ab/libhead.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r03a_bench_head_$i.json 2> gpurun_out/r03a_bench_head_$i.err; python -c "import json;d=json.load(open('gpurun_out/r03a_bench_head_$i.json'));print('HEAD',d['ms_per_step'],d['clocks'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r03a_bench_xm_$i.json 2> gpurun_out/r03a_bench_xm_$i.err; python -c "import json;d=json.load(open('gpurun_out/r03a_bench_xm_$i.json'));print('XM',d['ms_per_step'],d['clocks'])"
done
