set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "igev or gwc or conv3d or softargmin" > gpurun_out/r03b_pytest_igev.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r03b_pytest_igev.log
DKT_NATIVE_VOLUME=0 timeout 300 python tools/igev_preloop_breakdown.py > gpurun_out/r03b_preloop_torch.json 2> gpurun_out/r03b_preloop_torch.err; cat gpurun_out/r03b_preloop_torch.json
timeout 300 python tools/igev_preloop_breakdown.py > gpurun_out/r03b_preloop_native.json 2> gpurun_out/r03b_preloop_native.err; cat gpurun_out/r03b_preloop_native.json
timeout 400 python bench.py --model igev --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r03b_bench_igev.json 2> gpurun_out/r03b_bench_igev.err; python -c "import json;d=json.load(open('gpurun_out/r03b_bench_igev.json'));print('IGEV',d['ms_per_step'],d['value'],d['clocks'])"
