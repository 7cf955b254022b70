# final tree of round 2: GPU tests, smoke, both bench arms, IGEV line, compute-sanitizer memcheck of the kernels the last pass
# touched (pair conv kernel: 12 warps + setmaxnreg, new MMA / producer loops; InstanceNorm finalize)
set -x
cd /root/repo
TAG=${1:-r4z}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; wc -l gpurun_out/${TAG}_bench.json; cut -c1-330 gpurun_out/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_reference_arm.json
timeout 400 python bench.py --model igev --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench_igev.json 2> gpurun_out/${TAG}_bench_igev.err; echo "igev rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_igev.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -x -q -k "conv or encoder or update_block or instnorm or gru" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_memcheck.log
