TAG=${1:-r2j}
NCU="timeout 900 ncu --profile-from-start off --clock-control none"
$NCU --set full --import-source on -k regex:"lookup_tc" -c 1 -o gpurun_out/prof_${TAG}_lk_raft -f python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_a.log 2>&1
$NCU --set full --import-source on -k regex:"lookup_tc" -c 1 -o gpurun_out/prof_${TAG}_lk_igev -f python bench.py --model igev --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_b.log 2>&1
ls -la gpurun_out/*${TAG}*
