NCU="timeout 900 ncu --profile-from-start off --clock-control none"
$NCU --set full --import-source on -k regex:"lookup_tc" -c 1 -o gpurun_out/prof_r2j_lk_raft -f python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_r2j_a.log 2>&1
$NCU --set full --import-source on -k regex:"lookup_tc" -c 1 -o gpurun_out/prof_r2j_lk_igev -f python bench.py --model igev --ncu-step --warmup 3 > gpurun_out/ncu_r2j_b.log 2>&1
ls -la gpurun_out/*r2j*
