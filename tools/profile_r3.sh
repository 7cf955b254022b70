#!/bin/bash
# Round-2 final profiling pass (run under gpurun, 1 GPU): launch lists of one headline step and one IGEV step, `ncu --set
# full` of the gru08 z||r / q convs, K1 + the fused RAFT lookup, the IGEV TMA lookup; DRAM traffic into profiles/ncu_traffic.json.
set -u
TAG=${1:-r3c}
mkdir -p gpurun_out
NCU="timeout 900 ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${TAG}_raft.csv \
    python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_raft.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${TAG}_igev.csv \
    python bench.py --model igev --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_igev.log 2>&1
# first loop iteration's gru08 convs: z||r and q (GRU_ZR / GRU_Q instantiations at 1/4 resolution come after the coarse ones)
$NCU --set full --import-source on --kernel-name-base demangled -k regex:"conv_tc_pair_kernel<.int.[12]," -c 6 -o gpurun_out/prof_${TAG}_gru -f \
    python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_gru.log 2>&1
$NCU --set full --import-source on -k regex:"lookup_tc_kernel|corr1d_build" -c 2 -o gpurun_out/prof_${TAG}_corr -f \
    python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_corr.log 2>&1
$NCU --set full --import-source on -k regex:"geo_lookup_tma" -c 1 -o gpurun_out/prof_${TAG}_geo -f \
    python bench.py --model igev --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_geo.log 2>&1
ls -la gpurun_out | grep ${TAG}
