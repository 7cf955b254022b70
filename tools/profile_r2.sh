#!/bin/bash
# Round-2 profiling pass (run under gpurun, 1 GPU): launch list of one headline step, launch list of one IGEV step,
# `ncu --set full` of the gru08 z||r conv (2 MMAs per K step), of both fused lookups and of the IGEV 3-D convs.
set -u
TAG=${1:-r2g}
mkdir -p gpurun_out
NCU="timeout 900 ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${TAG}_raft.csv \
    python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_raft.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_${TAG}_igev.csv \
    python bench.py --model igev --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_igev.log 2>&1
# first loop iteration's gru08 convs: conv_tc launches 69 (z||r) and 70 (q) of a bench step
$NCU --set full --import-source on -k regex:"conv_tc" -s 69 -c 2 -o gpurun_out/prof_${TAG}_gru08 -f \
    python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_gru08.log 2>&1
$NCU --set full --import-source on -k regex:"corr1d_lookup|corr1d_build" -c 2 -o gpurun_out/prof_${TAG}_corr -f \
    python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_corr.log 2>&1
$NCU --set full --import-source on -k regex:"geo_lookup" -c 1 -o gpurun_out/prof_${TAG}_geo -f \
    python bench.py --model igev --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_geo.log 2>&1
ls -la gpurun_out | tail -12
