#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   launches_<tag>.csv       every kernel launch of ONE bench step with its device time (ncu, serialised)
#   prof_conv_tc_<tag>.ncu-rep  ncu --set full of one GRU iteration's conv launches (12 shapes)
#   prof_corr_<tag>.ncu-rep     ncu --set full of the volume build + first lookup
# Usage: tools/profile_gpu.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
mkdir -p gpurun_out
B="python bench.py --ncu-step --warmup 3 $*"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"conv_tc" -c 12 -o gpurun_out/prof_conv_tc_${TAG} -f $B > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"corr1d|split_nchw" -c 4 -o gpurun_out/prof_corr_${TAG} -f $B > gpurun_out/ncu_corr_${TAG}.log 2>&1
ls -la gpurun_out
