#!/bin/bash
# Run under gpurun (1 GPU).  Produces in gpurun_out/:
#   launches_<tag>.csv     every kernel launch of one bench step with its device time (ncu, serialised)
#   prof_<tag>.ncu-rep     ncu --set full of the dominant conv_tc launches (gru08 z||r, q, flow-head conv1)
# Usage: tools/profile_gpu.sh <tag>
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_launches_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 20 -c 3 \
    -o gpurun_out/prof_conv_tc_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"corr1d_build_tc_kernel|corr1d_lookup_kernel" -s 1 -c 2 \
    -o gpurun_out/prof_corr_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/ncu_corr_${TAG}.log 2>&1
ls -la gpurun_out
