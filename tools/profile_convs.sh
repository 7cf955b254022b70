#!/bin/bash
# Run under gpurun (1 GPU).  Per-shape view of the tensor-core conv kernel:
#   convs_<tag>.csv        key metrics for the first 100 conv launches of one step (encoder + 3 GRU iterations)
#   prof_iter_<tag>.ncu-rep  ncu --set full of the 8 conv launches of the 1/4-resolution part of one iteration
# Usage: tools/profile_convs.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
mkdir -p gpurun_out
B="python bench.py --ncu-step --warmup 3 $*"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg,launch__grid_size
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --page raw \
    -k regex:"conv_tc" -c ${NCONV:-100} --log-file gpurun_out/convs_${TAG}.csv $B > gpurun_out/ncu_convs_${TAG}.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"conv_tc" -s ${FULL_SKIP:-65} -c ${FULL_COUNT:-8} -o gpurun_out/prof_iter_${TAG} -f $B > gpurun_out/ncu_iter_${TAG}.log 2>&1
ls -la gpurun_out
