#!/bin/bash
# Run under gpurun (1 GPU): every kernel launch of ONE bench step with its device time -> gpurun_out/launches_<tag>.csv
# (summarise with tools/summarize_profiles.py <tag>)
set -u
TAG=${1:-r01}; shift || true
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --ncu-step --warmup 3 "$@" > gpurun_out/ncu_launches_${TAG}.log 2>&1
