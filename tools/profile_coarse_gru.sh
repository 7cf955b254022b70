#!/bin/bash
# ncu --set full of the first loop iteration's GRU convs (1/16, 1/8, 1/4 resolution: z||r and q) inside one headline step
set -u
TAG=${1:-r4d}
mkdir -p gpurun_out
NCU="timeout 600 ncu --profile-from-start off --clock-control none"
$NCU --set full --import-source on --kernel-name-base demangled -k regex:"conv_tc_pair_kernel<.int.[12]," -c 6 \
    -o gpurun_out/prof_${TAG}_gru -f python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_gru.log 2>&1
ls -la gpurun_out | grep ${TAG}
