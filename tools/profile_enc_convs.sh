#!/bin/bash
# ncu --set full of the encoder's full-resolution 64 -> 64 convs (fnet: fp32 + InstanceNorm statistics out; cnet: BatchNorm
# folded, ReLU, 16-bit pair out with / without the residual) inside one headline step
set -u
TAG=${1:-r4a}
mkdir -p gpurun_out
NCU="timeout 600 ncu --profile-from-start off --clock-control none"
$NCU --set full --import-source on --kernel-name-base demangled -k regex:"conv_tc_pair_kernel<.int.0, .int.[01], .int.64, .int.(33|2|10)>" -c 16 \
    -o gpurun_out/prof_${TAG}_enc -f python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_${TAG}_enc.log 2>&1
ls -la gpurun_out | grep ${TAG}
