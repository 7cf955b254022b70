#!/bin/bash
TAG=${1:-r2x}
NCU="timeout 600 ncu --clock-control none --set full --import-source on"
$NCU --kernel-name-base demangled -k regex:"geo_lookup_tma" -s 8 -c 1 -o gpurun_out/prof_${TAG}_geo_tma -f python tools/lookup_tc_bench.py > gpurun_out/ncu_${TAG}_geo_tma.log 2>&1
ls -la gpurun_out/*${TAG}*
