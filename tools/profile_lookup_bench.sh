#!/bin/bash
# ncu --set full of the stand-alone fused lookups (tools/lookup_tc_bench.py) under a DKT_LOOKUP_FLAGS value.
# usage: bash tools/profile_lookup_bench.sh TAG FLAGS
TAG=${1:-r2v}; FL=${2:-2}
NCU="timeout 600 ncu --clock-control none --set full --import-source on"
DKT_LOOKUP_FLAGS=$FL $NCU --kernel-name-base demangled -k regex:"lookup_tc_kernel<.int.4, .int.3, .int.1," -s 8 -c 1 -o gpurun_out/prof_${TAG}_igev_f$FL -f python tools/lookup_tc_bench.py > gpurun_out/ncu_${TAG}_igev_f$FL.log 2>&1
DKT_LOOKUP_FLAGS=$FL $NCU --kernel-name-base demangled -k regex:"lookup_tc_kernel<.int.4, .int.1, .int.1," -s 8 -c 1 -o gpurun_out/prof_${TAG}_raft_f$FL -f python tools/lookup_tc_bench.py > gpurun_out/ncu_${TAG}_raft_f$FL.log 2>&1
ls -la gpurun_out/*${TAG}*
