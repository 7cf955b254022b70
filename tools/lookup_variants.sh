#!/bin/bash
# A/B sweep of the tensor-core lookup's knobs (run under gpurun): flags (bit 0 window gather, bit 1 staged epilogue),
# CTAs per SM the grid is sized for, shared-memory carve-out (percent of 228 KB; -1 = driver default)
for f in 0 2; do for c in 1 2 3; do for cv in auto -1 100; do
  echo "=== flags=$f ctas=$c carveout=$cv"
  if [ $cv = auto ]; then DKT_LOOKUP_FLAGS=$f DKT_LOOKUP_CTAS=$c python tools/lookup_tc_bench.py 2>&1 | grep "tcgen05.*hi out"
  else DKT_LOOKUP_FLAGS=$f DKT_LOOKUP_CTAS=$c DKT_LOOKUP_CARVEOUT=$cv python tools/lookup_tc_bench.py 2>&1 | grep "tcgen05.*hi out"; fi
done; done; done
