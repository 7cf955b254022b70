# How much of the fused lookup + convc1 kernels is gather and how much is the fp32 encode, and do the two overlap across
# the CTAs resident on an SM?  DKT_LOOKUP_PHASE: 1 = gather skipped, 2 = encode skipped (profiling knob, csrc/corr.cu)
cd /root/repo
for ph in 0 1 2; do
  echo "== DKT_LOOKUP_PHASE=$ph"
  DKT_LOOKUP_PHASE=$ph timeout 200 python tools/geo_lookup_bench.py 2>/dev/null | head -1
  DKT_LOOKUP_PHASE=$ph timeout 200 python tools/lookup_bench.py 2>/dev/null | head -1
done
