# same-box A/B of the warp-specialised fused lookup kernels (DKT_LOOKUP_WS=0 -> the one-group kernels)
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ws in 0 1; do
  echo "== DKT_LOOKUP_WS=$ws"
  DKT_LOOKUP_WS=$ws timeout 200 python tools/geo_lookup_bench.py 2>/dev/null | head -1
  DKT_LOOKUP_WS=$ws timeout 200 python tools/lookup_bench.py 2>/dev/null | head -1
  DKT_LOOKUP_WS=$ws timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('raft', d['ms_per_step'], d['roofline_corr']['lookup_enc']['ms'], d['clocks']['sm_mhz'])"
  DKT_LOOKUP_WS=$ws timeout 300 python bench.py --model igev --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('igev', d['ms_per_step'], d['roofline']['ms_per_launch'], d['clocks']['sm_mhz'])"
done
