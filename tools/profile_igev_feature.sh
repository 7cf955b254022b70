ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3l_feature_launches.csv python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from argparse import Namespace
from bench import IGEV_CFG
from dkt_stereo_b200.igev_stereo import IGEVStereo
from dkt_stereo_b200.synthetic import synthetic_pair
dev = torch.device("cuda:0")
m = IGEVStereo(Namespace(mixed_precision=False, corr_implementation="b200", **IGEV_CFG)).eval().to(dev)
im1, im2 = (t.to(dev) for t in synthetic_pair(8, 544, 960, seed=1234))
both = torch.cat(((2 * (im1 / 255.0) - 1.0), (2 * (im2 / 255.0) - 1.0)), 0).contiguous()
with torch.no_grad():
    for _ in range(2):
        f = m.feature(both)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    f = m.feature(both)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
PY
