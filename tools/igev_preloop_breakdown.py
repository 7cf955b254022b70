#!/usr/bin/env python
"""Where the IGEV pre-loop (PyTorch) spends its time at cfg3: CUDA-event timing of each stage of
IGEVStereo.prepare().  Run on the GPU box: python tools/igev_preloop_breakdown.py"""
import os, sys, json
from argparse import Namespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from bench import IGEV_CFG
from dkt_stereo_b200.igev_stereo import IGEVStereo
from dkt_stereo_b200.igev_modules import build_gwc_volume, disparity_regression
from dkt_stereo_b200.raft_stereo import _fp32_math
from dkt_stereo_b200.synthetic import synthetic_pair

from dkt_stereo_b200 import ops
NATIVE = os.environ.get("DKT_NATIVE_VOLUME", "1") == "1"
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = IGEVStereo(Namespace(mixed_precision=False, corr_implementation="b200", **IGEV_CFG)).eval().to(dev)
im1, im2 = (t.to(dev) for t in synthetic_pair(8, 544, 960, seed=1234))
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
def run():
    marks.clear()
    a = m.args
    with torch.no_grad(), _fp32_math(True):
        i1 = (2 * (im1 / 255.0) - 1.0).contiguous(); i2 = (2 * (im2 / 255.0) - 1.0).contiguous()
        mark("start")
        fl, fr = m.feature(i1), m.feature(i2); mark("feature(MobileNetV2) x2")
        if m.native_match:
            _, s44 = m._stems_native(torch.cat((i1, i2), 0)); s4, s4y = s44[:8], s44[8:]; mark("stems (dkt kernels after the first conv)")
        else:
            s2 = m.stem_2(i1); s4 = m.stem_4(s2); s4y = m.stem_4(m.stem_2(i2)); mark("stems")
        fl[0] = torch.cat((fl[0], s4), 1); fr[0] = torch.cat((fr[0], s4y), 1)
        if m.native_match:
            mt = m._match_native(torch.cat((fl[0], fr[0]), 0)); ml, mr = mt[:8], mt[8:]; mark("conv+desc (dkt kernels)")
        else:
            ml = m.desc(m.conv(fl[0])); mr = m.desc(m.conv(fr[0])); mark("conv+desc")
        D = a.max_disp // 4
        if NATIVE:
            g = ops.gwc_volume(ml, mr, D, 8); mark("dkt_gwc_volume")
            bn = m.corr_stem.bn
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps); shift = bn.bias - bn.running_mean * scale
            att = m.corr_feature_att.feat_att(fl[0]); mark("feature_att logits (2-D convs)")
            vol = ops.conv3d_c8(g, m.corr_stem.conv.weight, scale, shift, 0.01, att); mark("dkt_conv3d_c8 corr_stem+BN+leaky+att")
        else:
            g = build_gwc_volume(ml, mr, D, 8); mark("build_gwc_volume")
            vol = m.corr_stem(g); mark("corr_stem (3D conv)")
            vol = m.corr_feature_att(vol, fl[0]); mark("feature_att")
        if NATIVE and m.cost_agg.native_ok(vol):
            gev = m.cost_agg.forward_native(vol, fl); mark("hourglass (3D, dkt kernels)")
        else:
            gev = m.cost_agg(vol, fl); mark("hourglass (3D)")
        if NATIVE:
            lg = ops.conv3d_c8(gev, m.classifier.weight); mark("dkt_conv3d_c8 classifier")
            d0 = ops.softargmin(lg.squeeze(1)); mark("dkt_softargmin")
        else:
            prob = F.softmax(m.classifier(gev).squeeze(1), dim=1); d0 = disparity_regression(prob, D); mark("classifier+regress")
        cl = m.cnet(i1, num_layers=a.n_gru_layers); mark("cnet")
        nl = [torch.tanh(x[0]) for x in cl]; cx = [c(torch.relu(x[1])) for x, c in zip(cl, m.context_zqr_convs)]; mark("ctx convs")
for _ in range(3):
    run()
torch.cuda.synchronize()
run(); torch.cuda.synchronize()
out = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 2) for i in range(1, len(marks))}
out["total"] = round(marks[0][1].elapsed_time(marks[-1][1]), 2)
print(json.dumps(out))
