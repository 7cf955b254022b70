"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz from the real reference.

Run in the build container (the only place /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

The reference is imported read-only with the recipe of SURVEY.md section 8(c): a namespace
stub for ``meta_arch`` (its __init__ pulls timm via cgi) and stubs for ``opt_einsum``/``timm``.
Nothing here is imported by the product, and the GPU box never runs this file (it only reads
the committed vectors).  Each vector stores inputs + reference outputs for one hot-path
function so that (a) the oracle restatement is pinned on CPU and (b) the CUDA path is checked
against the *reference's* numbers, not just against our own restatement.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import types
from argparse import Namespace

import numpy as np
import torch

REF = os.environ.get("DKT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference checkout not found at {REF}")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "meta_arch" not in sys.modules:
        pkg = types.ModuleType("meta_arch")
        pkg.__path__ = [os.path.join(REF, "meta_arch")]
        sys.modules["meta_arch"] = pkg
    if "opt_einsum" not in sys.modules:
        oe = types.ModuleType("opt_einsum")
        oe.contract = torch.einsum
        sys.modules["opt_einsum"] = oe
    if "timm" not in sys.modules:
        sys.modules["timm"] = types.ModuleType("timm")
    mods = dict(
        corr=importlib.import_module("meta_arch.raft_stereo.corr"),
        update=importlib.import_module("meta_arch.raft_stereo.update"),
        raft=importlib.import_module("meta_arch.raft_stereo.raft_stereo"),
        geometry=importlib.import_module("meta_arch.igev_stereo.geometry"),
        igev_update=importlib.import_module("meta_arch.igev_stereo.update"),
        igev_sub=importlib.import_module("meta_arch.igev_stereo.submodule"),
    )
    return mods


def raft_cfg():
    with open(os.path.join(REF, "configs/raft_stereo/base.json")) as f:
        return json.load(f)


def igev_cfg():
    with open(os.path.join(REF, "configs/igev_stereo/base.json")) as f:
        return json.load(f)


def save(name: str, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                  for k, v in arrays.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def golden_corr(m):
    """CorrBlock1D build + lookup incl. out-of-range coordinates and an odd W2."""
    g = torch.Generator().manual_seed(11)
    for tag, (B, D, H, W) in {"a": (2, 32, 5, 40), "b": (1, 64, 3, 53)}.items():
        f1 = torch.randn(B, D, H, W, generator=g)
        f2 = torch.randn(B, D, H, W, generator=g)
        blk = m["corr"].CorrBlock1D(f1, f2, num_levels=4, radius=4)
        xs = torch.arange(W).float().view(1, 1, 1, W).expand(B, 1, H, W)
        coords = torch.cat([xs + (torch.rand(B, 1, H, W, generator=g) * (W + 20) - W / 2 - 10),
                            torch.zeros(B, 1, H, W)], 1)
        coords[0, 0, 0, :6] = torch.tensor([-1.0, -0.5, 0.0, W - 1.0, W - 0.5, float(W)])
        out = blk(coords)
        pyr = [p.reshape(B, H, W, -1) for p in blk.corr_pyramid[:4]]
        save(f"corr1d_{tag}", fmap1=f1, fmap2=f2, coords=coords, out=out,
             **{f"pyr{i}": p for i, p in enumerate(pyr)})


def golden_geo(m):
    g = torch.Generator().manual_seed(12)
    B, D, H, W, C, Dd = 2, 24, 4, 36, 8, 12
    f1 = torch.randn(B, D, H, W, generator=g)
    f2 = torch.randn(B, D, H, W, generator=g)
    gev = torch.randn(B, C, Dd, H, W, generator=g)
    fn = m["geometry"].Combined_Geo_Encoding_Volume(f1, f2, gev, num_levels=2, radius=4)
    disp = torch.rand(B, 1, H, W, generator=g) * (Dd + 6) - 3
    disp[0, 0, 0, :4] = torch.tensor([0.0, -1.0, Dd - 1.0, float(Dd)])
    coords = torch.arange(W).float().reshape(1, 1, W, 1).repeat(B, H, 1, 1)
    out = fn(disp, coords)
    save("geo_a", fmap1=f1, fmap2=f2, gev=gev, disp=disp, out=out)


def _ns(cfg):
    return Namespace(mixed_precision=False, **cfg)


def golden_update(m):
    """One BasicMultiUpdateBlock step (RAFT and IGEV flavours) + upsampling."""
    from dkt_stereo_b200.synthetic import synthetic_state_dict, shapes_of
    g = torch.Generator().manual_seed(13)
    B, h, w = 1, 16, 24
    for tag, mod, cfg, cp, fc in (("raft", m["update"], raft_cfg(), 36, 2), ("igev", m["igev_update"], igev_cfg(), 162, 1)):
        blk = mod.BasicMultiUpdateBlock(_ns(cfg), hidden_dims=cfg["hidden_dims"]).eval()
        sd = synthetic_state_dict(shapes_of(blk.state_dict()), seed=3)
        blk.load_state_dict(sd)
        net = [torch.tanh(torch.randn(B, 128, h >> i, w >> i, generator=g)) for i in range(3)]
        inp = [[torch.randn(B, 128, h >> i, w >> i, generator=g) * 0.5 for _ in range(3)] for i in range(3)]
        corr = torch.randn(B, cp, h, w, generator=g)
        flow = torch.randn(B, fc, h, w, generator=g) * 3
        with torch.no_grad():
            net_o, mask, delta = blk([t.clone() for t in net], inp, corr, flow)
        arrays = dict(corr=corr, flow=flow, mask=mask, delta=delta, keys=np.array(sorted(sd.keys())),
                      key_shapes=np.array([json.dumps(list(sd[k].shape)) for k in sorted(sd.keys())]))
        for i in range(3):
            arrays[f"net{i}"] = net[i]
            arrays[f"net_out{i}"] = net_o[i]
            for j, n in enumerate("zrq"):
                arrays[f"c{n}{i}"] = inp[i][j]
        save(f"update_{tag}", **arrays)
    # convex upsample (raft_stereo.py:70-82) via an instance with only args
    R = m["raft"].RAFTStereo
    flow = torch.randn(2, 2, 6, 9, generator=g) * 5
    mask = torch.randn(2, 144, 6, 9, generator=g)
    dummy = types.SimpleNamespace(args=_ns(raft_cfg()))
    up = R.upsample_flow(dummy, flow, mask)
    save("convex_upsample", flow=flow, mask=mask, out=up)
    # IGEV context_upsample (submodule.py:242-254)
    disp = torch.rand(2, 1, 5, 7, generator=g) * 30
    wts = torch.softmax(torch.randn(2, 9, 20, 28, generator=g), 1)
    save("context_upsample", disp=disp, weights=wts, out=m["igev_sub"].context_upsample(disp, wts))


def golden_raft_forward(m, height=64, width=96, iters=4, tag="raft_fwd_small", batch=1, mode="noise",
                        wseed=0, iseed=1234, **cfg_over):
    """Full RAFTStereo.forward(test_mode=True) with name-seeded synthetic weights (weight seed ``wseed``, image seed
    ``iseed``; both are stored so the GPU test regenerates the same inputs without the reference)."""
    from dkt_stereo_b200.synthetic import synthetic_state_dict, shapes_of, synthetic_pair
    cfg = dict(raft_cfg(), **cfg_over)
    model = m["raft"].RAFTStereo(_ns(cfg)).eval()
    sd = synthetic_state_dict(shapes_of(model.state_dict()), seed=wseed)
    model.load_state_dict(sd, strict=True)
    im1, im2 = synthetic_pair(batch, height, width, seed=iseed, mode=mode)
    with torch.no_grad():
        flow_lr, flow_up = model(im1, im2, iters=iters, test_mode=True)
    save(tag, flow_lr=flow_lr, flow_up=flow_up, meta=np.array([batch, height, width, iters]),
         seeds=np.array([wseed, iseed]),
         mode=np.array(mode), keys=np.array(sorted(sd.keys())),
         key_shapes=np.array([json.dumps(list(sd[k].shape)) for k in sorted(sd.keys())]))


def _timm_stub():
    """timm 0.5.4 is absent offline; torchvision's MobileNetV2 is architecture-identical and
    supplies the attributes the reference slices (conv_stem, bn1, act1, blocks[0:7])."""
    import torchvision

    def create_model(name, pretrained=False, features_only=True):
        assert name == "mobilenetv2_100"
        f = torchvision.models.mobilenet_v2(weights=None).features
        net = types.SimpleNamespace()
        net.conv_stem, net.bn1, net.act1 = f[0][0], f[0][1], f[0][2]
        groups = [f[1:2], f[2:4], f[4:7], f[7:11], f[11:14], f[14:17], f[17:18]]
        net.blocks = [torch.nn.Sequential(*g) for g in groups]
        return net

    sys.modules["timm"].create_model = create_model


def golden_igev_forward(m, height=64, width=96, iters=4, tag="igev_fwd_small", batch=1, **cfg_over):
    """Full reference IGEVStereo.forward(test_mode=True) (timm stubbed by torchvision).  Stores the
    products of the pre-loop (the hot path's inputs, captured with hooks while the REAL forward runs)
    and the final disparity, so the engine's IGEV hot path is pinned against the reference's loop."""
    from dkt_stereo_b200.synthetic import synthetic_state_dict, shapes_of, synthetic_pair
    _timm_stub()
    igev = importlib.import_module("meta_arch.igev_stereo.igev_stereo")
    cfg = dict(igev_cfg(), **cfg_over)
    torch.manual_seed(0)
    model = igev.IGEVStereo(_ns(cfg)).eval()
    sd = synthetic_state_dict(shapes_of(model.state_dict()), seed=0)
    model.load_state_dict(sd, strict=True)
    cap = {}

    class Recorder(igev.Combined_Geo_Encoding_Volume):
        def __init__(self, f1, f2, gev, **kw):
            cap.update(match_left=f1.clone(), match_right=f2.clone(), gev=gev.clone())
            super().__init__(f1, f2, gev, **kw)

    def pre_hook(mod, args, kwargs):
        # first call of the update block: the initial hidden states / context terms (slow_fast_gru's extra coarse
        # updates come first and carry no disparity); first call WITH a disparity: the initial disparity
        if "net0" not in cap:
            net, inp = args[:2]
            for i in range(len(net)):
                cap[f"net{i}"] = net[i].clone()
                cap[f"ctx{i}"] = torch.cat(list(inp[i]), 1).clone()
        if "init_disp" not in cap and len(args) > 3 and args[3] is not None:
            cap["init_disp"] = args[3].clone()

    orig_up = model.upsample_disp

    def up(disp, mask_feat_4, stem_2x):
        cap.update(stem_2x=stem_2x.clone(), final_disp=disp.clone(), mask_feat_4=mask_feat_4.clone())
        return orig_up(disp, mask_feat_4, stem_2x)

    igev.Combined_Geo_Encoding_Volume = Recorder
    model.update_block.register_forward_pre_hook(pre_hook, with_kwargs=True)
    model.upsample_disp = up
    im1, im2 = synthetic_pair(batch, height, width, seed=1234, mode="noise")
    with torch.no_grad():
        _, disp_up = model(im1, im2, iters=iters, test_mode=True)
    hot_keys = sorted(k for k in sd if k.startswith(("update_block.", "spx_2_gru.", "spx_gru.")))
    save(tag, disp_up=disp_up, meta=np.array([batch, height, width, iters]), keys=np.array(hot_keys),
         key_shapes=np.array([json.dumps(list(sd[k].shape)) for k in hot_keys]), **cap)


def golden_igev_forward_full(m, height, width, iters, tag, batch=1, mode="noise", wseed=0, iseed=1234):
    """The REAL reference ``IGEVStereo.forward(image1, image2, iters, test_mode=True)`` (igev_stereo.py:151-226; timm
    stubbed by torchvision's MobileNetV2) at a BASELINE configuration: stores only ``disp_up`` plus the names / shapes of
    the reference's state dict (torchvision MobileNetV2 naming) and the two seeds, so that the GPU test rebuilds the same
    weights by name, maps them onto the drop-in's timm-named parameters and calls the PUBLIC forward()."""
    from dkt_stereo_b200.synthetic import synthetic_state_dict, shapes_of, synthetic_pair
    _timm_stub()
    igev = importlib.import_module("meta_arch.igev_stereo.igev_stereo")
    torch.manual_seed(0)
    model = igev.IGEVStereo(_ns(igev_cfg())).eval()
    sd = synthetic_state_dict(shapes_of(model.state_dict()), seed=wseed)
    model.load_state_dict(sd, strict=True)
    im1, im2 = synthetic_pair(batch, height, width, seed=iseed, mode=mode)
    with torch.no_grad():
        _, disp_up = model(im1, im2, iters=iters, test_mode=True)
    keys = sorted(sd.keys())
    save(tag, disp_up=disp_up, meta=np.array([batch, height, width, iters]), seeds=np.array([wseed, iseed]),
         mode=np.array(mode), keys=np.array(keys), key_shapes=np.array([json.dumps(list(sd[k].shape)) for k in keys]))


def golden_all_predictions(m, height=64, width=96, iters=3):
    """forward(test_mode=False) of both reference models under no_grad: every iteration's full-resolution prediction
    (raft_stereo.py:170-187) and, for IGEV, the upsampled initial disparity as well (igev_stereo.py:178-183,222-226)."""
    from dkt_stereo_b200.synthetic import synthetic_state_dict, shapes_of, synthetic_pair
    im1, im2 = synthetic_pair(1, height, width, seed=1234, mode="noise")
    model = m["raft"].RAFTStereo(_ns(raft_cfg())).eval()
    sd = synthetic_state_dict(shapes_of(model.state_dict()), seed=0)
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        res = model(im1, im2, iters=iters)
    keys = sorted(sd.keys())
    save("raft_all_preds", disp_preds=torch.stack(res["disp_preds"]), meta=np.array([1, height, width, iters]),
         seeds=np.array([0, 1234]), mode=np.array("noise"), keys=np.array(keys),
         key_shapes=np.array([json.dumps(list(sd[k].shape)) for k in keys]))
    _timm_stub()
    igev = importlib.import_module("meta_arch.igev_stereo.igev_stereo")
    torch.manual_seed(0)
    model = igev.IGEVStereo(_ns(igev_cfg())).eval()
    sd = synthetic_state_dict(shapes_of(model.state_dict()), seed=0)
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        res = model(im1, im2, iters=iters)
    keys = sorted(sd.keys())
    save("igev_all_preds", disp_preds=torch.stack(res["disp_preds"]), init_disp=res["init_disp"],
         meta=np.array([1, height, width, iters]), seeds=np.array([0, 1234]), mode=np.array("noise"), keys=np.array(keys),
         key_shapes=np.array([json.dumps(list(sd[k].shape)) for k in keys]))


def golden_igev_volume(m, B=1, C=96, H=9, W=37, D=20, tag="igev_volume"):
    """The volume stage of the IGEV pre-loop through the REAL reference modules: build_gwc_volume,
    corr_stem (BasicConv 3-D, eval-mode BatchNorm with non-trivial running statistics), FeatureAtt, the classifier
    convolution, softmax + disparity_regression (reference igev_stereo.py:169-176)."""
    sub = m["igev_sub"]
    g = torch.Generator().manual_seed(77)
    left = torch.randn(B, C, H, W, generator=g)
    right = torch.randn(B, C, H, W, generator=g)
    feat = torch.randn(B, C, H, W, generator=g)
    stem = sub.BasicConv(8, 8, is_3d=True, kernel_size=3, stride=1, padding=1).eval()
    fatt = sub.FeatureAtt(8, C).eval()
    classifier = torch.nn.Conv3d(8, 1, 3, 1, 1, bias=False)
    with torch.no_grad():
        for mod in (stem, fatt, classifier):
            for prm in mod.parameters():
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.3)
        for bn in (stem.bn, fatt.feat_att[0].bn):
            bn.running_mean.copy_(torch.randn(bn.running_mean.shape, generator=g) * 0.2)
            bn.running_var.copy_(torch.rand(bn.running_var.shape, generator=g) + 0.5)
        gwc = sub.build_gwc_volume(left, right, D, 8)
        att_logits = fatt.feat_att(feat)
        vol = fatt(stem(gwc), feat)
        logits = classifier(vol).squeeze(1)
        disp = sub.disparity_regression(torch.nn.functional.softmax(logits, dim=1), D)
    save(tag, left=left, right=right, gwc=gwc, stem_w=stem.conv.weight, bn_weight=stem.bn.weight, bn_bias=stem.bn.bias,
         bn_mean=stem.bn.running_mean, bn_var=stem.bn.running_var, bn_eps=np.array(stem.bn.eps), att_logits=att_logits,
         vol=vol, cls_w=classifier.weight, logits=logits, disp=disp, meta=np.array([B, C, H, W, D]))


def golden_igev_hourglass(m, B=1, D=16, H=16, W=40, tag="igev_hourglass"):
    """The REAL reference `hourglass(8)` (meta_arch/igev_stereo/igev_stereo.py:22-89) in eval mode with seeded random
    parameters and non-trivial BatchNorm statistics on a small volume; stores inputs, the state dict and the output."""
    _timm_stub()
    igev = importlib.import_module("meta_arch.igev_stereo.igev_stereo")
    g = torch.Generator().manual_seed(78)
    hg = igev.hourglass(8).eval()
    with torch.no_grad():
        for name, prm in hg.named_parameters():
            fan = prm[0].numel() if prm.dim() > 1 else 1
            prm.copy_(torch.randn(prm.shape, generator=g) * (1.5 / fan ** 0.5 if prm.dim() > 1 else 0.3))
            if name.endswith("bn.weight"):
                prm.add_(1.0)
        for name, buf in hg.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.2)
            elif name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
        x = torch.randn(B, 8, D, H, W, generator=g)
        feats = [torch.zeros(1), torch.randn(B, 64, H // 2, W // 2, generator=g), torch.randn(B, 192, H // 4, W // 4, generator=g),
                 torch.randn(B, 160, H // 8, W // 8, generator=g)]
        out = hg(x, feats)
    sd = {k: v for k, v in hg.state_dict().items() if not k.endswith("num_batches_tracked")}
    save(tag, x=x, feat1=feats[1], feat2=feats[2], feat3=feats[3], out=out, keys=np.array(sorted(sd)),
         **{"sd__" + k.replace(".", "__"): sd[k] for k in sd})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    m = import_reference()
    jobs = {
        "corr": lambda: golden_corr(m),
        "geo": lambda: golden_geo(m),
        "update": lambda: golden_update(m),
        "raft_small": lambda: golden_raft_forward(m),
        "raft_shift": lambda: golden_raft_forward(m, 64, 96, 6, "raft_fwd_shift", 1, "shift"),
        "raft_cfg1": lambda: golden_raft_forward(m, 256, 512, 12, "raft_fwd_cfg1", 1, "noise"),
        # the headline workload itself (BASELINE configs[1] resolution and iteration count), one pair
        "raft_cfg2": lambda: golden_raft_forward(m, 544, 960, 32, "raft_fwd_cfg2", 1, "noise"),
        # a second (weights, images) sample of the headline workload
        "raft_cfg2_s2": lambda: golden_raft_forward(m, 544, 960, 32, "raft_fwd_cfg2_s2", 1, "noise", wseed=1, iseed=4321),
        # BASELINE configs[3] resolution (w/4 = 320 > one UMMA N extent), one pair
        "raft_cfg4": lambda: golden_raft_forward(m, 736, 1280, 32, "raft_fwd_cfg4shape", 1, "noise"),
        # BASELINE configs[2] / configs[4] through the reference's public IGEVStereo.forward()
        "igev_cfg3": lambda: golden_igev_forward_full(m, 544, 960, 32, "igev_fwd_cfg3"),
        "igev_cfg3_shift": lambda: golden_igev_forward_full(m, 544, 960, 32, "igev_fwd_cfg3_shift", mode="shift"),
        "igev_cfg5": lambda: golden_igev_forward_full(m, 1024, 1536, 22, "igev_fwd_cfg5shape"),
        # slow_fast_gru=True (raft_stereo.py:157-160 / igev_stereo.py:201-204): extra coarse-GRU updates per iteration
        "raft_slowfast": lambda: golden_raft_forward(m, 64, 96, 4, "raft_fwd_slowfast", 1, "noise", slow_fast_gru=True),
        "igev_slowfast": lambda: golden_igev_forward(m, 64, 96, 4, "igev_fwd_slowfast", 1, slow_fast_gru=True),
        # n_gru_layers 1 / 2 (core/update.py:104-105,121-132; meta_arch/igev_stereo/update.py:111-112,126-135)
        "raft_gru1": lambda: golden_raft_forward(m, 64, 96, 4, "raft_fwd_gru1", 1, "noise", n_gru_layers=1),
        "raft_gru2": lambda: golden_raft_forward(m, 64, 96, 4, "raft_fwd_gru2", 1, "noise", n_gru_layers=2, slow_fast_gru=True),
        "igev_gru2": lambda: golden_igev_forward(m, 64, 96, 4, "igev_fwd_gru2", 1, n_gru_layers=2),
        # the upstream "real-time" RAFT-Stereo settings: shared backbone, 1/8 resolution, two GRU levels, slow-fast schedule
        "raft_realtime": lambda: golden_raft_forward(m, 128, 192, 5, "raft_fwd_realtime", 1, "noise", shared_backbone=True,
                                                     n_downsample=3, n_gru_layers=2, slow_fast_gru=True),
        "all_preds": lambda: golden_all_predictions(m),
        "igev_small": lambda: golden_igev_forward(m),
        "igev_mid": lambda: golden_igev_forward(m, 96, 160, 8, "igev_fwd_mid", 1),
        "igev_volume": lambda: golden_igev_volume(m),
        "igev_hourglass": lambda: golden_igev_hourglass(m),
    }
    for k, fn in jobs.items():
        if not args.only or k in args.only.split(","):
            fn()


if __name__ == "__main__":
    main()
