"""Container-only checks against the real reference checkout (skipped on the GPU box, where
/root/reference does not exist): the drop-in models expose exactly the reference's state-dict keys
and shapes, and the PyTorch-side IGEV pre-loop reproduces the reference's pre-loop products."""
import os
import sys
from argparse import Namespace

import pytest
import torch

from helpers import RAFT_CFG, IGEV_CFG, load_golden, stats, tv_to_timm

REF = os.environ.get("DKT_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _ref_modules():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import make_golden as MG
    m = MG.import_reference()
    MG._timm_stub()
    return MG, m


def test_raft_state_dict_matches_reference():
    MG, m = _ref_modules()
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    ref = m["raft"].RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG))
    mine = RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG))
    a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert a == b


def test_igev_state_dict_and_preloop_match_reference():
    import importlib
    MG, m = _ref_modules()
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    from dkt_stereo_b200.synthetic import synthetic_state_dict, shapes_of, synthetic_pair
    igev = importlib.import_module("meta_arch.igev_stereo.igev_stereo")
    ref = igev.IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval()
    mine = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval()
    sd_ref = synthetic_state_dict(shapes_of(ref.state_dict()), seed=0)
    mapped = {tv_to_timm(k): v for k, v in sd_ref.items()}
    a = {k: tuple(v.shape) for k, v in mapped.items()}
    b = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert a == b, (sorted(set(a) ^ set(b))[:10])
    mine.load_state_dict(mapped, strict=True)
    g = load_golden("igev_fwd_small")
    B, H, W, iters = [int(v) for v in g["meta"]]
    im1, im2 = synthetic_pair(B, H, W, seed=1234, mode="noise")
    with torch.no_grad():
        ml, mr, gev, init_disp, net, ctx, stem_2x = mine.prepare(im1, im2)
    for name, val in (("match_left", ml), ("match_right", mr), ("gev", gev), ("init_disp", init_disp), ("stem_2x", stem_2x)):
        assert stats(val, g[name])[1] < 2e-4, (name, stats(val, g[name]))
    for i in range(3):
        assert stats(net[i], g[f"net{i}"])[1] < 1e-4
        assert stats(ctx[i], g[f"ctx{i}"])[1] < 1e-4
