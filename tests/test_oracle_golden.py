"""CPU: the oracle restatement is pinned against vectors produced by the real reference
(oracle/make_golden.py).  Tolerances are fp32 round-off of a different summation order."""
import torch

from oracle import hotpath as O
from dkt_stereo_b200.synthetic import synthetic_state_dict, synthetic_pair
from helpers import load_golden, golden_shapes, golden_state_dict, stats, RAFT_CFG, IGEV_CFG


def test_corr1d_build_and_lookup():
    for tag in ("a", "b"):
        g = load_golden(f"corr1d_{tag}")
        pyr = O.corr1d_pyramid(g["fmap1"], g["fmap2"], 4)
        for i in range(4):
            assert pyr[i].shape == g[f"pyr{i}"].shape
            assert stats(pyr[i], g[f"pyr{i}"])[1] < 2e-5
        out = O.corr1d_lookup(pyr, g["coords"][:, 0], 4)
        assert out.shape == g["out"].shape
        assert stats(out, g["out"])[1] < 2e-5


def test_geo_lookup():
    g = load_golden("geo_a")
    geo, init = O.geo_pyramids(g["fmap1"], g["fmap2"], g["gev"], 2)
    out = O.geo_lookup(geo, init, g["disp"], 4)
    assert out.shape == g["out"].shape == (2, 162, 4, 36)
    assert stats(out, g["out"])[1] < 2e-5


def _update(tag, igev):
    g = load_golden(f"update_{tag}")
    sd = synthetic_state_dict(golden_shapes(g), seed=3)
    sd = {"update_block." + k: v for k, v in sd.items()}
    net = [g[f"net{i}"] for i in range(3)]
    inp = [[g[f"c{n}{i}"] for n in "zrq"] for i in range(3)]
    with torch.no_grad():
        net_o, mask, delta = O.update_block(sd, "update_block.", net, inp, g["corr"], g["flow"], igev=igev)
    for i in range(3):
        assert stats(net_o[i], g[f"net_out{i}"])[1] < 1e-5
    assert stats(mask, g["mask"])[1] < 1e-5
    assert stats(delta, g["delta"])[1] < 1e-5


def test_update_block_raft():
    _update("raft", False)


def test_update_block_igev():
    _update("igev", True)


def test_upsamplers():
    g = load_golden("convex_upsample")
    assert stats(O.convex_upsample(g["flow"], g["mask"], 4), g["out"])[1] < 1e-5
    g = load_golden("context_upsample")
    assert stats(O.context_upsample(g["disp"], g["weights"]), g["out"])[1] < 1e-5


def test_raft_forward_small_and_shift():
    for tag in ("raft_fwd_small", "raft_fwd_shift"):
        g = load_golden(tag)
        B, H, W, iters = [int(v) for v in g["meta"]]
        sd = synthetic_state_dict(golden_shapes(g), seed=0)
        im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
        lr, up = O.raft_forward(sd, im1, im2, iters, RAFT_CFG)
        assert up.shape == g["flow_up"].shape == (B, 1, H, W)
        mean, mx = stats(up, g["flow_up"])
        assert mean < 1e-4, (tag, mean, mx)
        assert stats(lr, g["flow_lr"])[0] < 1e-4


def test_slow_fast_gru_schedule():
    """slow_fast_gru=True (reference raft_stereo.py:157-160, igev_stereo.py:201-204): the extra coarse-GRU updates."""
    g = load_golden("raft_fwd_slowfast")
    B, H, W, iters = [int(v) for v in g["meta"]]
    sd = synthetic_state_dict(golden_shapes(g), seed=0)
    im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
    cfg = dict(RAFT_CFG, slow_fast_gru=True)
    lr, up = O.raft_forward(sd, im1, im2, iters, cfg)
    assert stats(up, g["flow_up"])[0] < 1e-4
    assert stats(O.raft_forward(sd, im1, im2, iters, RAFT_CFG)[1], g["flow_up"])[0] > 1e-2     # the flag matters
    g = load_golden("igev_fwd_slowfast")
    B, H, W, iters = [int(v) for v in g["meta"]]
    sd = synthetic_state_dict(golden_shapes(g), seed=0)
    net = [g[f"net{i}"] for i in range(3)]
    inp = [list(g[f"ctx{i}"].split(128, dim=1)) for i in range(3)]
    with torch.no_grad():
        up = O.igev_loop(sd, g["match_left"], g["match_right"], g["gev"], g["init_disp"], net, inp,
                         g["stem_2x"], iters, dict(IGEV_CFG, slow_fast_gru=True))
    assert stats(up, g["disp_up"])[0] < 1e-4


def test_n_gru_layers_one_and_two():
    """args.n_gru_layers = 1 / 2 (reference core/update.py:104-105,121-132; igev update.py:111-112,126-135): levels that do
    not run contribute no GRU input; n = 2 is generated with slow_fast_gru so that its mid-level-only extra update
    (raft_stereo.py:159-160) is pinned too."""
    for tag, over in (("raft_fwd_gru1", dict(n_gru_layers=1)), ("raft_fwd_gru2", dict(n_gru_layers=2, slow_fast_gru=True))):
        g = load_golden(tag)
        B, H, W, iters = [int(v) for v in g["meta"]]
        sd = synthetic_state_dict(golden_shapes(g), seed=0)
        im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
        lr, up = O.raft_forward(sd, im1, im2, iters, dict(RAFT_CFG, **over))
        assert stats(up, g["flow_up"])[0] < 1e-4, (tag, stats(up, g["flow_up"]))
    g = load_golden("igev_fwd_gru2")
    B, H, W, iters = [int(v) for v in g["meta"]]
    sd = synthetic_state_dict(golden_shapes(g), seed=0)
    net = [g[f"net{i}"] for i in range(2)]
    inp = [list(g[f"ctx{i}"].split(128, dim=1)) for i in range(2)]
    with torch.no_grad():
        up = O.igev_loop(sd, g["match_left"], g["match_right"], g["gev"], g["init_disp"], net, inp,
                         g["stem_2x"], iters, dict(IGEV_CFG, n_gru_layers=2))
    assert stats(up, g["disp_up"])[0] < 1e-4


def test_raft_forward_headline_config():
    """The oracle at the benchmark's own resolution and iteration count (544 x 960, 32 iterations, one pair) against
    the real reference's disparity map (tests/golden/raft_fwd_cfg2.npz, oracle/make_golden.py --only raft_cfg2)."""
    g = load_golden("raft_fwd_cfg2")
    B, H, W, iters = [int(v) for v in g["meta"]]
    assert (H, W, iters) == (544, 960, 32)
    sd = synthetic_state_dict(golden_shapes(g), seed=0)
    im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
    lr, up = O.raft_forward(sd, im1, im2, iters, RAFT_CFG)
    mean, mx = stats(up, g["flow_up"])
    assert mean < 1e-4, (mean, mx)


def test_igev_loop_against_reference_forward():
    """The oracle's IGEV hot loop, fed with the pre-loop products the REAL reference forward produced
    (captured by hooks in oracle/make_golden.py), reproduces the reference's final disparity."""
    for tag in ("igev_fwd_small", "igev_fwd_mid"):
        g = load_golden(tag)
        B, H, W, iters = [int(v) for v in g["meta"]]
        sd = synthetic_state_dict(golden_shapes(g), seed=0)
        net = [g[f"net{i}"] for i in range(3)]
        inp = [list(g[f"ctx{i}"].split(128, dim=1)) for i in range(3)]
        with torch.no_grad():
            up = O.igev_loop(sd, g["match_left"], g["match_right"], g["gev"], g["init_disp"], net, inp,
                             g["stem_2x"], iters, IGEV_CFG)
        assert up.shape == g["disp_up"].shape == (B, 1, H, W)
        mean, mx = stats(up, g["disp_up"])
        assert mean < 1e-4, (tag, mean, mx)


def test_igev_volume_stage():
    """GWC volume, 3-D corr_stem + BatchNorm + LeakyReLU + feature attention, classifier, soft-argmin against the real
    reference's modules (tests/golden/igev_volume.npz, oracle/make_golden.py --only igev_volume)."""
    g = load_golden("igev_volume")
    B, C, H, W, D = [int(v) for v in g["meta"]]
    gwc = O.gwc_volume(g["left"], g["right"], D, 8)
    assert gwc.shape == g["gwc"].shape == (B, 8, D, H, W)
    assert stats(gwc, g["gwc"])[1] < 2e-6
    bn = dict(weight=g["bn_weight"], bias=g["bn_bias"], running_mean=g["bn_mean"], running_var=g["bn_var"], eps=float(g["bn_eps"]))
    vol = O.conv3d_bn_leaky_att(g["gwc"], g["stem_w"], bn, 0.01, g["att_logits"])
    assert stats(vol, g["vol"])[1] < 2e-5
    logits = O.conv3d_bn_leaky_att(g["vol"], g["cls_w"], None, 1.0, None).squeeze(1)
    assert stats(logits, g["logits"])[1] < 2e-5
    disp = O.softargmin(g["logits"])
    assert disp.shape == g["disp"].shape == (B, 1, H, W)
    assert stats(disp, g["disp"])[1] < 2e-5


def test_igev_hourglass():
    """The oracle's 3-D hourglass against the real reference module's output (tests/golden/igev_hourglass.npz)."""
    g = load_golden("igev_hourglass")
    sd = {"hg." + k: v for k, v in golden_state_dict(g).items()}
    out = O.hourglass(sd, "hg", g["x"], [None, g["feat1"], g["feat2"], g["feat3"]])
    assert out.shape == g["out"].shape
    tol = 2e-5 * (float(g["out"].abs().max()) + 1.0)
    assert stats(out, g["out"])[1] < tol
    # the engine's module carries the reference's parameter names (strict load) and computes the same function
    from dkt_stereo_b200.igev_modules import Hourglass
    hg = Hourglass(8).eval()
    hg.load_state_dict(golden_state_dict(g), strict=True)
    with torch.no_grad():
        mine = hg(g["x"], [None, g["feat1"], g["feat2"], g["feat3"]])
    assert stats(mine, g["out"])[1] < tol
