"""CPU: host-side logic (checkpoint key compatibility, weight packing, padding, synthetic data)."""
from argparse import Namespace

import pytest
import torch

from helpers import load_golden, golden_shapes, RAFT_CFG, IGEV_CFG


def test_raft_state_dict_keys_match_reference():
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    g = load_golden("raft_fwd_small")
    ref = golden_shapes(g)                       # keys + shapes of the reference model's state_dict
    model = RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG))
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref
    # DataParallel-style checkpoints ("module." prefix, evaluate_stereo.py:361-370) load too
    wrapped = torch.nn.DataParallel(model)
    assert sorted(wrapped.state_dict()) == sorted("module." + k for k in ref)


def test_update_block_keys_match_reference():
    from dkt_stereo_b200.update import BasicMultiUpdateBlock
    for tag, cfg, igev in (("raft", RAFT_CFG, False), ("igev", IGEV_CFG, True)):
        ref = golden_shapes(load_golden(f"update_{tag}"))
        blk = BasicMultiUpdateBlock(Namespace(**cfg), hidden_dims=cfg["hidden_dims"], igev=igev)
        assert {k: tuple(v.shape) for k, v in blk.state_dict().items()} == ref


def test_training_mode_and_cpu_inputs_fail_loudly():
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    from dkt_stereo_b200._lib import DktError
    model = RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG)).eval()
    x = torch.zeros(1, 3, 32, 64)
    with pytest.raises(NotImplementedError):
        model(x, x, iters=1, test_mode=False)
    with pytest.raises(DktError):
        model(x, x, iters=1, test_mode=True)       # no CPU fallback
    with pytest.raises(NotImplementedError):
        RAFTStereo(Namespace(mixed_precision=False, **dict(RAFT_CFG, corr_implementation="alt")))


def test_pack_conv_layouts():
    from dkt_stereo_b200 import ops
    w = torch.randn(5, 7, 3, 3)
    b = torch.randn(5)
    p = ops.pack_conv(w, b, cin_pad=16, tc=True)
    assert p.w_simt.shape == (9, 16, 5) and p.w_hi.shape == (9, 16, 16) and p.cin == 16 and p.n == 5
    assert torch.equal(p.w_simt[4, :7, :], w[:, :, 1, 1].t())
    assert float(p.w_simt[:, 7:, :].abs().max()) == 0
    rec = p.w_hi.float() + p.w_lo.float()
    assert torch.allclose(rec[4, :5, :7], w[:, :, 1, 1], rtol=2e-5, atol=1e-6)
    assert float(rec[:, 5:, :].abs().max()) == 0
    hi, lo = ops.split_bf16(torch.tensor([1.2345678, -3.1415927e-3, 1e4]))
    err = (hi.float() + lo.float() - torch.tensor([1.2345678, -3.1415927e-3, 1e4])).abs()
    assert torch.all(err <= torch.tensor([1.2345678, 3.1415927e-3, 1e4]) * 2.0 ** -16)


def test_teacher_weight_updates_trigger_a_repack():
    """The drop-in as frozen / EMA teacher of the reference's fine-tuning loop (tools/ft_dkt.py:140-151,179-199): a
    DataParallel-wrapped model receives `load_state_dict` and, every step, the EMA assignment
    `t_params.data = ema * t_params.data + (1 - ema) * s_params.data`.  Both must be noticed by the engine (packed
    weights and captured CUDA graphs are stale afterwards); packing itself runs on the parameters' device, here the CPU."""
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    torch.manual_seed(0)
    teacher = torch.nn.DataParallel(RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG)))
    student = torch.nn.DataParallel(RAFTStereo(Namespace(mixed_precision=False, **RAFT_CFG)))
    for p in teacher.parameters():
        p.requires_grad = False
    teacher.eval()
    teacher.module.freeze_bn()
    eng = teacher.module.engine
    assert eng.pack_weights() is True and eng.pack_weights() is False
    w_before = eng.weights["zr0"].w_hi.clone()
    # EMA step exactly as the reference writes it
    for t_params, s_params in zip(teacher.parameters(), student.parameters()):
        t_params.data = (0.9 * t_params.data + 0.1 * s_params.data)
        t_params.requires_grad = False
    assert eng.pack_weights() is True
    assert not torch.equal(eng.weights["zr0"].w_hi, w_before)
    g = teacher.module.update_block.gru08
    ref = torch.cat([g.convz.weight, g.convr.weight]).detach().permute(2, 3, 0, 1).reshape(9, 256, 384)
    assert torch.allclose(eng.weights["zr0"].w_hi.float() + eng.weights["zr0"].w_lo.float(), ref, rtol=1e-5, atol=1e-7)
    assert eng.pack_weights() is False
    # checkpoint load into the wrapper ("module." keys, reference ft_dkt.py:136-151)
    teacher.load_state_dict(student.state_dict(), strict=True)
    assert eng.pack_weights() is True
    # in-place edits through .data bypass autograd's version counter: the explicit hook covers them
    with torch.no_grad():
        for p in teacher.parameters():
            p.data.mul_(0.5)
    teacher.module.invalidate_weights()
    assert eng.pack_weights() is True


def test_input_padder_roundtrip():
    from dkt_stereo_b200.utils import InputPadder, coords_grid
    x = torch.arange(2 * 3 * 37 * 50, dtype=torch.float32).view(2, 3, 37, 50)
    p = InputPadder(x.shape, divis_by=32)
    (y,) = p.pad(x)
    assert y.shape[-2] % 32 == 0 and y.shape[-1] % 32 == 0
    assert torch.equal(p.unpad(y), x)
    c = coords_grid(1, 3, 4)
    assert c[0, 0, 2, 3] == 3 and c[0, 1, 2, 3] == 2


def test_synthetic_is_deterministic():
    from dkt_stereo_b200.synthetic import synthetic_pair, synthetic_state_dict
    a, b = synthetic_pair(1, 8, 16)
    a2, b2 = synthetic_pair(1, 8, 16)
    assert torch.equal(a, a2) and torch.equal(b, b2) and not torch.equal(a, b)
    s1 = synthetic_state_dict({"x.weight": (4, 3, 3, 3), "y.norm3.weight": (4,), "y.downsample.1.weight": (4,)})
    assert torch.equal(s1["y.norm3.weight"], s1["y.downsample.1.weight"])


def test_registry_and_igev_hot_path_keys():
    import dkt_stereo_b200 as pkg
    assert "RAFTStereo" in pkg.__models__ and "IGEVStereo" in pkg.__models__
    model = pkg.__models__["IGEVStereo"](Namespace(mixed_precision=False, **IGEV_CFG))
    mine = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    ref = golden_shapes(load_golden("igev_fwd_small"))     # reference keys of update_block.*, spx_2_gru.*, spx_gru.*
    assert ref and all(mine.get(k) == v for k, v in ref.items())
    with pytest.raises(KeyError):
        pkg.__models__["PCVNet"]


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): exactly one stdout line, JSON, with
    the keys of the measurement contract; library chatter must not reach stdout.  Tiny shape so it runs in seconds."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "64", "--width", "96", "--iters", "2"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    # "reference" where oracle/_ref holds the reference's compiled modules (build container, GPU box), else the port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    # non-zero ranks of a torchrun launch exit quietly without work
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, cwd=root, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
