"""GPU parity tests: every kernel family is called through the C ABI (ctypes) and compared with
(a) the golden vectors produced by the real reference and (b) the CPU oracle on seeded inputs.

Tolerances (fp32 data, written next to each assert):
  * fp32 CUDA-core kernels: 2e-5 abs (different summation order only)
  * tensor-core kernels (3-term bf16 split, ~2^-16 relative per product): 1e-4 .. 1e-3 abs on
    O(1..10) values, stated per test
  * end to end: mean |disparity difference| <= 1e-3 px (the north-star gate)
"""
from argparse import Namespace

import pytest
import torch

from helpers import load_golden, golden_shapes, golden_state_dict, golden_seeds, stats, tv_to_timm, RAFT_CFG, IGEV_CFG

pytestmark = pytest.mark.gpu

IMPLS = ["simt", "tc"]


def dev():
    return torch.device("cuda:0")


def L16():
    """torch dtype of the library's 16-bit (hi, lo) planes (half by default)."""
    from dkt_stereo_b200 import _lib
    return _lib.split_dtype()


# ---------------------------------------------------------------------------------------------
# K1 / K2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_corr1d_build_golden(impl, tag):
    from dkt_stereo_b200 import ops
    g = load_golden(f"corr1d_{tag}")
    f1, f2 = g["fmap1"].to(dev()), g["fmap2"].to(dev())
    D = f1.shape[1]
    pyr = ops.corr1d_build(f1, f2, 4, 1.0 / D ** 0.5, impl=impl)
    tol = 2e-5 if impl == "simt" else 2e-4
    for i in range(4):
        ref = g[f"pyr{i}"]
        assert pyr[i].shape == ref.shape
        assert stats(pyr[i].cpu(), ref)[1] < tol, (impl, tag, i, stats(pyr[i].cpu(), ref))


@pytest.mark.parametrize("impl", IMPLS)
def test_corr1d_build_channels_last_and_sizes(impl):
    """Strided (channels_last) feature maps and the cfg-sized row width 240 with D=256."""
    from dkt_stereo_b200 import ops
    from oracle import hotpath as O
    g = torch.Generator().manual_seed(5)
    f1 = torch.randn(1, 256, 3, 240, generator=g)
    f2 = torch.randn(1, 256, 3, 240, generator=g)
    ref = O.corr1d_pyramid(f1, f2, 4)
    a = f1.to(dev()).contiguous(memory_format=torch.channels_last)
    b = f2.to(dev()).contiguous(memory_format=torch.channels_last)
    pyr = ops.corr1d_build(a, b, 4, 1.0 / 16.0, impl=impl)
    tol = 3e-5 if impl == "simt" else 5e-4
    for i in range(4):
        assert stats(pyr[i].cpu(), ref[i])[1] < tol, (impl, i, stats(pyr[i].cpu(), ref[i]))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("w", [320, 384, 272, 500])
def test_corr1d_build_wide_rows(impl, w):
    """Rows wider than one UMMA N extent (BASELINE configs[3]/[4]: w = 320 / 384): the tensor-core build cuts w2 into
    equal column blocks; every pyramid level must still match the oracle, including ragged last blocks."""
    from dkt_stereo_b200 import ops
    from oracle import hotpath as O
    g = torch.Generator().manual_seed(w)
    f1 = torch.randn(1, 128, 2, w, generator=g)
    f2 = torch.randn(1, 128, 2, w, generator=g)
    ref = O.corr1d_pyramid(f1, f2, 4)
    pyr = ops.corr1d_build(f1.to(dev()), f2.to(dev()), 4, 128 ** -0.5, impl=impl)
    tol = 3e-5 if impl == "simt" else 5e-4
    for i in range(4):
        assert pyr[i].shape == ref[i].shape
        assert stats(pyr[i].cpu(), ref[i])[1] < tol, (impl, w, i, stats(pyr[i].cpu(), ref[i]))


def test_corr1d_build_right_map_narrower_than_left():
    """W2 != W1 (the API and the C ABI carry both): the right map is read with its OWN strides (round-1 advisor finding:
    it used to be read with the left map's); the fp32 entry point, which has one stride set, must refuse instead."""
    from dkt_stereo_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(11)
    f1 = torch.randn(2, 64, 3, 48, generator=g)
    f2 = torch.randn(2, 64, 3, 32, generator=g)
    ref = torch.einsum("bdhi,bdhj->bhij", f1, f2) / 8.0
    pyr = ops.corr1d_build(f1.to(dev()), f2.to(dev()), 2, 1.0 / 8.0, impl="tc")
    assert pyr[0].shape == (2, 3, 48, 32) and pyr[1].shape == (2, 3, 48, 16)
    assert stats(pyr[0].cpu(), ref)[1] < 5e-4, stats(pyr[0].cpu(), ref)
    assert stats(pyr[1].cpu(), 0.5 * (ref[..., 0::2] + ref[..., 1::2]))[1] < 5e-4
    with pytest.raises(L.DktError):
        ops.corr1d_build(f1.to(dev()), f2.to(dev()), 2, 1.0 / 8.0, impl="simt")


def test_raft_forward_wide_image_runs_native():
    """736 x 1280 (BASELINE configs[3], w = 320 > 256) goes through the native tensor-core path end to end and agrees with
    the exact-fp32 CUDA-core path of the same engine within the end-to-end gate."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    tc, simt = _model("tc", g), _model("simt", g)
    im1, im2 = synthetic_pair(1, 736, 1280, seed=9, mode="shift")
    im1, im2 = im1.to(dev()), im2.to(dev())
    _, up_tc = tc(im1, im2, iters=4, test_mode=True)
    _, up_f32 = simt(im1, im2, iters=4, test_mode=True)
    mean, mx = stats(up_tc.cpu(), up_f32.cpu())
    print(f"[parity] 736x1280 tc vs fp32 engine: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
    assert up_tc.shape == (1, 1, 736, 1280) and mean <= 1e-3, (mean, mx)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_corr1d_lookup_golden(tag):
    from dkt_stereo_b200.corr import B200CorrBlock1D, corr_sampler_forward
    g = load_golden(f"corr1d_{tag}")
    blk = B200CorrBlock1D(g["fmap1"].to(dev()), g["fmap2"].to(dev()), num_levels=4, radius=4, impl="simt")
    out = blk(g["coords"].to(dev()))
    assert out.shape == g["out"].shape
    assert stats(out.cpu(), g["out"])[1] < 2e-5
    # plug-in ABI of the reference's corr_sampler extension, one level at a time
    lvl = 1
    (o1,) = corr_sampler_forward(g[f"pyr{lvl}"].to(dev()), g["coords"][:, :1].to(dev()) / 2 ** lvl, 4)
    assert stats(o1.cpu(), g["out"][:, 9 * lvl:9 * (lvl + 1)])[1] < 2e-5


@pytest.mark.parametrize("shape", [(2, 5, 40, 40, 4), (1, 3, 53, 26, 4), (1, 2, 33, 70, 2)])
def test_corr_sampler_backward_vs_autograd(shape):
    """corr_sampler.backward (reference core/corr.py:25-29): the gradient of the one-level lookup with respect to the
    volume, against autograd through the oracle's differentiable restatement of the same lookup (fp64)."""
    from dkt_stereo_b200.corr import CorrSampler, corr_sampler_backward
    from oracle import hotpath as O
    B, H, W1, W2, r = shape
    g = torch.Generator().manual_seed(W1 + W2)
    vol = torch.randn(B, H, W1, W2, generator=g)
    cx = torch.rand(B, 1, H, W1, generator=g) * (W2 + 12) - 6             # some taps out of range on both sides
    cx[0, 0, 0, :4] = torch.tensor([-1.0, 0.0, W2 - 1.0, float(W2)])
    gout = torch.randn(B, 2 * r + 1, H, W1, generator=g)
    v64 = vol.double().requires_grad_(True)
    out = O.corr1d_lookup([v64], cx[:, 0].double(), r)
    out.backward(gout.double())
    (gv,) = corr_sampler_backward(vol.to(dev()), cx.to(dev()), gout.to(dev()), r)
    assert gv.shape == vol.shape
    assert stats(gv.cpu(), v64.grad.float())[1] < 1e-5, stats(gv.cpu(), v64.grad.float())
    if r != 4:                 # the forward lookup is built for the shipped corr_radius = 4; the adjoint takes any radius
        return
    # the autograd wrapper the reference defines (CorrSampler.apply) end to end
    vd = vol.to(dev()).requires_grad_(True)
    o = CorrSampler.apply(vd, cx.to(dev()), r)
    assert stats(o.detach().cpu(), out.detach().float())[1] < 2e-5
    o.backward(gout.to(dev()))
    assert stats(vd.grad.cpu(), v64.grad.float())[1] < 1e-5


def test_lookup_fused_coordinate_update():
    """delta add + flow bookkeeping fused in front of the gather (raft_stereo.py:154-155,164-167)."""
    from dkt_stereo_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 4, 40
    cx = (torch.rand(B, H, W, generator=g) * W).to(dev())
    delta = torch.randn(B, H, W, 2, generator=g).to(dev())
    flow = torch.zeros(B, H, W, 2, device=dev())
    flow[..., 1] = 7.0
    cx0 = cx.clone()
    ops.corr1d_lookup([], cx, 4, None, delta=delta, flow=flow)
    xs = torch.arange(W, device=dev()).view(1, 1, W)
    assert torch.equal(cx, cx0 + delta[..., 0])
    assert torch.equal(flow[..., 0], cx - xs)
    assert torch.all(flow[..., 1] == 7.0)


def test_geo_golden():
    from dkt_stereo_b200 import ops
    g = load_golden("geo_a")
    f1, f2, gev, disp = (g[k].to(dev()) for k in ("fmap1", "fmap2", "gev", "disp"))
    init = ops.corr1d_build(f1, f2, 2, 1.0, impl="simt")
    geo = ops.geo_pool(gev)
    B, _, H, W = disp.shape
    out = torch.empty(B, 162, H, W, device=dev())
    ops.geo_lookup(geo, init, disp[:, 0].contiguous(), 4, out, out_layout="nchw")
    # init-corr values are unscaled dot products (|v| up to ~20 here): 1e-4 abs ~ 5e-6 relative
    assert stats(out.cpu(), g["out"])[1] < 1e-4
    assert stats(out.cpu(), g["out"])[0] < 5e-6


def test_corr1d_lookup_nhwc_fast_path_and_enc():
    """NHWC (padded, coalesced) output incl. bf16 hi/lo, ragged pixel count (not a multiple of 32), and
    the lookup fused with convc1 + ReLU (core/update.py:72,79) against the oracle."""
    from dkt_stereo_b200 import ops, _lib as L
    from oracle import hotpath as O
    g = torch.Generator().manual_seed(11)
    B, D, H, W = 2, 32, 5, 43                         # P = 430: last CTA has 14 pixels
    f1, f2 = torch.randn(B, D, H, W, generator=g), torch.randn(B, D, H, W, generator=g)
    pyr_ref = O.corr1d_pyramid(f1, f2, 4)
    cx = torch.arange(W).view(1, 1, W).float() + (torch.rand(B, H, W, generator=g) * (W + 16) - W / 2 - 8)
    delta = torch.randn(B, H, W, 2, generator=g)
    ref = O.corr1d_lookup(pyr_ref, cx + delta[..., 0], 4)              # (B,36,H,W)
    pyr = ops.corr1d_build(f1.to(dev()), f2.to(dev()), 4, 1.0 / D ** 0.5, impl="simt")
    cxd = cx.to(dev()).contiguous()
    out = torch.full((B, H, W, 64), 7.0, device=dev())
    hi = torch.zeros(B, H, W, 64, device=dev(), dtype=L16())
    lo = torch.zeros_like(hi)
    flow = torch.zeros(B, H, W, 2, device=dev())
    ops.corr1d_lookup(pyr, cxd, 4, out, "nhwc", out_hi=hi, out_lo=lo, delta=delta.to(dev()), flow=flow)
    assert torch.equal(cxd.cpu(), cx + delta[..., 0])
    assert stats(out[..., :36].permute(0, 3, 1, 2).cpu(), ref)[1] < 2e-5
    assert torch.all(out[..., 36:] == 0)                               # padding channels are zero-filled
    assert stats((hi.float() + lo.float())[..., :36].cpu(), out[..., :36].cpu())[1] < 2e-4   # 16-bit split
    # fused encoder
    wt = torch.randn(64, 36, 1, 1, generator=g) / 6.0
    bias = torch.randn(64, generator=g)
    enc_ref = torch.relu(torch.nn.functional.conv2d(ref, wt, bias))
    Wc = ops.pack_conv(wt.to(dev()), bias.to(dev()), cin_pad=64, tc=False)
    enc = torch.zeros(B, H, W, 64, device=dev())
    ehi = torch.zeros(B, H, W, 64, device=dev(), dtype=L16())
    elo = torch.zeros_like(ehi)
    cxd2 = cx.to(dev()).contiguous()
    ops.corr1d_lookup_enc(pyr, cxd2, 4, Wc, L.tensor_slice(enc, ehi, elo, 0, 64), delta=delta.to(dev()), flow=flow)
    assert stats(enc.permute(0, 3, 1, 2).cpu(), enc_ref)[1] < 5e-5
    assert stats((ehi.float() + elo.float()).cpu(), enc.cpu())[1] < 2e-4


def test_geo_lookup_update_nhwc_and_enc():
    """IGEV lookup with the fused `disp += delta` (igev_stereo.py:210), NHWC padded output and the
    fused 162 -> 64 convc1 (igev update.py:76,85)."""
    from dkt_stereo_b200 import ops, _lib as L
    from oracle import hotpath as O
    g = load_golden("geo_a")
    f1, f2, gev, disp = (g[k] for k in ("fmap1", "fmap2", "gev", "disp"))
    gen = torch.Generator().manual_seed(5)
    B, _, H, W = disp.shape
    delta = torch.randn(B, H, W, 1, generator=gen)
    geo_ref, init_ref = O.geo_pyramids(f1, f2, gev, 2)
    ref = O.geo_lookup(geo_ref, init_ref, disp + delta.permute(0, 3, 1, 2), 4)       # (B,162,H,W)
    init = ops.corr1d_build(f1.to(dev()), f2.to(dev()), 2, 1.0, impl="simt")
    geo = ops.geo_pool(gev.to(dev()))
    d = disp[:, 0].contiguous().to(dev())
    out = torch.full((B, H, W, 192), 3.0, device=dev())
    hi = torch.zeros(B, H, W, 192, device=dev(), dtype=L16())
    lo = torch.zeros_like(hi)
    ops.geo_lookup(geo, init, d, 4, out, "nhwc", out_hi=hi, out_lo=lo, delta=delta.to(dev()))
    assert torch.equal(d.cpu(), disp[:, 0] + delta[..., 0])
    assert stats(out[..., :162].permute(0, 3, 1, 2).cpu(), ref)[1] < 1e-4
    assert torch.all(out[..., 162:] == 0)
    wt = torch.randn(64, 162, 1, 1, generator=gen) / 12.0
    bias = torch.randn(64, generator=gen)
    enc_ref = torch.relu(torch.nn.functional.conv2d(ref, wt, bias))
    Wc = ops.pack_conv(wt.to(dev()), bias.to(dev()), cin_pad=192, tc=False)
    enc = torch.zeros(B, H, W, 64, device=dev())
    d2 = disp[:, 0].contiguous().to(dev())
    ops.geo_lookup_enc(geo, init, d2, 4, Wc, L.tensor_slice(enc, None, None, 0, 64), delta=delta.to(dev()))
    assert stats(enc.permute(0, 3, 1, 2).cpu(), enc_ref)[1] < 2e-4


@pytest.mark.parametrize("planes", [2, 1])
@pytest.mark.parametrize("shape", [(2, 32, 5, 43), (1, 64, 9, 240), (3, 16, 2, 130)])
def test_corr1d_lookup_enc_tensor_core(shape, planes):
    """The lookup fused with convc1 on tcgen05 (csrc/lookup_tc.cu): taps + coordinate bookkeeping identical to the
    fp32 kernel, the 36 -> 64 contraction as a 16-bit split with fp32 accumulation.  Ragged chunk (P % 128 != 0),
    several chunks per CTA, out-of-range coordinates; every output plane (fp32, hi, lo / hi only)."""
    from dkt_stereo_b200 import ops, _lib as L
    from oracle import hotpath as O
    B, D, H, W = shape
    g = torch.Generator().manual_seed(D + W)
    f1, f2 = torch.randn(B, D, H, W, generator=g), torch.randn(B, D, H, W, generator=g)
    pyr_ref = O.corr1d_pyramid(f1, f2, 4)
    cx = torch.arange(W).view(1, 1, W).float() + (torch.rand(B, H, W, generator=g) * (W + 16) - W / 2 - 8)
    delta = torch.randn(B, H, W, 2, generator=g)
    ref = O.corr1d_lookup(pyr_ref, cx + delta[..., 0], 4)              # (B,36,H,W)
    wt = torch.randn(64, 36, 1, 1, generator=g) / 6.0
    bias = torch.randn(64, generator=g)
    if planes == 1:
        ref = ref.half().float()                                      # hi-only taps: the half-rounded taps, exactly
    enc_ref = torch.relu(torch.nn.functional.conv2d(ref, wt, bias))
    pyr = ops.corr1d_build(f1.to(dev()), f2.to(dev()), 4, 1.0 / D ** 0.5, impl="simt")
    w_img, b_dev = ops.pack_lookup_tc(wt.to(dev()), bias.to(dev()))
    enc = torch.zeros(B, H, W, 128, device=dev())
    ehi = torch.zeros(B, H, W, 128, device=dev(), dtype=L16())
    elo = torch.zeros_like(ehi)
    flow = torch.zeros(B, H, W, 2, device=dev())
    cxd = cx.to(dev()).contiguous()
    ops.corr1d_lookup_enc_tc(pyr, cxd, 4, w_img, b_dev, L.tensor_slice(enc, ehi, elo, 64, 64), planes,
                             delta=delta.to(dev()), flow=flow)
    torch.cuda.synchronize()
    assert torch.equal(cxd.cpu(), cx + delta[..., 0])                  # coords1 += delta
    assert torch.equal(flow[..., 0].cpu(), cxd.cpu() - torch.arange(W).view(1, 1, W))
    got = enc[..., 64:].permute(0, 3, 1, 2).cpu()
    # O(1..10) outputs; 3-MMA split ~2^-22 per operand, dominated by the pyramid's own fp32 rounding
    assert stats(got, enc_ref)[1] < (2e-4 if planes == 2 else 1e-3), (shape, planes, stats(got, enc_ref))
    assert float(enc[..., :64].abs().max()) == 0 and float(ehi[..., :64].float().abs().max()) == 0
    assert stats((ehi.float() + elo.float())[..., 64:].cpu(), enc[..., 64:].cpu())[1] < 1e-5
    ehi2 = torch.zeros_like(ehi)
    cxd2 = cx.to(dev()).contiguous()
    ops.corr1d_lookup_enc_tc(pyr, cxd2, 4, w_img, b_dev, L.tensor_slice(None, ehi2, None, 64, 64), planes, delta=delta.to(dev()))
    assert torch.equal(ehi2, ehi)                                      # hi-only destination = the hi plane


@pytest.mark.parametrize("planes", [1, 2])
def test_geo_lookup_enc_tensor_core(planes):
    """IGEV combined lookup + 162 -> 64 convc1 on tcgen05, reading the geometry volume in the (B,H,W,D,C) layout of
    dkt_geo_pool_dc; channel order and `disp += delta` as the reference (geometry.py:34-58, igev_stereo.py:210)."""
    from dkt_stereo_b200 import ops, _lib as L
    from oracle import hotpath as O
    g = load_golden("geo_a")
    f1, f2, gev, disp = (g[k] for k in ("fmap1", "fmap2", "gev", "disp"))
    gen = torch.Generator().manual_seed(5)
    B, _, H, W = disp.shape
    delta = torch.randn(B, H, W, 1, generator=gen)
    geo_ref, init_ref = O.geo_pyramids(f1, f2, gev, 2)
    ref = O.geo_lookup(geo_ref, init_ref, disp + delta.permute(0, 3, 1, 2), 4)       # (B,162,H,W)
    if planes == 1:
        ref = ref.half().float()
    wt = torch.randn(64, 162, 1, 1, generator=gen) / 12.0
    bias = torch.randn(64, generator=gen)
    enc_ref = torch.relu(torch.nn.functional.conv2d(ref, wt, bias))
    init = ops.corr1d_build(f1.to(dev()), f2.to(dev()), 2, 1.0, impl="simt")
    geo = ops.geo_pool_dc(gev.to(dev()))
    old = ops.geo_pool(gev.to(dev()))
    assert torch.equal(geo[0], old[0].permute(0, 1, 2, 4, 3).contiguous())       # same numbers, (.., D, C) order
    assert torch.equal(geo[1], old[1].permute(0, 1, 2, 4, 3).contiguous())
    w_img, b_dev = ops.pack_lookup_tc(wt.to(dev()), bias.to(dev()))
    enc = torch.zeros(B, H, W, 64, device=dev())
    d = disp[:, 0].contiguous().to(dev())
    ops.geo_lookup_enc_tc(geo, init, d, 4, w_img, b_dev, L.tensor_slice(enc, None, None, 0, 64), planes, delta=delta.to(dev()))
    torch.cuda.synchronize()
    assert torch.equal(d.cpu(), disp[:, 0] + delta[..., 0])
    got = enc.permute(0, 3, 1, 2).cpu()
    # init-corr taps are unscaled dot products (|v| up to ~20): outputs O(10)
    assert stats(got, enc_ref)[1] < (5e-4 if planes == 2 else 5e-3), (planes, stats(got, enc_ref))


# ---------------------------------------------------------------------------------------------
# K3: single convs with each epilogue vs torch conv2d (fp32 reference of the same op)
# ---------------------------------------------------------------------------------------------
def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _slice_of(x_nhwc, impl, c0=0, cnt=None):
    from dkt_stereo_b200 import ops
    from dkt_stereo_b200._lib import tensor_slice
    if impl == "tc":
        hi, lo = ops.split_bf16(x_nhwc)
        keep = (x_nhwc, hi.contiguous(), lo.contiguous())
        return tensor_slice(None, keep[1], keep[2], c0, cnt), keep
    return tensor_slice(x_nhwc, None, None, c0, cnt), (x_nhwc,)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("shape", [(1, 64, 64, 8, 16, 3), (2, 128, 126, 19, 37, 3), (1, 64, 144, 9, 20, 1),
                                   (1, 256, 2, 11, 18, 3), (2, 384, 256, 17, 30, 3),
                                   # odd tile counts (the CTA-pair kernel's filler tile), resident / streamed filters
                                   (1, 64, 64, 24, 16, 3), (3, 64, 32, 8, 16, 3), (1, 128, 256, 24, 16, 3),
                                   (1, 192, 128, 17, 33, 3)])
def test_conv_linear(impl, shape):
    from dkt_stereo_b200 import ops, _lib as L
    B, Cin, N, H, W, k = shape
    g = torch.Generator().manual_seed(Cin + N)
    x = torch.randn(B, Cin, H, W, generator=g)
    wt = torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(N, generator=g)
    ref = torch.relu(torch.nn.functional.conv2d(x, wt, bias, padding=k // 2))
    xs, keep = _slice_of(_nhwc(x).to(dev()), impl)
    W_ = ops.pack_conv(wt.to(dev()), bias.to(dev()), tc=(impl == "tc"))
    Cout = (N + 3) // 4 * 4 if N >= 4 else N
    out = torch.zeros(B, H, W, Cout, device=dev())
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, None, None, 0, N), act=L.ACT_RELU, bias=W_.bias)
    ops.conv2d([xs], W_, e, B, H, W, impl)
    torch.cuda.synchronize()
    got = out[..., :N].permute(0, 3, 1, 2).cpu()
    tol = 2e-5 if impl == "simt" else 2e-4
    assert stats(got, ref)[1] < tol, (impl, shape, stats(got, ref))


@pytest.mark.parametrize("impl", IMPLS)
def test_conv_two_sources_and_tail(impl):
    """channel-concat of two slices (one of them a sub-range of a wider buffer) + flow tail."""
    from dkt_stereo_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(9)
    B, H, W = 1, 13, 21
    a = torch.randn(B, 64, H, W, generator=g)
    big = torch.randn(B, 192, H, W, generator=g)
    flow = torch.randn(B, 2, H, W, generator=g)
    wt = torch.randn(126, 128, 3, 3, generator=g) / 34.0
    bias = torch.randn(126, generator=g)
    ref = torch.relu(torch.nn.functional.conv2d(torch.cat([a, big[:, 64:128]], 1), wt, bias, padding=1))
    ref = torch.cat([ref, flow], 1)
    s0, k0 = _slice_of(_nhwc(a).to(dev()), impl)
    s1, k1 = _slice_of(_nhwc(big).to(dev()), impl, 64, 64)
    W_ = ops.pack_conv(wt.to(dev()), bias.to(dev()), tc=(impl == "tc"))
    out = torch.zeros(B, H, W, 384, device=dev())
    fl = _nhwc(flow).to(dev())
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, None, None, 128, 128), act=L.ACT_RELU, bias=W_.bias, tail=fl)
    ops.conv2d([s0, s1], W_, e, B, H, W, impl)
    got = out[..., 128:256].permute(0, 3, 1, 2).cpu()
    tol = 2e-5 if impl == "simt" else 2e-4
    assert stats(got, ref)[1] < tol, stats(got, ref)
    assert float(out[..., :128].abs().max()) == 0 and float(out[..., 256:].abs().max()) == 0


# ---------------------------------------------------------------------------------------------
# a6..a11: one full update-block step vs the reference's own output (golden)
# ---------------------------------------------------------------------------------------------
def _run_update(tag, igev, impl, terms=2, monkeypatch=None):
    """terms: MMAs per K step of the GRU / motion-encoder convs on the tensor-core path (3 = (hi, lo) activations,
    2 = one half value per activation; 1 = 2 plus single-plane weights for the two coarse GRUs, the default engine
    policy -- see UpdateEngine.gru2 / coarse1)."""
    from dkt_stereo_b200.update import BasicMultiUpdateBlock, UpdateEngine
    if monkeypatch is not None:
        monkeypatch.setenv("DKT_COARSE_GRU_TERMS", "1" if terms == 1 else "2")
        terms = max(terms, 2)
        monkeypatch.setenv("DKT_GRU_TERMS", str(terms))
        monkeypatch.setenv("DKT_MENC_TERMS", str(terms))
    from dkt_stereo_b200.synthetic import synthetic_state_dict
    from dkt_stereo_b200 import ops
    g = load_golden(f"update_{tag}")
    cfg = IGEV_CFG if igev else RAFT_CFG
    blk = BasicMultiUpdateBlock(Namespace(**cfg), hidden_dims=cfg["hidden_dims"], igev=igev)
    blk.load_state_dict(synthetic_state_dict(golden_shapes(g), seed=3), strict=True)
    blk = blk.to(dev())
    eng = UpdateEngine(blk, impl)
    assert impl != "tc" or monkeypatch is None or (eng.gru2, eng.menc2) == (terms == 2, terms == 2)
    coarse1 = eng.coarse1
    eng.pack_weights()
    B, _, h, w = g["net0"].shape
    eng.allocate(B, h, w, dev())
    eng.load_state([g[f"net{i}"].to(dev()) for i in range(3)],
                   [[g[f"c{n}{i}"].to(dev()) for n in "zrq"] for i in range(3)])
    corr, flow = g["corr"].to(dev()), g["flow"].to(dev())

    def lookup(e):
        e.CORR["f32"][..., :corr.shape[1]] = corr.permute(0, 2, 3, 1)
        if e.CORR["hi"] is not None:
            hi, lo = ops.split_bf16(e.CORR["f32"])
            e.CORR["hi"].copy_(hi)
            e.CORR["lo"].copy_(lo)
        e.FLOW["f32"].copy_(flow.permute(0, 2, 3, 1))

    eng.fused_enc = False          # this test injects the correlation features itself (no lookup kernel)
    eng.step(lookup, with_mask=True)
    torch.cuda.synchronize()
    net = eng.hidden_states()
    # one update step: fp32 kernels 3e-5; 3-MMA tensor-core path 3e-4; 2-MMA path 1.5e-3 max-abs on O(1) states (one
    # half-precision value per activation = 2^-12 relative per operand; the end-to-end gate is what bounds its use)
    tol = 3e-5 if impl == "simt" else (3e-4 if not eng.gru2 else (1.5e-3 if not coarse1 else 4e-3))
    for i in range(3):
        assert stats(net[i].cpu(), g[f"net_out{i}"])[1] < tol, (impl, i, stats(net[i].cpu(), g[f"net_out{i}"]))
    delta = eng.DELTA["f32"].permute(0, 3, 1, 2).cpu()
    # the tensor-core schedule only produces the consumed channel 0 (the reference zeroes delta_flow[:,1],
    # raft_stereo.py:164); the generic kernels (simt, or DKT_FAST_SMALL_CONVS=0) produce every channel
    nd = 1 if (impl == "tc" and eng.fast_small_convs) else delta.shape[1]
    assert stats(delta[:, :nd], g["delta"][:, :nd])[1] < tol * 3, stats(delta[:, :nd], g["delta"][:, :nd])
    mask = (eng.MH["f32"][..., :32] if igev else eng.MASK["f32"]).permute(0, 3, 1, 2).cpu()
    assert stats(mask, g["mask"])[1] < tol * 3, stats(mask, g["mask"])


@pytest.mark.parametrize("impl,terms", [("simt", 3), ("tc", 3), ("tc", 2), ("tc", 1)])
def test_update_block_raft(impl, terms, monkeypatch):
    _run_update("raft", False, impl, terms, monkeypatch)


def test_update_block_generic_small_convs(monkeypatch):
    """Same step with the 7x7 stem / head conv2 on the generic kernels (all delta channels checked)."""
    monkeypatch.setenv("DKT_FAST_SMALL_CONVS", "0")
    _run_update("raft", False, "tc", 3, monkeypatch)


@pytest.mark.parametrize("impl,terms", [("simt", 3), ("tc", 3), ("tc", 2), ("tc", 1)])
def test_update_block_igev(impl, terms, monkeypatch):
    _run_update("igev", True, impl, terms, monkeypatch)


@pytest.mark.parametrize("shape", [(1, 64, 64, 8, 16, 3), (2, 384, 256, 17, 30, 3), (1, 128, 126, 19, 37, 3), (3, 64, 32, 8, 16, 3)])
def test_conv_two_mma_mode(shape):
    """Sources WITHOUT a lo plane select the 2-MMA form x_hi * (w_hi + w_lo) and a destination without a lo plane gets
    its hi plane only: the result must equal the fp32 conv of the half-ROUNDED activations (weights at full precision)
    to 3-MMA accuracy, and the untouched lo plane must stay untouched."""
    from dkt_stereo_b200 import ops, _lib as L
    if L.split_dtype() != torch.float16:
        pytest.skip("bfloat16 build: hi-only activations are not used")
    B, Cin, N, H, W, k = shape
    g = torch.Generator().manual_seed(Cin * 7 + N)
    x = torch.randn(B, Cin, H, W, generator=g)
    wt = torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(N, generator=g)
    xh = x.half().float()
    ref = torch.relu(torch.nn.functional.conv2d(xh, wt, bias, padding=k // 2))
    hi = _nhwc(x).to(dev()).half().contiguous()
    W_ = ops.pack_conv(wt.to(dev()), bias.to(dev()), tc=True)
    Cout = (N + 7) // 8 * 8
    out = torch.zeros(B, H, W, Cout, device=dev())
    ohi = torch.zeros(B, H, W, Cout, device=dev(), dtype=torch.float16)
    olo = torch.full((B, H, W, Cout), 7.0, device=dev(), dtype=torch.float16)
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, ohi, None, 0, N), act=L.ACT_RELU, bias=W_.bias)
    ops.conv2d([L.tensor_slice(None, hi, None)], W_, e, B, H, W, "tc")
    torch.cuda.synchronize()
    got = out[..., :N].permute(0, 3, 1, 2).cpu()
    assert stats(got, ref)[1] < 2e-4, (shape, stats(got, ref))
    assert torch.equal(ohi[..., :N].cpu(), out[..., :N].half().cpu())         # hi = rn16(value)
    assert float((olo - 7.0).abs().max()) == 0.0
    # single-plane weights (w_lo = None): 1 MMA per K step = the conv of half-rounded activations AND weights
    import dataclasses
    W1 = dataclasses.replace(W_, w_lo=None)
    ops.conv2d([L.tensor_slice(None, hi, None)], W1, e, B, H, W, "tc")
    ref1 = torch.relu(torch.nn.functional.conv2d(xh, wt.half().float(), bias, padding=k // 2))
    assert stats(out[..., :N].permute(0, 3, 1, 2).cpu(), ref1)[1] < 2e-4
    assert stats(ref1, ref)[1] > 1e-4 or Cin * k * k < 1000          # ... which is a visibly different number
    # mixing sources with and without lo is refused
    lo = torch.zeros_like(hi)
    if Cin >= 128:
        with pytest.raises(L.DktError):
            ops.conv2d([L.tensor_slice(None, hi, lo, 0, 64), L.tensor_slice(None, hi, None, 64, Cin - 64)], W_, e, B, H, W, "tc")


# ---------------------------------------------------------------------------------------------
# K4
# ---------------------------------------------------------------------------------------------
def test_upsamplers_golden():
    from dkt_stereo_b200 import ops
    g = load_golden("convex_upsample")
    up = ops.convex_upsample(_nhwc(g["flow"]).to(dev()), _nhwc(g["mask"]).to(dev()), 4)
    assert stats(up.cpu(), g["out"][:, :1])[1] < 2e-5
    g = load_golden("context_upsample")
    up = ops.context_upsample(g["disp"][:, 0].contiguous().to(dev()), g["weights"].to(dev()), in_scale=1.0)
    assert stats(up[:, 0].cpu(), g["out"])[1] < 2e-5


# ---------------------------------------------------------------------------------------------
# a12: the whole forward(test_mode=True) vs the reference's disparity maps
# ---------------------------------------------------------------------------------------------
def _model(impl, g, **cfg_over):
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    from dkt_stereo_b200.synthetic import synthetic_state_dict
    cfg = dict(RAFT_CFG, corr_implementation="b200_fp32" if impl == "simt" else "b200", **cfg_over)
    model = RAFTStereo(Namespace(mixed_precision=False, **cfg)).eval()
    model.load_state_dict(synthetic_state_dict(golden_shapes(g), seed=golden_seeds(g)[0]), strict=True)
    if impl == "tc_torchenc":          # tensor-core hot path fed by the PyTorch (cuDNN fp32) encoders
        model.encoder = None
    else:
        assert (model.encoder is not None) == (impl == "tc")
    return model.to(dev())


@pytest.mark.parametrize("impl", ["simt", "tc", "tc_torchenc"])
@pytest.mark.parametrize("tag", ["raft_fwd_small", "raft_fwd_shift", "raft_fwd_cfg1", "raft_fwd_cfg2"])
def test_raft_forward_golden(impl, tag):
    """raft_fwd_cfg2 = the headline workload itself: 544 x 960, 32 iterations (one pair), the REAL reference's output."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden(tag)
    B, H, W, iters = [int(v) for v in g["meta"]]
    model = _model(impl, g)
    im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
    lr, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    assert up.shape == (B, 1, H, W) and lr.shape == (B, 2, H // 4, W // 4)
    mean, mx = stats(up.cpu(), g["flow_up"])
    print(f"[parity] {tag} impl={impl}: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
    assert mean <= 1e-3, (tag, impl, mean, mx)          # north-star gate
    assert stats(lr.cpu(), g["flow_lr"])[0] <= 1e-3
    # CUDA-graph replay (3rd call) must reproduce the eager result bit for bit
    model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    lr3, up3 = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    assert torch.equal(up3, up) and torch.equal(lr3, lr)


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("tag", ["raft_fwd_cfg2_s2", "raft_fwd_cfg4shape"])
def test_raft_forward_golden_large(impl, tag):
    """The REAL reference's ``flow_up`` at (a) a second (weights, images) sample of the headline workload and (b) the
    BASELINE configs[3] resolution 736 x 1280 (w/4 = 320: K1 runs in column blocks), 32 iterations, through forward()."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden(tag)
    B, H, W, iters = [int(v) for v in g["meta"]]
    model = _model(impl, g)
    im1, im2 = synthetic_pair(B, H, W, seed=golden_seeds(g)[1], mode=str(g["mode"]))
    lr, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    mean, mx = stats(up.cpu(), g["flow_up"])
    print(f"[parity] {tag} impl={impl}: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
    assert up.shape == (B, 1, H, W) and mean <= 1e-3, (tag, impl, mean, mx)          # north-star gate


def _igev_model_from_golden(g, impl, monkeypatch):
    """The drop-in IGEVStereo with the weights the golden was made with: every tensor is re-drawn from its REFERENCE-side
    name (torchvision MobileNetV2 naming of the generator's timm stub) and loaded under the timm name the drop-in (and
    real DKT checkpoints) use, strict=True."""
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    from dkt_stereo_b200.synthetic import synthetic_state_dict
    monkeypatch.setenv("DKT_IMPL", impl)
    model = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval()
    sd = synthetic_state_dict(golden_shapes(g), seed=golden_seeds(g)[0])
    model.load_state_dict({tv_to_timm(k): v for k, v in sd.items()}, strict=True)
    return model.to(dev())


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("tag", ["igev_fwd_cfg3", "igev_fwd_cfg3_shift", "igev_fwd_cfg5shape"])
def test_igev_forward_golden(tag, impl, monkeypatch):
    """Images in, ``disp_up`` out through the PUBLIC ``IGEVStereo.forward(test_mode=True)`` against the real reference's
    forward (igev_stereo.py:151-226) at BASELINE configs[2] (544 x 960, 32 iterations; noise pair and a pair with a true
    disparity ramp) and at the configs[4] resolution (1024 x 1536, 22 iterations)."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden(tag)
    B, H, W, iters = [int(v) for v in g["meta"]]
    model = _igev_model_from_golden(g, impl, monkeypatch)
    im1, im2 = synthetic_pair(B, H, W, seed=golden_seeds(g)[1], mode=str(g["mode"]))
    for rep in range(3):                                           # eager, graph capture, graph replay
        _, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
        mean, mx = stats(up.cpu(), g["disp_up"])
        print(f"[parity] {tag} impl={impl} rep={rep}: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
        assert up.shape == (B, 1, H, W) and mean <= 1e-3, (tag, impl, rep, mean, mx)   # north-star gate


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_raft_realtime_settings(impl):
    """The upstream "real-time" RAFT-Stereo settings the reference's switches allow (raft_stereo.py:37-49,96-100): shared
    backbone, n_downsample = 3 (1/8 resolution, 8x convex upsampling with a 576-channel mask), two GRU levels, slow-fast
    schedule.  The encoders of these settings run on the PyTorch extractor path; volume, loop and upsampling on the kernels."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_realtime")
    B, H, W, iters = [int(v) for v in g["meta"]]
    over = dict(shared_backbone=True, n_downsample=3, n_gru_layers=2, slow_fast_gru=True)
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    from dkt_stereo_b200.synthetic import synthetic_state_dict
    cfg = dict(RAFT_CFG, corr_implementation="b200_fp32" if impl == "simt" else "b200", **over)
    model = RAFTStereo(Namespace(mixed_precision=False, **cfg)).eval()
    model.load_state_dict(synthetic_state_dict(golden_shapes(g), seed=golden_seeds(g)[0]), strict=True)
    model = model.to(dev())
    im1, im2 = synthetic_pair(B, H, W, seed=golden_seeds(g)[1], mode=str(g["mode"]))
    for _ in range(3):                                           # eager, graph capture, graph replay
        lr, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    assert up.shape == (B, 1, H, W) and lr.shape == (B, 2, H // 8, W // 8)
    assert stats(up.cpu(), g["flow_up"])[0] <= 1e-3, stats(up.cpu(), g["flow_up"])


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_forward_all_predictions(impl, monkeypatch):
    """forward(test_mode=False) -- the reference's DEFAULT call (raft_stereo.py:85,185-187; igev_stereo.py:151,222-226):
    {'disp_preds': [prediction after every iteration]} (+ 'init_disp' for IGEV) against the real reference run under
    no_grad; with trainable parameters and autograd on, the engine refuses instead of returning graph-less tensors."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_all_preds")
    B, H, W, iters = [int(v) for v in g["meta"]]
    im1, im2 = synthetic_pair(B, H, W, seed=golden_seeds(g)[1], mode=str(g["mode"]))
    model = _model(impl, g)
    with pytest.raises(NotImplementedError):
        model(im1.to(dev()), im2.to(dev()), iters=iters)
    with torch.no_grad():
        res = model(im1.to(dev()), im2.to(dev()), iters=iters)
    assert len(res["disp_preds"]) == iters
    for i, p in enumerate(res["disp_preds"]):
        assert p.shape == (B, 1, H, W) and stats(p.cpu(), g["disp_preds"][i])[0] <= 1e-3, (i, stats(p.cpu(), g["disp_preds"][i]))
    _, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)            # and test mode still agrees
    assert stats(up.cpu(), g["disp_preds"][-1])[0] <= 1e-3
    g = load_golden("igev_all_preds")
    model = _igev_model_from_golden(g, impl, monkeypatch)
    for p_ in model.parameters():
        p_.requires_grad = False                                   # a frozen teacher may call it with autograd on
    res = model(im1.to(dev()), im2.to(dev()), iters=iters)
    assert stats(res["init_disp"].cpu(), g["init_disp"])[0] <= 1e-3, stats(res["init_disp"].cpu(), g["init_disp"])
    assert len(res["disp_preds"]) == iters
    for i, p in enumerate(res["disp_preds"]):
        assert p.shape == (B, 1, H, W) and stats(p.cpu(), g["disp_preds"][i])[0] <= 1e-3, (i, stats(p.cpu(), g["disp_preds"][i]))


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_slow_fast_gru(impl, monkeypatch):
    """args.slow_fast_gru=True (reference raft_stereo.py:157-160, igev_stereo.py:201-204): per iteration one extra update
    of the coarsest GRU and one of the two coarse GRUs; against the real reference's outputs with that flag."""
    from dkt_stereo_b200.synthetic import synthetic_pair, synthetic_state_dict
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    g = load_golden("raft_fwd_slowfast")
    B, H, W, iters = [int(v) for v in g["meta"]]
    model = _model(impl, g, slow_fast_gru=True)
    im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
    _, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    assert stats(up.cpu(), g["flow_up"])[0] <= 1e-3, stats(up.cpu(), g["flow_up"])
    _, up0 = _model(impl, g)(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
    assert stats(up0.cpu(), g["flow_up"])[0] > 1e-2                # without the flag the answer is a different one
    monkeypatch.setenv("DKT_IMPL", impl)
    g = load_golden("igev_fwd_slowfast")
    B, H, W, iters = [int(v) for v in g["meta"]]
    m = IGEVStereo(Namespace(mixed_precision=False, **dict(IGEV_CFG, slow_fast_gru=True))).eval()
    m.load_state_dict(synthetic_state_dict(golden_shapes(g), seed=0), strict=False)
    m = m.to(dev())
    d = lambda k: g[k].to(dev())
    with torch.no_grad():
        up = m.hot_path(d("match_left"), d("match_right"), d("gev"), d("init_disp"),
                        [d(f"net{i}") for i in range(3)], [d(f"ctx{i}") for i in range(3)], d("stem_2x"), iters)
    assert stats(up.cpu(), g["disp_up"])[0] <= 1e-3, stats(up.cpu(), g["disp_up"])


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_n_gru_layers_one_and_two(impl, monkeypatch):
    """args.n_gru_layers = 1 / 2 through the public forward() against the REAL reference's output (oracle/make_golden.py
    --only raft_gru1,raft_gru2,igev_gru2); n = 2 with slow_fast_gru (its extra update runs the mid level alone)."""
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    from dkt_stereo_b200.synthetic import synthetic_pair, synthetic_state_dict
    for tag, over in (("raft_fwd_gru1", dict(n_gru_layers=1)), ("raft_fwd_gru2", dict(n_gru_layers=2, slow_fast_gru=True))):
        g = load_golden(tag)
        B, H, W, iters = [int(v) for v in g["meta"]]
        model = _model(impl, g, **over)
        im1, im2 = synthetic_pair(B, H, W, seed=1234, mode=str(g["mode"]))
        for _ in range(3):                                       # eager, graph capture, graph replay
            _, up = model(im1.to(dev()), im2.to(dev()), iters=iters, test_mode=True)
        assert stats(up.cpu(), g["flow_up"])[0] <= 1e-3, (tag, stats(up.cpu(), g["flow_up"]))
    monkeypatch.setenv("DKT_IMPL", impl)
    g = load_golden("igev_fwd_gru2")
    B, H, W, iters = [int(v) for v in g["meta"]]
    m = IGEVStereo(Namespace(mixed_precision=False, **dict(IGEV_CFG, n_gru_layers=2))).eval()
    m.load_state_dict(synthetic_state_dict(golden_shapes(g), seed=0), strict=False)
    m = m.to(dev())
    d = lambda k: g[k].to(dev())
    with torch.no_grad():
        up = m.hot_path(d("match_left"), d("match_right"), d("gev"), d("init_disp"),
                        [d(f"net{i}") for i in range(2)], [d(f"ctx{i}") for i in range(2)], d("stem_2x"), iters)
    assert stats(up.cpu(), g["disp_up"])[0] <= 1e-3, stats(up.cpu(), g["disp_up"])


def test_serves_frozen_and_ema_teacher():
    """The two teacher passes of the reference's fine-tuning step (tools/ft_dkt.py:179-199): DataParallel-wrapped,
    frozen, `model_T(image1, image2, iters, test_mode=True)`; the EMA teacher's parameters are REASSIGNED every step.
    The engine must follow the weights (repack + new CUDA graph) and give exactly what a freshly loaded model gives."""
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    from dkt_stereo_b200.synthetic import synthetic_pair, synthetic_state_dict
    g = load_golden("raft_fwd_small")
    shapes = golden_shapes(g)
    ns = Namespace(mixed_precision=False, **dict(RAFT_CFG, corr_implementation="b200"))
    teacher = torch.nn.DataParallel(RAFTStereo(ns), device_ids=[0])
    teacher.load_state_dict({"module." + k: v for k, v in synthetic_state_dict(shapes, seed=0).items()}, strict=True)
    teacher.cuda()
    for p in teacher.parameters():
        p.requires_grad = False
    teacher.eval()
    teacher.module.freeze_bn()
    student_sd = synthetic_state_dict(shapes, seed=5)
    im1, im2 = synthetic_pair(2, 64, 96, seed=3)
    im1, im2 = im1.to(dev()), im2.to(dev())
    outs = [teacher(im1, im2, iters=3, test_mode=True)[1].clone() for _ in range(3)]     # eager, capture, replay
    assert torch.equal(outs[0], outs[2])
    ema = 0.5
    for (name, t_params) in teacher.module.named_parameters():
        t_params.data = (ema * t_params.data + (1 - ema) * student_sd[name].to(dev()))
        t_params.requires_grad = False
    after = [teacher(im1, im2, iters=3, test_mode=True)[1].clone() for _ in range(3)]
    assert float((after[0] - outs[0]).abs().mean()) > 1e-3                 # the new weights are in use ...
    assert torch.equal(after[0], after[2])                                 # ... also in the re-captured graph
    fresh = RAFTStereo(ns).eval()
    fresh.load_state_dict({k: v.detach().cpu() for k, v in teacher.module.state_dict().items()}, strict=True)
    _, want = fresh.to(dev())(im1, im2, iters=3, test_mode=True)
    assert torch.equal(after[0], want)


def test_igev_upsample_disp_native_vs_modules():
    """IGEV upsample_disp (reference igev_stereo.py:140-148) on the library's kernels -- both transposed convs as
    parity-grouped 3x3 tensor-core convs, folded BatchNorm + LeakyReLU, softmax + context_upsample fused -- against the
    PyTorch modules with the same (non-trivial) parameters, and the small kernels against their definitions."""
    from dkt_stereo_b200 import ops, _lib as L
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    g = torch.Generator().manual_seed(21)
    # (1) deconv-as-conv weights: F.conv2d + pixel shuffle == F.conv_transpose2d
    wt = torch.randn(6, 5, 4, 4, generator=g)
    x = torch.randn(2, 6, 7, 9, generator=g)
    ref = torch.nn.functional.conv_transpose2d(x, wt, None, stride=2, padding=1)
    y = torch.nn.functional.conv2d(x, ops.deconv4x4s2_as_conv3x3(wt), None, padding=1)           # (2, 4*5, 7, 9)
    y = y.view(2, 2, 2, 5, 7, 9).permute(0, 3, 4, 1, 5, 2).reshape(2, 5, 14, 18)
    assert stats(y, ref)[1] < 1e-5
    # (2) pixel shuffle kernel
    B, H, W, Cc = 2, 5, 7, 8
    src = torch.randn(B, H, W, 4 * Cc + 4, generator=g).to(dev())
    dst = torch.zeros(B, 2 * H, 2 * W, 16, device=dev())
    ops.pixel_shuffle2(src, Cc, L.tensor_slice(dst, None, None, 8, Cc), B, H, W)
    want = src[..., :4 * Cc].view(B, H, W, 2, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, Cc)
    assert torch.equal(dst[..., 8:], want) and float(dst[..., :8].abs().max()) == 0
    # (3) the whole tail through the model
    monkey_cfg = dict(IGEV_CFG)
    model = IGEVStereo(Namespace(mixed_precision=False, **monkey_cfg)).eval()
    with torch.no_grad():
        for name, prm in list(model.spx_2_gru.named_parameters()) + list(model.spx_gru.named_parameters()):
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.15 if prm.dim() > 1 else 0.3) + (1.0 if name.endswith("bn.weight") else 0.0))
        for name, buf in model.spx_2_gru.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.2)
            elif name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
    model = model.to(dev())
    assert model.native_upsample
    B, h, w = 2, 12, 20
    eng = model.engine
    eng.pack_weights()
    eng.allocate(B, h, w, dev())
    mf = torch.relu(torch.randn(B, 32, h, w, generator=g)).to(dev())
    stem = torch.randn(B, 32, 2 * h, 2 * w, generator=g).to(dev())
    disp = (torch.rand(B, h, w, generator=g) * 40).to(dev())
    nh = mf.permute(0, 2, 3, 1).contiguous()
    eng.MH["f32"][..., :32].copy_(nh)
    hi, lo = ops.split16(nh)
    eng.MH["hi"][..., :32].copy_(hi)
    eng.MH["lo"][..., :32].copy_(lo)
    with torch.no_grad():
        got = model.upsample_disp(disp, mf, stem)
        model.native_upsample = False
        want = model.upsample_disp(disp, mf, stem)
    assert got.shape == want.shape == (B, 1, 4 * h, 4 * w)
    assert stats(got.cpu(), want.cpu())[1] < 2e-3, stats(got.cpu(), want.cpu())        # disparities up to 160 px (4 x 40)
    assert stats(got.cpu(), want.cpu())[0] < 2e-4


def test_flow_init_and_batch_independence():
    """flow_init is honoured and per-sample results do not depend on the batch they ride in."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    model = _model("tc", g)
    im1, im2 = synthetic_pair(3, 64, 96, seed=77)
    im1, im2 = im1.to(dev()), im2.to(dev())
    _, up = model(im1, im2, iters=3, test_mode=True)
    _, up1 = model(im1[1:2], im2[1:2], iters=3, test_mode=True)
    assert stats(up[1:2].cpu(), up1.cpu())[0] < 1e-4      # cuDNN may pick another algorithm per batch size
    fi = torch.zeros(3, 2, 16, 24, device=dev())
    fi[:, 0] = -2.5
    lr, _ = model(im1, im2, iters=1, flow_init=fi, test_mode=True)
    lr0, _ = model(im1, im2, iters=1, test_mode=True)
    assert float((lr - lr0).abs().mean()) > 0.5


# ---------------------------------------------------------------------------------------------
# IGEV-Stereo hot path end to end: pre-loop products + final disparity of the REAL reference forward
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("tag", ["igev_fwd_small", "igev_fwd_mid"])
def test_igev_hot_path_golden(tag, impl, monkeypatch):
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    from dkt_stereo_b200.synthetic import synthetic_state_dict
    monkeypatch.setenv("DKT_IMPL", impl)
    g = load_golden(tag)
    B, H, W, iters = [int(v) for v in g["meta"]]
    model = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval()
    sd = synthetic_state_dict(golden_shapes(g), seed=0)           # update_block.*, spx_2_gru.*, spx_gru.*
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    assert not [k for k in missing if k.startswith(("update_block.", "spx_2_gru.", "spx_gru."))]
    model = model.to(dev())
    d = lambda k: g[k].to(dev())
    for rep in range(3):                                           # eager, graph capture, graph replay
        with torch.no_grad():
            up = model.hot_path(d("match_left"), d("match_right"), d("gev"), d("init_disp"),
                                [d(f"net{i}") for i in range(3)], [d(f"ctx{i}") for i in range(3)], d("stem_2x"), iters)
        torch.cuda.synchronize()
        mean, mx = stats(up.cpu(), g["disp_up"])
        print(f"[parity] {tag} impl={impl} rep={rep}: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
        assert up.shape == (B, 1, H, W)
        assert mean <= 1e-3, (tag, impl, rep, mean, mx)       # the north-star gate


# ---------------------------------------------------------------------------------------------
# encoder side: strided / 7x1 convs, residual epilogue, instance norm, whole encoders
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 64, 96, 19, 37, 3, 2), (1, 128, 128, 16, 32, 1, 2), (2, 64, 64, 18, 33, 3, 1),
                                   (1, 128, 256, 9, 21, 1, 1)])
def test_conv_ex_strided_and_residual(shape):
    """dkt_conv2d_tc_ex: 3x3 / 1x1, stride 1 / 2 (TMA element strides), bias + ReLU + residual tail."""
    from dkt_stereo_b200 import ops, _lib as L
    B, Cin, N, H, W, k, stride = shape
    g = torch.Generator().manual_seed(Cin * 7 + N + stride)
    x = torch.randn(B, Cin, H, W, generator=g)
    wt = torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(N, generator=g)
    y = torch.relu(torch.nn.functional.conv2d(x, wt, bias, stride=stride, padding=k // 2))
    Ho, Wo = y.shape[-2:]
    res = torch.randn(B, N, Ho, Wo, generator=g)
    ref = torch.relu(y + res)
    xs, keep = _slice_of(_nhwc(x).to(dev()), "tc")
    npad = (N + 63) // 64 * 64
    Wc = ops.pack_conv_general(wt.to(dev()), bias.to(dev()), stride=stride, n_pad=npad)
    out = torch.zeros(B, Ho, Wo, npad, device=dev())
    hi = torch.zeros(B, Ho, Wo, npad, device=dev(), dtype=L16())
    lo = torch.zeros_like(hi)
    resd = torch.zeros(B, Ho, Wo, npad, device=dev())
    resd[..., :N] = _nhwc(res).to(dev())
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, hi, lo, 0, npad), act=L.ACT_RELU, bias=Wc.bias, res=(resd, 0))
    assert ops.conv2d_ex([xs], Wc, e, B, H, W) == (Ho, Wo)
    torch.cuda.synchronize()
    got = out[..., :N].permute(0, 3, 1, 2).cpu()
    assert stats(got, ref)[1] < 3e-4, stats(got, ref)
    assert torch.all(out[..., N:] == 0)                          # padded output channels stay exactly zero
    assert stats((hi.float() + lo.float()).cpu(), out.cpu())[1] < 3e-4


def test_stem_rows_7x7():
    """Image normalisation + x-im2col + 7x1 tensor-core conv == conv2d(2*(img/255)-1, 7x7, pad 3)
    (reference raft_stereo.py:91-92, core/extractor.py:140)."""
    from dkt_stereo_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(21)
    B, H, W = 2, 21, 70
    img = torch.rand(B, 3, H, W, generator=g) * 255
    wt = torch.randn(64, 3, 7, 7, generator=g) / 12.0
    bias = torch.randn(64, generator=g)
    ref = torch.nn.functional.conv2d(2 * (img / 255.0) - 1.0, wt, bias, padding=3)
    hi = torch.zeros(B, H, W, 64, device=dev(), dtype=L16())
    lo = torch.zeros_like(hi)
    ops.stem_rows(img.to(dev()), hi, lo)
    w7 = wt.permute(0, 3, 1, 2).reshape(64, 21, 7, 1)
    Wc = ops.pack_conv_general(w7.to(dev()), bias.to(dev()), cin_pad=64)
    out = torch.zeros(B, H, W, 64, device=dev())
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, None, None, 0, 64), bias=Wc.bias)
    ops.conv2d_ex([L.tensor_slice(None, hi, lo, 0, 64)], Wc, e, B, H, W)
    torch.cuda.synchronize()
    assert stats(out.permute(0, 3, 1, 2).cpu(), ref)[1] < 2e-4


def test_instnorm():
    from dkt_stereo_b200 import ops, _lib as L
    g = torch.Generator().manual_seed(8)
    B, Cc, H, W = 3, 128, 37, 29
    x = torch.randn(B, Cc, H, W, generator=g) * 3 + 1.5
    x[:, 100:] = 0                                             # zero-padded channels must stay zero
    res = torch.randn(B, Cc, H, W, generator=g)
    y = torch.relu(torch.nn.functional.instance_norm(x))
    ref = torch.relu(res + y)
    xd, rd = _nhwc(x).to(dev()), _nhwc(res).to(dev())
    st = torch.zeros(B, Cc, 2, device=dev())
    ws = ops.instnorm_workspace(B, Cc, dev())
    out = torch.zeros(B, H, W, Cc, device=dev())
    hi = torch.zeros(B, H, W, Cc, device=dev(), dtype=L16())
    lo = torch.zeros_like(hi)
    ops.instnorm_stats(L.tensor_slice(xd, None, None, 0, Cc), ws, st, B, H, W)
    ops.instnorm_apply(L.tensor_slice(xd, None, None, 0, Cc), st, L.tensor_slice(out, hi, lo, 0, Cc), B, H, W,
                       relu=True, res=L.tensor_slice(rd, None, None, 0, Cc))
    torch.cuda.synchronize()
    assert stats(out.permute(0, 3, 1, 2).cpu(), ref)[1] < 2e-5
    assert stats(st[:, :100, 0].cpu(), x[:, :100].mean(dim=(2, 3)))[1] < 1e-5
    plain = torch.zeros(B, H, W, Cc, device=dev())
    ops.instnorm_apply(L.tensor_slice(xd, None, None, 0, Cc), st, L.tensor_slice(plain, None, None, 0, Cc), B, H, W, relu=False)
    assert torch.all(plain[..., 100:] == 0)
    assert stats(plain.permute(0, 3, 1, 2).cpu(), torch.nn.functional.instance_norm(x))[1] < 2e-5


def test_encoder_engine_vs_pytorch_extractor():
    """fnet / cnet / context convs on libdkt kernels vs the same modules on cuDNN fp32."""
    from dkt_stereo_b200.raft_stereo import RAFTStereo
    from dkt_stereo_b200.synthetic import synthetic_state_dict, synthetic_pair, shapes_of
    model = RAFTStereo(Namespace(mixed_precision=False, **dict(RAFT_CFG, corr_implementation="b200"))).eval()
    model.load_state_dict(synthetic_state_dict(shapes_of(model.state_dict()), seed=0), strict=True)
    model = model.to(dev())
    assert model.encoder is not None
    im1, im2 = synthetic_pair(2, 96, 160, seed=77)
    im1, im2 = im1.to(dev()), im2.to(dev())
    with torch.no_grad():
        fmap1, fmap2, net_list, ctx_list = model.extract(im1, im2)
        model.encoder.run(im1, im2)
    torch.cuda.synchronize()
    enc, eng = model.encoder, model.engine
    f = (enc.FMAP.hi.float() + enc.FMAP.lo.float()).permute(0, 3, 1, 2)
    ref = torch.cat([fmap1, fmap2], 0)
    m, mx = stats(f.cpu(), ref.cpu())
    print(f"[encoder] fmap: mean-abs {m:.3e} max-abs {mx:.3e} (|ref| mean {float(ref.abs().mean()):.3f})")
    assert mx < 2e-3 and m < 1e-4
    for i in range(3):
        h = eng.X[i]["f32"][..., :128].permute(0, 3, 1, 2)
        assert stats(h.cpu(), net_list[i].cpu())[1] < 1e-3, (i, stats(h.cpu(), net_list[i].cpu()))
        c = eng.CTX[i]["f32"].permute(0, 3, 1, 2)
        cref = ctx_list[i] + eng.gru_bias[i].view(1, -1, 1, 1)
        assert stats(c.cpu(), cref.cpu())[1] < 2e-3, (i, stats(c.cpu(), cref.cpu()))


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 128, 256, 19, 37), (2, 64, 96, 8, 16), (1, 128, 250, 24, 48)])
def test_conv_proj_epilogue(shape):
    """DKT_EPI_PROJ: conv1 + ReLU + the channel half of a one-output 3x3 conv2 in one kernel, then tapsum3x3
    == F.conv2d(relu(F.conv2d(x, w1, b1)), w2, b2)[:, 0] (FlowHead, reference core/update.py:13-14)."""
    from dkt_stereo_b200 import ops, _lib as L
    B, Cin, N, H, W = shape
    g = torch.Generator().manual_seed(N + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    w1 = torch.randn(N, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b1 = torch.randn(N, generator=g)
    w2 = torch.randn(2, N, 3, 3, generator=g) / (N * 9) ** 0.5
    b2 = torch.randn(2, generator=g)
    ref = torch.nn.functional.conv2d(torch.relu(torch.nn.functional.conv2d(x, w1, b1, padding=1)), w2, b2, padding=1)[:, 0]
    xs, keep = _slice_of(_nhwc(x).to(dev()), "tc")
    W1 = ops.pack_conv(w1.to(dev()), b1.to(dev()), tc=True)
    proj = ops.pack_proj3x3(w2.to(dev()), 0)
    taps = torch.full((B, H, W, 16), float("nan"), device=dev())
    e = ops.make_epilogue(L.EPI_PROJ, L.tensor_slice(taps, None, None, 0, 16), act=L.ACT_RELU, bias=W1.bias, proj=proj)
    ops.conv2d([xs], W1, e, B, H, W, "tc")
    out = torch.zeros(B, H, W, 2, device=dev())
    ops.tapsum3x3(taps, float(b2[0]), out)
    torch.cuda.synchronize()
    assert not torch.isnan(taps[..., :12]).any()
    assert stats(out[..., 0].cpu(), ref)[1] < 3e-4, stats(out[..., 0].cpu(), ref)


def test_conv_pair_matches_single_cta():
    """The CTA-pair (cta_group::2) kernel with its x-major halo patch and the single-CTA row-patch kernel compute the
    same products; only the order in which the taps enter the fp32 accumulator differs, so the results agree to
    fp32 round-off (tolerance 2e-5 on O(1) outputs).  DKT_CONV_PAIR is read once per process, so this test drives
    the choice through a subprocess."""
    import subprocess, sys, os
    code = r"""
import torch, sys
sys.path.insert(0, %r)
from dkt_stereo_b200 import ops, _lib as L
torch.manual_seed(3)
dev = torch.device('cuda:0')
outs = []
for (B, Cin, N, H, W) in [(2, 384, 256, 17, 30), (1, 64, 64, 24, 16), (1, 128, 126, 40, 56)]:
    x = torch.randn(B, H, W, Cin, device=dev)
    hi, lo = ops.split_bf16(x)
    wt = torch.randn(N, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5
    bias = torch.randn(N, device=dev)
    Wp = ops.pack_conv(wt, bias, tc=True)
    out = torch.zeros(B, H, W, (N + 3) // 4 * 4, device=dev)
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, None, None, 0, N), act=L.ACT_RELU, bias=Wp.bias)
    ops.conv2d([L.tensor_slice(None, hi.contiguous(), lo.contiguous())], Wp, e, B, H, W, 'tc')
    torch.cuda.synchronize()
    outs.append(out.cpu())
torch.save(outs, sys.argv[1])
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    res = {}
    for mode in ("0", "1"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            env = dict(os.environ, DKT_CONV_PAIR=mode)
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300)
            res[mode] = torch.load(f.name)
    for a, b in zip(res["0"], res["1"]):
        assert float((a - b).abs().max()) <= 2e-5, float((a - b).abs().max())


@pytest.mark.parametrize("shape", [(2, 64, 64, 19, 37, 1), (3, 64, 128, 40, 50, 2), (1, 128, 128, 8, 16, 1)])
def test_conv_fused_instnorm_stats(shape):
    """dkt_epilogue.stats_partial + dkt_instnorm_finalize_tiles == mean / rstd of the conv output
    (nn.InstanceNorm2d statistics, reference core/extractor.py:16-33), ragged tiles and strided convs included."""
    from dkt_stereo_b200 import ops, _lib as L
    B, Cin, N, H, W, stride = shape
    g = torch.Generator().manual_seed(N + H)
    x = torch.randn(B, Cin, H, W, generator=g) * 2 + 0.5
    wt = torch.randn(N, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    bias = torch.randn(N, generator=g)
    ref = torch.nn.functional.conv2d(x, wt, bias, padding=1, stride=stride)
    Ho, Wo = ref.shape[2:]
    xs, keep = _slice_of(_nhwc(x).to(dev()), "tc")
    Wp = ops.pack_conv_general(wt.to(dev()), bias.to(dev()), stride=stride)
    out = torch.zeros(B, Ho, Wo, N, device=dev())
    part = torch.full((B * ops.conv_tiles(Ho, Wo) * 2 * N,), float("nan"), device=dev())
    e = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(out, None, None, 0, N), bias=Wp.bias, stats_partial=part)
    ops.conv2d_ex([xs], Wp, e, B, H, W)
    stats = torch.zeros(B, N, 2, device=dev())
    ops.instnorm_finalize_tiles(part, ops.instnorm_tiles_workspace(B, N, dev()), stats, B, N, Ho, Wo)
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).double().cpu()
    mean = got.mean(dim=(2, 3))
    rstd = 1.0 / torch.sqrt(got.var(dim=(2, 3), unbiased=False) + 1e-5)
    assert not torch.isnan(part).any()
    assert float((stats[..., 0].cpu().double() - mean).abs().max()) < 1e-5
    assert float(((stats[..., 1].cpu().double() - rstd) / rstd).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------
# BASELINE sizes (544 x 960, configs[1]): size-independent properties, no oracle needed
# ---------------------------------------------------------------------------------------------
def test_full_size_properties():
    """At the benchmark resolution: (1) the volume pyramid is self-consistent (level l+1 = pairwise mean of level l,
    level 0 = scaled dot products of the feature maps); (2) looking the volume up at integer coordinates returns
    the volume's own entries; (3) a sample's disparity map does not depend on the batch it rides in (bit exact:
    every kernel reduces in a fixed, batch-independent order) and a repeated forward is bit-identical."""
    from dkt_stereo_b200 import ops
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    model = _model("tc", g)
    H, W = 544, 960
    im1, im2 = synthetic_pair(2, H, W, seed=5)
    im1, im2 = im1.to(dev()), im2.to(dev())
    lr, up = model(im1, im2, iters=4, test_mode=True)
    assert up.shape == (2, 1, H, W) and torch.isfinite(up).all()
    pyr = [p.clone() for p in model._pyr]
    # (1) pyramid consistency
    for l in range(3):
        w2 = pyr[l + 1].shape[-1]
        ref = 0.5 * (pyr[l][..., 0:2 * w2:2] + pyr[l][..., 1:2 * w2:2])
        assert float((pyr[l + 1] - ref).abs().max()) < 1e-5
    f = model.encoder.FMAP
    f1 = (f.hi[:2].float() + f.lo[:2].float())[0, 7]               # image 0, row 7: (W1, D)
    f2 = (f.hi[2:].float() + f.lo[2:].float())[0, 7]
    ref = (f1.double() @ f2.double().T / 16.0).float()
    assert float((pyr[0][0, 7] - ref).abs().max()) < 2e-4
    # (2) identity lookup: coords = pixel x -> tap k of level 0 is volume[.., x, x + k - 4]
    B, h, w = 2, H // 4, W // 4
    cx = torch.arange(w, device=dev(), dtype=torch.float32).view(1, 1, w).expand(B, h, w).contiguous()
    out = torch.zeros(B, h, w, 36, device=dev())
    ops.corr1d_lookup(pyr, cx, 4, out, "nhwc")
    xs = torch.arange(w, device=dev())
    for k in (0, 4, 8):
        j = xs + k - 4
        ok = (j >= 0) & (j < w)
        ref = pyr[0][:, :, xs[ok], j[ok]]
        assert torch.equal(out[:, :, ok, k], ref)
    # (3) batch independence + determinism, bit exact
    lr_b, up_b = model(im1, im2, iters=4, test_mode=True)
    assert torch.equal(up_b, up) and torch.equal(lr_b, lr)
    _, up0 = model(im1[:1], im2[:1], iters=4, test_mode=True)
    assert torch.equal(up0, up[:1]), float((up0 - up[:1]).abs().max())
    _, up1 = model(im1[1:], im2[1:], iters=4, test_mode=True)
    assert torch.equal(up1, up[1:]), float((up1 - up[1:]).abs().max())


def test_host_pipeline_matches_direct_calls():
    """HostPipeline (pinned upload overlapped with compute, pinned read-back) returns exactly what direct forward
    calls return, batch after batch, including across the eager -> capture -> replay transitions of the CUDA graph."""
    from dkt_stereo_b200.pipeline import HostPipeline
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    model = _model("tc", g)
    batches = [tuple(t.pin_memory() for t in synthetic_pair(2, 64, 96, seed=100 + i)) for i in range(5)]
    want = []
    for a, b in batches:
        _, up = model(a.to(dev()), b.to(dev()), iters=3, test_mode=True)
        want.append(up.cpu())
    pipe = HostPipeline(model, iters=3)
    pipe.prefetch(*batches[0])
    for i in range(5):
        got = pipe.step(batches[i + 1] if i + 1 < 5 else None).clone()
        assert torch.equal(got, want[i]), (i, float((got - want[i]).abs().max()))


def test_host_pipeline_async_matches_direct_calls():
    """HostPipeline.step_async (the throughput form bench.py's e2e runs: forward + read-back enqueued, the previous call's
    result returned) delivers exactly the direct forward's maps, one call late, and drain() the last one; seven batches so
    that every input slot, staging buffer and pinned result is reused while its predecessor's copies may still be in flight."""
    from dkt_stereo_b200.pipeline import HostPipeline
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    model = _model("tc", g)
    n = 7
    batches = [tuple(t.pin_memory() for t in synthetic_pair(2, 64, 96, seed=300 + i)) for i in range(n)]
    want = []
    for a, b in batches:
        _, up = model(a.to(dev()), b.to(dev()), iters=3, test_mode=True)
        want.append(up.cpu())
    pipe = HostPipeline(model, iters=3)
    pipe.prefetch(*batches[0])
    got = []
    for i in range(n):
        out = pipe.step_async(batches[i + 1] if i + 1 < n else None)
        assert (out is None) == (i == 0)
        if out is not None:
            got.append(out.clone())
    got.append(pipe.drain().clone())
    assert pipe.drain() is None
    assert len(got) == n
    for i in range(n):
        assert torch.equal(got[i], want[i]), (i, float((got[i] - want[i]).abs().max()))
    # the blocking form still works on the same pipeline afterwards
    pipe.prefetch(*batches[2])
    assert torch.equal(pipe.step(None), want[2])


def test_host_pipeline_async_across_shape_changes():
    """The evaluator's use of step_async: consecutive batches of DIFFERENT shapes (ragged datasets).  Input slots, staging
    buffers and pinned results are re-created per shape while the previous batch's read-back is still owed to the caller."""
    from dkt_stereo_b200.pipeline import HostPipeline
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    model = _model("tc", g)
    shapes = [(2, 64, 96), (1, 96, 128), (1, 96, 128), (2, 64, 96), (3, 64, 64)]
    batches = [tuple(t.pin_memory() for t in synthetic_pair(b, h, w, seed=400 + i)) for i, (b, h, w) in enumerate(shapes)]
    want = []
    for a, b in batches:
        _, up = model(a.to(dev()), b.to(dev()), iters=3, test_mode=True)
        want.append(up.cpu())
    pipe = HostPipeline(model, iters=3)
    pipe.prefetch(*batches[0])
    got = []
    for i in range(len(batches)):
        out = pipe.step_async(batches[i + 1] if i + 1 < len(batches) else None)
        if out is not None:
            got.append(out.clone())
    got.append(pipe.drain().clone())
    for i, (gt, wt) in enumerate(zip(got, want)):
        assert gt.shape == wt.shape and torch.equal(gt, wt), (i, tuple(gt.shape), tuple(wt.shape))


def test_igev_context_encoder_on_engine(monkeypatch):
    """IGEV-Stereo's cnet + context convs on the tensor-core EncoderEngine (fnet-less mode) vs the same modules in
    PyTorch fp32: same weights, same images -> same disparity within the end-to-end gate."""
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    from dkt_stereo_b200.synthetic import synthetic_pair
    torch.manual_seed(11)
    monkeypatch.setenv("DKT_NATIVE_ENCODER", "1")
    a = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval().to(dev())
    assert a.encoder is not None
    monkeypatch.setenv("DKT_NATIVE_ENCODER", "0")
    b = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval().to(dev())
    assert b.encoder is None
    b.load_state_dict(a.state_dict(), strict=True)
    im1, im2 = synthetic_pair(2, 128, 160, seed=21, mode="shift")
    im1, im2 = im1.to(dev()), im2.to(dev())
    for rep in range(3):                       # eager, graph capture, graph replay
        _, ua = a(im1, im2, iters=6, test_mode=True)
        _, ub = b(im1, im2, iters=6, test_mode=True)
        mean, mx = stats(ua.cpu(), ub.cpu())
        print(f"[parity] igev cnet engine vs torch rep={rep}: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
        assert mean <= 1e-3, (rep, mean, mx)


def test_shape_changes_invalidate_graphs():
    """A -> B -> A input shapes: the engines re-allocate their buffers per shape, so CUDA graphs captured for an
    earlier shape must not be replayed; every call must equal a fresh model's eager result."""
    from dkt_stereo_b200.synthetic import synthetic_pair
    g = load_golden("raft_fwd_small")
    model, fresh = _model("tc", g), _model("tc", g)
    fresh.use_cuda_graph = False
    shapes = [(1, 64, 96), (2, 96, 128), (1, 64, 96), (1, 64, 96), (1, 64, 96), (2, 96, 128), (2, 96, 128), (2, 96, 128)]
    for k, (B, H, W) in enumerate(shapes):
        im1, im2 = synthetic_pair(B, H, W, seed=40 + k)
        im1, im2 = im1.to(dev()), im2.to(dev())
        _, up = model(im1, im2, iters=3, test_mode=True)
        _, ref = fresh(im1, im2, iters=3, test_mode=True)
        assert torch.equal(up, ref), (k, float((up - ref).abs().max()))


# ---------------------------------------------------------------------------------------------
# IGEV pre-loop volume stage: exact-fp32 kernels, tolerance = fp32 round-off of a different
# summation order (2e-5 abs on O(1..10) values; the soft-argmin's disparities are O(D): 1e-4)
# ---------------------------------------------------------------------------------------------
def _bn_fold(g):
    scale = g["bn_weight"] / torch.sqrt(g["bn_var"] + float(g["bn_eps"]))
    return scale, g["bn_bias"] - g["bn_mean"] * scale


def test_igev_volume_stage_golden():
    """dkt_gwc_volume / dkt_conv3d_c8 / dkt_softargmin against the real reference's modules (igev_volume.npz)."""
    from dkt_stereo_b200 import ops
    g = load_golden("igev_volume")
    B, C, H, W, D = [int(v) for v in g["meta"]]
    d = dev()
    gwc = ops.gwc_volume(g["left"].to(d), g["right"].to(d), D, 8)
    assert stats(gwc.cpu(), g["gwc"])[1] < 2e-6, stats(gwc.cpu(), g["gwc"])
    scale, shift = _bn_fold(g)
    vol = ops.conv3d_c8(g["gwc"].to(d), g["stem_w"].to(d), scale.to(d), shift.to(d), 0.01, g["att_logits"].to(d))
    assert stats(vol.cpu(), g["vol"])[1] < 2e-5, stats(vol.cpu(), g["vol"])
    logits = ops.conv3d_c8(g["vol"].to(d), g["cls_w"].to(d))
    assert logits.shape == (B, 1, D, H, W)
    assert stats(logits.cpu().squeeze(1), g["logits"])[1] < 2e-5
    disp = ops.softargmin(g["logits"].to(d))
    assert stats(disp.cpu(), g["disp"])[1] < 1e-4
    # chained, as IGEVStereo.prepare runs them
    disp2 = ops.softargmin(ops.conv3d_c8(vol, g["cls_w"].to(d)).squeeze(1))
    assert stats(disp2.cpu(), g["disp"])[1] < 1e-4


@pytest.mark.parametrize("shape", [(2, 96, 5, 150, 48), (1, 32, 3, 64, 7), (1, 96, 2, 33, 48)])
def test_gwc_volume_vs_oracle(shape):
    """Several x tiles, W < D (disparities that never see a valid pixel stay 0), ragged widths."""
    from dkt_stereo_b200 import ops
    from oracle import hotpath as O
    B, C, H, W, D = shape
    g = torch.Generator().manual_seed(W + D)
    left, right = torch.randn(B, C, H, W, generator=g), torch.randn(B, C, H, W, generator=g)
    ref = O.gwc_volume(left, right, D, 8)
    got = torch.full((B, 8, D, H, W), float("nan"), device=dev())
    ops.gwc_volume(left.to(dev()), right.to(dev()), D, 8, out=got)
    assert not torch.isnan(got).any()
    assert stats(got.cpu(), ref)[1] < 2e-6, stats(got.cpu(), ref)


@pytest.mark.parametrize("shape", [(2, 8, 48, 11, 70), (1, 8, 5, 4, 32), (1, 1, 48, 9, 45), (2, 1, 17, 6, 33)])
def test_conv3d_c8_vs_oracle(shape):
    """Ragged tiles in d / y / x, both channel variants, every epilogue term on and off."""
    from dkt_stereo_b200 import ops
    from oracle import hotpath as O
    B, CO, D, H, W = shape
    g = torch.Generator().manual_seed(D * 7 + W)
    x = torch.randn(B, 8, D, H, W, generator=g)
    w = torch.randn(CO, 8, 3, 3, 3, generator=g) * 0.2
    bn = dict(weight=torch.randn(CO, generator=g), bias=torch.randn(CO, generator=g), running_mean=torch.randn(CO, generator=g) * 0.1,
              running_var=torch.rand(CO, generator=g) + 0.5, eps=1e-5)
    att = torch.randn(B, CO, H, W, generator=g)
    scale = bn["weight"] / torch.sqrt(bn["running_var"] + bn["eps"])
    shift = bn["bias"] - bn["running_mean"] * scale
    d = dev()
    for use_bn, slope, use_att in ((True, 0.01, True), (False, 1.0, False), (True, 1.0, False)):
        ref = O.conv3d_bn_leaky_att(x, w, bn if use_bn else None, slope, att if use_att else None)
        got = torch.full((B, CO, D, H, W), float("nan"), device=d)
        ops.conv3d_c8(x.to(d), w.to(d), scale.to(d) if use_bn else None, shift.to(d) if use_bn else None, slope,
                      att.to(d) if use_att else None, out=got)
        assert not torch.isnan(got).any()
        assert stats(got.cpu(), ref)[1] < 3e-5, (use_bn, slope, use_att, stats(got.cpu(), ref))


def test_softargmin_vs_oracle():
    from dkt_stereo_b200 import ops
    from oracle import hotpath as O
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(3, 48, 7, 61, generator=g) * 4
    ref = O.softargmin(logits)
    got = ops.softargmin(logits.to(dev()))
    assert got.shape == ref.shape
    assert stats(got.cpu(), ref)[1] < 1e-4


def test_igev_native_volume_stage_end_to_end(monkeypatch):
    """IGEVStereo with the volume stage on libdkt kernels vs the same model with that stage in PyTorch fp32
    (DKT_NATIVE_VOLUME=0): geometry volume and initial disparity agree to fp32 round-off, the final disparity within the
    end-to-end gate (1e-3 px mean-abs)."""
    from dkt_stereo_b200.igev_stereo import IGEVStereo
    from dkt_stereo_b200.synthetic import synthetic_pair
    torch.manual_seed(12)
    monkeypatch.setenv("DKT_NATIVE_VOLUME", "1")
    a = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval().to(dev())
    monkeypatch.setenv("DKT_NATIVE_VOLUME", "0")
    b = IGEVStereo(Namespace(mixed_precision=False, **IGEV_CFG)).eval().to(dev())
    assert a.native_volume and not b.native_volume
    with torch.no_grad():                      # non-trivial BatchNorm statistics in the folded layer
        a.corr_stem.bn.running_mean.normal_(0, 0.2)
        a.corr_stem.bn.running_var.uniform_(0.5, 1.5)
        a.corr_stem.bn.weight.normal_(1, 0.2)
        a.corr_stem.bn.bias.normal_(0, 0.2)
    b.load_state_dict(a.state_dict(), strict=True)
    im1, im2 = synthetic_pair(2, 128, 224, seed=22, mode="shift")
    im1, im2 = im1.to(dev()), im2.to(dev())
    with torch.no_grad():
        pa, pb = a.prepare(im1, im2), b.prepare(im1, im2)
    for name, i in (("gev", 2), ("init_disp", 3)):
        mean, mx = stats(pa[i].cpu(), pb[i].cpu())
        print(f"[parity] igev native volume stage {name}: mean-abs {mean:.3e}, max-abs {mx:.3e}")
        assert mx < 5e-4, (name, mean, mx)
    _, ua = a(im1, im2, iters=6, test_mode=True)
    _, ub = b(im1, im2, iters=6, test_mode=True)
    mean, mx = stats(ua.cpu(), ub.cpu())
    print(f"[parity] igev native volume stage, final disparity: mean-abs {mean:.3e} px, max-abs {mx:.3e} px")
    assert mean <= 1e-3, (mean, mx)


@pytest.mark.parametrize("shape", [(2, 16, 16, 9, 7, 45, 1), (1, 16, 32, 12, 10, 70, 2), (1, 32, 48, 6, 9, 33, 2),
                                   (1, 8, 16, 11, 9, 37, 2), (1, 48, 48, 5, 6, 30, 1), (1, 6, 8, 4, 5, 20, 1)])
def test_conv3d_k3_vs_torch(shape):
    """Every channel blocking (16 / 8 per thread, partial channel blocks), both strides, ragged tiles; fp32 reference =
    torch conv3d on CPU (the oracle's primitive)."""
    from dkt_stereo_b200 import ops
    from oracle import hotpath as O
    B, CI, CO, D, H, W, stride = shape
    g = torch.Generator().manual_seed(CI * CO + D)
    x = torch.randn(B, CI, D, H, W, generator=g)
    w = torch.randn(CO, CI, 3, 3, 3, generator=g) / (CI * 27) ** 0.5
    scale, shift = torch.randn(CO, generator=g), torch.randn(CO, generator=g)
    ref = torch.nn.functional.conv3d(x, w, None, stride=stride, padding=1)
    att = torch.randn(B, CO, ref.shape[3], ref.shape[4], generator=g)
    ref = torch.nn.functional.leaky_relu(ref * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1), 0.01) * torch.sigmoid(att.unsqueeze(2))
    d = dev()
    got = ops.conv3d_k3(x.to(d), w.to(d), scale.to(d), shift.to(d), 0.01, att.to(d), stride)
    assert got.shape == ref.shape
    assert stats(got.cpu(), ref)[1] < 3e-5, stats(got.cpu(), ref)


@pytest.mark.parametrize("shape", [(2, 16, 8, 5, 6, 37), (1, 48, 32, 3, 9, 15), (1, 32, 16, 6, 4, 32), (1, 8, 6, 2, 3, 5)])
def test_deconv3d_k4s2_vs_torch(shape):
    from dkt_stereo_b200 import ops
    B, CI, CO, D, H, W = shape
    g = torch.Generator().manual_seed(CI + CO + W)
    x = torch.randn(B, CI, D, H, W, generator=g)
    w = torch.randn(CI, CO, 4, 4, 4, generator=g) / (CI * 8) ** 0.5
    scale, shift = torch.randn(CO, generator=g), torch.randn(CO, generator=g)
    ref = torch.nn.functional.conv_transpose3d(x, w, None, stride=2, padding=1)
    d = dev()
    got = ops.deconv3d_k4s2(x.to(d), w.to(d))
    assert got.shape == ref.shape == (B, CO, 2 * D, 2 * H, 2 * W)
    assert stats(got.cpu(), ref)[1] < 3e-5, stats(got.cpu(), ref)
    ref2 = torch.nn.functional.leaky_relu(ref * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1), 0.01)
    got2 = ops.deconv3d_k4s2(x.to(d), w.to(d), scale.to(d), shift.to(d), 0.01)
    assert stats(got2.cpu(), ref2)[1] < 3e-5


@pytest.mark.parametrize("shape", [(2, 32, 32, 32, 5, 6, 37), (1, 16, 16, 16, 4, 9, 70), (1, 24, 0, 20, 3, 5, 11)])
def test_conv3d_k1_concat_vs_torch(shape):
    from dkt_stereo_b200 import ops
    B, C0, C1, CO, D, H, W = shape
    g = torch.Generator().manual_seed(C0 + CO)
    a = torch.randn(B, C0, D, H, W, generator=g)
    b = torch.randn(B, C1, D, H, W, generator=g) if C1 else None
    w = torch.randn(CO, C0 + C1, 1, 1, 1, generator=g) / (C0 + C1) ** 0.5
    scale, shift = torch.randn(CO, generator=g), torch.randn(CO, generator=g)
    xin = torch.cat((a, b), 1) if C1 else a
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv3d(xin, w) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1), 0.01)
    d = dev()
    got = ops.conv3d_k1(a.to(d), b.to(d) if C1 else None, w.to(d), scale.to(d), shift.to(d), 0.01)
    assert stats(got.cpu(), ref)[1] < 2e-5, stats(got.cpu(), ref)


def test_igev_hourglass_golden():
    """Hourglass.forward_native (libdkt 3-D kernels) against the REAL reference hourglass's output
    (tests/golden/igev_hourglass.npz): same state dict, fp32 round-off only."""
    from dkt_stereo_b200.igev_modules import Hourglass
    g = load_golden("igev_hourglass")
    hg = Hourglass(8).eval()
    hg.load_state_dict(golden_state_dict(g), strict=True)
    hg = hg.to(dev())
    x = g["x"].to(dev())
    feats = [None, g["feat1"].to(dev()), g["feat2"].to(dev()), g["feat3"].to(dev())]
    assert hg.native_ok(x)
    from dkt_stereo_b200.raft_stereo import _fp32_math
    with torch.no_grad(), _fp32_math(True):    # the 2-D attention convs stay in PyTorch: no TF32, as in IGEVStereo.prepare
        out = hg.forward_native(x, feats)
        out_t = hg(x, feats)
    scale = float(g["out"].abs().max()) + 1.0
    mean, mx = stats(out.cpu(), g["out"])
    print(f"[parity] igev hourglass native vs reference: mean-abs {mean:.3e}, max-abs {mx:.3e} (|out| max {scale - 1:.2f})")
    assert mx < 3e-5 * scale, (mean, mx)
    assert stats(out_t.cpu(), g["out"])[1] < 3e-5 * scale


@pytest.mark.parametrize("stride", [1, 2])
def test_dwconv3x3_vs_torch(stride):
    """dkt_dwconv3x3 (MobileNetV2 conv_dw + folded BatchNorm + ReLU6, input clamp) against F.conv2d(groups=C): odd sizes,
    channel slice of a wider source, fp32 and 16-bit (hi, lo) destinations."""
    import ctypes
    from dkt_stereo_b200 import _lib as L
    torch.manual_seed(stride)
    B, Cs, C, H, W = 2, 40, 24, 13, 21
    src = torch.randn(B, H, W, Cs, device=dev()) * 4
    w = torch.randn(C, 1, 3, 3, device=dev())
    bias = torch.randn(C, device=dev())
    x = src[..., 8:8 + C].clamp(max=6.0).permute(0, 3, 1, 2)
    ref = (torch.nn.functional.conv2d(x, w, bias, stride=stride, padding=1, groups=C)).clamp(0, 6).permute(0, 2, 3, 1)
    Ho, Wo = ref.shape[1:3]
    of = torch.zeros(B, Ho, Wo, C, device=dev())
    oh = torch.zeros(B, Ho, Wo, C, device=dev(), dtype=L.split_dtype())
    ol = torch.zeros_like(oh)
    wt = w[:, 0].permute(1, 2, 0).reshape(9, C).contiguous()
    s_t, d_t = L.tensor_slice(src, None, None, 8, C), L.tensor_slice(of, oh, ol, 0, C)
    L.check(L.load().dkt_dwconv3x3(ctypes.byref(s_t), wt.data_ptr(), bias.data_ptr(), 6.0, 0.0, 6.0, ctypes.byref(d_t),
                                   B, H, W, stride, L.stream_ptr()), "dwconv3x3")
    assert stats(of.cpu(), ref.cpu())[1] < 1e-5, stats(of.cpu(), ref.cpu())
    assert stats((oh.float() + ol.float()).cpu(), ref.cpu())[1] < 1e-4


def test_igev_feature_pyramid_native_vs_modules(monkeypatch):
    """Feature.forward on the library's kernels (MobileNetV2 encoder: folded BatchNorm, depthwise kernel, residual on the
    project conv; decoder: parity-grouped transposed convs, InstanceNorm kernels) against the same module's PyTorch path
    (fp32, TF32 off) with non-trivial BatchNorm statistics; the end-to-end goldens pin it against the reference itself."""
    from dkt_stereo_b200.igev_modules import Feature
    from dkt_stereo_b200.raft_stereo import _fp32_math
    torch.manual_seed(3)
    f = Feature().eval().to(dev())
    with torch.no_grad():
        for m in f.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.1)
    x = torch.rand(2, 3, 96, 160, device=dev()) * 2 - 1
    with torch.no_grad(), _fp32_math(True):
        monkeypatch.setenv("DKT_NATIVE_MBV2", "0"); monkeypatch.setenv("DKT_NATIVE_FEATUP", "0")
        ref = f(x)
        monkeypatch.setenv("DKT_NATIVE_MBV2", "1"); monkeypatch.setenv("DKT_NATIVE_FEATUP", "1")
        for _ in range(2):
            out = f(x)
    assert [tuple(o.shape) for o in out] == [tuple(r.shape) for r in ref]
    for i, (o, r) in enumerate(zip(out, ref)):
        rng = float(r.abs().max())
        assert stats(o.cpu(), r.cpu())[1] < 2e-4 * rng, (i, stats(o.cpu(), r.cpu()), rng)


@pytest.mark.parametrize("C,shape", [(32, (2, 32, 12, 34, 60)), (48, (1, 48, 6, 17, 30)), (16, (1, 16, 8, 20, 44))])
def test_hourglass_layer_on_tensor_cores_vs_fp32_kernel(C, shape):
    """Hourglass._k3_tc -- a stride-1 3x3x3 BasicConv + FeatureAtt as one launch of the 2-D tcgen05 conv over the depth planes
    of a depth-padded NDHWC copy (dkt_ncdhw_to_ndhwc_pad / dkt_ndhwc_pad_to_ncdhw) -- against the exact-fp32 3-D kernel:
    ragged tiles, zero padding in depth at both ends of every sample, non-trivial BatchNorm statistics.  Gate 2e-5 of the
    output range (16-bit (hi, lo) operands, 3 MMAs per K step)."""
    from dkt_stereo_b200 import ops
    from dkt_stereo_b200.igev_modules import ConvNormAct, Hourglass
    torch.manual_seed(C)
    hg = Hourglass(8).eval().to(dev())
    m = ConvNormAct(C, C, is_3d=True, kernel_size=3, padding=1, stride=1).eval().to(dev())
    with torch.no_grad():
        m.bn.running_mean.normal_(0, 0.3); m.bn.running_var.uniform_(0.5, 2.0); m.bn.weight.uniform_(0.5, 1.5); m.bn.bias.normal_(0, 0.2)
    v = torch.randn(*shape, device=dev())
    att = torch.randn(shape[0], C, shape[3], shape[4], device=dev())
    scale = (m.bn.weight / torch.sqrt(m.bn.running_var + m.bn.eps)).detach()
    shift = (m.bn.bias - m.bn.running_mean * scale).detach()
    with torch.no_grad():
        ref = ops.conv3d_k3(v, m.conv.weight, scale, shift, 0.01, att, 1)
        for _ in range(2):                                          # second call: cached packs / buffers
            out = hg._k3_tc(m, v, att)
        from dkt_stereo_b200.raft_stereo import _fp32_math
        with _fp32_math(True):                                      # cuDNN would run the module in TF32 otherwise
            ref_t = torch.sigmoid(att).unsqueeze(2) * m(v)
    rng = float(ref.abs().max())
    assert out.shape == ref.shape and stats(out.cpu(), ref.cpu())[1] < 2e-5 * rng, stats(out.cpu(), ref.cpu())
    assert stats(out.cpu(), ref_t.cpu())[1] < 5e-5 * rng


# ---------------------------------------------------------------------------------------------
# a8: pool2x / interp between the GRU scales (reference core/update.py:87-95), called directly
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 128, 136, 240), (1, 128, 17, 31), (3, 64, 8, 9)])
def test_pool2x_and_interp_vs_torch(shape):
    """dkt_pool2x = F.avg_pool2d(x, 3, stride=2, padding=1) (count_include_pad: border windows divide by 9) and
    dkt_interp = F.interpolate(bilinear, align_corners=True) to the finer grid; fp32 source, destination written in
    all three precisions into a channel slice of a wider buffer (the way the update engine uses them).
    Tolerances: fp32 output 2e-6 (same arithmetic, different order); hi + lo reconstructs fp32 to 2^-16 relative."""
    import torch.nn.functional as F
    from dkt_stereo_b200 import ops
    from dkt_stereo_b200._lib import tensor_slice
    B, Cc, H, W = shape
    g = torch.Generator().manual_seed(H * W)
    x = torch.randn(B, Cc, H, W, generator=g)
    xs = _nhwc(x).to(dev())
    src = tensor_slice(xs, None, None, 0, Cc)
    # pool to the coarser grid, into channels [Cc, 2Cc) of a 3Cc-wide buffer
    Hd, Wd = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    f32 = torch.full((B, Hd, Wd, 3 * Cc), 7.0, device=dev())
    hi = torch.zeros(B, Hd, Wd, 3 * Cc, device=dev(), dtype=L16())
    lo = torch.zeros_like(hi)
    ops.pool2x(src, tensor_slice(f32, hi, lo, Cc, Cc), B, H, W)
    ref = _nhwc(F.avg_pool2d(x, 3, stride=2, padding=1))
    got = f32[..., Cc:2 * Cc].cpu()
    assert stats(got, ref)[1] < 2e-6, stats(got, ref)
    rec = (hi.float() + lo.float())[..., Cc:2 * Cc].cpu()
    assert stats(rec, ref)[1] < 1e-4
    assert bool((f32[..., :Cc] == 7.0).all()) and bool((f32[..., 2 * Cc:] == 7.0).all())      # neighbours untouched
    # interpolate the pooled map back to (H, W)
    coarse = f32[..., Cc:2 * Cc].contiguous()
    up32 = torch.zeros(B, H, W, 2 * Cc, device=dev())
    uph = torch.zeros(B, H, W, 2 * Cc, device=dev(), dtype=L16())
    upl = torch.zeros_like(uph)
    ops.interp(tensor_slice(coarse, None, None, 0, Cc), tensor_slice(up32, uph, upl, Cc, Cc), B, Hd, Wd, H, W)
    ref_up = _nhwc(F.interpolate(coarse.permute(0, 3, 1, 2).cpu(), (H, W), mode="bilinear", align_corners=True))
    got_up = up32[..., Cc:].cpu()
    assert stats(got_up, ref_up)[1] < 5e-6, stats(got_up, ref_up)
    assert stats((uph.float() + upl.float())[..., Cc:].cpu(), ref_up)[1] < 1e-4
    assert bool((up32[..., :Cc] == 0).all())
