"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU (PyTorch fp32) restatement of the DKT-Stereo inference hot path, written as
pure functions over a flat ``state_dict`` so that it shares no code with the
engine in ``dkt_stereo_b200/``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; the product must never route through it.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so these functions are pinned against the reference
itself, imported read-only in the build container by ``oracle/make_golden.py``;
the resulting vectors live in ``tests/golden/`` and are re-checked on CPU by
``tests/test_oracle_golden.py``.

Every function cites the reference lines it follows (paths are relative to the
reference checkout).  All arithmetic is fp32 and uses the same third-party
primitives as the reference (torch conv2d / einsum / softmax on CPU).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def _conv(sd: SD, name: str, x: Tensor, stride: int = 1, padding: int = 0) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _grid_sample_x(x: Tensor, width: int) -> Tensor:
    """Pixel coordinate after the reference's normalise -> grid_sample(align_corners=True)
    round trip: core/utils/utils.py:63 (2x/(W-1)-1) then ATen's ((g+1)/2)*(W-1)."""
    g = 2 * x / (width - 1) - 1
    return ((g + 1) / 2) * (width - 1)


def sample_rows_linear(rows: Tensor, x: Tensor) -> Tensor:
    """1-D linear interpolation with zero padding.

    rows: (P, C, W) values; x: (P, T) sample positions in pixels -> (P, C, T).
    Follows ``bilinear_sampler`` core/utils/utils.py:59-74 for an H==1 image:
    F.grid_sample(bilinear, padding_mode='zeros', align_corners=True).
    """
    P, C, W = rows.shape
    ix = _grid_sample_x(x, W)
    x0 = torch.floor(ix)
    a = (ix - x0).unsqueeze(1)                      # (P,1,T)
    i0 = x0.long()
    i1 = i0 + 1

    def tap(idx: Tensor) -> Tensor:
        ok = ((idx >= 0) & (idx < W)).unsqueeze(1).to(rows.dtype)
        g = torch.gather(rows, 2, idx.clamp(0, W - 1).unsqueeze(1).expand(P, C, idx.shape[1]))
        return g * ok

    return (1 - a) * tap(i0) + a * tap(i1)


# ----------------------------------------------------------------------------
# a1/a2  correlation volume + pyramid (RAFT-Stereo)
# ----------------------------------------------------------------------------
def corr1d_all_pairs(fmap1: Tensor, fmap2: Tensor, scale: bool = True) -> Tensor:
    """corr[b,h,i,j] = sum_d f1[b,d,h,i] f2[b,d,h,j] (/ sqrt(D)).

    core/corr.py:148-156 (CorrBlock1D.corr; scale=True) and
    meta_arch/igev_stereo/geometry.py:61-69 (no scaling).  Returns (B,H,W1,W2).
    """
    D = fmap1.shape[1]
    corr = torch.einsum("bdhi,bdhj->bhij", fmap1.float(), fmap2.float())
    if scale:
        corr = corr / torch.sqrt(torch.tensor(D).float())
    return corr


def pool_w2(v: Tensor) -> Tensor:
    """avg_pool2d(kernel [1,2], stride [1,2]) along the last axis; an odd tail
    element is dropped (core/corr.py:123-125)."""
    w = v.shape[-1] // 2
    return (v[..., 0:2 * w:2] + v[..., 1:2 * w:2]) * 0.5


def corr1d_pyramid(fmap1: Tensor, fmap2: Tensor, num_levels: int = 4) -> List[Tensor]:
    """core/corr.py:111-125 -- only the ``num_levels`` levels that the lookup
    reads are returned (the reference stores one more, never read)."""
    pyr = [corr1d_all_pairs(fmap1, fmap2, scale=True)]
    for _ in range(num_levels - 1):
        pyr.append(pool_w2(pyr[-1]))
    return pyr


# ----------------------------------------------------------------------------
# a3  indexed lookup (RAFT-Stereo)
# ----------------------------------------------------------------------------
def corr1d_lookup(pyr: Sequence[Tensor], coords_x: Tensor, radius: int = 4) -> Tensor:
    """core/corr.py:127-146.  pyr[i]: (B,H,W1,W2>>i); coords_x: (B,H,W1) ->
    (B, L*(2r+1), H, W1) with channel = level*(2r+1) + tap."""
    B, H, W1 = coords_x.shape
    dx = torch.linspace(-radius, radius, 2 * radius + 1)
    out = []
    for i, vol in enumerate(pyr):
        rows = vol.reshape(B * H * W1, 1, vol.shape[-1])
        x = coords_x.reshape(-1, 1) / 2 ** i + dx.view(1, -1)
        out.append(sample_rows_linear(rows, x).reshape(B, H, W1, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# ----------------------------------------------------------------------------
# a4/a5  IGEV combined geometry encoding volume
# ----------------------------------------------------------------------------
def geo_pyramids(fmap1: Tensor, fmap2: Tensor, geo_volume: Tensor, num_levels: int = 2):
    """meta_arch/igev_stereo/geometry.py:7-29.  Returns (geo_pyr, init_pyr):
    geo_pyr[i]: (B,H,W,C,D>>i) from geo_volume (B,C,D,H,W); init_pyr[i]: (B,H,W,W2>>i)."""
    init = [corr1d_all_pairs(fmap1, fmap2, scale=False)]
    geo = [geo_volume.float().permute(0, 3, 4, 1, 2).contiguous()]
    for _ in range(num_levels - 1):
        geo.append(pool_w2(geo[-1]))
        init.append(pool_w2(init[-1]))
    return geo, init


def geo_lookup(geo_pyr, init_pyr, disp: Tensor, radius: int = 4) -> Tensor:
    """meta_arch/igev_stereo/geometry.py:34-58.  disp: (B,1,H,W) -> (B, L*(C+1)*(2r+1), H, W);
    per level the C*(2r+1) geometry taps (channel-major) then the (2r+1) init-corr taps
    sampled at (x - disp)/2^i + dx."""
    B, _, H, W = disp.shape
    P = B * H * W
    dx = torch.linspace(-radius, radius, 2 * radius + 1).view(1, -1)
    d = disp.reshape(P, 1)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(B, H, W).reshape(P, 1)
    out = []
    for i, (g, c) in enumerate(zip(geo_pyr, init_pyr)):
        C = g.shape[3]
        rows = g.reshape(P, C, g.shape[-1])
        out.append(sample_rows_linear(rows, d / 2 ** i + dx).reshape(B, H, W, -1))
        rows = c.reshape(P, 1, c.shape[-1])
        out.append(sample_rows_linear(rows, xs / 2 ** i - d / 2 ** i + dx).reshape(B, H, W, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# ----------------------------------------------------------------------------
# a6..a11  update block
# ----------------------------------------------------------------------------
def motion_encoder(sd: SD, pre: str, flow: Tensor, corr: Tensor, igev: bool = False) -> Tensor:
    """core/update.py:64-85 (RAFT, flow has 2 ch) / meta_arch/igev_stereo/update.py:73-92
    (IGEV, disp has 1 ch, convd1/convd2)."""
    f1, f2 = ("convd1", "convd2") if igev else ("convf1", "convf2")
    cor = F.relu(_conv(sd, pre + "convc1", corr))
    cor = F.relu(_conv(sd, pre + "convc2", cor, padding=1))
    flo = F.relu(_conv(sd, pre + f1, flow, padding=3))
    flo = F.relu(_conv(sd, pre + f2, flo, padding=1))
    out = F.relu(_conv(sd, pre + "conv", torch.cat([cor, flo], 1), padding=1))
    return torch.cat([out, flow], 1)


def conv_gru(sd: SD, pre: str, h: Tensor, cz: Tensor, cr: Tensor, cq: Tensor, *xs: Tensor) -> Tensor:
    """core/update.py:16-32."""
    x = torch.cat(xs, 1)
    hx = torch.cat([h, x], 1)
    z = torch.sigmoid(_conv(sd, pre + "convz", hx, padding=1) + cz)
    r = torch.sigmoid(_conv(sd, pre + "convr", hx, padding=1) + cr)
    q = torch.tanh(_conv(sd, pre + "convq", torch.cat([r * h, x], 1), padding=1) + cq)
    return (1 - z) * h + z * q


def pool2x(x: Tensor) -> Tensor:
    """core/update.py:87-88."""
    return F.avg_pool2d(x, 3, stride=2, padding=1)


def interp_to(x: Tensor, dest: Tensor) -> Tensor:
    """core/update.py:93-95."""
    return F.interpolate(x, dest.shape[2:], mode="bilinear", align_corners=True)


def update_block(sd: SD, pre: str, net: List[Tensor], inp, corr: Optional[Tensor], flow: Optional[Tensor],
                 igev: bool = False, n_gru_layers: int = 3, with_mask: bool = True,
                 iter_fine: bool = True, iter_mid: bool = True, iter_coarse: bool = True, update: bool = True):
    """core/update.py:115-138 / meta_arch/igev_stereo/update.py:121-142.  ``iter_*`` / ``update`` are the reference's
    iter08/iter16/iter32 (iter04/08/16 for IGEV) and update flags, used by the slow_fast_gru schedule.
    Returns (net, mask, delta), or net alone when ``update`` is False."""
    g_fine, g_mid, g_coarse = ("gru04", "gru08", "gru16") if igev else ("gru08", "gru16", "gru32")
    net = list(net)
    if n_gru_layers == 3 and iter_coarse:
        net[2] = conv_gru(sd, pre + g_coarse + ".", net[2], *inp[2], pool2x(net[1]))
    if n_gru_layers >= 2 and iter_mid:
        if n_gru_layers > 2:
            net[1] = conv_gru(sd, pre + g_mid + ".", net[1], *inp[1], pool2x(net[0]), interp_to(net[2], net[1]))
        else:
            net[1] = conv_gru(sd, pre + g_mid + ".", net[1], *inp[1], pool2x(net[0]))
    if iter_fine:
        motion = motion_encoder(sd, pre + "encoder.", flow, corr, igev)
        if n_gru_layers > 1:
            net[0] = conv_gru(sd, pre + g_fine + ".", net[0], *inp[0], motion, interp_to(net[1], net[0]))
        else:
            net[0] = conv_gru(sd, pre + g_fine + ".", net[0], *inp[0], motion)
    if not update:
        return net
    head = "disp_head." if igev else "flow_head."
    delta = _conv(sd, pre + head + "conv2", F.relu(_conv(sd, pre + head + "conv1", net[0], padding=1)), padding=1)
    mask = None
    if with_mask:
        if igev:
            mask = F.relu(_conv(sd, pre + "mask_feat_4.0", net[0], padding=1))
        else:
            mask = 0.25 * _conv(sd, pre + "mask.2", F.relu(_conv(sd, pre + "mask.0", net[0], padding=1)))
    return net, mask, delta


# ----------------------------------------------------------------------------
# a13  final upsampling
# ----------------------------------------------------------------------------
def convex_upsample(flow: Tensor, mask: Tensor, factor: int = 4) -> Tensor:
    """meta_arch/raft_stereo/raft_stereo.py:70-82."""
    N, D, H, W = flow.shape
    m = torch.softmax(mask.view(N, 1, 9, factor, factor, H, W), dim=2)
    up = F.unfold(factor * flow, [3, 3], padding=1).view(N, D, 9, 1, 1, H, W)
    up = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, D, factor * H, factor * W)


def context_upsample(disp_low: Tensor, up_weights: Tensor) -> Tensor:
    """meta_arch/igev_stereo/submodule.py:242-254.  disp_low (B,1,h,w), up_weights (B,9,4h,4w)."""
    b, c, h, w = disp_low.shape
    unf = F.unfold(disp_low, 3, 1, 1).reshape(b, -1, h, w)
    unf = F.interpolate(unf, (h * 4, w * 4), mode="nearest").reshape(b, 9, h * 4, w * 4)
    return (unf * up_weights).sum(1)


# ----------------------------------------------------------------------------
# extractors (L1; produce the hot path's inputs) -- functional restatement
# ----------------------------------------------------------------------------
def _norm(sd: SD, name: str, x: Tensor, kind: str) -> Tensor:
    if kind == "instance":
        return F.instance_norm(x, eps=1e-5)
    if kind == "batch":
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                            sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)
    if kind == "none":
        return x
    raise ValueError(kind)


def _res_block(sd: SD, pre: str, x: Tensor, kind: str, stride: int) -> Tensor:
    """core/extractor.py:6-60."""
    y = F.relu(_norm(sd, pre + "norm1", _conv(sd, pre + "conv1", x, stride, 1), kind))
    y = F.relu(_norm(sd, pre + "norm2", _conv(sd, pre + "conv2", y, 1, 1), kind))
    if (pre + "downsample.0.weight") in sd:
        x = _norm(sd, pre + "norm3", _conv(sd, pre + "downsample.0", x, stride, 0), kind)
    return F.relu(x + y)


def _stage(sd: SD, pre: str, x: Tensor, kind: str, stride: int) -> Tensor:
    return _res_block(sd, pre + "1.", _res_block(sd, pre + "0.", x, kind, stride), kind, 1)


def _trunk(sd: SD, pre: str, x: Tensor, kind: str, downsample: int) -> Tensor:
    x = F.relu(_norm(sd, pre + "norm1", _conv(sd, pre + "conv1", x, 1 + (downsample > 2), 3), kind))
    x = _stage(sd, pre + "layer1.", x, kind, 1)
    x = _stage(sd, pre + "layer2.", x, kind, 1 + (downsample > 1))
    return _stage(sd, pre + "layer3.", x, kind, 1 + (downsample > 0))


def basic_encoder(sd: SD, pre: str, x: Tensor, kind: str = "instance", downsample: int = 2) -> Tensor:
    """core/extractor.py:173-197 (fnet)."""
    return _conv(sd, pre + "conv2", _trunk(sd, pre, x, kind, downsample))


def multi_basic_encoder(sd: SD, pre: str, x: Tensor, kind: str = "batch", downsample: int = 2,
                        names=("outputs08", "outputs16", "outputs32")):
    """core/extractor.py:274-300 (cnet, num_layers=3, two heads per scale)."""
    x = _trunk(sd, pre, x, kind, downsample)
    y = _stage(sd, pre + "layer4.", x, kind, 2)
    z = _stage(sd, pre + "layer5.", y, kind, 2)

    def heads(name: str, t: Tensor, with_block: bool):
        outs = []
        for j in range(2):
            p = f"{pre}{name}.{j}."
            if with_block:
                outs.append(_conv(sd, p + "1", _res_block(sd, p + "0.", t, kind, 1), 1, 1))
            else:
                outs.append(_conv(sd, p[:-1], t, 1, 1))
        return outs

    return heads(names[0], x, True), heads(names[1], y, True), heads(names[2], z, False)


# ----------------------------------------------------------------------------
# a12  RAFT-Stereo forward (test_mode=True)
# ----------------------------------------------------------------------------
def raft_prepare(sd: SD, image1: Tensor, image2: Tensor, cfg: dict):
    """meta_arch/raft_stereo/raft_stereo.py:91-116: everything before the volume."""
    ds = cfg.get("n_downsample", 2)
    image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
    image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
    # num_layers = n_gru_layers (raft_stereo.py:101, core/extractor.py:288-295): only the scales that have a GRU
    cnet = multi_basic_encoder(sd, "cnet.", image1, cfg.get("context_norm", "batch"), ds)[:cfg.get("n_gru_layers", 3)]
    fmaps = basic_encoder(sd, "fnet.", torch.cat([image1, image2], 0), "instance", ds)
    fmap1, fmap2 = fmaps.split(image1.shape[0], 0)
    net = [torch.tanh(s[0]) for s in cnet]
    inp = [torch.relu(s[1]) for s in cnet]
    inp = [list(_conv(sd, f"context_zqr_convs.{i}", t, 1, 1).split(t.shape[1], dim=1)) for i, t in enumerate(inp)]
    return fmap1.float(), fmap2.float(), net, inp


def slow_fast_updates(sd: SD, net, inp, cfg: dict, igev: bool):
    """The extra coarse-GRU updates of ``slow_fast_gru`` (raft_stereo.py:157-160, igev_stereo.py:201-204)."""
    n = cfg.get("n_gru_layers", 3)
    if not cfg.get("slow_fast_gru", False):
        return net
    if n == 3:
        net = update_block(sd, "update_block.", net, inp, None, None, igev=igev, n_gru_layers=n,
                           iter_fine=False, iter_mid=False, iter_coarse=True, update=False)
    if n >= 2:
        net = update_block(sd, "update_block.", net, inp, None, None, igev=igev, n_gru_layers=n,
                           iter_fine=False, iter_mid=True, iter_coarse=(n == 3), update=False)
    return net


def raft_loop(sd: SD, fmap1: Tensor, fmap2: Tensor, net, inp, iters: int, cfg: dict,
              flow_init: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """meta_arch/raft_stereo/raft_stereo.py:118-183 with corr_implementation='reg', test_mode=True."""
    L, r = cfg.get("corr_levels", 4), cfg.get("corr_radius", 4)
    B, _, h, w = fmap1.shape
    pyr = corr1d_pyramid(fmap1, fmap2, L)
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    coords0 = torch.stack([xs, ys], 0).float()[None].repeat(B, 1, 1, 1)
    coords1 = coords0.clone()
    if flow_init is not None:
        coords1 = coords1 + flow_init
    mask = None
    for it in range(iters):
        corr = corr1d_lookup(pyr, coords1[:, 0], r)
        flow = coords1 - coords0
        net = slow_fast_updates(sd, net, inp, cfg, igev=False)
        net, mask, delta = update_block(sd, "update_block.", net, inp, corr, flow,
                                        igev=False, n_gru_layers=cfg.get("n_gru_layers", 3),
                                        with_mask=(it == iters - 1))
        delta = delta.clone()
        delta[:, 1] = 0.0
        coords1 = coords1 + delta
    flow_lr = coords1 - coords0
    flow_up = convex_upsample(flow_lr, mask, 2 ** cfg.get("n_downsample", 2))[:, :1]
    return flow_lr, flow_up


def raft_forward(sd: SD, image1: Tensor, image2: Tensor, iters: int, cfg: dict,
                 flow_init: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """RAFTStereo.forward(test_mode=True): meta_arch/raft_stereo/raft_stereo.py:85-183."""
    with torch.no_grad():
        fmap1, fmap2, net, inp = raft_prepare(sd, image1, image2, cfg)
        return raft_loop(sd, fmap1, fmap2, net, inp, iters, cfg, flow_init)


# ----------------------------------------------------------------------------
# IGEV hot loop (a12, IGEV flavour).  The pre-loop (MobileNetV2 features, GWC volume,
# 3-D hourglass) is an SURVEY section 8(f) "next" row: the loop takes its products as inputs.
# ----------------------------------------------------------------------------
def igev_upsample_disp(sd: SD, disp: Tensor, mask_feat_4: Tensor, stem_2x: Tensor) -> Tensor:
    """meta_arch/igev_stereo/igev_stereo.py:140-148: Conv2x(deconv) -> ConvTranspose2d -> softmax ->
    context_upsample.  BasicConv = conv(no bias) + BatchNorm(eval) + LeakyReLU(0.01)
    (meta_arch/igev_stereo/submodule.py:10-36,39-78)."""
    def bconv(pre: str, x: Tensor, deconv: bool, **kw) -> Tensor:
        w = sd[pre + "conv.weight"]
        x = F.conv_transpose2d(x, w, None, **kw) if deconv else F.conv2d(x, w, None, **kw)
        x = F.batch_norm(x, sd[pre + "bn.running_mean"], sd[pre + "bn.running_var"],
                         sd[pre + "bn.weight"], sd[pre + "bn.bias"], False, 0.0, 1e-5)
        return F.leaky_relu(x, 0.01)

    x = bconv("spx_2_gru.conv1.", mask_feat_4, True, stride=2, padding=1)
    if x.shape != stem_2x.shape:
        x = F.interpolate(x, size=stem_2x.shape[-2:], mode="nearest")
    x = bconv("spx_2_gru.conv2.", torch.cat((x, stem_2x), 1), False, stride=1, padding=1)
    spx = F.conv_transpose2d(x, sd["spx_gru.0.weight"], sd["spx_gru.0.bias"], stride=2, padding=1)
    spx = F.softmax(spx, 1)
    return context_upsample(disp * 4.0, spx).unsqueeze(1)


def igev_loop(sd: SD, match_left: Tensor, match_right: Tensor, geo_volume: Tensor, init_disp: Tensor,
              net, inp, stem_2x: Tensor, iters: int, cfg: dict) -> Tensor:
    """meta_arch/igev_stereo/igev_stereo.py:192-220 (test_mode=True) -> disp_up (B,1,H,W), negated."""
    L, r = cfg.get("corr_levels", 2), cfg.get("corr_radius", 4)
    geo_pyr, init_pyr = geo_pyramids(match_left.float(), match_right.float(), geo_volume.float(), L)
    disp = init_disp
    mask_feat = None
    for it in range(iters):
        feat = geo_lookup(geo_pyr, init_pyr, disp, r)
        net = slow_fast_updates(sd, net, inp, cfg, igev=True)
        net, mask_feat, delta = update_block(sd, "update_block.", net, inp, feat, disp, igev=True,
                                             n_gru_layers=cfg.get("n_gru_layers", 3),
                                             with_mask=(it == iters - 1))
        disp = disp + delta
    return -igev_upsample_disp(sd, disp, mask_feat, stem_2x)


# ----------------------------------------------------------------------------
# IGEV pre-loop volume stage
# ----------------------------------------------------------------------------
def gwc_volume(left: Tensor, right: Tensor, maxdisp: int, groups: int) -> Tensor:
    """Group-wise correlation volume (B,groups,maxdisp,H,W): reference
    meta_arch/igev_stereo/submodule.py:152-170 (build_gwc_volume + groupwise_correlation)."""
    B, C, H, W = left.shape
    cpg = C // groups
    vol = left.new_zeros(B, groups, maxdisp, H, W)
    for d in range(maxdisp):
        if d >= W:
            break
        prod = left[:, :, :, d:] * right[:, :, :, : W - d]
        vol[:, :, d, :, d:] = prod.view(B, groups, cpg, H, W - d).mean(dim=2)
    return vol


def conv3d_bn_leaky_att(x: Tensor, weight: Tensor, bn: Optional[Dict[str, Tensor]], slope: float,
                        att_logits: Optional[Tensor]) -> Tensor:
    """BasicConv(is_3d, eval-mode BatchNorm3d, LeakyReLU(0.01)) followed by FeatureAtt's product: reference
    meta_arch/igev_stereo/submodule.py:10-36 and :227-240 (call sites igev_stereo.py:170-171).
    ``bn`` = dict(weight, bias, running_mean, running_var, eps) or None; ``att_logits`` (B,C,H,W) or None."""
    y = F.conv3d(x, weight, None, stride=1, padding=1)
    if bn is not None:
        y = F.batch_norm(y, bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"], False, 0.0, float(bn["eps"]))
    if slope != 1.0:
        y = F.leaky_relu(y, slope)
    if att_logits is not None:
        y = torch.sigmoid(att_logits.unsqueeze(2)) * y
    return y


def softargmin(logits: Tensor) -> Tensor:
    """F.softmax over the disparity axis + disparity_regression: reference igev_stereo.py:175-176 and
    submodule.py:220-224.  logits (B,D,H,W) -> (B,1,H,W)."""
    prob = F.softmax(logits, dim=1)
    D = logits.shape[1]
    values = torch.arange(0, D, dtype=logits.dtype).view(1, D, 1, 1)
    return torch.sum(prob * values, 1, keepdim=True)


def _basic_conv3d(sd: SD, pre: str, x: Tensor, stride: int = 1, padding: int = 1, deconv: bool = False,
                  bn: bool = True, relu: bool = True) -> Tensor:
    """BasicConv(is_3d=True) in eval mode: conv | deconv (no bias) -> BatchNorm3d(running stats) -> LeakyReLU(0.01)
    (reference meta_arch/igev_stereo/submodule.py:10-36)."""
    w = sd[pre + ".conv.weight"]
    y = F.conv_transpose3d(x, w, None, stride=stride, padding=padding) if deconv else F.conv3d(x, w, None, stride=stride, padding=padding)
    if bn:
        y = F.batch_norm(y, sd[pre + ".bn.running_mean"], sd[pre + ".bn.running_var"], sd[pre + ".bn.weight"],
                         sd[pre + ".bn.bias"], False, 0.0, 1e-5)
    return F.leaky_relu(y, 0.01) if relu else y


def _feature_att(sd: SD, pre: str, cv: Tensor, feat: Tensor) -> Tensor:
    """FeatureAtt: cv * sigmoid(conv1x1(BasicConv1x1(feat))) broadcast over disparity (submodule.py:227-240)."""
    y = F.conv2d(feat, sd[pre + ".feat_att.0.conv.weight"])
    y = F.batch_norm(y, sd[pre + ".feat_att.0.bn.running_mean"], sd[pre + ".feat_att.0.bn.running_var"],
                     sd[pre + ".feat_att.0.bn.weight"], sd[pre + ".feat_att.0.bn.bias"], False, 0.0, 1e-5)
    y = F.conv2d(F.leaky_relu(y, 0.01), sd[pre + ".feat_att.1.weight"], sd[pre + ".feat_att.1.bias"])
    return torch.sigmoid(y.unsqueeze(2)) * cv


def hourglass(sd: SD, pre: str, x: Tensor, feats: Sequence[Tensor]) -> Tensor:
    """The 3-D hourglass `cost_agg` (reference meta_arch/igev_stereo/igev_stereo.py:22-89): x (B,8,D,H,W),
    feats[1..3] the 1/8, 1/16, 1/32 feature maps (64 / 192 / 160 channels)."""
    def down(name, v):
        return _basic_conv3d(sd, f"{pre}.{name}.1", _basic_conv3d(sd, f"{pre}.{name}.0", v, stride=2))

    def agg(name, v):
        v = _basic_conv3d(sd, f"{pre}.{name}.0", v, padding=0)
        return _basic_conv3d(sd, f"{pre}.{name}.2", _basic_conv3d(sd, f"{pre}.{name}.1", v))

    c1 = _feature_att(sd, pre + ".feature_att_8", down("conv1", x), feats[1])
    c2 = _feature_att(sd, pre + ".feature_att_16", down("conv2", c1), feats[2])
    c3 = _feature_att(sd, pre + ".feature_att_32", down("conv3", c2), feats[3])
    u3 = _basic_conv3d(sd, pre + ".conv3_up", c3, stride=2, deconv=True)
    c2 = _feature_att(sd, pre + ".feature_att_up_16", agg("agg_0", torch.cat((u3, c2), 1)), feats[2])
    u2 = _basic_conv3d(sd, pre + ".conv2_up", c2, stride=2, deconv=True)
    c1 = _feature_att(sd, pre + ".feature_att_up_8", agg("agg_1", torch.cat((u2, c1), 1)), feats[1])
    return _basic_conv3d(sd, pre + ".conv1_up", c1, stride=2, deconv=True, bn=False, relu=False)
