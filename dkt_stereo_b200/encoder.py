"""Feature / context encoders of RAFT-Stereo executed on the B200 kernels.

The PyTorch modules in ``extractor.py`` stay the *parameter containers* (reference state-dict keys,
reference core/extractor.py:122-300); ``EncoderEngine`` runs their arithmetic with the same
tensor-core convolution kernel as the GRU loop (3-term bf16 split, fp32 accumulate), NHWC end to end:

* stem            image normalisation + x-im2col (``dkt_stem_rows_bf16x2``), then the 7x7 conv as a 7x1 conv
* residual stages ``dkt_conv2d_tc_ex`` (3x3 / strided 3x3 / strided 1x1 shortcut)
    - fnet (InstanceNorm): conv -> raw fp32 -> ``dkt_instnorm_stats`` -> ``dkt_instnorm_apply``
      (normalise + ReLU + residual in one pass, writes fp32 + bf16 hi/lo)
    - cnet (BatchNorm, eval): the norm is folded into weight/bias at pack time; ReLU and the
      residual ``relu(x + y)`` run in the conv epilogue
* outputs         fnet's 1x1 conv writes the matching features directly as NHWC bf16 (hi, lo) -- the
                  operand layout of the correlation build, so no split / transpose pass exists; cnet's
                  heads write tanh(hidden) straight into the GRU state buffers and the context convs
                  (``context_zqr_convs`` + folded GRU gate biases, reference raft_stereo.py:110-114)
                  straight into the per-scale context buffers of ``UpdateEngine``.

96-channel stages run as 96 channels (N = 96; the second 64-channel K block of a 96-channel source is half
used, conv_tc.cu `klast`).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from ._lib import tensor_slice as TS
from .extractor import BasicEncoder, MultiBasicEncoder, ResidualBlock
from .update import UpdateEngine


# 96-channel stages (layer2) run as 96 channels: N = 96 is a legal UMMA shape and a source's last 64-channel K block
# may be partial (conv_tc.cu `klast`).  DKT_ENC_PAD64=1 restores the zero-padding to 128 (1.78x the MMA work of those layers).
_PAD = 64 if os.environ.get("DKT_ENC_PAD64", "0") == "1" else 16


def _pad64(c: int) -> int:
    return (c + _PAD - 1) // _PAD * _PAD


class _Act:
    """An NHWC activation (B,H,W,C) held as fp32 and/or bf16 (hi, lo)."""

    def __init__(self, B, H, W, Cc, device, f32=True, split=True):
        self.B, self.H, self.W, self.C = B, H, W, Cc
        self.f32 = torch.zeros(B, H, W, Cc, device=device) if f32 else None
        self.hi = torch.zeros(B, H, W, Cc, device=device, dtype=L.split_dtype()) if split else None
        self.lo = torch.zeros(B, H, W, Cc, device=device, dtype=L.split_dtype()) if split else None

    def view(self, B: int) -> "_Act":
        """The first B images (cnet reuses fnet's larger scratch buffers)."""
        v = object.__new__(_Act)
        v.B, v.H, v.W, v.C = B, self.H, self.W, self.C
        v.f32 = self.f32[:B] if self.f32 is not None else None
        v.hi = self.hi[:B] if self.hi is not None else None
        v.lo = self.lo[:B] if self.lo is not None else None
        return v

    def s(self, f32=True, split=True, c0=0, cnt=None):
        return TS(self.f32 if f32 else None, self.hi if split else None, self.lo if split else None, c0, cnt)


class EncoderEngine:
    def __init__(self, fnet: Optional[BasicEncoder], cnet: MultiBasicEncoder, zqr_convs: nn.ModuleList, update: UpdateEngine):
        """``fnet=None``: context encoder only (IGEV-Stereo, whose matching features come from its own pyramid)."""
        assert (fnet is None or fnet.norm_fn == "instance") and cnet.norm_fn == "batch", \
            "engine serves configs/raft_stereo/base.json and configs/igev_stereo/base.json"
        self.fnet, self.cnet, self.zqr, self.update = fnet, cnet, zqr_convs, update
        self.w: Optional[Dict[str, ops.ConvWeights]] = None
        self._sig = None
        self.shape = None

    # ---- weights -------------------------------------------------------------------------------
    def _signature(self):
        mods = (list(self.fnet.parameters()) if self.fnet is not None else []) + list(self.cnet.parameters()) + \
            list(self.cnet.buffers()) + list(self.zqr.parameters()) + list(self.update.block.parameters())
        return tuple((p.data_ptr(), p._version) for p in mods)

    @staticmethod
    def _bn(norm: nn.Module):
        return (norm.weight, norm.bias, norm.running_mean, norm.running_var) if isinstance(norm, nn.BatchNorm2d) else None

    def _pack_block(self, w, key: str, blk: ResidualBlock):
        cin, cout = blk.conv1.in_channels, blk.conv1.out_channels
        stride = blk.conv1.stride[0]
        w[key + ".conv1"] = ops.pack_conv_general(blk.conv1.weight, blk.conv1.bias, stride=stride, cin_pad=_pad64(cin),
                                                  n_pad=_pad64(cout), bn=self._bn(blk.norm1))
        w[key + ".conv2"] = ops.pack_conv_general(blk.conv2.weight, blk.conv2.bias, cin_pad=_pad64(cout),
                                                  n_pad=_pad64(cout), bn=self._bn(blk.norm2))
        if blk.downsample is not None:
            ds = blk.downsample[0]
            w[key + ".ds"] = ops.pack_conv_general(ds.weight, ds.bias, stride=stride, cin_pad=_pad64(cin),
                                                   n_pad=_pad64(cout), bn=self._bn(blk.norm3))

    def pack_weights(self) -> bool:
        """Repack when a parameter changed; returns True if it did (captured CUDA graphs are stale then)."""
        sig = self._signature()
        if self.w is not None and sig == self._sig:
            return False
        self.update.pack_weights()
        w: Dict[str, ops.ConvWeights] = {}
        for tag, net in (("f", self.fnet), ("c", self.cnet)):
            if net is None:
                continue
            c1 = net.conv1
            assert c1.kernel_size == (7, 7) and c1.stride == (1, 1) and c1.in_channels == 3, "n_downsample <= 2 stem"
            # 7x7x3 -> 7x1 over the x-im2col channels kx*3 + c (dkt_stem_rows_bf16x2)
            w7 = c1.weight.detach().permute(0, 3, 1, 2).reshape(c1.out_channels, 21, 7, 1)
            w[tag + ".conv1"] = ops.pack_conv_general(w7, c1.bias, cin_pad=32, bn=self._bn(net.norm1))
            for lname in ("layer1", "layer2", "layer3") + (("layer4", "layer5") if tag == "c" else ()):
                for i, blk in enumerate(getattr(net, lname)):
                    self._pack_block(w, f"{tag}.{lname}.{i}", blk)
        if self.fnet is not None:
            c2 = self.fnet.conv2                  # 128 -> 256 as two N = 128 halves (second accumulator for the lo products)
            for g in range(2):
                w[f"f.conv2.{g}"] = ops.pack_conv_general(c2.weight[128 * g:128 * (g + 1)], c2.bias[128 * g:128 * (g + 1)])
        heads = [getattr(self.cnet, n) for n in self.cnet.head_names][:len(self.zqr)]     # n_gru_layers levels
        for i, hl in enumerate(heads):
            for j, head in enumerate(hl):
                if isinstance(head, nn.Sequential):
                    self._pack_block(w, f"c.head{i}.{j}.block", head[0])
                    conv = head[1]
                else:
                    conv = head
                w[f"c.head{i}.{j}.conv"] = ops.pack_conv_general(conv.weight, conv.bias)
            z = self.zqr[i]
            bias = z.bias.detach() + self.update.gru_bias[i].to(z.bias.device)      # fold the GRU gate biases
            # three N = 128 convs, not N = 256 + 128: with N <= 128 the lo products get their own TMEM accumulator (the
            # context terms are constant over the iterations, so a bias in them is the most expensive error of the whole
            # forward: 1.3e-3 px for a half-precision activation here, profiles/r2_precision_study_encoder_layers.txt)
            for g, name in enumerate("zrq"):
                w[f"c.zqr{i}.{name}"] = ops.pack_conv_general(z.weight[128 * g:128 * (g + 1)], bias[128 * g:128 * (g + 1)])
        self.w, self._sig = w, sig
        return True

    # ---- buffers -------------------------------------------------------------------------------
    def allocate(self, B: int, H: int, W: int, device) -> None:
        shape = (B, H, W, str(device))
        if self.shape == shape:
            return
        self.device, self.B, self.H, self.W = device, B, H, W
        nd = self.cnet.downsample
        assert nd == 2, "engine serves n_downsample = 2 (configs/*/base.json)"
        dims = [(H, W)]
        for _ in range(4):
            h, w_ = dims[-1]
            dims.append(((h - 1) // 2 + 1, (w_ - 1) // 2 + 1))
        self.dims = dims                                   # full, 1/2, 1/4, 1/8, 1/16
        B2 = 2 * B if self.fnet is not None else B
        chans = [64, _pad64(self.cnet.layer2[0].conv1.out_channels), 128, 128, 128]
        nimg = [B2, B2, B2, B, B]
        # block inputs / outputs live as bf16 (hi, lo) only: the residual x of relu(x + y) is read back as hi + lo
        # (16 mantissa bits, the precision every conv already consumes), so no fp32 copy of them is written or read.
        # DKT_SPLIT_RES=0 keeps fp32 copies and fp32 residuals.
        self.split_res = os.environ.get("DKT_SPLIT_RES", "1") == "1"
        kf = not self.split_res
        self.lvl = []
        for (h, w_), c, n in zip(dims, chans, nimg):
            self.lvl.append(dict(A=_Act(n, h, w_, c, device, f32=kf), Bb=_Act(n, h, w_, c, device, f32=kf),
                                 Y=_Act(n, h, w_, c, device, f32=False), RAW=_Act(n, h, w_, c, device, split=False),
                                 RAWD=_Act(n, h, w_, c, device, split=False)))
        self.STEM = _Act(B2, H, W, 32, device, f32=False)     # 7 x 3 = 21 x-im2col channels in a 32-channel K block
        self.FMAP = _Act(B2, dims[2][0], dims[2][1], 256, device, f32=False) if self.fnet is not None else None
        self.HEAD = [_Act(B, *dims[2 + i], 128, device, f32=kf) for i in range(3)]   # head residual-block output
        self.HY = [_Act(B, *dims[2 + i], 128, device, f32=False) for i in range(3)]
        self.CIN = [_Act(B, *dims[2 + i], 128, device, f32=False) for i in range(3)]  # relu(context head)
        self.stats = torch.zeros(B2, 128, 2, device=device)
        self.stats2 = torch.zeros(B2, 128, 2, device=device)
        self.ws = ops.instnorm_workspace(B2, 128, device)
        # InstanceNorm statistics come out of the producing conv's epilogue as per-tile sums (DKT_FUSED_STATS=0:
        # separate statistics pass over the raw conv output)
        self.fused_stats = os.environ.get("DKT_FUSED_STATS", "1") == "1"
        self.tile_part = torch.zeros(B2 * max(ops.conv_tiles(h, w_) * 2 * c for (h, w_), c in zip(dims, chans)),
                                     device=device)
        self.tile_ws = ops.instnorm_tiles_workspace(B2, 128, device)
        self.shape = shape

    # ---- building blocks -------------------------------------------------------------------------
    def _o(self, a: _Act):
        """destination slice of a block input / output activation"""
        return a.s(f32=not self.split_res)

    def _res_t(self, x: _Act):
        """residual operand of instnorm_apply (dkt_tensor)"""
        return x.s(f32=False) if self.split_res else x.s(split=False)

    def _res_e(self, x: _Act):
        """residual operand of a conv epilogue"""
        return ((x.hi, x.lo), 0) if self.split_res else (x.f32, 0)

    def _conv(self, key: str, src: _Act, out_slice, Hin: int, Win: int, act=L.ACT_NONE, res=None,
              stats: bool = False) -> Tuple[int, int]:
        """``stats``: the output feeds an InstanceNorm -- the epilogue also writes its per-tile channel sums."""
        w = self.w[key]
        part = self.tile_part if (stats and self.fused_stats) else None
        e = ops.make_epilogue(L.EPI_LINEAR, out_slice, act=act, bias=w.bias, res=res, stats_partial=part)
        return ops.conv2d_ex([src.s(f32=False)], w, e, src.B, Hin, Win)

    def _in_norm(self, raw: _Act, out_slice, stats, relu=True, res=None):
        """InstanceNorm2d (+ReLU, + residual) of a raw conv output produced by ``_conv(..., stats=True)``."""
        if self.fused_stats:
            ops.instnorm_finalize_tiles(self.tile_part, self.tile_ws, stats, raw.B, raw.C, raw.H, raw.W)
        else:
            ops.instnorm_stats(raw.s(split=False), self.ws, stats, raw.B, raw.H, raw.W)
        ops.instnorm_apply(raw.s(split=False), stats, out_slice, raw.B, raw.H, raw.W, relu=relu, res=res)

    def _block(self, key: str, x: _Act, out: _Act, scratch: dict, inst: bool) -> None:
        """ResidualBlock (reference core/extractor.py:6-60): x -> out (possibly at half resolution)."""
        n = x.B
        Y, RAW, RAWD = scratch["Y"].view(n), scratch["RAW"].view(n), scratch["RAWD"].view(n)
        has_ds = (key + ".ds") in self.w
        if inst:
            self._conv(key + ".conv1", x, RAW.s(split=False), x.H, x.W, stats=True)
            self._in_norm(RAW, Y.s(f32=False), self.stats, relu=True)
            if has_ds:
                self._conv(key + ".ds", x, RAWD.s(split=False), x.H, x.W, stats=True)
                self._in_norm(RAWD, RAWD.s(split=False), self.stats2, relu=False)
                res = RAWD.s(split=False)
            else:
                res = self._res_t(x)
            self._conv(key + ".conv2", Y, RAW.s(split=False), Y.H, Y.W, stats=True)
            self._in_norm(RAW, self._o(out), self.stats, relu=True, res=res)
        else:
            self._conv(key + ".conv1", x, Y.s(f32=False), x.H, x.W, act=L.ACT_RELU)
            if has_ds:
                self._conv(key + ".ds", x, RAWD.s(split=False), x.H, x.W)
                res = (RAWD.f32, 0)
            else:
                res = self._res_e(x)
            self._conv(key + ".conv2", Y, self._o(out), Y.H, Y.W, act=L.ACT_RELU, res=res)

    def _trunk(self, tag: str, net, n: int) -> _Act:
        """conv1 + norm + relu, layer1..3 (reference core/extractor.py:173-190 / 274-281) on the first n
        images of the stem buffer -> 128-channel activation at 1/4 resolution."""
        inst = net.norm_fn == "instance"
        l0 = self.lvl[0]
        x = l0["A"].view(n)
        stem = self.STEM.view(n)
        if inst:
            raw = l0["RAW"].view(n)
            self._conv(tag + ".conv1", stem, raw.s(split=False), self.H, self.W, stats=True)
            self._in_norm(raw, self._o(x), self.stats, relu=True)
        else:
            self._conv(tag + ".conv1", stem, self._o(x), self.H, self.W, act=L.ACT_RELU)
        cur, cur_lvl = x, 0
        for lname, lvl in (("layer1", 0), ("layer2", 1), ("layer3", 2)):
            for i in range(2):
                scratch = self.lvl[lvl]
                cand = scratch["A"].view(n)
                out = scratch["Bb"].view(n) if cand.hi.data_ptr() == cur.hi.data_ptr() else cand
                self._block(f"{tag}.{lname}.{i}", cur, out, scratch, inst)
                cur = out
        return cur

    # ---- forward -----------------------------------------------------------------------------------
    def run(self, image1: torch.Tensor, image2: Optional[torch.Tensor] = None) -> None:
        """image (B,3,H,W) fp32 in [0,255].  Fills self.FMAP (bf16 hi/lo, images [0,B) = left, [B,2B) = right; only
        with an fnet), and the update engine's hidden-state slices X[i][:, :128] and context buffers CTX[i]."""
        B, _, H, W = image1.shape
        L.require_device(image1)
        self.pack_weights()
        self.allocate(B, H, W, image1.device)
        eng = self.update
        eng.allocate(B, self.dims[2][0], self.dims[2][1], image1.device)
        im1 = image1.contiguous().float()
        ops.stem_rows(im1, self.STEM.hi[:B], self.STEM.lo[:B])
        if self.fnet is not None:
            # ---- fnet on [left; right] (reference raft_stereo.py:102) ----
            im2 = image2.contiguous().float()
            ops.stem_rows(im2, self.STEM.hi[B:], self.STEM.lo[B:])
            x = self._trunk("f", self.fnet, 2 * B)
            for g in range(2):
                self._conv(f"f.conv2.{g}", x, self.FMAP.s(f32=False, c0=128 * g, cnt=128), x.H, x.W)
        # ---- cnet on left (reference raft_stereo.py:101); the stem rows of the left images are still there ----
        x = self._trunk("c", self.cnet, B)
        feats = [x]
        nl = len(self.zqr)                                # args.n_gru_layers (reference core/extractor.py:288-295)
        for li, lname in ((3, "layer4"), (4, "layer5"))[:nl - 1]:
            cur = feats[-1]
            for i in range(2):
                scratch = self.lvl[li]
                cand = scratch["A"].view(B)
                out = scratch["Bb"].view(B) if cand.hi.data_ptr() == cur.hi.data_ptr() else cand
                self._block(f"c.{lname}.{i}", cur, out, scratch, False)
                cur = out
            feats.append(cur)
        for i in range(nl):
            f = feats[i]
            scratch = dict(Y=self.HY[i], RAW=self.lvl[2 + i]["RAW"], RAWD=self.lvl[2 + i]["RAWD"])
            for j in range(2):
                src = f
                if (f"c.head{i}.{j}.block.conv1") in self.w:
                    self._block(f"c.head{i}.{j}.block", f, self.HEAD[i], scratch, False)
                    src = self.HEAD[i]
                if j == 0:      # hidden state: tanh -> GRU state buffer (fp32 + hi/lo)
                    split = eng.impl == "tc"
                    self._conv(f"c.head{i}.0.conv", src, eng._slice(eng.X[i], 0, 128, True, split), src.H, src.W, act=L.ACT_TANH)
                else:           # context: relu -> zqr conv (+ folded gate biases) -> CTX
                    self._conv(f"c.head{i}.1.conv", src, self.CIN[i].s(f32=False), src.H, src.W, act=L.ACT_RELU)
                    ctx = eng.CTX[i]
                    for g, name in enumerate("zrq"):
                        self._conv(f"c.zqr{i}.{name}", self.CIN[i], TS(ctx["f32"], None, None, 128 * g, 128), src.H, src.W)
