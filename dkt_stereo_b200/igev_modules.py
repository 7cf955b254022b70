"""PyTorch-side modules of IGEV-Stereo that run BEFORE the hot loop (feature pyramid, group-wise
correlation volume, 3-D hourglass regularisation, soft-argmin initial disparity) and the small
up-sampling convs after it.  They stay on cuDNN (SURVEY section 8f rank 2 is the "next" row that
moves them); only their *products* -- matching features, geometry encoding volume, initial
disparity, context terms -- enter the B200 kernels.

Parameter names reproduce the reference so that IGEV / DKT-IGEV checkpoints load with
``strict=True``:

* ``ConvNormAct``  <- ``BasicConv`` / ``BasicConv_IN`` (meta_arch/igev_stereo/submodule.py:10-36, 80-106):
  attributes ``conv`` + ``bn`` or ``IN``; LeakyReLU(0.01).
* ``UpFuse``       <- ``Conv2x`` / ``Conv2x_IN`` (submodule.py:39-78, 109-148): ``conv1``, ``conv2``.
* ``FeatureAtt``   <- submodule.py:227-240: ``feat_att.{0,1}``.
* ``Hourglass``    <- ``hourglass`` (meta_arch/igev_stereo/igev_stereo.py:22-89).
* ``Feature``      <- meta_arch/igev_stereo/extractor.py:327-361.  The reference builds it from
  ``timm.create_model('mobilenetv2_100')`` (timm 0.5.4, not available offline); the MobileNetV2
  below is written out with timm's attribute names (``conv_stem``, ``bn1``, ``conv_pw``,
  ``conv_dw``, ``conv_pwl``, ``bn1..3``) so the state-dict keys match.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# conv + norm + LeakyReLU in 2-D / 3-D, conv or transposed conv
# ---------------------------------------------------------------------------------------------
class ConvNormAct(nn.Module):
    def __init__(self, cin: int, cout: int, norm: str = "bn", use_norm: bool = True, act: bool = True,
                 deconv: bool = False, is_3d: bool = False, **conv_kw):
        super().__init__()
        assert norm in ("bn", "in")
        conv_cls = {(False, False): nn.Conv2d, (False, True): nn.ConvTranspose2d,
                    (True, False): nn.Conv3d, (True, True): nn.ConvTranspose3d}[(is_3d, deconv)]
        self.conv = conv_cls(cin, cout, bias=False, **conv_kw)
        if norm == "bn":
            self.bn = (nn.BatchNorm3d if is_3d else nn.BatchNorm2d)(cout)
        else:
            self.IN = (nn.InstanceNorm3d if is_3d else nn.InstanceNorm2d)(cout)
        self._norm_name = "bn" if norm == "bn" else "IN"
        self.use_norm = use_norm
        self.act = act

    def forward(self, x):
        x = self.conv(x)
        if self.use_norm:
            x = getattr(self, self._norm_name)(x)
        return F.leaky_relu(x, 0.01) if self.act else x


class UpFuse(nn.Module):
    """Strided (de)conv to the skip tensor's resolution, concatenate with it, 3x3 conv."""

    def __init__(self, cin: int, cout: int, deconv: bool = False, norm: str = "bn", keep_concat: bool = True):
        super().__init__()
        k = 4 if deconv else 3
        self.conv1 = ConvNormAct(cin, cout, norm, deconv=deconv, kernel_size=k, stride=2, padding=1)
        self.conv2 = ConvNormAct(cout * 2, cout * (2 if keep_concat else 1), norm, kernel_size=3, stride=1, padding=1)

    def forward(self, x, skip):
        x = self.conv1(x)
        if x.shape != skip.shape:
            x = F.interpolate(x, size=skip.shape[-2:], mode="nearest")
        return self.conv2(torch.cat((x, skip), 1))


class FeatureAtt(nn.Module):
    """Channel attention of a cost volume from a 2-D feature map: cv * sigmoid(conv(feat))."""

    def __init__(self, cv_chan: int, feat_chan: int):
        super().__init__()
        self.feat_att = nn.Sequential(ConvNormAct(feat_chan, feat_chan // 2, kernel_size=1, stride=1, padding=0),
                                      nn.Conv2d(feat_chan // 2, cv_chan, 1))

    def forward(self, cv, feat):
        return torch.sigmoid(self.feat_att(feat).unsqueeze(2)) * cv


class Hourglass(nn.Module):
    def __init__(self, c: int):
        super().__init__()

        def down(cin, cout):
            return nn.Sequential(ConvNormAct(cin, cout, is_3d=True, kernel_size=3, padding=1, stride=2, dilation=1),
                                 ConvNormAct(cout, cout, is_3d=True, kernel_size=3, padding=1, stride=1, dilation=1))

        def up(cin, cout, **kw):
            return ConvNormAct(cin, cout, deconv=True, is_3d=True, kernel_size=(4, 4, 4), padding=(1, 1, 1),
                               stride=(2, 2, 2), **kw)

        def agg(cin, cout):
            return nn.Sequential(ConvNormAct(cin, cout, is_3d=True, kernel_size=1, padding=0, stride=1),
                                 ConvNormAct(cout, cout, is_3d=True, kernel_size=3, padding=1, stride=1),
                                 ConvNormAct(cout, cout, is_3d=True, kernel_size=3, padding=1, stride=1))

        self.conv1, self.conv2, self.conv3 = down(c, 2 * c), down(2 * c, 4 * c), down(4 * c, 6 * c)
        self.conv3_up, self.conv2_up = up(6 * c, 4 * c), up(4 * c, 2 * c)
        self.conv1_up = up(2 * c, 8, use_norm=False, act=False)
        self.agg_0, self.agg_1 = agg(8 * c, 4 * c), agg(4 * c, 2 * c)
        self.feature_att_8 = FeatureAtt(2 * c, 64)
        self.feature_att_16 = FeatureAtt(4 * c, 192)
        self.feature_att_32 = FeatureAtt(6 * c, 160)
        self.feature_att_up_16 = FeatureAtt(4 * c, 192)
        self.feature_att_up_8 = FeatureAtt(2 * c, 64)

    def forward_native(self, x, feats: Sequence[torch.Tensor]):
        """The same graph on libdkt's exact-fp32 3-D convolution kernels (csrc/igev_preloop.cu): every BasicConv is one
        launch with its eval-mode BatchNorm folded, LeakyReLU and the following FeatureAtt product in the epilogue; the
        two torch.cat are read in place by the 1x1x1 kernel.  Reference igev_stereo.py:66-89."""
        from . import ops

        def fold(m: ConvNormAct):
            if not m.use_norm:
                return None, None, (0.01 if m.act else 1.0)
            bn = m.bn
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            return scale, bn.bias - bn.running_mean * scale, (0.01 if m.act else 1.0)

        def k3(m, v, stride=1, att=None):
            sc, sh, sl = fold(m)
            return ops.conv3d_k3(v, m.conv.weight, sc, sh, sl, att, stride)

        def up(m, v):
            sc, sh, sl = fold(m)
            return ops.deconv3d_k4s2(v, m.conv.weight, sc, sh, sl)

        def k1(m, a, b):
            sc, sh, sl = fold(m)
            return ops.conv3d_k1(a, b, m.conv.weight, sc, sh, sl)

        # stride-1 3x3x3 layers with 32 / 48 channels on the tensor-core conv kernel (see _k3_tc): 405 -> 256 us and
        # 165 -> 88 us per layer at cfg3 (tools/hourglass_tc_probe.py, profiles/r3j_hourglass_tc_probe.txt).  The
        # 16-channel layers stay on the exact-fp32 kernel: a 16-channel source fills a quarter of a 64-channel K block and
        # N = 16 a sixteenth of the MMA width (1105 us against 628); so do the strided and the 8-channel layers, the
        # transposed convs and the 1x1x1 convs.
        tc = self.tc_ok()

        def k3s(m, v, att=None):
            return self._k3_tc(m, v, att) if (tc and m.conv.in_channels >= 32) else k3(m, v, 1, att)

        att = lambda fa, f: fa.feat_att(f).float()
        c1 = k3s(self.conv1[1], k3(self.conv1[0], x, 2), att=att(self.feature_att_8, feats[1]))
        c2 = k3s(self.conv2[1], k3(self.conv2[0], c1, 2), att=att(self.feature_att_16, feats[2]))
        c3 = k3s(self.conv3[1], k3(self.conv3[0], c2, 2), att=att(self.feature_att_32, feats[3]))
        a = k3s(self.agg_0[1], k1(self.agg_0[0], up(self.conv3_up, c3), c2))
        c2 = k3s(self.agg_0[2], a, att=att(self.feature_att_up_16, feats[2]))
        a = k3s(self.agg_1[1], k1(self.agg_1[0], up(self.conv2_up, c2), c1))
        c1 = k3s(self.agg_1[2], a, att=att(self.feature_att_up_8, feats[1]))
        return up(self.conv1_up, c1)

    # ---- 3x3x3 convolutions on tcgen05 ------------------------------------------------------------------------------
    @staticmethod
    def tc_ok() -> bool:
        import os
        from . import _lib as L
        return os.environ.get("DKT_HOURGLASS_TC", "1") == "1" and L.split_dtype() == torch.float16

    def _k3_tc(self, m: "ConvNormAct", v: torch.Tensor, att=None) -> torch.Tensor:
        """A stride-1 3x3x3 BasicConv (eval-mode BatchNorm folded, LeakyReLU; reference submodule.py:10-36) as ONE launch of
        the 2-D tensor-core conv (csrc/conv_tc.cu): the depth planes of a depth-padded NDHWC copy of the volume are the
        images of the batch, and the three kz taps are three channel-concatenated SOURCES -- the planes d-1, d, d+1, i.e. the
        same buffer at three plane offsets.  K = 3 x 9 x CI with 16-bit (hi, lo) operands, 3 MMAs per K step (22 mantissa
        bits: 1e-6 relative against the exact-fp32 kernel).  The attention product of FeatureAtt (submodule.py:227-240) rides
        on the layout pass on the way out (dkt_ndhwc_pad_to_ncdhw).  v (B,CI,D,H,W) fp32 -> (B,CO,D,H,W) fp32."""
        from . import _lib as L, ops
        B, CI, D, H, W = v.shape
        CO = m.conv.out_channels
        cache = self.__dict__.setdefault("_tc_cache", {})
        wkey = ("w", id(m))
        sig = tuple((t.data_ptr(), t._version) for t in (m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var))
        if wkey not in cache or cache[wkey][0] != sig:
            # (CO, CI, kz, ky, kx) -> 2-D filter (CO, kz*CI + ci, ky, kx): source s of the conv carries the taps kz = s
            w2 = m.conv.weight.detach().permute(0, 2, 1, 3, 4).reshape(CO, 3 * CI, 3, 3)
            cache[wkey] = (sig, ops.pack_conv_general(w2, None, bn=(m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var),
                                                      bn_eps=m.bn.eps))
        pack = cache[wkey][1]
        bkey = ("b", B, CI, CO, D, H, W, str(v.device))
        if bkey not in cache:
            for k in [k for k in cache if k[0] == "b" and k[2:4] == (CI, CO)]:      # a new input shape replaces the old buffers
                del cache[k]
            dt = L.split_dtype()
            cache[bkey] = (torch.zeros(B, D + 2, H, W, CI, device=v.device, dtype=dt),      # zero planes at d = -1 and d = D
                           torch.zeros(B, D + 2, H, W, CI, device=v.device, dtype=dt),
                           torch.empty(B, D + 2, H, W, CO, device=v.device, dtype=torch.float32))
        xh, xl, of = cache[bkey]
        lib = L.load()
        v = v.contiguous()
        L.check(lib.dkt_ncdhw_to_ndhwc_pad(v.data_ptr(), xh.data_ptr(), xl.data_ptr(), B, CI, D, H, W, L.stream_ptr()),
                "ncdhw_to_ndhwc_pad")
        ni = B * (D + 2) - 2         # images: every plane but the first and the last (inter-sample planes: results unused)
        fh, fl, fo = xh.view(B * (D + 2), H, W, CI), xl.view(B * (D + 2), H, W, CI), of.view(B * (D + 2), H, W, CO)
        srcs = [L.tensor_slice(None, fh[sft:sft + ni], fl[sft:sft + ni], 0, CI) for sft in range(3)]
        epi = ops.make_epilogue(L.EPI_LINEAR, L.tensor_slice(fo[1:1 + ni], None, None, 0, CO),
                                act=L.ACT_LEAKY if m.act else L.ACT_NONE, bias=pack.bias)
        ops.conv2d_ex(srcs, pack, epi, ni, H, W)
        out = torch.empty(B, CO, D, H, W, device=v.device, dtype=torch.float32)
        att = att.contiguous() if att is not None else None
        L.check(lib.dkt_ndhwc_pad_to_ncdhw(of.data_ptr(), L.ptr(att), out.data_ptr(), B, CO, D, H, W, L.stream_ptr()),
                "ndhwc_pad_to_ncdhw")
        return out

    def native_ok(self, x: torch.Tensor) -> bool:
        """Even sizes down the three stride-2 levels (the deconvolutions then return exactly the skip shapes)."""
        return x.is_cuda and x.dtype == torch.float32 and not self.training and all(v % 8 == 0 for v in x.shape[2:])

    def forward(self, x, feats: Sequence[torch.Tensor]):
        c1 = self.feature_att_8(self.conv1(x), feats[1])
        c2 = self.feature_att_16(self.conv2(c1), feats[2])
        c3 = self.feature_att_32(self.conv3(c2), feats[3])
        c2 = self.feature_att_up_16(self.agg_0(torch.cat((self.conv3_up(c3), c2), 1)), feats[2])
        c1 = self.feature_att_up_8(self.agg_1(torch.cat((self.conv2_up(c2), c1), 1)), feats[1])
        return self.conv1_up(c1)


# ---------------------------------------------------------------------------------------------
# MobileNetV2-1.0 feature pyramid with timm's module / parameter names
# ---------------------------------------------------------------------------------------------
class _DSConv(nn.Module):
    """timm DepthwiseSeparableConv: 3x3 depthwise + BN + ReLU6, 1x1 project + BN."""

    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, 3, stride, 1, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.act1 = nn.ReLU6(inplace=True)
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = self.bn2(self.conv_pw(self.act1(self.bn1(self.conv_dw(x)))))
        return x + y if self.skip else y


class _InvertedResidual(nn.Module):
    """timm InvertedResidual: 1x1 expand, 3x3 depthwise, 1x1 linear projection."""

    def __init__(self, cin: int, cout: int, stride: int, expand: int = 6):
        super().__init__()
        mid = cin * expand
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.act1 = nn.ReLU6(inplace=True)
        self.conv_dw = nn.Conv2d(mid, mid, 3, stride, 1, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.act2 = nn.ReLU6(inplace=True)
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = self.act1(self.bn1(self.conv_pw(x)))
        y = self.act2(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(y))
        return x + y if self.skip else y


# (repeats, out channels, first stride) of MobileNetV2-1.0 stages 1..5 (stage 0 is the DS conv, stage 6 unused)
_MBV2_STAGES = ((2, 24, 2), (3, 32, 2), (4, 64, 2), (3, 96, 1), (3, 160, 2))


def _mbv2_stage(cin: int, reps: int, cout: int, stride: int) -> nn.Sequential:
    return nn.Sequential(*[_InvertedResidual(cin if i == 0 else cout, cout, stride if i == 0 else 1) for i in range(reps)])


class Feature(nn.Module):
    """1/4, 1/8, 1/16, 1/32 features: MobileNetV2 encoder + U-Net style decoder with instance norm."""

    def __init__(self):
        super().__init__()
        chans = [16, 24, 32, 96, 160]
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.act1 = nn.ReLU6(inplace=True)
        stages = [nn.Sequential(_DSConv(32, 16, 1))]
        cin = 16
        for reps, cout, stride in _MBV2_STAGES:
            stages.append(_mbv2_stage(cin, reps, cout, stride))
            cin = cout
        # the reference regroups timm's stage list as blocks[0:1], [1:2], [2:3], [3:5], [5:6]
        self.block0 = nn.Sequential(stages[0])
        self.block1 = nn.Sequential(stages[1])
        self.block2 = nn.Sequential(stages[2])
        self.block3 = nn.Sequential(stages[3], stages[4])
        self.block4 = nn.Sequential(stages[5])
        self.deconv32_16 = UpFuse(chans[4], chans[3], deconv=True, norm="in")
        self.deconv16_8 = UpFuse(chans[3] * 2, chans[2], deconv=True, norm="in")
        self.deconv8_4 = UpFuse(chans[2] * 2, chans[1], deconv=True, norm="in")
        self.conv4 = ConvNormAct(chans[1] * 2, chans[1] * 2, "in", kernel_size=3, stride=1, padding=1)

    def forward(self, x) -> List[torch.Tensor]:
        if self.native_encoder_ok(x):
            x4, x8, x16, x32 = self.encode_native(x)
        else:
            x2 = self.block0(self.act1(self.bn1(self.conv_stem(x))))
            x4 = self.block1(x2)
            x8 = self.block2(x4)
            x16 = self.block3(x8)
            x32 = self.block4(x16)
        if self.native_decoder_ok(x4, x8, x16, x32):
            return self.decode_native(x4, x8, x16, x32)
        x16 = self.deconv32_16(x32, x16)
        x8 = self.deconv16_8(x16, x8)
        x4 = self.conv4(self.deconv8_4(x8, x4))
        return [x4, x8, x16, x32]

    # ---- MobileNetV2 encoder (reference meta_arch/igev_stereo/extractor.py:331-361) on the library's kernels --------------
    def native_encoder_ok(self, x) -> bool:
        import os
        from . import _lib as L
        return (os.environ.get("DKT_NATIVE_MBV2", "1") == "1" and not self.training and x.is_cuda and x.dtype == torch.float32
                and os.environ.get("DKT_IMPL", "tc") == "tc" and L.split_dtype() == torch.float16
                and x.shape[-2] % 32 == 0 and x.shape[-1] % 32 == 0)

    def encode_native(self, x):
        """block0 .. block4 after the 3-channel stem conv (which stays on cuDNN with its BatchNorm + ReLU6): every 1x1 conv
        is a tensor-core conv over NHWC 16-bit pairs with its eval-mode BatchNorm folded into weights and bias, every
        depthwise conv one pass of dkt_dwconv3x3 (BatchNorm folded, ReLU6), the linear-bottleneck residual rides on the
        project conv's epilogue as its additive per-pixel term.  ReLU6 of an expand conv = ReLU in its epilogue + min(., 6) on
        the depthwise kernel's loads.  24-channel maps live in 32-channel buffers (zero upper channels, zero weight columns);
        expand convs wider than 256 channels run as several launches.  -> x4 (24), x8 (32), x16 (96), x32 (160), NCHW fp32."""
        from . import _lib as L, ops
        TS, E = L.tensor_slice, ops.make_epilogue
        lib = L.load()
        dev, Bt = x.device, x.shape[0]
        cache = self.__dict__.setdefault("_enc_cache", {})
        blocks = [b for stage in (self.block0, self.block1, self.block2, self.block3, self.block4) for seq in stage for b in seq]
        sig = tuple((p.data_ptr(), p._version) for b in blocks for p in list(b.parameters()) + list(b.buffers()))
        pad16 = lambda c: (c + 15) // 16 * 16                  # noqa: E731

        def bn4(bn):
            return (bn.weight, bn.bias, bn.running_mean, bn.running_var)

        def pack_dw(conv, bn):
            s_ = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().double()
            wd = (conv.weight.detach().double()[:, 0] * s_.view(-1, 1, 1)).permute(1, 2, 0).reshape(9, -1)     # [tap][C]
            bd = (bn.bias.detach().double() - bn.running_mean.detach().double() * s_)
            return wd.float().contiguous(), bd.float().contiguous()

        def pack_pw(conv, bn, cin_pad):
            w_ = conv.weight.detach()
            N = w_.shape[0]
            chunks = [(0, N)] if N <= 256 else [(i, min(i + 192, N)) for i in range(0, N, 192)] if N % 192 == 0 else \
                [(i, min(i + 240, N)) for i in range(0, N, 240)]
            g, bta, mu, var = bn4(bn)
            return [(c0, c1, ops.pack_conv_general(w_[c0:c1], None, cin_pad=cin_pad, bn=(g[c0:c1], bta[c0:c1], mu[c0:c1], var[c0:c1]),
                                                   bn_eps=bn.eps)) for c0, c1 in chunks]

        if cache.get("sig") != sig:
            packs = []
            for b in blocks:
                if isinstance(b, _DSConv):
                    packs.append(dict(dw=pack_dw(b.conv_dw, b.bn1), pwl=pack_pw(b.conv_pw, b.bn2, pad16(b.conv_pw.in_channels))))
                else:
                    packs.append(dict(pw=pack_pw(b.conv_pw, b.bn1, pad16(b.conv_pw.in_channels)), dw=pack_dw(b.conv_dw, b.bn2),
                                      pwl=pack_pw(b.conv_pwl, b.bn3, pad16(b.conv_pwl.in_channels))))
            # block0's projection (32 -> 16, linear: BatchNorm, no activation, no skip) is consumed by block1's first expand
            # conv alone: the two 1x1 convs are ONE 32 -> 96 conv, composed in fp64 here -- the 16-channel map at half
            # resolution (a quarter K block for the conv that read it: 0.72 ms) never exists
            b0, b1 = blocks[0], blocks[1]
            if isinstance(b0, _DSConv) and not b0.skip and not b1.skip:
                def fold(conv, bn):
                    s_ = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().double()
                    return conv.weight.detach().double()[:, :, 0, 0] * s_.view(-1, 1), (bn.bias.detach().double() - bn.running_mean.detach().double() * s_)
                w1, c1 = fold(b0.conv_pw, b0.bn2)
                w2, c2 = fold(b1.conv_pw, b1.bn1)
                wc, cc = (w2 @ w1).float(), (w2 @ c1 + c2).float()
                packs[1]["pw"] = [(0, wc.shape[0], ops.pack_conv_general(wc.view(wc.shape[0], wc.shape[1], 1, 1), cc))]
                packs[0]["fused_next"] = True
            cache.clear()
            cache.update(sig=sig, packs=packs)
        packs = cache["packs"]
        dt = L.split_dtype()
        if cache.get("shape") != (tuple(x.shape), str(dev)):              # one input shape's buffers at a time (datasets
            for k in [k for k in cache if isinstance(k, tuple)]:          # with many image sizes must not pile them up)
                del cache[k]
            cache["shape"] = (tuple(x.shape), str(dev))

        def buf(key, *shape, d=torch.float32):
            k = (key,) + shape + (str(dev), d)
            if k not in cache:
                cache[k] = torch.zeros(*shape, device=dev, dtype=d)
            return cache[k]

        x2 = self.act1(self.bn1(self.conv_stem(x)))                       # (Bt,32,H/2,W/2) fp32, cuDNN
        h, w = x2.shape[-2:]
        cur_f = buf("in", Bt, h, w, 32)
        ops.nchw_to_nhwc(x2, TS(cur_f, None, None, 0, 32))
        cur = dict(f=cur_f, hi=None, lo=None, C=32)                      # running map: fp32 (+ 16-bit pair) NHWC
        outs, bi = [], 0
        for si, stage in enumerate((self.block0, self.block1, self.block2, self.block3, self.block4)):
            for seq in stage:
                for b in seq:
                    pk = packs[bi]
                    tag = f"b{bi}"
                    bi += 1
                    stride = b.conv_dw.stride[0]
                    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
                    if isinstance(b, _DSConv):
                        mid, mid_f, in_max = b.conv_dw.in_channels, cur["f"], 3.0e38
                    else:
                        mid = b.conv_pw.out_channels
                        mid_f = buf(tag + "m", Bt, h, w, mid)
                        for c0, c1, pw in pk["pw"]:                       # expand 1x1 + BN + ReLU (the min(., 6) is the dw's)
                            ops.conv2d_ex([TS(None, cur["hi"], cur["lo"], 0, cur["C"])], pw,
                                          E(L.EPI_LINEAR, TS(mid_f, None, None, c0, c1 - c0), act=L.ACT_RELU, bias=pw.bias), Bt, h, w)
                        in_max = 6.0
                    dh, dl = buf(tag + "dh", Bt, ho, wo, mid, d=dt), buf(tag + "dl", Bt, ho, wo, mid, d=dt)
                    wd, bd = pk["dw"]
                    src_t, dst_t = TS(mid_f, None, None, 0, mid), TS(None, dh, dl, 0, mid)
                    L.check(lib.dkt_dwconv3x3(L.C.byref(src_t), wd.data_ptr(), bd.data_ptr(), in_max, 0.0, 6.0, L.C.byref(dst_t),
                                              Bt, h, w, stride, L.stream_ptr()), "dwconv3x3")
                    if pk.get("fused_next"):                          # projection composed into the next block's expand conv
                        cur = dict(f=None, hi=dh, lo=dl, C=mid)
                        h, w = ho, wo
                        continue
                    pwl_conv = b.conv_pw if isinstance(b, _DSConv) else b.conv_pwl
                    co = pwl_conv.out_channels
                    cop = pad16(co)
                    of, oh, ol = buf(tag + "of", Bt, ho, wo, cop), buf(tag + "oh", Bt, ho, wo, cop, d=dt), buf(tag + "ol", Bt, ho, wo, cop, d=dt)
                    (_, _, pwl), = pk["pwl"]
                    ops.conv2d_ex([TS(None, dh, dl, 0, mid)], pwl,
                                  E(L.EPI_LINEAR, TS(of, oh, ol, 0, co), bias=pwl.bias, ctx=cur["f"] if b.skip else None), Bt, ho, wo)
                    cur = dict(f=of, hi=oh, lo=ol, C=cop)
                    h, w = ho, wo
            if si >= 1:                                                   # block1..4 -> x4, x8, x16, x32
                outs.append(ops.nhwc_to_nchw(TS(cur["f"], None, None, 0, co), Bt, h, w, dev))
        return outs

    # ---- U-Net decoder (FeatUp, reference meta_arch/igev_stereo/extractor.py:296-325) on the library's kernels ----------
    def native_decoder_ok(self, x4, x8, x16, x32) -> bool:
        import os
        from . import _lib as L
        if os.environ.get("DKT_NATIVE_FEATUP", "1") != "1" or self.training or not x4.is_cuda or x4.dtype != torch.float32:
            return False
        if os.environ.get("DKT_IMPL", "tc") != "tc" or L.split_dtype() != torch.float16:
            return False
        return all(tuple(a.shape[-2:]) == (2 * b.shape[-2], 2 * b.shape[-1]) for a, b in ((x4, x8), (x8, x16), (x16, x32)))

    def decode_native(self, x4, x8, x16, x32) -> List[torch.Tensor]:
        """Three UpFuse steps + conv4: every transposed conv (kernel 4, stride 2, padding 1) is a 3x3 tensor-core conv that
        produces the four output parities as channel groups (ops.deconv4x4s2_as_conv3x3) + a pixel shuffle; InstanceNorm +
        LeakyReLU through the library's statistics / apply kernels; ``torch.cat`` never materialises (the skip tensor and the
        upsampled tensor are written into channel slices of the next conv's source).  NHWC 16-bit pairs in between; the four
        returned maps are fp32 NCHW like the module's."""
        from . import _lib as L, ops
        TS, E = L.tensor_slice, ops.make_epilogue
        dev, Bt = x4.device, x4.shape[0]
        steps = (("deconv32_16", x32, x16), ("deconv16_8", None, x8), ("deconv8_4", None, x4))
        cache = self.__dict__.setdefault("_dec_cache", {})
        mods = [getattr(self, n) for n, _, _ in steps]
        sig = tuple((p.data_ptr(), p._version) for m in mods + [self.conv4] for p in m.parameters())
        if cache.get("sig") != sig:
            w = {}
            for (name, _, _), m in zip(steps, mods):
                w3 = ops.deconv4x4s2_as_conv3x3(m.conv1.conv.weight.detach())          # (4*CO, CI, 3, 3)
                co4 = w3.shape[0]
                nl = 2 if co4 > 256 else 1                                             # N <= 256 per launch
                w[name + ".d"] = [ops.pack_conv_general(w3[i * co4 // nl:(i + 1) * co4 // nl], None) for i in range(nl)]
                w[name + ".c"] = ops.pack_conv_general(m.conv2.conv.weight.detach(), None)
            w["conv4"] = ops.pack_conv_general(self.conv4.conv.weight.detach(), None)
            cache.clear()
            cache.update(sig=sig, w=w)
        w = cache["w"]
        dt = L.split_dtype()
        if cache.get("shape") != (tuple(x4.shape), str(dev)):
            for k in [k for k in cache if isinstance(k, tuple)]:
                del cache[k]
            cache["shape"] = (tuple(x4.shape), str(dev))

        def buf(key, *shape, d=torch.float32):
            k = (key,) + shape + (str(dev), d)
            if k not in cache:
                cache[k] = torch.zeros(*shape, device=dev, dtype=d)
            return cache[k]

        def inorm(raw, Cc, h, wd, dst, tag):
            st = buf("st" + tag, Bt, Cc, 2)
            ops.instnorm_stats(TS(raw, None, None, 0, Cc), buf("ws" + tag, ops.instnorm_workspace(Bt, Cc, dev).numel()), st, Bt, h, wd)
            ops.instnorm_apply(TS(raw, None, None, 0, Cc), st, dst, Bt, h, wd, relu="leaky")

        outs = []
        cur = None                                   # (hi, lo) NHWC of the running coarse map
        for (name, first, skip), m in zip(steps, mods):
            CO = m.conv1.conv.out_channels
            hs, ws_ = skip.shape[-2:]
            hc, wc = hs // 2, ws_ // 2
            if cur is None:                          # coarsest input comes from MobileNetV2 as NCHW fp32
                CI = first.shape[1]
                cur = (buf(name + "ih", Bt, hc, wc, CI, d=dt), buf(name + "il", Bt, hc, wc, CI, d=dt))
                ops.nchw_to_nhwc(first, TS(None, cur[0], cur[1], 0, CI))
            CI = cur[0].shape[-1]
            # transposed conv as a parity-grouped 3x3 conv, shuffled to the skip's resolution
            u = buf(name + "u", Bt, hc, wc, 4 * CO)
            packs = w[name + ".d"]
            for i, pk in enumerate(packs):
                n_i = 4 * CO // len(packs)
                ops.conv2d_ex([TS(None, cur[0], cur[1], 0, CI)], pk, E(L.EPI_LINEAR, TS(u, None, None, i * n_i, n_i)), Bt, hc, wc)
            up_raw = buf(name + "ur", Bt, hs, ws_, CO)
            ops.pixel_shuffle2(u, CO, TS(up_raw, None, None, 0, CO), Bt, hc, wc)
            cat_h, cat_l = buf(name + "ch", Bt, hs, ws_, 2 * CO, d=dt), buf(name + "cl", Bt, hs, ws_, 2 * CO, d=dt)
            inorm(up_raw, CO, hs, ws_, TS(None, cat_h, cat_l, 0, CO), name + "a")              # cat((x, skip), 1): x first
            ops.nchw_to_nhwc(skip, TS(None, cat_h, cat_l, CO, CO))
            # conv2: 3x3 over the concatenation, InstanceNorm, LeakyReLU
            keep = name != "deconv8_4"               # deconv8_4.conv2 keeps 2*CO channels -> conv4 follows; others are outputs
            C2 = m.conv2.conv.out_channels
            raw = buf(name + "r", Bt, hs, ws_, C2)
            ops.conv2d_ex([TS(None, cat_h, cat_l, 0, 2 * CO)], w[name + ".c"], E(L.EPI_LINEAR, TS(raw, None, None, 0, C2)), Bt, hs, ws_)
            o_f = buf(name + "of", Bt, hs, ws_, C2) if keep else None
            o_h, o_l = buf(name + "oh", Bt, hs, ws_, C2, d=dt), buf(name + "ol", Bt, hs, ws_, C2, d=dt)
            inorm(raw, C2, hs, ws_, TS(o_f, o_h, o_l, 0, C2), name + "b")
            if keep:
                outs.append(ops.nhwc_to_nchw(TS(o_f, None, None, 0, C2), Bt, hs, ws_, dev))
            cur = (o_h, o_l)
        # conv4 on the 1/4 map
        C4 = self.conv4.conv.out_channels
        h4, w4 = x4.shape[-2:]
        raw = buf("c4r", Bt, h4, w4, C4)
        ops.conv2d_ex([TS(None, cur[0], cur[1], 0, cur[0].shape[-1])], w["conv4"], E(L.EPI_LINEAR, TS(raw, None, None, 0, C4)), Bt, h4, w4)
        o4 = buf("c4o", Bt, h4, w4, C4)
        inorm(raw, C4, h4, w4, TS(o4, None, None, 0, C4), "c4")
        x4o = ops.nhwc_to_nchw(TS(o4, None, None, 0, C4), Bt, h4, w4, dev)
        x16o, x8o = outs
        return [x4o, x8o, x16o, x32]


# ---------------------------------------------------------------------------------------------
# volume helpers
# ---------------------------------------------------------------------------------------------
def build_gwc_volume(ref: torch.Tensor, tgt: torch.Tensor, maxdisp: int, groups: int) -> torch.Tensor:
    """Group-wise correlation volume (B, groups, maxdisp, H, W): for disparity d, the mean over each
    channel group of ref[..., x] * tgt[..., x - d], zero where x < d (reference submodule.py:152-170).
    One fused expression per disparity on views (no Python-side masking)."""
    B, C, H, W = ref.shape
    cpg = C // groups
    vol = ref.new_zeros(B, groups, maxdisp, H, W)
    r = ref.view(B, groups, cpg, H, W)
    t = tgt.view(B, groups, cpg, H, W)
    for d in range(min(maxdisp, W)):
        vol[:, :, d, :, d:] = (r[..., d:] * t[..., : W - d]).mean(dim=2)
    return vol


def disparity_regression(prob: torch.Tensor, maxdisp: int) -> torch.Tensor:
    """Soft-argmin: sum_d d * p(d) (reference submodule.py:220-224)."""
    d = torch.arange(maxdisp, dtype=prob.dtype, device=prob.device).view(1, maxdisp, 1, 1)
    return (prob * d).sum(1, keepdim=True)
