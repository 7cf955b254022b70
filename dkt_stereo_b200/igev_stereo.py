"""IGEV-Stereo with the hot path (Combined Geometry Encoding Volume, lookup, GRU loop, learned
upsampling) on B200 kernels.

Drop-in for reference meta_arch/igev_stereo/igev_stereo.py: same constructor (``IGEVStereo(args)``),
``forward(image1, image2, iters, flow_init, test_mode)`` and state-dict keys.  Only the inference
contract (``test_mode=True`` -> ``(None, disp_up)``, disp_up = -disparity, reference :215-220) is
served.

What runs where:
  PyTorch (cuDNN)  : the two 3-channel stem convs (MobileNetV2's conv_stem, stem_2's first conv), the 2-D attention convs
                     of FeatureAtt (reference igev_stereo.py:154-168)
  libdkt kernels   : the feature pyramid (MobileNetV2 encoder + UpFuse x3 + conv4), the rest of stem_2 / stem_4, the matching-feature
                     head conv + desc (tensor-core convs, InstanceNorm kernels), cnet + context convs (EncoderEngine),
                     GWC volume, corr_stem (3-D conv + BN + LeakyReLU + feature attention), the 3-D hourglass (32 / 48-channel
                     stride-1 layers on tcgen05), classifier + soft-argmin init disparity (igev_preloop.cu), upsample_disp,
                     all-pairs init-corr pyramid (K1, scale 1), geometry-volume pyramid
                     (dkt_geo_pool), per-iteration combined lookup + `disp += delta` + convc1
                     (dkt_geo_lookup_enc), motion encoder + 3 ConvGRUs + disp head (K3),
                     mask_feat_4 head, context_upsample (K4)   (reference igev_stereo.py:192-216)
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops
from .extractor import MultiBasicEncoder
from .igev_modules import ConvNormAct, Feature, FeatureAtt, Hourglass, UpFuse, build_gwc_volume, disparity_regression
from .raft_stereo import _fp32_math
from .update import BasicMultiUpdateBlock, UpdateEngine


class IGEVStereo(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        hd = args.hidden_dims
        self.impl = os.environ.get("DKT_IMPL") or {"b200_fp32": "simt", "b200_simt": "simt"}.get(
            getattr(args, "corr_implementation", "b200"), "tc")
        self.cnet = MultiBasicEncoder(output_dim=[hd, hd], norm_fn="batch", downsample=args.n_downsample,
                                      head_names=("outputs04", "outputs08", "outputs16"))
        self.update_block = BasicMultiUpdateBlock(args, hidden_dims=hd, igev=True)
        self.context_zqr_convs = nn.ModuleList([nn.Conv2d(hd[i], hd[i] * 3, 3, padding=1) for i in range(args.n_gru_layers)])
        self.feature = Feature()

        def stem(cin, cout):
            return nn.Sequential(ConvNormAct(cin, cout, "in", kernel_size=3, stride=2, padding=1),
                                 nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.InstanceNorm2d(cout), nn.ReLU())

        self.stem_2 = stem(3, 32)
        self.stem_4 = stem(32, 48)
        self.spx = nn.Sequential(nn.ConvTranspose2d(2 * 32, 9, kernel_size=4, stride=2, padding=1))
        self.spx_2 = UpFuse(24, 32, deconv=True, norm="in")
        self.spx_4 = nn.Sequential(ConvNormAct(96, 24, "in", kernel_size=3, stride=1, padding=1),
                                   nn.Conv2d(24, 24, 3, 1, 1, bias=False), nn.InstanceNorm2d(24), nn.ReLU())
        self.spx_2_gru = UpFuse(32, 32, deconv=True, norm="bn")
        self.spx_gru = nn.Sequential(nn.ConvTranspose2d(2 * 32, 9, kernel_size=4, stride=2, padding=1))
        self.conv = ConvNormAct(96, 96, "in", kernel_size=3, padding=1, stride=1)
        self.desc = nn.Conv2d(96, 96, kernel_size=1, padding=0, stride=1)
        self.corr_stem = ConvNormAct(8, 8, is_3d=True, kernel_size=3, stride=1, padding=1)
        self.corr_feature_att = FeatureAtt(8, 96)
        self.cost_agg = Hourglass(8)
        self.classifier = nn.Conv3d(8, 1, 3, 1, 1, bias=False)

        self.engine = UpdateEngine(self.update_block, self.impl)
        # context encoder + context convs on the library's tensor-core conv kernel (same EncoderEngine as RAFT-Stereo's
        # cnet; 152 of the 235 ms the PyTorch pre-loop took at cfg3 were this fp32 cuDNN encoder)
        self.encoder = None
        if (os.environ.get("DKT_NATIVE_ENCODER", "1") == "1" and self.impl == "tc" and args.n_downsample == 2
                and not getattr(args, "mixed_precision", False)):
            from .encoder import EncoderEngine
            self.encoder = EncoderEngine(None, self.cnet, self.context_zqr_convs, self.engine)
        # volume stage of the pre-loop on libdkt kernels (GWC volume, corr_stem, classifier, soft-argmin)
        self.native_volume = os.environ.get("DKT_NATIVE_VOLUME", "1") == "1" and not getattr(args, "mixed_precision", False)
        self.use_cuda_graph = os.environ.get("DKT_CUDA_GRAPH", "1") == "1"
        self.extractor_fp32 = not getattr(args, "extractor_tf32", False)
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._seen = set()
        # whole-forward CUDA graphs (feature side + volume stage + context encoder + loop + upsampling), per input shape and
        # iteration count: one replay per call instead of ~1300 launches through Python.  On one GPU the device time is the
        # same (the host runs ahead of the GPU either way); with eight ranks on 16 host cores the eager form starves the
        # GPUs (r4z, 8 GPUs: 590 pairs/s end to end against 690 device-resident).  DKT_FULL_GRAPH=0: only the loop is a graph.
        self.full_graph = os.environ.get("DKT_FULL_GRAPH", "1") == "1"
        self._full: Dict[tuple, tuple] = {}
        self._full_seen = set()
        self._full_sig = None
        self._capturing = False
        self._vol = None
        self._vol_key = None
        # upsample_disp (spx_2_gru deconv + conv, spx_gru deconv, softmax, context_upsample) on the library's kernels
        self.native_upsample = (os.environ.get("DKT_NATIVE_UPSAMPLE", "1") == "1" and self.impl == "tc"
                                and not getattr(args, "mixed_precision", False))
        self._up_w = None
        self._up_sig = None
        self._up_buf = None
        # matching-feature head (conv 3x3 + InstanceNorm + LeakyReLU, desc 1x1) on the library's kernels
        self.native_match = (os.environ.get("DKT_NATIVE_MATCH", "1") == "1" and self.impl == "tc"
                             and not getattr(args, "mixed_precision", False))
        self._match_w = None
        self._match_buf = None
        self._stem_w = None
        self._stem_buf = None

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def invalidate_weights(self) -> None:
        """Force a repack of the engine's weights (and with it the re-capture of the CUDA graphs) on the next forward.
        `load_state_dict`, optimizer steps and `p.data = ...` assignments (the EMA teacher of reference
        tools/ft_dkt.py:179-181) are noticed automatically through (data_ptr, version) of every parameter; in-place
        edits THROUGH `.data` (`p.data.mul_(...)`) change neither and need this call."""
        self.engine._wsig = None
        if self.encoder is not None:
            self.encoder._sig = None

    # ---- pre-loop (PyTorch): reference igev_stereo.py:154-189 -------------------------------------
    def prepare(self, image1: torch.Tensor, image2: torch.Tensor):
        """-> match_left, match_right (B,96,h,w), geo_encoding_volume (B,8,D,h,w), init_disp (B,1,h,w),
        net_list, ctx_list (cz|cr|cq concatenated per scale), stem_2x (B,32,H/2,W/2)."""
        args = self.args
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        with _fp32_math(self.extractor_fp32), torch.autocast("cuda", enabled=bool(getattr(args, "mixed_precision", False))):
            # left and right images go through the feature side as ONE batch (reference igev_stereo.py:154-168 runs the two
            # halves separately): BatchNorm is in eval mode and InstanceNorm is per sample, so the result is the same; the
            # 1/16 and 1/32-resolution layers are launch-bound at batch 8 and half as many launches are issued
            B = image1.shape[0]
            both = torch.cat((image1, image2), 0)
            feats = self.feature(both)
            if self.native_match and both.is_cuda and both.dtype == torch.float32:
                stem_2, stem_4 = self._stems_native(both)
            else:
                stem_2 = self.stem_2(both)
                stem_4 = self.stem_4(stem_2)
            stem_2x = stem_2[:B]
            feats[0] = torch.cat((feats[0], stem_4), 1)
            self._feat4 = feats[0][:B]            # left 1/4 features: input of spx_4 (test_mode=False only, reference :179)
            if self.native_match and both.is_cuda and feats[0].dtype == torch.float32:
                match = self._match_native(feats[0])
            else:
                match = self.desc(self.conv(feats[0]))
            match_left, match_right = match[:B], match[B:]
            fl = [f[:B] for f in feats]
            D = args.max_disp // 4
            native_vol = (self.native_volume and image1.is_cuda and match_left.dtype == torch.float32
                          and not self.corr_stem.bn.training)
            if native_vol:
                # libdkt (igev_preloop.cu): GWC volume; corr_stem's 3-D conv with its eval-mode BatchNorm folded, LeakyReLU
                # and the feature-attention product in the epilogue (reference igev_stereo.py:169-171)
                bn = self.corr_stem.bn
                scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                shift = bn.bias - bn.running_mean * scale
                att = self.corr_feature_att.feat_att(fl[0]).float()
                vol = ops.conv3d_c8(ops.gwc_volume(match_left, match_right, D, 8), self.corr_stem.conv.weight,
                                    scale, shift, 0.01, att)
            else:
                vol = self.corr_stem(build_gwc_volume(match_left, match_right, D, 8))
                vol = self.corr_feature_att(vol, fl[0])
            if native_vol and self.cost_agg.native_ok(vol):
                gev = self.cost_agg.forward_native(vol, fl)          # 3-D hourglass, reference igev_stereo.py:22-89,172
            else:
                gev = self.cost_agg(vol, fl)
            if native_vol and gev.dtype == torch.float32:
                # classifier conv + softmax + disparity regression (reference igev_stereo.py:175-176)
                init_disp = ops.softargmin(ops.conv3d_c8(gev, self.classifier.weight).squeeze(1))
            else:
                prob = F.softmax(self.classifier(gev).squeeze(1), dim=1)
                init_disp = disparity_regression(prob, D)
            if self.encoder is not None and image1.is_cuda:
                net_list = ctx_list = None       # written straight into the update engine's buffers by forward()
            else:
                cnet_list = self.cnet(image1, num_layers=args.n_gru_layers)
                net_list = [torch.tanh(x[0]).float() for x in cnet_list]
                ctx_list = [conv(torch.relu(x[1])).float() for x, conv in zip(cnet_list, self.context_zqr_convs)]
        return (match_left.float(), match_right.float(), gev.float(), init_disp.float(), net_list, ctx_list, stem_2x.float())

    def _stems_native(self, both: torch.Tensor):
        """stem_2 / stem_4 (reference igev_stereo.py:96-107,159-160): the first conv (3 -> 32, stride 2, + InstanceNorm +
        LeakyReLU) stays on cuDNN (3 input channels); everything after it -- conv 32 -> 32 + IN + ReLU, conv 32 -> 48 stride 2
        + IN + LeakyReLU, conv 48 -> 48 + IN + ReLU -- runs on the library's kernels in NHWC 16-bit pairs.
        -> (stem_2 (Bt,32,H/2,W/2), stem_4 (Bt,48,H/4,W/4)) fp32 NCHW."""
        TS, E = L.tensor_slice, ops.make_epilogue
        x2 = self.stem_2[0](both)
        Bt, _, h2, w2 = x2.shape
        c2b, c4a, c4b = self.stem_2[1], self.stem_4[0].conv, self.stem_4[1]
        sig = tuple((p.data_ptr(), p._version) for p in (c2b.weight, c4a.weight, c4b.weight))
        if self._stem_w is None or self._stem_w[0] != sig:
            self._stem_w = (sig, ops.pack_conv_general(c2b.weight, None), ops.pack_conv_general(c4a.weight, None, stride=2),
                            ops.pack_conv_general(c4b.weight, None))
        _, w2b, w4a, w4b = self._stem_w
        h4, w4 = (h2 - 1) // 2 + 1, (w2 - 1) // 2 + 1
        key = (Bt, h2, w2, str(both.device))
        if self._stem_buf is None or self._stem_buf[0] != key:
            dt, dev = L.split_dtype(), both.device
            z = lambda *sh, d=torch.float32: torch.zeros(*sh, device=dev, dtype=d)       # noqa: E731
            self._stem_buf = (key, dict(
                ah=z(Bt, h2, w2, 32, d=dt), al=z(Bt, h2, w2, 32, d=dt), raw2=z(Bt, h2, w2, 32),
                s2=z(Bt, h2, w2, 32), s2h=z(Bt, h2, w2, 32, d=dt), s2l=z(Bt, h2, w2, 32, d=dt),
                raw4=z(Bt, h4, w4, 48), th=z(Bt, h4, w4, 48, d=dt), tl=z(Bt, h4, w4, 48, d=dt), raw4b=z(Bt, h4, w4, 48), u=z(Bt, h4, w4, 48),
                part=z(Bt * ops.conv_tiles(h2, w2) * 2 * 32), tws=ops.instnorm_tiles_workspace(Bt, 32, dev),
                ws=ops.instnorm_workspace(Bt, 48, dev), st32=z(Bt, 32, 2), st48=z(Bt, 48, 2)))
        b = self._stem_buf[1]
        ops.nchw_to_nhwc(x2, TS(None, b["ah"], b["al"], 0, 32))
        # stem_2: conv 32 -> 32, InstanceNorm (statistics from the conv epilogue), ReLU
        ops.conv2d_ex([TS(None, b["ah"], b["al"], 0, 32)], w2b, E(L.EPI_LINEAR, TS(b["raw2"], None, None, 0, 32), stats_partial=b["part"]),
                      Bt, h2, w2)
        ops.instnorm_finalize_tiles(b["part"], b["tws"], b["st32"], Bt, 32, h2, w2)
        ops.instnorm_apply(TS(b["raw2"], None, None, 0, 32), b["st32"], TS(b["s2"], b["s2h"], b["s2l"], 0, 32), Bt, h2, w2, relu=True)
        stem_2 = ops.nhwc_to_nchw(TS(b["s2"], None, None, 0, 32), Bt, h2, w2, both.device)
        # stem_4: conv 32 -> 48 stride 2, InstanceNorm, LeakyReLU
        ops.conv2d_ex([TS(None, b["s2h"], b["s2l"], 0, 32)], w4a, E(L.EPI_LINEAR, TS(b["raw4"], None, None, 0, 48)), Bt, h2, w2)
        ops.instnorm_stats(TS(b["raw4"], None, None, 0, 48), b["ws"], b["st48"], Bt, h4, w4)
        ops.instnorm_apply(TS(b["raw4"], None, None, 0, 48), b["st48"], TS(None, b["th"], b["tl"], 0, 48), Bt, h4, w4, relu="leaky")
        #         conv 48 -> 48, InstanceNorm, ReLU
        ops.conv2d_ex([TS(None, b["th"], b["tl"], 0, 48)], w4b, E(L.EPI_LINEAR, TS(b["raw4b"], None, None, 0, 48)), Bt, h4, w4)
        ops.instnorm_stats(TS(b["raw4b"], None, None, 0, 48), b["ws"], b["st48"], Bt, h4, w4)
        ops.instnorm_apply(TS(b["raw4b"], None, None, 0, 48), b["st48"], TS(b["u"], None, None, 0, 48), Bt, h4, w4, relu=True)
        return stem_2, ops.nhwc_to_nchw(TS(b["u"], None, None, 0, 48), Bt, h4, w4, both.device)

    def _match_native(self, x: torch.Tensor) -> torch.Tensor:
        """Matching-feature head ``desc(conv(x))`` (reference igev_stereo.py:127-128,166-167: BasicConv_IN 3x3 96 -> 96 =
        conv + InstanceNorm + LeakyReLU, then a 1x1 conv with bias) on the library's kernels: NCHW fp32 -> NHWC 16-bit pairs,
        the 3x3 conv on tcgen05 (K = 96 = one and a half K blocks, N = 96) with the InstanceNorm statistics out of its epilogue,
        normalise + LeakyReLU, the 1x1 conv, back to NCHW for the volume kernels.  4.1 -> 0.9 ms at cfg3 (both images)."""
        Bt, Cc, h, w = x.shape
        c3, d1 = self.conv.conv, self.desc
        sig = tuple((p.data_ptr(), p._version) for p in (c3.weight, d1.weight, d1.bias))
        if self._match_w is None or self._match_w[0] != sig:
            self._match_w = (sig, ops.pack_conv_general(c3.weight, None), ops.pack_conv_general(d1.weight, d1.bias))
        _, w3, w1 = self._match_w
        key = (Bt, Cc, h, w, str(x.device))
        if self._match_buf is None or self._match_buf[0] != key:
            dt, dev = L.split_dtype(), x.device
            z = lambda *sh, d=torch.float32: torch.zeros(*sh, device=dev, dtype=d)       # noqa: E731
            self._match_buf = (key, dict(xh=z(Bt, h, w, Cc, d=dt), xl=z(Bt, h, w, Cc, d=dt), raw=z(Bt, h, w, Cc),
                                         yh=z(Bt, h, w, Cc, d=dt), yl=z(Bt, h, w, Cc, d=dt), o=z(Bt, h, w, Cc),
                                         part=z(Bt * ops.conv_tiles(h, w) * 2 * Cc), ws=ops.instnorm_tiles_workspace(Bt, Cc, dev),
                                         stats=z(Bt, Cc, 2)))
        b, TS, E = self._match_buf[1], L.tensor_slice, ops.make_epilogue
        ops.nchw_to_nhwc(x, TS(None, b["xh"], b["xl"], 0, Cc))
        ops.conv2d_ex([TS(None, b["xh"], b["xl"], 0, Cc)], w3,
                      E(L.EPI_LINEAR, TS(b["raw"], None, None, 0, Cc), stats_partial=b["part"]), Bt, h, w)
        ops.instnorm_finalize_tiles(b["part"], b["ws"], b["stats"], Bt, Cc, h, w)
        ops.instnorm_apply(TS(b["raw"], None, None, 0, Cc), b["stats"], TS(None, b["yh"], b["yl"], 0, Cc), Bt, h, w, relu="leaky")
        ops.conv2d_ex([TS(None, b["yh"], b["yl"], 0, Cc)], w1, E(L.EPI_LINEAR, TS(b["o"], None, None, 0, Cc), bias=w1.bias), Bt, h, w)
        return ops.nhwc_to_nchw(TS(b["o"], None, None, 0, Cc), Bt, h, w, x.device)

    # ---- hot path (B200 kernels): reference igev_stereo.py:192-216 --------------------------------
    def _lookup(self, eng: UpdateEngine) -> None:
        geo, init = self._vol
        r = self.args.corr_radius
        disp = eng.FLOW["f32"].view(eng.B, *eng.hw[0])            # (B,h,w,1) -> (B,h,w)
        if eng.lookup_tc:
            ops.geo_lookup_enc_tc(geo, init, disp, r, *eng.lookup_tc_w, eng.cor1_slice(), eng.lookup_tap_planes,
                                  delta=eng.DELTA["f32"])
            return
        if eng.fused_enc:
            ops.geo_lookup_enc(geo, init, disp, r, eng.weights["convc1"], eng.cor1_slice(), delta=eng.DELTA["f32"])
        else:
            ops.geo_lookup(geo, init, disp, r, eng.CORR["f32"], "nhwc", out_hi=eng.CORR["hi"], out_lo=eng.CORR["lo"],
                           delta=eng.DELTA["f32"])

    def _run_loop(self, iters: int) -> None:
        for _ in range(iters):
            self.engine.step(self._lookup, with_mask=False)

    # ---- upsample_disp on libdkt kernels ------------------------------------------------------------
    def _pack_upsample(self) -> None:
        """spx_2_gru.conv1 (ConvTranspose2d 32->32, k4 s2 p1, + BN + LeakyReLU), spx_2_gru.conv2 (3x3 64->64 + BN + LeakyReLU)
        and spx_gru (ConvTranspose2d 64->9, k4 s2 p1, bias) as three 3x3 tensor-core convs: a k4 s2 p1 transposed conv is a
        3x3 conv producing the four output parities as channel groups (ops.deconv4x4s2_as_conv3x3); eval-mode BatchNorm is
        folded into weights and bias."""
        mods = (self.spx_2_gru, self.spx_gru)
        sig = tuple((p.data_ptr(), p._version) for m in mods for p in list(m.parameters()) + list(m.buffers()))
        if self._up_w is not None and sig == self._up_sig:
            return
        c1, c2, dg = self.spx_2_gru.conv1, self.spx_2_gru.conv2, self.spx_gru[0]
        assert c1.conv.weight.shape == (32, 32, 4, 4) and c2.conv.weight.shape == (64, 64, 3, 3) and dg.weight.shape == (64, 9, 4, 4)

        def bn4(bn, rep):
            return tuple(t.detach().repeat(rep) for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))

        w = {}
        w["d1"] = ops.pack_conv_general(ops.deconv4x4s2_as_conv3x3(c1.conv.weight.detach()), None, cin_pad=64,
                                        bn=bn4(c1.bn, 4), bn_eps=c1.bn.eps)
        w["c2"] = ops.pack_conv_general(c2.conv.weight.detach(), None, bn=bn4(c2.bn, 1), bn_eps=c2.bn.eps)
        w["d3"] = ops.pack_conv_general(ops.deconv4x4s2_as_conv3x3(dg.weight.detach()), dg.bias.detach().repeat(4))
        self._up_w, self._up_sig = w, sig

    def _upsample_native(self, disp: torch.Tensor, stem_2x: torch.Tensor) -> torch.Tensor:
        """reference igev_stereo.py:140-148 on kernels; mask_feat_4 is read from the engine's MH buffer (NHWC, 64 channels of
        which 32 are real), where the mask head left it as a 16-bit (hi, lo) pair."""
        eng = self.engine
        B, (h, w), dev = eng.B, eng.hw[0], eng.device
        assert stem_2x.shape == (B, 32, 2 * h, 2 * w), "upsample_disp: stem_2x must be at twice the mask features' resolution"
        self._pack_upsample()
        key = (B, h, w, str(dev))
        if self._up_buf is None or self._up_buf[0] != key:
            dt = L.split_dtype()
            z = lambda *sh, d=torch.float32: torch.zeros(*sh, device=dev, dtype=d)       # noqa: E731
            self._up_buf = (key, dict(u1=z(B, h, w, 128), x2h=z(B, 2 * h, 2 * w, 64, d=dt), x2l=z(B, 2 * h, 2 * w, 64, d=dt),
                                      yh=z(B, 2 * h, 2 * w, 64, d=dt), yl=z(B, 2 * h, 2 * w, 64, d=dt), zl=z(B, 2 * h, 2 * w, 36)))
        b, Wt, TS, E = self._up_buf[1], self._up_w, L.tensor_slice, ops.make_epilogue
        mh = eng.MH
        # spx_2_gru.conv1: deconv + BN + LeakyReLU -> four parities as 4 x 32 channels at 1/4 resolution, then to 1/2
        ops.conv2d_ex([TS(None, mh["hi"], mh["lo"], 0, 64)], Wt["d1"],
                      E(L.EPI_LINEAR, TS(b["u1"], None, None, 0, 128), act=L.ACT_LEAKY, bias=Wt["d1"].bias), B, h, w)
        ops.pixel_shuffle2(b["u1"], 32, TS(None, b["x2h"], b["x2l"], 0, 32), B, h, w)
        ops.nchw_to_nhwc(stem_2x, TS(None, b["x2h"], b["x2l"], 32, 32))                  # torch.cat((x, rem), 1)
        # spx_2_gru.conv2: 3x3 64 -> 64 + BN + LeakyReLU at 1/2 resolution
        ops.conv2d_ex([TS(None, b["x2h"], b["x2l"], 0, 64)], Wt["c2"],
                      E(L.EPI_LINEAR, TS(None, b["yh"], b["yl"], 0, 64), act=L.ACT_LEAKY, bias=Wt["c2"].bias), B, 2 * h, 2 * w)
        # spx_gru: deconv 64 -> 9 (+ bias) as 4 x 9 logits per half-resolution pixel; softmax + context_upsample fused
        ops.conv2d_ex([TS(None, b["yh"], b["yl"], 0, 64)], Wt["d3"],
                      E(L.EPI_LINEAR, TS(b["zl"], None, None, 0, 36), bias=Wt["d3"].bias), B, 2 * h, 2 * w)
        return ops.context_upsample_logits(disp, b["zl"], in_scale=4.0, out_scale=-1.0)  # reference :216 negates

    def upsample_disp(self, disp: torch.Tensor, mask_feat_4: torch.Tensor, stem_2x: torch.Tensor) -> torch.Tensor:
        """reference igev_stereo.py:140-148.  disp (B,h,w) fp32, mask_feat_4 (B,32,h,w), stem_2x (B,32,2h,2w)
        -> (B,1,4h,4w) = -context_upsample(4 * disp, softmax(spx_gru(spx_2_gru(...))))."""
        if self.native_upsample and disp.is_cuda and not self.spx_2_gru.conv1.bn.training:
            return self._upsample_native(disp, stem_2x.float())
        with _fp32_math(self.extractor_fp32):
            spx = F.softmax(self.spx_gru(self.spx_2_gru(mask_feat_4, stem_2x)), 1)
        return ops.context_upsample(disp, spx, in_scale=4.0, out_scale=-1.0)      # reference :216 negates

    def hot_path(self, match_left, match_right, gev, init_disp, net_list, ctx_list, stem_2x, iters: int,
                 all_preds: bool = False):
        """-> disp_up (B,1,H,W), already negated like the reference's return value (``all_preds``: the list of every
        iteration's disp_up, reference igev_stereo.py:213-217 with test_mode=False)."""
        args, eng = self.args, self.engine
        L.require_device(match_left)
        assert args.corr_levels == 2, "IGEV configs use a 2-level geometry pyramid (configs/igev_stereo/base.json)"
        B, D, h, w = match_left.shape
        dev = match_left.device
        if eng.pack_weights():                    # weights changed: the captured loop graph points into the old packs
            self._graphs.clear()
            self._seen.clear()
        eng.allocate(B, h, w, dev)
        dc = eng.lookup_tc               # the tensor-core lookup reads the geometry volume as (B,h,w,D,C)
        key = (B, D, h, w, tuple(gev.shape), str(dev), dc)
        if self._vol_key != key:
            self._graphs.clear()
            self._seen.clear()
            self._vol_key = key
            self._init_pyr = ops.alloc_pyramid(B, h, w, w, 2, dev)
            Cg, Dg = gev.shape[1], gev.shape[2]
            self._geo_pyr = ((torch.empty(B, h, w, Dg, Cg, device=dev), torch.empty(B, h, w, Dg // 2, Cg, device=dev)) if dc else
                             (torch.empty(B, h, w, Cg, Dg, device=dev), torch.empty(B, h, w, Cg, Dg // 2, device=dev)))
        # volumes: init-corr pyramid = K1 with scale 1 (geometry.py:14,61-69), GEV pyramid (geometry.py:17-26);
        # the buffers persist per shape because the captured loop graph holds their addresses
        init = ops.corr1d_build(match_left, match_right, 2, 1.0, impl=self.impl, pyr=self._init_pyr)
        geo = (ops.geo_pool_dc if dc else ops.geo_pool)(gev, out=self._geo_pyr)
        self._vol = (geo, init)
        if net_list is not None:
            eng.load_state(net_list, ctx_list)   # else: EncoderEngine.run already filled X[i][:, :128] and CTX[i]
        eng.DELTA["f32"].zero_()
        eng.FLOW["f32"].copy_(init_disp.permute(0, 2, 3, 1))
        if all_preds:
            preds = []
            for _ in range(iters):
                eng.step(self._lookup, with_mask=False)
                dsp = eng.FLOW["f32"].view(B, h, w)
                ops.corr1d_lookup([], dsp, args.corr_radius, None, delta=eng.DELTA["f32"])      # disp += delta_disp
                eng.DELTA["f32"].zero_()          # applied: the next iteration's lookup must not add it again
                eng.mask_head()
                preds.append(self.upsample_disp(dsp, eng.MH["f32"][..., :32].permute(0, 3, 1, 2), stem_2x).clone())
            return preds
        gkey = (iters,)
        if self._capturing:                      # inside the whole-forward capture: the loop joins that graph
            self._run_loop(iters)
        elif self.use_cuda_graph and gkey in self._graphs:
            self._graphs[gkey].replay()
        elif self.use_cuda_graph and gkey in self._seen:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_loop(iters)
            self._graphs[gkey] = g
            g.replay()
        else:
            self._run_loop(iters)
            self._seen.add(gkey)
        # disp += delta of the last iteration (igev_stereo.py:210); mask features of the last state (:208)
        disp = eng.FLOW["f32"].view(B, h, w)
        ops.corr1d_lookup([], disp, args.corr_radius, None, delta=eng.DELTA["f32"])
        eng.mask_head()
        mask_feat_4 = eng.MH["f32"][..., :32].permute(0, 3, 1, 2)
        return self.upsample_disp(disp, mask_feat_4, stem_2x)

    def _forward_test(self, image1, image2, iters: int, pre=None) -> torch.Tensor:
        """test_mode=True forward, launched eagerly: pre-loop, context encoder, hot path."""
        if pre is None:
            pre = self.prepare(image1, image2)
        if self.encoder is not None:
            if self.encoder.pack_weights():
                self._graphs.clear()
                self._seen.clear()
            self.encoder.run(image1)             # cnet + context convs -> hidden states / context terms (NHWC, in place)
        return self.hot_path(*pre, iters)

    def _forward_graphed(self, image1, image2, iters: int) -> torch.Tensor:
        """First call with a shape: eager (allocates buffers, packs weights, lets cuDNN choose); second: captured into
        one CUDA graph on static input buffers; from then on one replay per call.  The graph holds raw pointers into
        packed weights and activation buffers: any parameter / buffer change ((data_ptr, version) signature over the whole
        module) or a new input shape drops it."""
        sig = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if sig != self._full_sig:
            self._full_sig = sig
            self._full.clear()
            self._full_seen.clear()
        key = (tuple(image1.shape), str(image1.device), iters)
        if any(k[:2] != key[:2] for k in list(self._full) + list(self._full_seen)):    # another shape: its buffers are gone
            self._full.clear()
            self._full_seen.clear()
        ent = self._full.get(key)
        if ent is None:
            if key not in self._full_seen:
                self._full_seen.add(key)
                return self._forward_test(image1, image2, iters)
            in1 = torch.empty(image1.shape, device=image1.device, dtype=torch.float32)
            in2 = torch.empty_like(in1)
            in1.copy_(image1)
            in2.copy_(image2)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            self._capturing = True
            try:
                with torch.cuda.graph(g):
                    up = self._forward_test(in1, in2, iters)
            except RuntimeError as e:            # a module of this configuration cannot be captured: keep serving, eagerly
                import warnings
                warnings.warn(f"IGEVStereo: whole-forward CUDA graph capture failed ({e}); running eagerly")
                self.full_graph = False
                torch.cuda.synchronize()
                return self._forward_test(image1, image2, iters)
            finally:
                self._capturing = False
            ent = self._full[key] = (g, in1, in2, up)
        g, in1, in2, up = ent
        in1.copy_(image1, non_blocking=True)     # device or pinned-host source; same stream as the replay
        in2.copy_(image2, non_blocking=True)
        g.replay()
        return up.clone()

    def forward(self, image1, image2, iters=12, flow_init=None, test_mode=False):
        """Estimate disparity between a stereo pair; returns (None, -disparity) like the reference."""
        if not test_mode and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "test_mode=False with trainable parameters asks for the autograd graph: train with the reference graph and "
                "load the checkpoint here (under torch.no_grad() this call returns every iteration's prediction)")
        if not image1.is_cuda:
            raise L.DktError("IGEVStereo (B200 engine) needs CUDA inputs; there is no CPU fallback")
        with torch.no_grad():
            if test_mode and self.use_cuda_graph and self.full_graph and self.impl == "tc":
                return None, self._forward_graphed(image1, image2, iters)
            pre = self.prepare(image1, image2)
            if not test_mode:
                # reference igev_stereo.py:178-189,222-226: {'init_disp': ..., 'disp_preds': [...]}, no autograd graph
                if self.encoder is not None:
                    if self.encoder.pack_weights():
                        self._graphs.clear()
                        self._seen.clear()
                    self.encoder.run(image1)
                init_disp, stem_2x = pre[3], pre[6]
                xspx = self.spx_4(self._feat4)
                spx_pred = F.softmax(self.spx(self.spx_2(xspx, stem_2x)), 1)
                init_up = -ops.context_upsample(init_disp.squeeze(1).contiguous(), spx_pred.float().contiguous(),
                                                in_scale=4.0, out_scale=1.0)
                return {"init_disp": init_up, "disp_preds": self.hot_path(*pre, iters, all_preds=True)}
            return None, self._forward_test(image1, image2, iters, pre)
