"""Update block of RAFT-Stereo / IGEV-Stereo served by the B200 kernels.

``BasicMultiUpdateBlock`` below is a *parameter container* with exactly the reference's
state-dict keys (reference core/update.py:97-113 and meta_arch/igev_stereo/update.py:104-119),
so DKT checkpoints load unchanged.  The arithmetic of ``forward`` (reference core/update.py:115-138)
is executed by ``UpdateEngine``: NHWC buffers resident in HBM for the whole GRU loop and a fixed
sequence of fused kernels per iteration ("Data layout" and "Iteration schedule").

Channel-concatenations of the reference (``torch.cat``) never materialise: producers write
directly into channel slices of the per-scale buffers

    X0 = [ h0 | motion(126)+flow(2) | up(h1) ]   384 ch   (gru08 input "hx")
    X1 = [ h1 | pool(h0)            | up(h2) ]   384 ch   (gru16)
    X2 = [ h2 | pool(h1)                     ]   256 ch   (gru32)
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from ._lib import tensor_slice as TS


# ---------------------------------------------------------------------------------------------
# parameter containers (names == reference)
# ---------------------------------------------------------------------------------------------
class _Head(nn.Module):
    def __init__(self, input_dim=128, hidden_dim=256, output_dim=2):
        super().__init__()
        self.conv1 = nn.Conv2d(input_dim, hidden_dim, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden_dim, output_dim, 3, padding=1)


class ConvGRU(nn.Module):
    def __init__(self, hidden_dim, input_dim, kernel_size=3):
        super().__init__()
        for n in ("convz", "convr", "convq"):
            setattr(self, n, nn.Conv2d(hidden_dim + input_dim, hidden_dim, kernel_size, padding=kernel_size // 2))


class BasicMotionEncoder(nn.Module):
    def __init__(self, args, igev: bool):
        super().__init__()
        taps = 2 * args.corr_radius + 1
        cor_planes = args.corr_levels * taps * (9 if igev else 1)
        stem = ("convd1", "convd2") if igev else ("convf1", "convf2")
        nflow = 1 if igev else 2
        self.convc1 = nn.Conv2d(cor_planes, 64, 1)
        self.convc2 = nn.Conv2d(64, 64, 3, padding=1)
        setattr(self, stem[0], nn.Conv2d(nflow, 64, 7, padding=3))
        setattr(self, stem[1], nn.Conv2d(64, 64, 3, padding=1))
        self.conv = nn.Conv2d(128, 128 - nflow, 3, padding=1)


class BasicMultiUpdateBlock(nn.Module):
    """Same parameters as the reference block; forward runs on the engine."""

    def __init__(self, args, hidden_dims: Sequence[int] = (), igev: bool = False):
        super().__init__()
        self.args = args
        self.igev = igev
        hd = list(hidden_dims)
        n = args.n_gru_layers
        assert hd == [128, 128, 128] and n in (1, 2, 3), \
            "the B200 engine serves 128 hidden channels (the shipped configs) with 1, 2 or 3 GRU levels"
        self.encoder = BasicMotionEncoder(args, igev)
        names = ("gru04", "gru08", "gru16") if igev else ("gru08", "gru16", "gru32")
        # input widths as in the reference (core/update.py:104-106): a level that does not run contributes no input
        setattr(self, names[0], ConvGRU(hd[2], 128 + hd[1] * (n > 1)))
        setattr(self, names[1], ConvGRU(hd[1], hd[0] * (n == 3) + hd[2]))
        setattr(self, names[2], ConvGRU(hd[0], hd[1]))
        self.gru_names = names
        if igev:
            self.disp_head = _Head(hd[2], 256, 1)
            self.mask_feat_4 = nn.Sequential(nn.Conv2d(hd[2], 32, 3, padding=1), nn.ReLU(inplace=True))
        else:
            self.flow_head = _Head(hd[2], 256, 2)
            factor = 2 ** args.n_downsample
            self.mask = nn.Sequential(nn.Conv2d(hd[2], 256, 3, padding=1), nn.ReLU(inplace=True),
                                      nn.Conv2d(256, factor ** 2 * 9, 1, padding=0))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("BasicMultiUpdateBlock.forward is executed by UpdateEngine (B200 kernels); "
                           "there is no PyTorch fallback")


# ---------------------------------------------------------------------------------------------
# engine
# ---------------------------------------------------------------------------------------------
def LaunchProfilerActive():
    return ops.LaunchProfiler.active


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


class UpdateEngine:
    """Holds packed weights + per-shape activation buffers and runs update-block iterations."""

    def __init__(self, block: BasicMultiUpdateBlock, impl: str = "tc"):
        assert impl in ("tc", "simt")
        self.block = block
        self.impl = impl
        self.igev = block.igev
        # lookup kernels apply convc1 (1x1 + ReLU) on chip and write COR1 directly; set False to run the
        # unfused pair (lookup -> CORR buffer -> convc1 conv) for A/B measurements
        self.fused_enc = os.environ.get("DKT_FUSED_ENC", "1") == "1"
        # 7x7 flow stem and the head's last conv in their tensor-core forms (tc only); 0 = generic kernels
        self.fast_small_convs = os.environ.get("DKT_FAST_SMALL_CONVS", "1") == "1"
        # head conv1 + ReLU + the channel half of conv2 in ONE kernel (DKT_EPI_PROJ): the 256-channel hidden map of
        # the flow / disparity head never reaches HBM; 0 = conv1 -> FH -> 1x1 tap conv
        self.fused_head = os.environ.get("DKT_FUSED_HEAD", "1") == "1"
        self.two_streams = os.environ.get("DKT_TWO_STREAMS", "1") == "1"
        # args.slow_fast_gru (reference raft_stereo.py:157-160, igev_stereo.py:201-204): per iteration one extra
        # update of the coarsest GRU, then one of the two coarse GRUs, before the full update
        self.slow_fast = bool(getattr(block.args, "slow_fast_gru", False))
        # GRU levels that run (args.n_gru_layers, reference core/update.py:117-132): level i lives at 1 / (4 * 2^i) resolution
        self.n = int(block.args.n_gru_layers)
        # MMAs per K step of the GRU convs and of the motion encoder's 64/128-channel convs (tensor-core path): 3 = every
        # operand a 16-bit (hi, lo) pair; 2 = activations as ONE half value (hi plane only) against (hi, lo) weights.
        # The per-group error study on the headline workload (profiles/r2_precision_study_*)
        # puts 2-MMA GRUs + motion encoder at +1.3e-4 px over the 3-MMA engine; encoders and heads keep 3 (they cost
        # 3e-3 / 3e-4 px with 2).  bfloat16 builds (DKT_SPLIT_FP16=0) always run 3: one bf16 value has 8 bits.
        half = impl == "tc" and L.split_dtype() == torch.float16
        self.gru2 = half and os.environ.get("DKT_GRU_TERMS", "2") == "2"
        self.menc2 = half and os.environ.get("DKT_MENC_TERMS", "2") == "2"
        # the two COARSE GRUs (1/8 and 1/16 resolution) with ONE MMA per K step: half activations x half weights (the
        # study's h_x1_w1 row: +5e-5 px for this group; every other group needs its weights as a pair).  Measured on
        # B200 against the reference goldens (profiles/r2f_parity_accsplit.txt): +0.2..0.5e-4 px, -3 % step time.
        self.coarse1 = self.gru2 and os.environ.get("DKT_COARSE_GRU_TERMS", "1") == "1"
        # fused lookup + convc1 with the 1x1 contraction on tcgen05 (csrc/lookup_tc.cu); 0 = the exact-fp32 CUDA-core form.
        # The taps follow the motion encoder's policy: hi plane only (2 MMAs per K step) when menc2 -- the study's convc1
        # row: +6.8e-5 px; RAFT 43.7 vs 51.6 us per launch, IGEV 130 vs 231 us (a (hi, lo) tile leaves one CTA per SM).
        self.lookup_tc = impl == "tc" and self.fused_enc and os.environ.get("DKT_LOOKUP_TC", "1") == "1"
        taps = os.environ.get("DKT_LOOKUP_TAP_PLANES")
        self.lookup_tap_planes = int(taps) if taps else (1 if self.menc2 else 2)
        self.lookup_tc_w = None
        self.side_stream = None
        self.weights: Optional[Dict[str, ops.ConvWeights]] = None
        self._wsig = None
        self.shape = None

    # ---- weights -------------------------------------------------------------------------------
    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.block.parameters())

    def pack_weights(self) -> bool:
        """Repack when a parameter changed; returns True if it did (captured CUDA graphs are stale then)."""
        sig = self._signature()
        if self.weights is not None and sig == self._wsig:
            return False
        b, tc = self.block, self.impl == "tc"
        enc = b.encoder
        w: Dict[str, ops.ConvWeights] = {}
        self.corr_planes = enc.convc1.in_channels
        self.corr_pad = _pad64(self.corr_planes)
        w["convc1"] = ops.pack_conv(enc.convc1.weight, enc.convc1.bias, cin_pad=self.corr_pad, tc=tc)
        if self.lookup_tc:
            self.lookup_tc_w = ops.pack_lookup_tc(enc.convc1.weight, enc.convc1.bias)
        w["convc2"] = ops.pack_conv(enc.convc2.weight, enc.convc2.bias, tc=tc)
        stem1 = enc.convd1 if self.igev else enc.convf1
        stem2 = enc.convd2 if self.igev else enc.convf2
        w["stem1"] = ops.pack_conv(stem1.weight, stem1.bias, tc=False)           # 7x7, 1-2 input ch: CUDA cores
        w["stem2"] = ops.pack_conv(stem2.weight, stem2.bias, tc=tc)
        w["conv"] = ops.pack_conv(enc.conv.weight, enc.conv.bias, tc=tc)
        self.gru_bias = []
        for i, name in enumerate(b.gru_names[:self.n]):
            g = getattr(b, name)
            w[f"zr{i}"] = ops.pack_conv_cat([g.convz.weight, g.convr.weight], tc=tc)
            w[f"q{i}"] = ops.pack_conv(g.convq.weight, None, tc=tc)
            if self.coarse1 and i > 0:
                w[f"zr{i}"].w_lo = w[f"q{i}"].w_lo = None
            self.gru_bias.append(torch.cat([g.convz.bias, g.convr.bias, g.convq.bias]).detach().float().contiguous())
        head = b.disp_head if self.igev else b.flow_head
        w["head1"] = ops.pack_conv(head.conv1.weight, head.conv1.bias, tc=tc)
        w["head2"] = ops.pack_conv(head.conv2.weight, head.conv2.bias, tc=tc)
        if tc:
            # tensor-core forms of the two awkward convs (see DESIGN.md "Iteration schedule"):
            #  * the 7x7 flow stem as x-im2col rows + a 7x1 conv over 64 channels
            #  * the head's 3x3 conv to ONE used channel as a 1x1 conv to 9 tap responses + a spatial tap sum
            nfl = stem1.in_channels
            w7 = stem1.weight.detach().permute(0, 3, 1, 2).reshape(stem1.out_channels, 7 * nfl, 7, 1)
            w["stem1t"] = ops.pack_conv_general(w7, stem1.bias, cin_pad=32)
            c2 = head.conv2
            w9 = c2.weight.detach()[0].permute(1, 2, 0).reshape(9, c2.in_channels, 1, 1)
            w["head2t"] = ops.pack_conv(w9, None, tc=True)
            self.head2_proj = ops.pack_proj3x3(c2.weight, 0)
            self.head2_bias0 = float(c2.bias.detach()[0])
        if self.igev:
            w["mask0"] = ops.pack_conv(b.mask_feat_4[0].weight, b.mask_feat_4[0].bias, tc=tc)
        else:
            w["mask0"] = ops.pack_conv(b.mask[0].weight, b.mask[0].bias, tc=tc)
            # 9 * factor^2 output channels: 144 at n_downsample = 2, 576 at 3 -- more than one conv launch holds (N <= 256)
            m2 = b.mask[2]
            self.mask2_chunks = [(c0, min(c0 + 192, m2.out_channels)) for c0 in range(0, m2.out_channels, 192)] \
                if m2.out_channels > 256 else [(0, m2.out_channels)]
            for j, (c0, c1) in enumerate(self.mask2_chunks):
                w[f"mask2.{j}"] = ops.pack_conv(m2.weight[c0:c1], m2.bias[c0:c1], tc=tc)
        self.weights, self._wsig = w, sig
        return True

    # ---- buffers -------------------------------------------------------------------------------
    def _buf(self, B, H, W, Cc, f32=True, split=None):
        split = (self.impl == "tc") if split is None else split
        dev = self.device
        d = {"C": Cc, "f32": None, "hi": None, "lo": None}
        if f32:
            d["f32"] = torch.zeros(B, H, W, Cc, device=dev, dtype=torch.float32)
        if split:
            d["hi"] = torch.zeros(B, H, W, Cc, device=dev, dtype=L.split_dtype())
            d["lo"] = torch.zeros(B, H, W, Cc, device=dev, dtype=L.split_dtype())
        return d

    def cor1_slice(self):
        simt, split = self.impl == "simt", self.impl == "tc"
        return self._slice(self.COR1, 0, 64, simt, split, lo=not self.menc2)

    @staticmethod
    def _slice(buf, c0=0, cnt=None, f32=True, split=True, lo=True):
        """``lo=False``: the hi plane only -- as a conv SOURCE this selects the 2-MMA form, as a destination it skips
        the lo store (the buffer keeps its lo plane allocated; nobody reads it)."""
        return TS(buf["f32"] if f32 else None, buf["hi"] if split else None, buf["lo"] if (split and lo) else None, c0, cnt)

    def allocate(self, B: int, h: int, w: int, device) -> None:
        shape = (B, h, w, str(device))
        if self.shape == shape:
            return
        self.device = device
        self.B = B
        self.hw = [(h, w)]
        for _ in range(self.n - 1):
            ph, pw = self.hw[-1]
            self.hw.append(((ph - 1) // 2 + 1, (pw - 1) // 2 + 1))
        h0, w0 = self.hw[0]
        n = self.n
        simt = self.impl == "simt"
        nflow = 1 if self.igev else 2
        self.nflow = nflow
        # in "tc" mode fp32 copies are only kept where an elementwise consumer needs them
        # X[i] = [h_i | x]: level 0 x = motion (+ up(h1)), level 1 x = pool(h0) (+ up(h2)), level 2 x = pool(h1)
        self.xc = [128 + (128 if n > 1 else 0), 128 + (128 if n == 3 else 0), 128][:n]
        self.X = [self._buf(B, *self.hw[i], 128 + self.xc[i]) for i in range(n)]
        self.RH = [self._buf(B, *self.hw[i], 128, f32=simt) for i in range(n)]
        self.Z = [self._buf(B, *self.hw[i], 128, split=False) for i in range(n)]
        self.CTX = [self._buf(B, *self.hw[i], 384, split=False) for i in range(n)]
        self.CORR = self._buf(B, h0, w0, self.corr_pad)
        self.COR1 = self._buf(B, h0, w0, 64, f32=simt)
        self.FLO1 = self._buf(B, h0, w0, 64, f32=simt)
        self.CF = self._buf(B, h0, w0, 128, f32=simt)
        self.FH = self._buf(B, h0, w0, 256, f32=simt)
        self.FLOW = self._buf(B, h0, w0, nflow, split=False)     # flow (x,y) or disparity
        self.DELTA = self._buf(B, h0, w0, nflow, split=False)
        # mask-head features; IGEV's 32 mask_feat_4 channels sit in a 64-channel buffer (upper half zero) because they
        # feed the 64-channel K blocks of the upsampling convs (igev_stereo.py::_upsample_native)
        self.MH = self._buf(B, h0, w0, 256 if not self.igev else 64, f32=True)
        if not self.igev:
            factor = 2 ** self.block.args.n_downsample
            self.MASK = self._buf(B, h0, w0, 9 * factor * factor, split=False)
        if self.impl == "tc":
            self.FROWS = self._buf(B, h0, w0, 32, f32=False)        # x-im2col of the flow field (7 taps x nflow <= 14 ch)
            self.TAPS = self._buf(B, h0, w0, 16, split=False)       # 9 tap responses of the head's last conv
        self.coords_x = torch.zeros(B, h0, w0, device=device, dtype=torch.float32)
        self.shape = shape

    # ---- per-forward state ---------------------------------------------------------------------
    def load_state(self, net_list: Sequence[torch.Tensor], inp_list: Sequence[Sequence[torch.Tensor]]) -> None:
        """net_list[i] (B,128,h_i,w_i) NCHW hidden states, inp_list[i] = (cz,cr,cq) NCHW context
        terms (reference raft_stereo.py:110-114).  Conv biases of the gates are folded into the
        context term once here."""
        split = self.impl == "tc"
        for i in range(self.n):
            ops.nchw_to_nhwc(net_list[i], self._slice(self.X[i], 0, 128, True, split))
            ctx = inp_list[i]
            ctx = ctx if torch.is_tensor(ctx) else torch.cat(list(ctx), dim=1)
            ops.nchw_to_nhwc(ctx, self._slice(self.CTX[i], 0, 384, True, False), bias=self.gru_bias[i])

    def hidden_states(self) -> List[torch.Tensor]:
        return [ops.nhwc_to_nchw(self._slice(self.X[i], 0, 128, True, False), self.B, *self.hw[i], self.device)
                for i in range(self.n)]

    # ---- one GRU ---------------------------------------------------------------------------------
    def _gru(self, i: int, x_cnt: int) -> None:
        """ConvGRU at scale i on X[i] = [h | x]; reference core/update.py:23-32."""
        B, (H, W), impl, split, simt = self.B, self.hw[i], self.impl, self.impl == "tc", self.impl == "simt"
        X, RH, Z, CTX = self.X[i], self.RH[i], self.Z[i], self.CTX[i]
        lo = not self.gru2                   # the GRU convs read (hi, lo) pairs (3 MMAs) or hi planes (2 MMAs)
        hx = self._slice(X, 0, 128 + x_cnt, simt, split, lo)
        h = self._slice(X, 0, 128, True, False)
        z = self._slice(Z, 0, 128, True, False)
        e = ops.make_epilogue(L.EPI_GRU_ZR, out=self._slice(RH, 0, 128, simt, split, lo), ctx=CTX["f32"], ctx_c0=0, z=z, h=h)
        ops.conv2d([hx], self.weights[f"zr{i}"], e, B, H, W, impl)
        # h' of the finest scale also feeds the flow / mask heads, which always run 3 MMAs: it keeps its lo plane
        e = ops.make_epilogue(L.EPI_GRU_Q, out=self._slice(X, 0, 128, True, split, lo or i == 0), ctx=CTX["f32"], ctx_c0=256, z=z, h=h)
        ops.conv2d([self._slice(RH, 0, 128, simt, split, lo), self._slice(X, 128, x_cnt, simt, split, lo)],
                   self.weights[f"q{i}"], e, B, H, W, impl)

    def _coarsest_gru(self) -> None:
        """gru32 alone: update_block(iter32=True, iter16=False, iter08=False, update=False), reference core/update.py:117-118
        (only with three levels)."""
        if self.n < 3:
            return
        B, split, simt = self.B, self.impl == "tc", self.impl == "simt"
        h1, w1 = self.hw[1]
        S = self._slice
        ops.pool2x(S(self.X[1], 0, 128, True, False), S(self.X[2], 128, 128, simt, split, not self.gru2), B, h1, w1)
        self._gru(2, 128)

    def _coarse_grus(self) -> None:
        """gru32 then gru16 (reference core/update.py:118-128): coarse -> fine; every GRU sees the OLD finer state
        pooled and the NEW coarser state upsampled.  Two levels: gru16 sees pool(h0) alone; one level: nothing to do."""
        if self.n < 2:
            return
        B, split, simt = self.B, self.impl == "tc", self.impl == "simt"
        (h0, w0), (h1, w1) = self.hw[:2]
        X0, X1 = self.X[:2]
        S = self._slice
        self._coarsest_gru()
        ops.pool2x(S(X0, 0, 128, True, False), S(X1, 128, 128, simt, split, not self.gru2), B, h0, w0)
        if self.n == 3:
            h2, w2 = self.hw[2]
            ops.interp(S(self.X[2], 0, 128, True, False), S(X1, 256, 128, simt, split, not self.gru2), B, h2, w2, h1, w1)
        self._gru(1, self.xc[1])

    # ---- one update-block call (reference core/update.py:115-138) -------------------------------
    def step(self, lookup, with_mask: bool = False) -> None:
        """lookup(engine) does the coords / disparity bookkeeping of this iteration and fills self.COR1
        (fused_enc: lookup + convc1 in one kernel) or self.CORR (unfused)."""
        B, impl, split, simt = self.B, self.impl, self.impl == "tc", self.impl == "simt"
        h0, w0 = self.hw[0]
        X0 = self.X[0]
        Wt = self.weights
        S = self._slice
        E = ops.make_epilogue
        # The two coarse GRUs and the motion encoder are independent until gru08: they run on two streams (fork /
        # join with events, also inside a CUDA-graph capture) so that the tail of one branch's persistent kernels
        # (1/8 and 1/16 resolution fill 3.4 and 0.9 waves of CTA pairs) overlaps the other branch's work.
        fork = self.two_streams and LaunchProfilerActive() is None and self.n > 1
        main = torch.cuda.current_stream()
        if self.slow_fast:
            self._coarsest_gru()
            self._coarse_grus()
        if fork:
            if self.side_stream is None or self.side_stream.device != self.device:
                self.side_stream = torch.cuda.Stream(device=self.device)
            ev = torch.cuda.Event()
            ev.record(main)
            self.side_stream.wait_event(ev)
            with torch.cuda.stream(self.side_stream):
                self._coarse_grus()
                ev2 = torch.cuda.Event()
                ev2.record(self.side_stream)
        else:
            self._coarse_grus()
        # motion encoder (reference core/update.py:77-85)
        lookup(self)
        mlo, glo = not self.menc2, not self.gru2   # lo planes of the motion encoder's / the GRUs' activations in use?
        if not self.fused_enc:
            ops.conv2d([S(self.CORR, 0, self.corr_pad, simt, split)], Wt["convc1"],
                       E(L.EPI_LINEAR, S(self.COR1, 0, 64, simt, split, mlo), act=L.ACT_RELU, bias=Wt["convc1"].bias), B, h0, w0, impl)
        ops.conv2d([S(self.COR1, 0, 64, simt, split, mlo)], Wt["convc2"],
                   E(L.EPI_LINEAR, S(self.CF, 0, 64, simt, split, mlo), act=L.ACT_RELU, bias=Wt["convc2"].bias), B, h0, w0, impl)
        if split and self.fast_small_convs:
            # the 7x7 stem sees the flow / disparity VALUES (hundreds of pixels at large widths): always a (hi, lo) pair
            ops.stem_rows(self.FLOW["f32"], self.FROWS["hi"], self.FROWS["lo"], kw=7, scale=1.0, shift=0.0, layout="nhwc")
            ops.conv2d_ex([S(self.FROWS, 0, 32, False, True)], Wt["stem1t"],
                          E(L.EPI_LINEAR, S(self.FLO1, 0, 64, False, True, mlo), act=L.ACT_RELU, bias=Wt["stem1t"].bias), B, h0, w0)
        else:
            ops.conv2d([S(self.FLOW, 0, self.nflow, True, False)], Wt["stem1"],
                       E(L.EPI_LINEAR, S(self.FLO1, 0, 64, simt, split, mlo), act=L.ACT_RELU, bias=Wt["stem1"].bias), B, h0, w0, "simt")
        ops.conv2d([S(self.FLO1, 0, 64, simt, split, mlo)], Wt["stem2"],
                   E(L.EPI_LINEAR, S(self.CF, 64, 64, simt, split, mlo), act=L.ACT_RELU, bias=Wt["stem2"].bias), B, h0, w0, impl)
        ops.conv2d([S(self.CF, 0, 128, simt, split, mlo)], Wt["conv"],
                   E(L.EPI_LINEAR, S(X0, 128, 128, simt, split, glo), act=L.ACT_RELU, bias=Wt["conv"].bias,
                     tail=self.FLOW["f32"]), B, h0, w0, impl)
        if fork:
            main.wait_event(ev2)
        if self.n > 1:
            h1, w1 = self.hw[1]
            ops.interp(S(self.X[1], 0, 128, True, False), S(X0, 256, 128, simt, split, glo), B, h1, w1, h0, w0)
        self._gru(0, self.xc[0])
        # flow / disparity head (reference core/update.py:13-14)
        if split and self.fast_small_convs and self.fused_head:
            ops.conv2d([S(X0, 0, 128, False, True)], Wt["head1"],
                       E(L.EPI_PROJ, S(self.TAPS, 0, 16, True, False), act=L.ACT_RELU, bias=Wt["head1"].bias,
                         proj=self.head2_proj), B, h0, w0, "tc")
            ops.tapsum3x3(self.TAPS["f32"], self.head2_bias0, self.DELTA["f32"])
            if with_mask:
                self.mask_head()
            return
        ops.conv2d([S(X0, 0, 128, simt, split)], Wt["head1"],
                   E(L.EPI_LINEAR, S(self.FH, 0, 256, simt, split), act=L.ACT_RELU, bias=Wt["head1"].bias), B, h0, w0, impl)
        if split and self.fast_small_convs:
            ops.conv2d([S(self.FH, 0, 256, False, True)], Wt["head2t"],
                       E(L.EPI_LINEAR, S(self.TAPS, 0, 9, True, False)), B, h0, w0, "tc")
            ops.tapsum3x3(self.TAPS["f32"], self.head2_bias0, self.DELTA["f32"])
        else:
            ops.conv2d([S(self.FH, 0, 256, simt, split)], Wt["head2"],
                       E(L.EPI_LINEAR, S(self.DELTA, 0, self.nflow, True, False), bias=Wt["head2"].bias), B, h0, w0, impl)
        if with_mask:
            self.mask_head()

    def mask_head(self) -> None:
        """reference core/update.py:110-113,137 (RAFT) / igev update.py:117-119,141 (IGEV)."""
        B, impl, split, simt = self.B, self.impl, self.impl == "tc", self.impl == "simt"
        h0, w0 = self.hw[0]
        S, E, Wt = self._slice, ops.make_epilogue, self.weights
        if self.igev:
            ops.conv2d([S(self.X[0], 0, 128, simt, split)], Wt["mask0"],
                       E(L.EPI_LINEAR, S(self.MH, 0, 32, True, split), act=L.ACT_RELU, bias=Wt["mask0"].bias), B, h0, w0, impl)
        else:
            ops.conv2d([S(self.X[0], 0, 128, simt, split)], Wt["mask0"],
                       E(L.EPI_LINEAR, S(self.MH, 0, 256, True, split), act=L.ACT_RELU, bias=Wt["mask0"].bias), B, h0, w0, impl)
            for j, (c0, c1) in enumerate(self.mask2_chunks):
                wj = Wt[f"mask2.{j}"]
                ops.conv2d([S(self.MH, 0, 256, simt, split)], wj,
                           E(L.EPI_LINEAR, S(self.MASK, c0, c1 - c0, True, False), scale=0.25, bias=wj.bias), B, h0, w0, impl)
