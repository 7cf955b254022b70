"""RAFT-Stereo with the hot path (volume, lookup, GRU loop, upsampling) on B200 kernels.

Drop-in for reference meta_arch/raft_stereo/raft_stereo.py: same constructor
(``RAFTStereo(args)`` with the merged JSON/CLI namespace), same ``forward(image1, image2,
iters, flow_init, test_mode)`` and the same ``state_dict`` keys, so DKT checkpoints load with
``strict=True``.  Only the inference contract (``test_mode=True``) is served; the training
graph (autograd through the loop) is outside this engine's scope and raises.

What runs where:
  libdkt kernels   : correlation pyramid (K1), per-iteration lookup + coordinate update (K2),
                     motion encoder + 3 ConvGRUs + flow head (K3), mask head + convex
                     upsampling (K4)                          (reference raft_stereo.py:118-183)
                     and -- default on the tensor-core path -- cnet / fnet / context_zqr_convs on the
                     same conv kernel (``encoder.EncoderEngine``, reference raft_stereo.py:91-114)
  PyTorch (cuDNN)  : the encoders when ``DKT_NATIVE_ENCODER=0``, ``corr_implementation="b200_fp32"``,
                     mixed precision, or a configuration the encoder engine does not serve
The GRU loop is captured into a CUDA graph per (shape, iters) and replayed.
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .extractor import BasicEncoder, MultiBasicEncoder, ResidualBlock
from .update import BasicMultiUpdateBlock, UpdateEngine

# corr_implementation values of the reference config that this engine serves, -> kernel family
_IMPLS = {"reg": "tc", "reg_cuda": "tc", "b200": "tc", "b200_tc": "tc", "b200_fp32": "simt", "b200_simt": "simt"}


@contextlib.contextmanager
def _fp32_math(enabled: bool):
    """Run the PyTorch extractors in true fp32 (parity target) unless TF32 is explicitly allowed."""
    if not enabled:
        yield
        return
    c, m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32 = c
        torch.backends.cuda.matmul.allow_tf32 = m


class RAFTStereo(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        impl = getattr(args, "corr_implementation", "reg")
        if impl not in _IMPLS:
            raise NotImplementedError(
                f"corr_implementation={impl!r} is not served by the B200 engine (supported: {sorted(_IMPLS)})")
        self.impl = _IMPLS[impl]
        if os.environ.get("DKT_IMPL"):
            self.impl = os.environ["DKT_IMPL"]
        context_dims = args.hidden_dims
        self.cnet = MultiBasicEncoder(output_dim=[args.hidden_dims, context_dims], norm_fn=args.context_norm,
                                      downsample=args.n_downsample)
        self.update_block = BasicMultiUpdateBlock(args, hidden_dims=args.hidden_dims, igev=False)
        self.context_zqr_convs = nn.ModuleList(
            [nn.Conv2d(context_dims[i], args.hidden_dims[i] * 3, 3, padding=1) for i in range(args.n_gru_layers)])
        backbone = getattr(args, "backbone_type", "default")
        if backbone != "default":
            raise NotImplementedError("only backbone_type='default' is served (configs/raft_stereo/base.json)")
        if getattr(args, "shared_backbone", False):
            self.conv2 = nn.Sequential(ResidualBlock(128, 128, "instance", stride=1), nn.Conv2d(128, 256, 3, padding=1))
        else:
            self.fnet = BasicEncoder(output_dim=256, norm_fn="instance", downsample=args.n_downsample)
        self.engine = UpdateEngine(self.update_block, self.impl)
        self.encoder = None
        if (os.environ.get("DKT_NATIVE_ENCODER", "1") == "1" and self.impl == "tc" and not getattr(args, "shared_backbone", False)
                and args.context_norm == "batch" and args.n_downsample == 2 and not getattr(args, "mixed_precision", False)):
            from .encoder import EncoderEngine
            self.encoder = EncoderEngine(self.fnet, self.cnet, self.context_zqr_convs, self.engine)
        self.use_cuda_graph = os.environ.get("DKT_CUDA_GRAPH", "1") == "1"
        self.extractor_fp32 = not getattr(args, "extractor_tf32", False)
        self.channels_last = os.environ.get("DKT_CHANNELS_LAST", "0") == "1"
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._seen = set()
        # whole-forward CUDA graphs of the native path (encoders + volume + loop + upsampling), per input shape:
        # one replay per call instead of ~700 launches through Python (DKT_FULL_GRAPH=0: only the loop is a graph)
        self.full_graph = os.environ.get("DKT_FULL_GRAPH", "1") == "1"
        self._full: Dict[tuple, tuple] = {}
        self._full_seen = set()
        self._capturing = False
        self._native_shape = None
        self._pyr = None
        self._pyr_key = None

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

    # ---- L1: extractors (PyTorch) ---------------------------------------------------------------
    def extract(self, image1: torch.Tensor, image2: torch.Tensor):
        """reference raft_stereo.py:91-116 -> fmap1, fmap2, net_list, ctx_list (cz|cr|cq concatenated)."""
        args = self.args
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        if self.channels_last:
            image1 = image1.contiguous(memory_format=torch.channels_last)
            image2 = image2.contiguous(memory_format=torch.channels_last)
        with _fp32_math(self.extractor_fp32), torch.autocast("cuda", enabled=bool(getattr(args, "mixed_precision", False))):
            if getattr(args, "shared_backbone", False):
                *cnet_list, x = self.cnet(torch.cat((image1, image2), dim=0), dual_inp=True, num_layers=args.n_gru_layers)
                fmap1, fmap2 = self.conv2(x).split(dim=0, split_size=x.shape[0] // 2)
            else:
                cnet_list = self.cnet(image1, num_layers=args.n_gru_layers)
                fmap1, fmap2 = self.fnet([image1, image2])
            net_list = [torch.tanh(x[0]) for x in cnet_list]
            ctx_list = [conv(torch.relu(x[1])) for x, conv in zip(cnet_list, self.context_zqr_convs)]
        return fmap1.float(), fmap2.float(), [n.float() for n in net_list], [c.float() for c in ctx_list]

    # ---- L2: hot path (B200 kernels) ----------------------------------------------------------------
    def _lookup(self, eng: UpdateEngine) -> None:
        if eng.lookup_tc:
            ops.corr1d_lookup_enc_tc(self._pyr, eng.coords_x, self.args.corr_radius, *eng.lookup_tc_w, eng.cor1_slice(),
                                     eng.lookup_tap_planes, delta=eng.DELTA["f32"], flow=eng.FLOW["f32"])
            return
        if eng.fused_enc:
            ops.corr1d_lookup_enc(self._pyr, eng.coords_x, self.args.corr_radius, eng.weights["convc1"],
                                  eng.cor1_slice(), delta=eng.DELTA["f32"], flow=eng.FLOW["f32"])
            return
        ops.corr1d_lookup(self._pyr, eng.coords_x, self.args.corr_radius, eng.CORR["f32"], "nhwc",
                          out_hi=eng.CORR["hi"], out_lo=eng.CORR["lo"], delta=eng.DELTA["f32"], flow=eng.FLOW["f32"])

    def _run_loop(self, iters: int) -> None:
        for _ in range(iters):
            self.engine.step(self._lookup, with_mask=False)

    def hot_path(self, fmap1, fmap2, net_list, ctx_list, iters: int, flow_init: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Volume build -> ``iters`` update iterations -> convex upsampling, from the PyTorch encoders'
        NCHW products.  Returns (flow_lowres (B,2,h,w), flow_up (B,1,H,W)) as reference raft_stereo.py:182-183."""
        args, eng = self.args, self.engine
        L.require_device(fmap1)
        B, D, h, w = fmap1.shape
        if eng.pack_weights():                    # weights changed: the captured loop graph points into the old packs
            self._graphs.clear()
            self._seen.clear()
        eng.allocate(B, h, w, fmap1.device)
        self._ensure_volume(B, D, h, w, fmap1.device)
        ops.corr1d_build(fmap1, fmap2, args.corr_levels, 1.0 / (D ** 0.5), impl=self.impl, pyr=self._pyr)      # K1
        eng.load_state(net_list, ctx_list)
        return self._loop_and_upsample(B, h, w, iters, flow_init)

    def _ensure_volume(self, B, D, h, w, dev) -> None:
        key = (B, D, h, w, self.args.corr_levels, str(dev))
        if self._pyr_key != key:
            self._pyr = ops.alloc_pyramid(B, h, w, w, self.args.corr_levels, dev)
            self._pyr_key = key
            self._graphs.clear()
            self._seen.clear()

    def _loop_and_upsample(self, B, h, w, iters, flow_init, all_preds: bool = False):
        args, eng = self.args, self.engine
        dev = eng.device
        # loop state: coords0 = pixel grid; coords1 = coords0 (+ flow_init); flow = coords1 - coords0
        eng.DELTA["f32"].zero_()
        xs = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, w).expand(B, h, w)
        eng.FLOW["f32"].zero_()
        if flow_init is not None:
            eng.coords_x.copy_(xs + flow_init[:, 0].float())
            eng.FLOW["f32"][..., 1].copy_(flow_init[:, 1].float())
        else:
            eng.coords_x.copy_(xs)
        if all_preds:
            # test_mode=False (reference raft_stereo.py:170-187): the mask head and the convex upsampling run after EVERY
            # iteration and every prediction is returned; launched eagerly (a training-time validation path, no graph)
            preds = []
            for _ in range(iters):
                eng.step(self._lookup, with_mask=True)
                ops.corr1d_lookup([], eng.coords_x, args.corr_radius, None, delta=eng.DELTA["f32"], flow=eng.FLOW["f32"])
                eng.DELTA["f32"].zero_()          # applied: the next iteration's lookup must not add it again
                preds.append(ops.convex_upsample(eng.FLOW["f32"], eng.MASK["f32"], 2 ** args.n_downsample))
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
        # coords1 += delta of the last iteration; flow = coords1 - coords0
        ops.corr1d_lookup([], eng.coords_x, args.corr_radius, None, delta=eng.DELTA["f32"], flow=eng.FLOW["f32"])
        eng.mask_head()
        flow_up = ops.convex_upsample(eng.FLOW["f32"], eng.MASK["f32"], 2 ** args.n_downsample)
        flow_lr = eng.FLOW["f32"].permute(0, 3, 1, 2).contiguous()
        return flow_lr, flow_up

    def _forward_native_eager(self, image1, image2, iters: int, flow_init=None, all_preds: bool = False):
        args, enc = self.args, self.encoder
        B = image1.shape[0]
        enc.run(image1, image2)
        h, w = enc.dims[2]
        D = enc.FMAP.C
        self._ensure_volume(B, D, h, w, image1.device)
        f = enc.FMAP
        ops.corr1d_build_split(f.hi[:B], f.lo[:B], f.hi[B:], f.lo[B:], args.corr_levels, 1.0 / (D ** 0.5), self._pyr)
        return self._loop_and_upsample(B, h, w, iters, flow_init, all_preds)

    def forward_native(self, image1, image2, iters: int, flow_init=None):
        """Whole forward on libdkt kernels: encoders (EncoderEngine) -> K1 from the bf16 (hi, lo) feature maps
        the encoder wrote -> loop -> upsampling.  No NCHW <-> NHWC conversion and no fp32 -> bf16 split pass.
        From the third call with the same input shape the whole thing is ONE CUDA-graph replay: the images are
        copied into static device buffers, the results are returned as fresh tensors."""
        shape_key = (tuple(image1.shape), str(image1.device))
        # captured graphs hold raw pointers: into the packed weights (stale after a parameter update) and into the
        # engines' activation buffers (re-allocated when the input shape changes) -- drop them in both cases
        if self.encoder.pack_weights() or shape_key != self._native_shape:
            self._native_shape = shape_key
            self._graphs.clear()
            self._seen.clear()
            self._full.clear()
            self._full_seen.clear()
        if not (self.use_cuda_graph and self.full_graph and flow_init is None):
            return self._forward_native_eager(image1, image2, iters, flow_init)
        key = (tuple(image1.shape), str(image1.device), iters)
        ent = self._full.get(key)
        if ent is None:
            if key not in self._full_seen:       # first call: allocates buffers, packs weights, warms up
                self._full_seen.add(key)
                return self._forward_native_eager(image1, image2, iters, None)
            in1 = torch.empty(image1.shape, device=image1.device, dtype=torch.float32)
            in2 = torch.empty_like(in1)
            in1.copy_(image1)
            in2.copy_(image2)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            self._capturing = True
            try:
                with torch.cuda.graph(g):
                    lr, up = self._forward_native_eager(in1, in2, iters, None)
            finally:
                self._capturing = False
            ent = self._full[key] = (g, in1, in2, lr, up)
        g, in1, in2, lr, up = ent
        in1.copy_(image1, non_blocking=True)     # device or pinned-host source; same stream as the replay
        in2.copy_(image2, non_blocking=True)
        g.replay()
        return lr.clone(), up.clone()

    def forward(self, image1, image2, iters=12, flow_init=None, test_mode=False):
        """Estimate disparity (returned as negative flow, like the reference) between a stereo pair."""
        # test_mode=False, reference raft_stereo.py:185-187: {'disp_preds': [one full-resolution prediction per iteration]}.
        # The engine computes them without an autograd graph: fine for validation / pseudo-labelling under
        # torch.no_grad() or with frozen parameters, refused where a backward pass would silently get no gradients.
        if not test_mode and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "test_mode=False with trainable parameters asks for the autograd graph: train with the reference "
                "graph and load the checkpoint here (under torch.no_grad() this call returns every iteration's prediction)")
        if not image1.is_cuda:
            raise L.DktError("RAFTStereo (B200 engine) needs CUDA inputs; there is no CPU fallback")
        if not test_mode:
            with torch.no_grad():
                if self.encoder is not None:
                    if self.encoder.pack_weights():
                        self._graphs.clear(); self._seen.clear(); self._full.clear(); self._full_seen.clear()
                    return {"disp_preds": self._forward_native_eager(image1, image2, iters, flow_init, all_preds=True)}
                fmap1, fmap2, net_list, ctx_list = self.extract(image1, image2)
                args, eng = self.args, self.engine
                B, D, h, w = fmap1.shape
                if eng.pack_weights():
                    self._graphs.clear(); self._seen.clear()
                eng.allocate(B, h, w, fmap1.device)
                self._ensure_volume(B, D, h, w, fmap1.device)
                ops.corr1d_build(fmap1, fmap2, args.corr_levels, 1.0 / (D ** 0.5), impl=self.impl, pyr=self._pyr)
                eng.load_state(net_list, ctx_list)
                return {"disp_preds": self._loop_and_upsample(B, h, w, iters, flow_init, all_preds=True)}
        with torch.no_grad():
            if self.encoder is not None:
                return self.forward_native(image1, image2, iters, flow_init)
            fmap1, fmap2, net_list, ctx_list = self.extract(image1, image2)
            return self.hot_path(fmap1, fmap2, net_list, ctx_list, iters, flow_init)
