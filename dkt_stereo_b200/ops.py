"""Functional wrappers over the C ABI (one Python function per entry point family).

Tensors are NHWC unless a wrapper says otherwise.  ``impl`` selects between the two CUDA
implementations that exist for the GEMM-shaped kernels: ``"tc"`` (tcgen05 tensor cores,
3-term bf16 split) and ``"simt"`` (exact fp32 CUDA cores).  Neither is a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from ._lib import DktEpilogue, DktTensor, tensor_slice, null_tensor


def split16(x: torch.Tensor):
    """x (fp32) -> (hi, lo) in the library's 16-bit format (``_lib.split_dtype()``: half by default) with
    hi = rn16(x), lo = rn16(x - hi) -- the same rule as csrc/common.cuh."""
    dt = L.split_dtype()
    hi = x.to(dt)
    lo = (x - hi.float()).to(dt)
    return hi, lo


split_bf16 = split16          # round-1 name


# ---------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------
def pyramid_widths(w2: int, levels: int) -> List[int]:
    out = []
    for _ in range(levels):
        out.append(w2)
        w2 //= 2
    return out


def alloc_pyramid(B: int, H: int, W1: int, W2: int, levels: int, device) -> List[torch.Tensor]:
    return [torch.empty(B, H, W1, w, device=device, dtype=torch.float32) for w in pyramid_widths(W2, levels)]


def corr1d_build(fmap1: torch.Tensor, fmap2: torch.Tensor, levels: int, scale: float,
                 impl: str = "tc", pyr: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
    """fmap (B,D,H,W) fp32 logical NCHW with any strides -> pyramid [(B,H,W1,W2>>l)]."""
    L.require_device(fmap1)
    lib = L.load()
    assert fmap1.dtype == torch.float32 and fmap2.dtype == torch.float32
    B, D, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    if fmap2.stride() != fmap1.stride():
        fmap2 = fmap2.contiguous()
        fmap1 = fmap1.contiguous()
    if pyr is None:
        pyr = alloc_pyramid(B, H, W1, W2, levels, fmap1.device)
    ptrs = L.pointer_array(pyr)
    sb, sd, sh, sw = fmap1.stride()
    sb2, sd2, sh2, sw2 = fmap2.stride()        # W2 != W1: the right map has its own batch / channel / row strides
    if impl == "tc" and D % 8 == 0:
        hi1 = torch.empty(B, H, W1, D, device=fmap1.device, dtype=L.split_dtype())
        lo1, hi2, lo2 = torch.empty_like(hi1), torch.empty(B, H, W2, D, device=fmap1.device, dtype=L.split_dtype()), None
        lo2 = torch.empty_like(hi2)
        s = L.stream_ptr()
        L.check(lib.dkt_split_nchw_to_nhwc_bf16x2(fmap1.data_ptr(), sb, sd, sh, sw, hi1.data_ptr(), lo1.data_ptr(),
                                                  B, D, H, W1, s), "split fmap1")
        L.check(lib.dkt_split_nchw_to_nhwc_bf16x2(fmap2.data_ptr(), sb2, sd2, sh2, sw2, hi2.data_ptr(), lo2.data_ptr(),
                                                  B, D, H, W2, s), "split fmap2")
        L.check(lib.dkt_corr1d_build_tc(hi1.data_ptr(), lo1.data_ptr(), hi2.data_ptr(), lo2.data_ptr(), ptrs,
                                        B, D, H, W1, W2, levels, float(scale), s), "corr1d_build_tc")
    else:
        if (sb2, sd2, sh2, sw2) != (sb, sd, sh, sw):
            raise L.DktError("corr1d_build (fp32 kernel): dkt_corr1d_build_f32 takes ONE stride set for both feature maps; "
                             "W2 != W1 is served by the tensor-core path (impl='tc', D % 8 == 0)")
        L.check(lib.dkt_corr1d_build_f32(fmap1.data_ptr(), fmap2.data_ptr(), sb, sd, sh, sw, ptrs,
                                         B, D, H, W1, W2, levels, float(scale), L.stream_ptr()), "corr1d_build_f32")
    return pyr


def corr1d_build_split(hi1: torch.Tensor, lo1: torch.Tensor, hi2: torch.Tensor, lo2: torch.Tensor,
                       levels: int, scale: float, pyr: List[torch.Tensor]) -> List[torch.Tensor]:
    """K1 straight from NHWC (B,H,W,D) bf16 (hi, lo) feature maps (what the encoder engine writes)."""
    lib = L.load()
    B, H, W1, D = hi1.shape
    W2 = hi2.shape[2]
    for t in (hi1, lo1, hi2, lo2):
        assert t.is_contiguous() and t.dtype == L.split_dtype()
    L.check(lib.dkt_corr1d_build_tc(hi1.data_ptr(), lo1.data_ptr(), hi2.data_ptr(), lo2.data_ptr(), L.pointer_array(pyr),
                                    B, D, H, W1, W2, levels, float(scale), L.stream_ptr()), "corr1d_build_tc")
    return pyr


# ---------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------
def corr1d_lookup(pyr: Sequence[torch.Tensor], coords_x: torch.Tensor, radius: int,
                  out: Optional[torch.Tensor], out_layout: str = "nhwc",
                  out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None,
                  delta: Optional[torch.Tensor] = None, flow: Optional[torch.Tensor] = None) -> None:
    """coords_x (B,H,W) fp32, updated in place by delta[...,0] when given.  out: NHWC (B,H,W,Cpad)
    or NCHW (B,C,H,W); None -> only the coordinate / flow bookkeeping runs."""
    lib = L.load()
    L.require_device(coords_x)
    B, H, W1 = coords_x.shape
    levels = len(pyr) if out is not None else 0
    W2 = pyr[0].shape[-1] if levels else 1
    if out is not None:
        if out_layout == "nhwc":
            Cp = out.shape[-1]
            ob, oc, op = H * W1 * Cp, 1, Cp
        else:
            Cc = out.shape[1]
            ob, oc, op = Cc * H * W1, H * W1, 1
    else:
        ob = oc = op = 0
    ptrs = L.pointer_array(pyr) if levels else None
    L.check(lib.dkt_corr1d_lookup(ptrs, levels, radius, coords_x.data_ptr(),
                                  L.ptr(delta), delta.shape[-1] if delta is not None else 0, L.ptr(flow),
                                  L.ptr(out), L.ptr(out_hi), L.ptr(out_lo), ob, oc, op,
                                  B, H, W1, W2, L.stream_ptr()), "corr1d_lookup")


def geo_pool(gev: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
    """(B,C,D,H,W) -> (B,H,W,C,D), (B,H,W,C,D//2); ``out`` reuses existing level buffers."""
    lib = L.load()
    L.require_device(gev)
    gev = gev.contiguous().float()
    B, Cc, D, H, W = gev.shape
    if out is not None:
        g0, g1 = out
        assert g0.shape == (B, H, W, Cc, D) and g1.shape == (B, H, W, Cc, D // 2)
    else:
        g0 = torch.empty(B, H, W, Cc, D, device=gev.device, dtype=torch.float32)
        g1 = torch.empty(B, H, W, Cc, D // 2, device=gev.device, dtype=torch.float32)
    L.check(lib.dkt_geo_pool(gev.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, Cc, D, H, W, L.stream_ptr()), "geo_pool")
    return g0, g1


def gwc_volume(left: torch.Tensor, right: torch.Tensor, maxdisp: int, groups: int,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Group-wise correlation volume (B,groups,maxdisp,H,W) of two (B,C,H,W) fp32 feature maps
    (reference meta_arch/igev_stereo/submodule.py:152-170)."""
    lib = L.load()
    L.require_device(left)
    assert left.shape == right.shape and left.dtype == torch.float32 and right.dtype == torch.float32
    left, right = left.contiguous(), right.contiguous()
    B, C, H, W = left.shape
    if out is None:
        out = torch.empty(B, groups, maxdisp, H, W, device=left.device, dtype=torch.float32)
    assert out.shape == (B, groups, maxdisp, H, W) and out.is_contiguous()
    L.check(lib.dkt_gwc_volume(left.data_ptr(), right.data_ptr(), out.data_ptr(), B, C, groups, maxdisp, H, W,
                               L.stream_ptr()), "gwc_volume")
    return out


def conv3d_c8(x: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor] = None,
              shift: Optional[torch.Tensor] = None, slope: float = 1.0, att: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3x3 Conv3d, 8 -> 8 or 1 channels, padding 1, no bias, then ``leaky(acc * scale + shift, slope) *
    sigmoid(att)``: x (B,8,D,H,W), weight (CO,8,3,3,3), scale/shift (CO,), att (B,CO,H,W) logits
    (reference BasicConv / FeatureAtt / classifier, meta_arch/igev_stereo/submodule.py:10-36,227-240)."""
    lib = L.load()
    L.require_device(x)
    B, Cin, D, H, W = x.shape
    CO = weight.shape[0]
    assert Cin == 8 and tuple(weight.shape) == (CO, 8, 3, 3, 3) and x.dtype == torch.float32
    x, weight = x.contiguous(), weight.contiguous().float()
    if out is None:
        out = torch.empty(B, CO, D, H, W, device=x.device, dtype=torch.float32)
    assert out.shape == (B, CO, D, H, W) and out.is_contiguous()
    keep = [t.contiguous().float() if t is not None else None for t in (scale, shift, att)]
    if keep[2] is not None:
        assert keep[2].shape == (B, CO, H, W)
    ptr = [t.data_ptr() if t is not None else None for t in keep]
    L.check(lib.dkt_conv3d_c8(x.data_ptr(), weight.data_ptr(), ptr[0], ptr[1], ptr[2], float(slope), out.data_ptr(),
                              B, CO, D, H, W, L.stream_ptr()), "conv3d_c8")
    return out


def _opt_ptrs(*ts):
    keep = [t.contiguous().float() if t is not None else None for t in ts]
    return keep, [t.data_ptr() if t is not None else None for t in keep]


def conv3d_k3(x: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor] = None,
              shift: Optional[torch.Tensor] = None, slope: float = 1.0, att: Optional[torch.Tensor] = None,
              stride: int = 1) -> torch.Tensor:
    """3x3x3 Conv3d (padding 1, stride 1 | 2, no bias) + ``leaky(acc * scale + shift, slope) * sigmoid(att)``:
    x (B,CI,D,H,W) fp32, weight (CO,CI,3,3,3), att (B,CO,Ho,Wo) logits (reference BasicConv / FeatureAtt,
    meta_arch/igev_stereo/submodule.py:10-36,227-240)."""
    lib = L.load()
    L.require_device(x)
    B, CI, D, H, W = x.shape
    CO = weight.shape[0]
    assert tuple(weight.shape) == (CO, CI, 3, 3, 3) and x.dtype == torch.float32
    x, weight = x.contiguous(), weight.contiguous().float()
    Do, Ho, Wo = ((v - 1) // stride + 1 for v in (D, H, W))
    out = torch.empty(B, CO, Do, Ho, Wo, device=x.device, dtype=torch.float32)
    keep, ptr = _opt_ptrs(scale, shift, att)
    if keep[2] is not None:
        assert keep[2].shape == (B, CO, Ho, Wo)
    L.check(lib.dkt_conv3d_k3(x.data_ptr(), weight.data_ptr(), ptr[0], ptr[1], ptr[2], float(slope), out.data_ptr(),
                              B, CI, CO, D, H, W, stride, L.stream_ptr()), "conv3d_k3")
    return out


def deconv3d_k4s2(x: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor] = None,
                  shift: Optional[torch.Tensor] = None, slope: float = 1.0) -> torch.Tensor:
    """ConvTranspose3d(kernel 4, stride 2, padding 1, no bias) + ``leaky(acc * scale + shift, slope)``:
    x (B,CI,D,H,W), weight (CI,CO,4,4,4) -> (B,CO,2D,2H,2W) (reference hourglass conv*_up, igev_stereo.py:42-49)."""
    lib = L.load()
    L.require_device(x)
    B, CI, D, H, W = x.shape
    CO = weight.shape[1]
    assert tuple(weight.shape) == (CI, CO, 4, 4, 4) and x.dtype == torch.float32
    x, weight = x.contiguous(), weight.contiguous().float()
    out = torch.empty(B, CO, 2 * D, 2 * H, 2 * W, device=x.device, dtype=torch.float32)
    keep, ptr = _opt_ptrs(scale, shift)
    L.check(lib.dkt_deconv3d_k4s2(x.data_ptr(), weight.data_ptr(), ptr[0], ptr[1], float(slope), out.data_ptr(),
                                  B, CI, CO, D, H, W, L.stream_ptr()), "deconv3d_k4s2")
    return out


def conv3d_k1(x0: torch.Tensor, x1: Optional[torch.Tensor], weight: torch.Tensor, scale: Optional[torch.Tensor] = None,
              shift: Optional[torch.Tensor] = None, slope: float = 1.0, att: Optional[torch.Tensor] = None) -> torch.Tensor:
    """1x1x1 Conv3d over cat(x0, x1) (x1 may be None) with the epilogue of ``conv3d_k3``: weight (CO, C0+C1[,1,1,1])."""
    lib = L.load()
    L.require_device(x0)
    B, C0, D, H, W = x0.shape
    C1 = 0 if x1 is None else x1.shape[1]
    CO = weight.shape[0]
    weight = weight.reshape(CO, -1).contiguous().float()
    assert weight.shape[1] == C0 + C1 and x0.dtype == torch.float32
    x0 = x0.contiguous()
    if x1 is not None:
        assert x1.shape == (B, C1, D, H, W) and x1.dtype == torch.float32
        x1 = x1.contiguous()
    out = torch.empty(B, CO, D, H, W, device=x0.device, dtype=torch.float32)
    keep, ptr = _opt_ptrs(scale, shift, att)
    L.check(lib.dkt_conv3d_k1(x0.data_ptr(), C0, x1.data_ptr() if x1 is not None else None, C1, weight.data_ptr(),
                              ptr[0], ptr[1], ptr[2], float(slope), out.data_ptr(), B, CO, D, H, W, L.stream_ptr()),
            "conv3d_k1")
    return out


def softargmin(logits: torch.Tensor) -> torch.Tensor:
    """(B,D,H,W) fp32 -> (B,1,H,W) = sum_d d * softmax_d (reference submodule.py:220-224 after F.softmax)."""
    lib = L.load()
    L.require_device(logits)
    assert logits.dtype == torch.float32
    logits = logits.contiguous()
    B, D, H, W = logits.shape
    out = torch.empty(B, 1, H, W, device=logits.device, dtype=torch.float32)
    L.check(lib.dkt_softargmin(logits.data_ptr(), out.data_ptr(), B, D, H, W, L.stream_ptr()), "softargmin")
    return out


def corr1d_lookup_enc(pyr: Sequence[torch.Tensor], coords_x: torch.Tensor, radius: int, w: "ConvWeights",
                      enc_out: DktTensor, delta: Optional[torch.Tensor] = None,
                      flow: Optional[torch.Tensor] = None) -> None:
    """Lookup fused with the motion encoder's 1x1 ``convc1`` + ReLU (exact fp32); ``w`` is the packed
    convc1 (``w_simt`` [1][Cin_pad][64], bias [64]); ``enc_out`` a 64-channel NHWC slice."""
    lib = L.load()
    L.require_device(coords_x)
    B, H, W1 = coords_x.shape
    assert w.ksize == 1 and w.n == 64 and w.bias is not None
    L.check(lib.dkt_corr1d_lookup_enc(L.pointer_array(pyr), len(pyr), radius, coords_x.data_ptr(),
                                      L.ptr(delta), delta.shape[-1] if delta is not None else 0, L.ptr(flow),
                                      w.w_simt.data_ptr(), w.bias.data_ptr(), C.byref(enc_out),
                                      B, H, W1, pyr[0].shape[-1], L.stream_ptr()), "corr1d_lookup_enc")


def pack_lookup_tc(weight: torch.Tensor, bias: torch.Tensor):
    """convc1 (64, C, 1, 1) -> (w_img, bias): the tensor-core lookup's shared-memory image of the weights,
    [2 planes (hi, lo)][KB][64 rows x 64 k] 16-bit in K-major SWIZZLE_128B order (csrc/lookup_tc.cu sw128_off)."""
    dev, weight, bias = _pack_src(weight, bias)
    N, C = weight.shape[0], weight.shape[1]
    assert N == 64 and C % 9 == 0
    w = weight.detach().float().reshape(N, C)
    hi, lo = split16(w)
    n = torch.arange(N, device=w.device).view(N, 1).expand(N, C)
    c = torch.arange(C, device=w.device).view(1, C).expand(N, C)
    k = (c // 9) * 10 + c % 9              # the kernel gives every row sample 10 K slots: its 9 taps + one zero (LT_TS)
    KB = ((C // 9) * 10 + 63) // 64
    off = (k >> 6) * (64 * 128) + (n >> 3) * 1024 + (n & 7) * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2
    idx = (off // 2).reshape(-1)
    img = torch.zeros(2, KB * 64 * 64, dtype=hi.dtype, device=w.device)
    img[0, idx] = hi.reshape(-1)
    img[1, idx] = lo.reshape(-1)
    return _to(img.contiguous(), dev), _to(bias.detach().float().contiguous(), dev)


def corr1d_lookup_enc_tc(pyr: Sequence[torch.Tensor], coords_x: torch.Tensor, radius: int, w_img: torch.Tensor,
                         bias: torch.Tensor, enc_out: DktTensor, tap_planes: int = 2,
                         delta: Optional[torch.Tensor] = None, flow: Optional[torch.Tensor] = None) -> None:
    """corr1d_lookup_enc with the 1x1 ``convc1`` on tensor cores (csrc/lookup_tc.cu)."""
    lib = L.load()
    L.require_device(coords_x)
    B, H, W1 = coords_x.shape
    L.check(lib.dkt_corr1d_lookup_enc_tc(L.pointer_array(pyr), len(pyr), radius, coords_x.data_ptr(),
                                         L.ptr(delta), delta.shape[-1] if delta is not None else 0, L.ptr(flow),
                                         w_img.data_ptr(), bias.data_ptr(), C.byref(enc_out), tap_planes,
                                         B, H, W1, pyr[0].shape[-1], L.stream_ptr()), "corr1d_lookup_enc_tc")


def geo_pool_dc(gev: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
    """(B,C,D,H,W) -> (B,H,W,D,C), (B,H,W,D//2,C): the layout ``geo_lookup_enc_tc`` reads."""
    lib = L.load()
    L.require_device(gev)
    gev = gev.contiguous().float()
    B, Cc, D, H, W = gev.shape
    if out is not None:
        g0, g1 = out
        assert g0.shape == (B, H, W, D, Cc) and g1.shape == (B, H, W, D // 2, Cc)
    else:
        g0 = torch.empty(B, H, W, D, Cc, device=gev.device, dtype=torch.float32)
        g1 = torch.empty(B, H, W, D // 2, Cc, device=gev.device, dtype=torch.float32)
    L.check(lib.dkt_geo_pool_dc(gev.data_ptr(), g0.data_ptr(), g1.data_ptr(), B, Cc, D, H, W, L.stream_ptr()), "geo_pool_dc")
    return g0, g1


def geo_lookup_enc_tc(geo: Sequence[torch.Tensor], init: Sequence[torch.Tensor], disp: torch.Tensor, radius: int,
                      w_img: torch.Tensor, bias: torch.Tensor, enc_out: DktTensor, tap_planes: int = 1,
                      delta: Optional[torch.Tensor] = None) -> None:
    """geo_lookup_enc on tensor cores; ``geo`` in the (B,H,W,D,C) layout of ``geo_pool_dc``."""
    lib = L.load()
    B, H, W, D, Cc = geo[0].shape
    L.check(lib.dkt_geo_lookup_enc_tc(geo[0].data_ptr(), geo[1].data_ptr(), init[0].data_ptr(), init[1].data_ptr(),
                                      disp.data_ptr(), L.ptr(delta), delta.shape[-1] if delta is not None else 0,
                                      radius, Cc, D, w_img.data_ptr(), bias.data_ptr(), C.byref(enc_out), tap_planes,
                                      B, H, W, L.stream_ptr()), "geo_lookup_enc_tc")


def geo_lookup(geo: Sequence[torch.Tensor], init: Sequence[torch.Tensor], disp: torch.Tensor, radius: int,
               out: torch.Tensor, out_layout: str = "nhwc",
               out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None,
               delta: Optional[torch.Tensor] = None) -> None:
    """disp (B,H,W) fp32, updated in place by delta[...,0] (NHWC) when given."""
    lib = L.load()
    B, H, W, Cc, D = geo[0].shape
    if out_layout == "nhwc":
        Cp = out.shape[-1]
        ob, oc, op = H * W * Cp, 1, Cp
    else:
        ob, oc, op = out.shape[1] * H * W, H * W, 1
    L.check(lib.dkt_geo_lookup(geo[0].data_ptr(), geo[1].data_ptr(), init[0].data_ptr(), init[1].data_ptr(),
                               disp.data_ptr(), L.ptr(delta), delta.shape[-1] if delta is not None else 0,
                               radius, Cc, D, out.data_ptr(), L.ptr(out_hi), L.ptr(out_lo),
                               ob, oc, op, B, H, W, L.stream_ptr()), "geo_lookup")


def geo_lookup_enc(geo: Sequence[torch.Tensor], init: Sequence[torch.Tensor], disp: torch.Tensor, radius: int,
                   w: "ConvWeights", enc_out: DktTensor, delta: Optional[torch.Tensor] = None) -> None:
    lib = L.load()
    B, H, W, Cc, D = geo[0].shape
    assert w.ksize == 1 and w.n == 64 and w.bias is not None
    L.check(lib.dkt_geo_lookup_enc(geo[0].data_ptr(), geo[1].data_ptr(), init[0].data_ptr(), init[1].data_ptr(),
                                   disp.data_ptr(), L.ptr(delta), delta.shape[-1] if delta is not None else 0,
                                   radius, Cc, D, w.w_simt.data_ptr(), w.bias.data_ptr(), C.byref(enc_out),
                                   B, H, W, L.stream_ptr()), "geo_lookup_enc")


# ---------------------------------------------------------------------------------------------
# K3
# ---------------------------------------------------------------------------------------------
@dataclass
class ConvWeights:
    """A conv layer repacked for the engine (done once per checkpoint load)."""
    ksize: int               # kh (== kw for the square filters of the update block)
    cin: int                 # padded input channels the kernel iterates over
    n: int                   # real output channels
    w_simt: torch.Tensor     # fp32 [taps][cin][n]
    w_hi: Optional[torch.Tensor]   # bf16 [taps][npad][cin]
    w_lo: Optional[torch.Tensor]
    bias: Optional[torch.Tensor]   # fp32 [n] (None when folded into a context term)
    kw: Optional[int] = None       # filter width when it differs from ksize (7x1 stem)
    stride: int = 1


def _pack_on_host() -> bool:
    """DKT_PACK_ON_HOST=1: repack weights with CPU arithmetic and upload the packed tensors (no GPU kernels at all
    for a checkpoint load; the default packs on the parameters' own device, which is what a per-step EMA-teacher
    update wants)."""
    return os.environ.get("DKT_PACK_ON_HOST", "0") == "1"


def _pack_src(*ts):
    """-> (device the packs must end up on, detached sources on the device the packing arithmetic runs on)"""
    dev = next(t for t in ts if t is not None).device
    if _pack_on_host():
        ts = tuple(t.detach().cpu() if t is not None else None for t in ts)
    else:
        ts = tuple(t.detach() if t is not None else None for t in ts)
    return (dev,) + ts


def _to(t: Optional[torch.Tensor], dev) -> Optional[torch.Tensor]:
    return t if t is None or t.device == dev else t.to(dev)


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], cin_pad: Optional[int] = None,
              tc: bool = True, keep_bias: bool = True) -> ConvWeights:
    """weight (N, Cin, k, k) fp32 (PyTorch OIHW) -> engine layouts.  cin_pad zero-pads the
    reduction dimension (e.g. 36 correlation channels carried in a 64-channel buffer)."""
    dev, weight, bias = _pack_src(weight, bias)
    N, Cin, k, _ = weight.shape
    w = weight.detach().float()
    if cin_pad is not None and cin_pad > Cin:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin_pad - Cin))
        Cin = cin_pad
    taps = k * k
    w_tkn = w.permute(2, 3, 1, 0).reshape(taps, Cin, N).contiguous()          # [tap][cin][n]
    w_hi = w_lo = None
    if tc:
        npad = (N + 15) // 16 * 16
        w_tnk = w.permute(2, 3, 0, 1).reshape(taps, N, Cin)
        if npad != N:
            w_tnk = torch.nn.functional.pad(w_tnk, (0, 0, 0, npad - N))
        w_tnk = w_tnk.contiguous()
        w_hi, w_lo = split16(w_tnk)
        w_hi, w_lo = w_hi.contiguous(), w_lo.contiguous()
    b = bias.detach().float().contiguous() if (bias is not None and keep_bias) else None
    return ConvWeights(k, Cin, N, _to(w_tkn, dev), _to(w_hi, dev), _to(w_lo, dev), _to(b, dev))


def pack_conv_cat(weights: Sequence[torch.Tensor], tc: bool = True) -> ConvWeights:
    """Stack several convs with identical input along N (convz || convr -> one N=256 GEMM)."""
    return pack_conv(torch.cat([w.detach() for w in weights], dim=0), None, tc=tc)


def make_epilogue(kind: int, out: DktTensor, act: int = L.ACT_NONE, scale: float = 1.0,
                  bias: Optional[torch.Tensor] = None, ctx: Optional[torch.Tensor] = None, ctx_c0: int = 0,
                  z: Optional[DktTensor] = None, h: Optional[DktTensor] = None,
                  tail: Optional[torch.Tensor] = None, res: Optional[tuple] = None,
                  proj: Optional[torch.Tensor] = None, stats_partial: Optional[torch.Tensor] = None) -> DktEpilogue:
    e = DktEpilogue()
    e.kind, e.act, e.scale = kind, act, scale
    e.bias = L.ptr(bias)
    e.ctx = L.ptr(ctx)
    e.ctx_C = ctx.shape[-1] if ctx is not None else 0
    e.ctx_c0 = ctx_c0
    e.out = out
    e.z = z if z is not None else null_tensor()
    e.h = h if h is not None else null_tensor()
    e.tail = L.ptr(tail)
    e.tail_C = tail.shape[-1] if tail is not None else 0
    if res is not None:          # (tensor NHWC fp32, first channel) or ((hi, lo) NHWC bf16, first channel)
        if isinstance(res[0], tuple):
            e.res_hi, e.res_lo, e.res_C, e.res_c0 = L.ptr(res[0][0]), L.ptr(res[0][1]), res[0][0].shape[-1], res[1]
        else:
            e.res, e.res_C, e.res_c0 = L.ptr(res[0]), res[0].shape[-1], res[1]
    if proj is not None:         # fp32 [N][PROJ_LD], see pack_proj3x3
        assert proj.dtype == torch.float32 and proj.is_contiguous() and proj.shape[-1] == L.PROJ_LD
        e.proj = proj.data_ptr()
    if stats_partial is not None:    # fp32 [tiles][2][N]: per-tile sums written next to the conv output
        e.stats_partial = stats_partial.data_ptr()
    return e


def pack_proj3x3(weight: torch.Tensor, out_channel: int = 0) -> torch.Tensor:
    """conv weight (Nout, Cin, 3, 3) -> fp32 [Cin][PROJ_LD] with [c][ky*3+kx] = weight[out_channel, c, ky, kx]:
    the channel half of a one-output-channel 3x3 conv, applied by the DKT_EPI_PROJ epilogue of the conv that
    produces its input; the spatial half is ``tapsum3x3``."""
    dev, weight = _pack_src(weight)
    w = weight.detach().float()[out_channel]                     # (Cin, 3, 3)
    out = torch.zeros(w.shape[0], L.PROJ_LD, device=w.device, dtype=torch.float32)
    out[:, :9] = w.reshape(w.shape[0], 9)
    return _to(out.contiguous(), dev)


def pack_conv_general(weight: torch.Tensor, bias: Optional[torch.Tensor], *, stride: int = 1,
                      cin_pad: Optional[int] = None, n_pad: Optional[int] = None,
                      bn: Optional[Sequence[torch.Tensor]] = None, bn_eps: float = 1e-5) -> ConvWeights:
    """Encoder conv (N, Cin, kh, kw) -> tensor-core layout [kh*kw][Npad][Cin_pad] bf16 hi/lo.
    ``bn`` = (gamma, beta, running_mean, running_var) folds an eval-mode BatchNorm that follows the conv
    into weight and bias; ``cin_pad`` / ``n_pad`` zero-pad channels (96-channel stages run as 128)."""
    dev, weight, bias = _pack_src(weight, bias)
    w = weight.detach().double()
    N, Cin, kh, kw = w.shape
    b = bias.detach().double() if bias is not None else torch.zeros(N, dtype=torch.float64, device=w.device)
    if bn is not None:
        gamma, beta, mean, var = (t.detach().double().to(w.device) for t in bn)
        s = gamma / torch.sqrt(var + bn_eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (b - mean) * s + beta
    if cin_pad is not None and cin_pad > Cin:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin_pad - Cin))
        Cin = cin_pad
    if n_pad is not None and n_pad > N:
        w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, n_pad - N))
        b = torch.nn.functional.pad(b, (0, n_pad - N))
        N = n_pad
    w = w.float()
    npad = (N + 15) // 16 * 16
    w_tnk = w.permute(2, 3, 0, 1).reshape(kh * kw, N, Cin)
    if npad != N:
        w_tnk = torch.nn.functional.pad(w_tnk, (0, 0, 0, npad - N))
    w_hi, w_lo = split16(w_tnk.contiguous())
    return ConvWeights(kh, Cin, N, None, _to(w_hi.contiguous(), dev), _to(w_lo.contiguous(), dev),
                       _to(b.float().contiguous(), dev), kw=kw, stride=stride)


def conv2d_ex(srcs: Sequence[DktTensor], w: ConvWeights, epi: DktEpilogue, B: int, Hin: int, Win: int) -> tuple:
    """Tensor-core conv with kh x kw filter and stride; returns the output extent (H, W)."""
    lib = L.load()
    kh, kw, s = w.ksize, (w.kw or w.ksize), w.stride
    H = (Hin + 2 * (kh // 2) - kh) // s + 1
    W = (Win + 2 * (kw // 2) - kw) // s + 1
    n = len(srcs)
    arr = (DktTensor * n)(*srcs)
    assert sum(t.c_count for t in srcs) == w.cin, (sum(t.c_count for t in srcs), w.cin)
    L.check(lib.dkt_conv2d_tc_ex(arr, n, w.w_hi.data_ptr(), L.ptr(w.w_lo), kh, kw, s, w.n, C.byref(epi),
                                 B, Hin, Win, H, W, L.stream_ptr()), "conv2d_tc_ex")
    return H, W


def stem_rows(img: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor, kw: int = 7,
              scale: float = 2.0 / 255.0, shift: float = -1.0, layout: str = "nchw") -> None:
    """img fp32, (B,Cin,H,W) or NHWC (B,H,W,Cin) -> hi/lo (B,H,W,Cpad) bf16, channel kx*Cin + c (see the header)."""
    assert img.is_contiguous() and img.dtype == torch.float32
    if layout == "nchw":
        B, Cin, H, W = img.shape
        sb, sc, sy, sx = Cin * H * W, H * W, W, 1
    else:
        B, H, W, Cin = img.shape
        sb, sc, sy, sx = H * W * Cin, 1, W * Cin, Cin
    assert hi.shape == (B, H, W, hi.shape[-1])
    L.check(L.load().dkt_stem_rows_bf16x2(img.data_ptr(), sb, sc, sy, sx, scale, shift, hi.data_ptr(), lo.data_ptr(),
                                          B, Cin, H, W, kw, hi.shape[-1], L.stream_ptr()), "stem_rows")


def tapsum3x3(taps: torch.Tensor, bias: float, out: torch.Tensor) -> None:
    """taps NHWC (B,H,W,>=9) fp32 -> out[..., 0] (NHWC, any channel count) = bias + 3x3 spatial tap sum."""
    B, H, W, TCc = taps.shape
    L.check(L.load().dkt_tapsum3x3(taps.data_ptr(), TCc, float(bias), out.data_ptr(), out.shape[-1], B, H, W,
                                   L.stream_ptr()), "tapsum3x3")


def instnorm_workspace(B: int, Cc: int, device) -> torch.Tensor:
    return torch.empty(L.load().dkt_instnorm_workspace_floats(B, Cc), device=device, dtype=torch.float32)


def instnorm_stats(x: DktTensor, workspace: torch.Tensor, stats: torch.Tensor, B: int, H: int, W: int,
                   eps: float = 1e-5) -> None:
    L.check(L.load().dkt_instnorm_stats(C.byref(x), workspace.data_ptr(), stats.data_ptr(), eps, B, H, W,
                                        L.stream_ptr()), "instnorm_stats")


def instnorm_tiles_workspace(B: int, Cc: int, device) -> torch.Tensor:
    return torch.empty(L.load().dkt_instnorm_tiles_workspace_floats(B, Cc), device=device, dtype=torch.float32)


def conv_tiles(H: int, W: int) -> int:
    """Entries of dkt_epilogue.stats_partial per image: 8 x 16 output tiles of the tensor-core conv x 4 quarters."""
    return ((H + 7) // 8) * ((W + 15) // 16) * 4


def instnorm_finalize_tiles(partial: torch.Tensor, workspace: torch.Tensor, stats: torch.Tensor, B: int, Cc: int,
                            H: int, W: int, eps: float = 1e-5) -> None:
    """per-tile sums (conv epilogue, ``stats_partial``) -> stats (B,C,2) = (mean, rstd)."""
    L.check(L.load().dkt_instnorm_finalize_tiles(partial.data_ptr(), workspace.data_ptr(), stats.data_ptr(), eps,
                                                 B, Cc, H, W, L.stream_ptr()), "instnorm_finalize_tiles")


def instnorm_apply(x: DktTensor, stats: torch.Tensor, out: DktTensor, B: int, H: int, W: int,
                   relu=True, res: Optional[DktTensor] = None) -> None:
    """``relu``: False / True, or "leaky" for LeakyReLU(0.01)."""
    act = 2 if relu == "leaky" else int(bool(relu))
    L.check(L.load().dkt_instnorm_apply(C.byref(x), stats.data_ptr(), C.byref(res) if res is not None else None,
                                        C.byref(out), act, B, H, W, L.stream_ptr()), "instnorm_apply")


def conv2d(srcs: Sequence[DktTensor], w: ConvWeights, epi: DktEpilogue, B: int, H: int, W: int,
           impl: str = "tc") -> None:
    lib = L.load()
    n = len(srcs)
    arr = (DktTensor * n)(*srcs)
    cin = sum(s.c_count for s in srcs)
    assert cin == w.cin, (cin, w.cin)
    if impl == "tc":
        # w.w_lo None = single-plane weights: the x * w_lo MMA is dropped
        L.check(lib.dkt_conv2d_tc(arr, n, w.w_hi.data_ptr(), L.ptr(w.w_lo), w.ksize, w.n, C.byref(epi),
                                  B, H, W, L.stream_ptr()), "conv2d_tc")
    else:
        L.check(lib.dkt_conv2d_simt(arr, n, w.w_simt.data_ptr(), w.ksize, w.n, C.byref(epi),
                                    B, H, W, L.stream_ptr()), "conv2d_simt")


# ---------------------------------------------------------------------------------------------
# a8 / K4 / layout
# ---------------------------------------------------------------------------------------------
def pool2x(src: DktTensor, dst: DktTensor, B: int, Hs: int, Ws: int) -> None:
    Hd, Wd = (Hs - 1) // 2 + 1, (Ws - 1) // 2 + 1
    L.check(L.load().dkt_pool2x(C.byref(src), C.byref(dst), B, Hs, Ws, Hd, Wd, L.stream_ptr()), "pool2x")


def interp(src: DktTensor, dst: DktTensor, B: int, Hs: int, Ws: int, Hd: int, Wd: int) -> None:
    L.check(L.load().dkt_interp(C.byref(src), C.byref(dst), B, Hs, Ws, Hd, Wd, L.stream_ptr()), "interp")


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor, factor: int) -> torch.Tensor:
    """flow NHWC (B,H,W,Cf) (channel 0 used), mask NHWC (B,H,W,9*f*f) -> (B,1,f*H,f*W)."""
    B, H, W, Cf = flow.shape
    out = torch.empty(B, 1, H * factor, W * factor, device=flow.device, dtype=torch.float32)
    L.check(L.load().dkt_convex_upsample(flow.data_ptr(), Cf, mask.data_ptr(), out.data_ptr(), B, H, W, factor,
                                         L.stream_ptr()), "convex_upsample")
    return out


def context_upsample(disp: torch.Tensor, weights: torch.Tensor, in_scale: float = 4.0,
                     out_scale: float = 1.0) -> torch.Tensor:
    """disp (B,H,W) fp32, weights (B,9,4H,4W) NCHW (softmaxed) -> (B,1,4H,4W)."""
    B, H, W = disp.shape
    weights = weights.contiguous().float()
    out = torch.empty(B, 1, 4 * H, 4 * W, device=disp.device, dtype=torch.float32)
    L.check(L.load().dkt_context_upsample(disp.data_ptr(), weights.data_ptr(), out.data_ptr(), in_scale, out_scale,
                                          B, H, W, L.stream_ptr()), "context_upsample")
    return out


def deconv4x4s2_as_conv3x3(weight: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(kernel 4, stride 2, padding 1) weight (Cin, Cout, 4, 4) -> Conv2d(3x3, padding 1) weight
    (4*Cout, Cin, 3, 3) whose output channel (py*2+px)*Cout + co is the deconv's output at (2y+py, 2x+px):
    out[2y+py] = sum_iy in[iy] W[2y+py+1-2iy], i.e. per parity a 2-tap subset of the 4 kernel rows."""
    Cin, Cout = weight.shape[:2]
    tap = ({0: 3, 1: 1}, {1: 2, 2: 0})          # parity -> {dy (3x3 row, input row y+dy-1): ky (4x4 row)}
    w3 = torch.zeros(4 * Cout, Cin, 3, 3, dtype=weight.dtype, device=weight.device)
    for py in range(2):
        for px in range(2):
            g = (py * 2 + px) * Cout
            for dy, ky in tap[py].items():
                for dx, kx in tap[px].items():
                    w3[g:g + Cout, :, dy, dx] = weight[:, :, ky, kx].t()
    return w3


def pixel_shuffle2(src: torch.Tensor, group: int, dst: DktTensor, B: int, H: int, W: int) -> None:
    """src fp32 NHWC (B,H,W,Csrc), channel (py*2+px)*group + c  ->  dst slice at (B,2H,2W)."""
    L.check(L.load().dkt_pixel_shuffle2(src.data_ptr(), src.shape[-1], group, C.byref(dst), B, H, W, L.stream_ptr()),
            "pixel_shuffle2")


def context_upsample_logits(disp: torch.Tensor, logits: torch.Tensor, in_scale: float = 4.0,
                            out_scale: float = 1.0) -> torch.Tensor:
    """disp (B,H,W) fp32, logits fp32 NHWC (B,2H,2W,>=36) parity-major (see the header) -> (B,1,4H,4W)."""
    B, H, W = disp.shape
    assert logits.shape[:3] == (B, 2 * H, 2 * W) and logits.is_contiguous() and logits.dtype == torch.float32
    out = torch.empty(B, 1, 4 * H, 4 * W, device=disp.device, dtype=torch.float32)
    L.check(L.load().dkt_context_upsample_logits(disp.data_ptr(), logits.data_ptr(), logits.shape[-1], out.data_ptr(),
                                                 in_scale, out_scale, B, H, W, L.stream_ptr()), "context_upsample_logits")
    return out


def nchw_to_nhwc(src: torch.Tensor, dst: DktTensor, bias: Optional[torch.Tensor] = None) -> None:
    src = src.contiguous().float()
    B, Cc, H, W = src.shape
    L.check(L.load().dkt_nchw_to_nhwc(src.data_ptr(), L.ptr(bias), C.byref(dst), B, Cc, H, W, L.stream_ptr()),
            "nchw_to_nhwc")


def nhwc_to_nchw(src: DktTensor, B: int, H: int, W: int, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty(B, src.c_count, H, W, device=device, dtype=torch.float32)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == B * src.c_count * H * W
    L.check(L.load().dkt_nhwc_to_nchw(C.byref(src), out.data_ptr(), B, src.c_count, H, W, L.stream_ptr()),
            "nhwc_to_nchw")
    return out


# ---------------------------------------------------------------------------------------------
# optional per-launch CUDA-event profiler (used by bench.py for the per-kernel breakdown)
# ---------------------------------------------------------------------------------------------
class LaunchProfiler:
    """with LaunchProfiler() as p: ... ; p.summary() -> {name: (count, total_ms)} after a sync."""

    active: Optional["LaunchProfiler"] = None

    def __init__(self):
        self.records = []

    def __enter__(self):
        LaunchProfiler.active = self
        return self

    def __exit__(self, *exc):
        LaunchProfiler.active = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.records:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + a.elapsed_time(b))
        return out


def _profiled(namer):
    def deco(fn):
        def wrapper(*args, **kwargs):
            prof = LaunchProfiler.active
            if prof is None:
                return fn(*args, **kwargs)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = fn(*args, **kwargs)
            b.record()
            prof.records.append((namer(*args, **kwargs), a, b))
            return r
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


def _conv_name(srcs, w, epi, B, H, W, impl="tc"):
    kind = {0: "lin", 1: "gru_zr", 2: "gru_q", 3: "proj"}[epi.kind]
    return f"conv{w.ksize}x{w.ksize}_{w.cin}to{w.n}_{H}x{W}_{kind}_{impl}"


conv2d = _profiled(_conv_name)(conv2d)
conv2d_ex = _profiled(lambda srcs, w, epi, B, Hin, Win: f"enc_conv{w.ksize}x{w.kw or w.ksize}s{w.stride}_{w.cin}to{w.n}_{Hin}x{Win}")(conv2d_ex)
stem_rows = _profiled(lambda *a, **k: "stem_rows")(stem_rows)
tapsum3x3 = _profiled(lambda *a, **k: "tapsum3x3")(tapsum3x3)
instnorm_stats = _profiled(lambda *a, **k: "instnorm_stats")(instnorm_stats)
instnorm_apply = _profiled(lambda *a, **k: "instnorm_apply")(instnorm_apply)
instnorm_finalize_tiles = _profiled(lambda *a, **k: "instnorm_finalize_tiles")(instnorm_finalize_tiles)
corr1d_build = _profiled(lambda *a, **k: f"corr1d_build_{k.get('impl', a[4] if len(a) > 4 else 'tc')}")(corr1d_build)
corr1d_build_split = _profiled(lambda *a, **k: "corr1d_build_tc")(corr1d_build_split)
corr1d_lookup = _profiled(lambda pyr, *a, **k: "corr1d_lookup" if len(pyr) else "coords_update")(corr1d_lookup)
pool2x = _profiled(lambda *a, **k: "pool2x")(pool2x)
interp = _profiled(lambda *a, **k: "interp")(interp)
convex_upsample = _profiled(lambda *a, **k: "convex_upsample")(convex_upsample)
nchw_to_nhwc = _profiled(lambda *a, **k: "nchw_to_nhwc")(nchw_to_nhwc)
pixel_shuffle2 = _profiled(lambda *a, **k: "pixel_shuffle2")(pixel_shuffle2)
context_upsample_logits = _profiled(lambda *a, **k: "context_upsample_logits")(context_upsample_logits)
geo_lookup = _profiled(lambda *a, **k: "geo_lookup")(geo_lookup)
corr1d_lookup_enc = _profiled(lambda *a, **k: "corr1d_lookup_enc")(corr1d_lookup_enc)
corr1d_lookup_enc_tc = _profiled(lambda *a, **k: "corr1d_lookup_enc_tc")(corr1d_lookup_enc_tc)
geo_lookup_enc_tc = _profiled(lambda *a, **k: "geo_lookup_enc_tc")(geo_lookup_enc_tc)
geo_lookup_enc = _profiled(lambda *a, **k: "geo_lookup_enc")(geo_lookup_enc)
