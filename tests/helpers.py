"""Shared test helpers (golden loading, state-dict synthesis)."""
import json
import os
import re

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RAFT_CFG = dict(model="RAFTStereo", loss_func="sequence_loss_raft", backbone_type="default",
                corr_implementation="reg", shared_backbone=False, corr_levels=4, corr_radius=4,
                n_downsample=2, context_norm="batch", slow_fast_gru=False, n_gru_layers=3,
                hidden_dims=[128, 128, 128])
IGEV_CFG = dict(model="IGEVStereo", loss_func="sequence_loss_raft", corr_levels=2, corr_radius=4,
                n_downsample=2, context_norm="batch", slow_fast_gru=False, n_gru_layers=3,
                hidden_dims=[128, 128, 128], max_disp=192)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        out = {}
        for k in z.files:
            a = z[k]
            out[k] = torch.from_numpy(a) if a.dtype.kind == "f" else a
        return out


def golden_shapes(g):
    return {str(k): tuple(json.loads(str(s))) for k, s in zip(g["keys"], g["key_shapes"])}


def stats(a, b):
    d = (a.double() - b.double()).abs()
    return float(d.mean()), float(d.max())


def golden_state_dict(g, prefix="sd__"):
    """state dict stored in a golden file as sd__<key with '.' -> '__'> arrays."""
    return {k[len(prefix):].replace("__", "."): v for k, v in g.items() if k.startswith(prefix)}


def golden_seeds(g):
    """(weight seed, image seed) a forward golden was made with (older files: 0, 1234)."""
    return tuple(int(v) for v in g["seeds"]) if "seeds" in g else (0, 1234)


# torchvision MobileNetV2 parameter names (what oracle/make_golden.py's timm stub yields, and therefore the names the
# IGEV forward goldens were seeded by) -> timm 0.5.4 names (what the reference checkpoints and the drop-in use)
_TV_DS = {"conv.0.0": "conv_dw", "conv.0.1": "bn1", "conv.1": "conv_pw", "conv.2": "bn2"}
_TV_IR = {"conv.0.0": "conv_pw", "conv.0.1": "bn1", "conv.1.0": "conv_dw", "conv.1.1": "bn2", "conv.2": "conv_pwl", "conv.3": "bn3"}


def tv_to_timm(key: str) -> str:
    mt = re.match(r"(feature\.block(\d)\.\d+\.\d+\.)(conv\.\d(?:\.\d)?)\.(.*)", key)
    if not mt:
        return key
    table = _TV_DS if mt.group(2) == "0" else _TV_IR
    return mt.group(1) + table[mt.group(3)] + "." + mt.group(4)
