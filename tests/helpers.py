"""Shared test helpers (golden loading, state-dict synthesis)."""
import json
import os

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
