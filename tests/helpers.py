"""Shared test helpers: golden fixture loading and comparison utilities."""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    with np.load(os.path.join(GOLDEN_DIR, name)) as z:
        return {k: z[k] for k in z.files}


def photometric_golden(name):
    """Rebuild the reference-keyed `inputs` / network tensors `t` from a golden file."""
    g = load_npz(name)
    tt = lambda k: torch.from_numpy(g[k].copy())
    inputs = {("color", f, 0): tt(f"in_color_{f}") for f in (0, -1, 1)}
    inputs[("K", 0)], inputs[("inv_K", 0)] = tt("in_K"), tt("in_inv_K")
    t = {("cam_T_cam", 0, f): tt(f"in_T_{f}") for f in (-1, 1)}
    for f in (-1, 1):
        t[("syn", f, 0)] = tt(f"in_syn_{f}")
    t[("mono_disp", 0)], t[("multi_disp", 0)] = tt("in_mono_disp"), tt("in_multi_disp")
    t["noise"] = [tt("in_noise_mono"), tt("in_noise_main")]
    t["consistency_mask"], t["augmentation_mask"] = tt("in_consistency_mask"), tt("in_augmentation_mask")
    t["lowest_cost"] = tt("in_lowest_cost")
    ref = {k[4:]: v for k, v in g.items() if k.startswith("ref_")}
    return inputs, t, ref


def rel_err(a, b):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
