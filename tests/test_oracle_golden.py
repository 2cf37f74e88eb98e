"""The oracle restatement against the committed outputs of the reference itself.

CPU-only.  tests/golden/*.npz were produced by tests/golden/make_golden.py from the
unmodified reference code; here oracle/mal_oracle.py must reproduce them: selection
indices exactly, floating-point fields to 1e-6 relative (they are bit-identical on the
machine that generated them; another CPU's BLAS may differ in the last ulp).
"""
import numpy as np
import pytest
import torch

from oracle import mal_oracle as O
from tests.helpers import load_npz, photometric_golden, rel_err

CASES = ["photometric_smooth.npz", "photometric_noise.npz"]


def _mono(inputs, t, temporal):
    H, W = inputs[("color", 0, 0)].shape[-2:]
    Ts = {f: t[("cam_T_cam", 0, f)].clone().requires_grad_(True) for f in (-1, 1)}
    out = {("disp", 0): t[("mono_disp", 0)].clone().requires_grad_(True)}
    for f in (-1, 1):
        out[("cam_T_cam", 0, f)] = Ts[f]
        out[("syn", f, 0)] = t[("syn", f, 0)]
    O.images_pred(inputs, out, height=H, width=W)
    losses, mono_reproj, aux = O.mono_losses(inputs, out, temporal, True, noise=t["noise"][0])
    return out, Ts, losses, mono_reproj, aux


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("temporal,tag", [(False, "plain"), (True, "temporal")])
def test_mono_losses_match_reference(name, temporal, tag):
    inputs, t, ref = photometric_golden(name)
    out, Ts, losses, mono_reproj, aux = _mono(inputs, t, temporal)
    assert np.array_equal(aux["frame_idx"].numpy().astype(np.uint8), ref[f"mono_{tag}_frame_idx"])
    assert np.array_equal(aux["automask"].numpy().astype(np.uint8), ref[f"mono_{tag}_automask"])
    assert rel_err(losses["loss"], ref[f"mono_{tag}_loss"]) < 1e-6
    assert rel_err(mono_reproj, ref[f"mono_{tag}_min_reproj"]) < 1e-6
    g = torch.autograd.grad(losses["loss"], [out[("disp", 0)], Ts[-1], Ts[1]])
    assert rel_err(g[0], ref[f"mono_{tag}_grad_disp"]) < 1e-5
    assert rel_err(g[1], ref[f"mono_{tag}_grad_T_-1"]) < 1e-5
    assert rel_err(g[2], ref[f"mono_{tag}_grad_T_1"]) < 1e-5
    for f in (-1, 1):
        assert rel_err(out[("color", f, 0)], ref[f"mono_color_{f}"]) < 1e-6


@pytest.mark.parametrize("name", CASES)
def test_main_losses_match_reference(name):
    inputs, t, ref = photometric_golden(name)
    H, W = inputs[("color", 0, 0)].shape[-2:]
    B = inputs[("color", 0, 0)].shape[0]
    mono_out, _, _, mono_reproj, _ = _mono(inputs, t, True)
    disp_ens = (t[("mono_disp", 0)] + t[("multi_disp", 0)]) / 2.0
    ens = O.images_pred_ensemble(inputs, t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)], disp_ens,
                                 height=H, width=W)
    assert rel_err(ens, ref["ensemble_reproj"]) < 1e-6
    multi = {("disp", 0): t[("multi_disp", 0)].clone().requires_grad_(True),
             "consistency_mask": t["consistency_mask"], "augmentation_mask": t["augmentation_mask"],
             ("mono_depth", 0, 0): mono_out[("depth", 0, 0)].detach(), "lowest_cost": t["lowest_cost"]}
    for f in (-1, 1):
        multi[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
        multi[("syn", f, 0)] = t[("syn", f, 0)]
    O.images_pred(inputs, multi, height=H, width=W, is_multi=True)
    assert np.array_equal(O.matching_mask(multi).numpy().astype(np.uint8), ref["matching_mask"])
    for ens_t, has_ins, blc, tag in ((ens, False, True, "ens_blc"), (None, True, False, "noens_ins")):
        losses, _, ll, aux = O.main_losses(inputs, multi, mono_reproj.detach(), ens_t, batch_size=B,
                                           multi_has_ins=has_ins, loss_blc=blc, noise=t["noise"][1])
        for k in ("loss", "distil_loss", "reproj_loss/0", "consistency_loss/0"):
            assert rel_err(losses[k], ref[f"main_{tag}_{k.replace('/', '_')}"]) < 1e-6, k
        total = losses["loss"] + (0.25 * ll[1] if blc else 0)
        g, = torch.autograd.grad(total, multi[("disp", 0)], retain_graph=True)
        assert rel_err(g, ref[f"main_{tag}_grad_disp"]) < 1e-5
        assert rel_err(aux["consistency_target"], ref[f"main_{tag}_consistency_target"]) < 1e-6


def test_cost_volume_matches_reference():
    g = load_npz("cost_volume.npz")
    tt = lambda k: torch.from_numpy(g[k].copy())
    vol, miss = O.match_features(tt("in_current_feats"), tt("in_lookup_feats"), tt("in_relative_poses"),
                                 tt("in_K"), tt("in_inv_K"), tt("in_bins"))
    assert np.array_equal(miss.numpy().astype(np.uint8), g["ref_missing"])
    assert rel_err(vol, g["ref_cost_volume"]) < 1e-6
    conf = O.confidence_mask(vol * (1 - miss))
    assert np.array_equal(conf.numpy(), g["ref_confidence"])
    low, idx = O.lowest_cost(vol, tt("in_bins"))
    assert np.array_equal(idx.numpy().astype(np.int32), g["ref_argmin"])
    assert rel_err(low, g["ref_lowest_cost"]) < 1e-6


def test_forward_warp_matches_reference():
    """dynamicdepth/rigid_warp.forward_warp (third-party coalesce restated, see make_golden.py)."""
    g = load_npz("forward_warp.npz")
    tt = lambda k: torch.from_numpy(g[k].copy())
    img_w, depth_w, valid = O.forward_warp(tt("in_img"), tt("in_depth"), tt("in_pose"), tt("in_K"), 3)
    assert np.array_equal(valid.numpy().astype(np.uint8), g["ref_valid"])
    assert rel_err(depth_w, g["ref_depth_w"]) < 1e-6
    assert rel_err(img_w, g["ref_img_w"]) < 1e-6
    mats = O.forward_warp_matrices(tt("in_pose"), tt("in_K"), 3)
    for m, k in zip(mats, ("in_Ku_inv", "in_K_inv", "in_proj")):
        assert rel_err(m, g[k]) < 1e-6
