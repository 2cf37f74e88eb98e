"""BASELINE.json configs 3-5 at their real sizes on the GPU, against the oracle (each oracle call
takes a few seconds on the host):

  config 3  ManyDepth+MAL Cityscapes 192x512 with synthetic Mask2Former-shaped motion masks:
            warps -> temporal-hint synthesis -> compute_mono_losses (classic path, syn gradients)
  config 4  DualRefine+MAL KITTI 192x640: per-(scale, deq_iter) losses, half-pixel convention
  config 5  DynamicDepth+MAL Cityscapes 192x512: forward_warp (upscale 3), the DynamicDepth cost
            volume (cv_min + 3-D max-pool fill) at 96 bins x 64 channels x 48x128
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mal_b200 import dyn_utils, layers, loss_utils, matching, rigid_warp, trainer_ops
from mal_b200.utils.synthetic import (CITYSCAPES_K, make_cost_volume_inputs, make_instance_masks,
                                      make_photometric_inputs, to_device)
from oracle import mal_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _gerr(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    s = float(b.abs().max())
    return float((a - b).abs().max()) / (s if s > 0 else 1.0)


def test_config3_cityscapes_temporal_hint_step():
    B, H, W = 2, 192, 512
    inputs, t = make_photometric_inputs(B, H, W, seed=301, normalised_K=CITYSCAPES_K, translation_scale=0.2)
    masks = [make_instance_masks(7, H, W, seed=310 + b, max_shift=12) for b in range(B)]
    res = []
    for who, dv in (("oracle", torch.device("cpu")), ("ours", DEV)):
        inp = {k: v.to(dv) for k, v in inputs.items()}
        disp = t[("mono_disp", 0)].clone().to(dv).requires_grad_(True)
        Ts = {f: t[("cam_T_cam", 0, f)].clone().to(dv).requires_grad_(True) for f in (-1, 1)}
        o = {("disp", 0): disp, ("cam_T_cam", 0, -1): Ts[-1], ("cam_T_cam", 0, 1): Ts[1]}
        if who == "oracle":
            O.images_pred(inp, o, height=H, width=W)
            syn = [[], []]
            for b in range(B):
                sl, sn, _ = O.generate_dynamic_instance(masks[b][0], masks[b][1], o[("color", -1, 0)][b],
                                                        o[("color", 1, 0)][b], False)
                syn[0].append(sl)
                syn[1].append(sn)
            o[("syn", -1, 0)], o[("syn", 1, 0)] = torch.stack(syn[0]), torch.stack(syn[1])
            losses, mono_reproj, aux = O.mono_losses(inp, o, True, True, noise=t["noise"][0])
            idx = aux["frame_idx"]
        else:
            opt = SimpleNamespace(height=H, width=W, min_depth=0.1, max_depth=100.0, frame_ids=[0, -1, 1], sclm=0)
            trainer_ops.generate_images_pred(inp, o, opt, materialize=True)
            del o[("warp_spec", 0)]                       # classic path: score the materialised warps
            syn = [[], []]
            for b in range(B):
                sl, sn = dyn_utils.generate_dynamic_instance(None, None, masks[b][0].to(dv), masks[b][1].to(dv),
                                                             o[("color", -1, 0)][b], o[("color", 1, 0)][b], False)
                syn[0].append(sl)
                syn[1].append(sn)
            o[("syn", -1, 0)], o[("syn", 1, 0)] = torch.stack(syn[0]), torch.stack(syn[1])
            losses, mono_reproj = loss_utils.compute_mono_losses(layers.SSIM(), inp, o, True, True,
                                                                 noise=t["noise"][0].to(dv))
            idx = o[("mal_selection", 0)] & 0x7F
        g = torch.autograd.grad(losses["loss"], [disp, Ts[-1], Ts[1]])
        res.append((float(losses["loss"].detach()), idx.cpu().to(torch.uint8), mono_reproj.detach().cpu(),
                    [x.cpu() for x in g]))
    (l0, i0, r0, g0), (l1, i1, r1, g1) = res
    assert abs(l0 - l1) <= 1e-5 * abs(l0)
    # the materialised warps come from mal_b200's own grid_sample (CPU rounding): everything is exact
    assert torch.equal(i0, i1)
    assert torch.equal(r0, r1)
    assert float((i0 >= 2).float().mean()) > 1e-4
    for a, b in zip(g1, g0):
        assert _gerr(a, b) < 1e-4


def test_config4_dualrefine_full_size():
    B, H, W = 1, 192, 640
    scales = [0, 1, 2, 3]
    inputs, t = make_photometric_inputs(B, H, W, num_scales=4, seed=401, translation_scale=0.2)
    opt = SimpleNamespace(height=H, width=W, scales=scales, n_losses=1, min_depth=0.1, max_depth=100.0,
                          disparity_smoothness=1e-3)
    keys = [(s, it) for s in scales if s != 1 for it in range(2 if s in (0, 1, 2) else 1)]
    noises = [torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(470 + i)) for i in range(len(keys))]

    def build(dv):
        leaves = {(s, it): t[("mono_disp" if it == 0 else "multi_disp", s)].clone().to(dv).requires_grad_(True)
                  for s, it in keys}
        o = {("disp", s, it): leaves[(s, it)] for s, it in keys}
        for k, f in (((0, -1), -1), ((0, 1), 1), ((0, -1, 1), -1)):
            o[("cam_T_cam",) + k] = t[("cam_T_cam", 0, f)].to(dv)
        o["consistency_mask"] = t["consistency_mask"].unsqueeze(1).to(dv)
        return to_device(inputs, dv), o, list(leaves.values())

    inp_c, o_c, lv_c = build("cpu")
    O.dualrefine_images_pred(inp_c, o_c, scales, 1, H, W)
    want, aux = O.dualrefine_compute_losses(inp_c, o_c, scales, 1, noises=noises)
    want_g = torch.autograd.grad(want["loss"], lv_c)
    inp_d, o_d, lv_d = build(DEV)
    trainer_ops.generate_images_pred_dualrefine(inp_d, o_d, opt)
    got = trainer_ops.compute_losses_dualrefine(inp_d, o_d, opt, noises=[n.to(DEV) for n in noises])
    for k in want:
        assert abs(float(got[k]) - float(want[k])) <= 1e-5 * abs(float(want[k])), k
    for s, it in keys:
        sel = o_d[("mal_selection", s, it)].cpu().numpy()
        assert np.array_equal(sel & 0x7F, aux[("frame_idx", s, it)].numpy().astype(np.uint8)), (s, it)
    for a, b in zip(torch.autograd.grad(got["loss"], lv_d), want_g):
        assert _gerr(a, b) < 1e-4


def test_config5_dynamicdepth_forward_warp_and_cost_volume():
    B, H, W = 2, 192, 512
    inputs, t = make_photometric_inputs(B, H, W, seed=501, normalised_K=CITYSCAPES_K, translation_scale=0.3)
    img = inputs[("color", 0, 0)].clone()
    img[:, :, :60] = 0                                        # doj_mask-style masked image
    depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    pose = t[("cam_T_cam", 0, -1)][:, :3, :].contiguous()
    K = inputs[("K", 0)][:, :3, :3].contiguous()
    mats = O.forward_warp_matrices(pose, K, 3)
    want = O.forward_warp(img, depth, pose, K, 3, matrices=mats)
    got = rigid_warp.forward_warp(img.to(DEV), depth.to(DEV), pose.to(DEV), K.to(DEV), upscale=3,
                                  matrices=[m.to(DEV) for m in mats])
    for a, b, n in zip(got, want, ("img_w", "depth_w", "valid")):
        assert torch.equal(a.cpu(), b), n
    assert 0.2 < float(want[2].mean()) < 0.95

    cv = make_cost_volume_inputs(B, H, W, channels=64, num_lookup=2, num_bins=96, seed=502, normalised_K=CITYSCAPES_K,
                                 min_bin=0.5, max_bin=6.0, translation_scale=0.5)
    look_img = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(503))
    look_img[:, :, 60:130, 100:260] = 0.0
    aug = torch.zeros(B, 1, 1, 1)
    aug[1] = 1
    want_vol, want_miss = O.match_features_dynamic(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"],
                                                   cv["K"], cv["inv_K"], cv["bins"], look_img, True, aug, False, True,
                                                   1, 0.7)
    m = matching.DynamicCostVolumeMatcher(num_depth_bins=96, min_depth_bin=0.5, max_depth_bin=6.0)
    d = to_device(cv, DEV)
    vol, miss = m.match_features(d["current_feats"], d["lookup_feats"], d["relative_poses"], d["K"], d["inv_K"],
                                 look_img.to(DEV), True, aug.to(DEV), False, True, 1, 0.7)
    assert torch.equal(miss.cpu(), want_miss)
    assert torch.equal(vol.cpu(), want_vol)


@pytest.mark.parametrize("is_multi", [False, True])
def test_config5_dynamicdepth_fused_losses_full_size(is_multi):
    """Config 5's loss half at 192x512 on the fused path (one photo_kernel<DD> pass per scale) against the oracle's
    compute_losses over materialised warps (dynamicdepth/trainer.py:906-975, :1006-1128): losses within 1e-5, the
    automask bit-exact, the zeroed target identical, gradients within 1e-4."""
    B, H, W = 2, 192, 512
    res = []
    for who, dv in (("oracle", torch.device("cpu")), ("ours", DEV)):
        inputs, t = make_photometric_inputs(B, H, W, num_scales=4, seed=511, normalised_K=CITYSCAPES_K,
                                            translation_scale=0.3)
        inputs[("color", -1, 0)][:, :, 40:100, 60:200] = 0.0     # DOMD-style holes reach the warps through the sources
        inputs[("color", 1, 0)][:, :, 70:150, 150:330] = 0.0
        inputs = {k: v.to(dv) for k, v in inputs.items()}
        name = "multi" if is_multi else "mono"
        disps = [t[(name + "_disp", s)].clone().to(dv).requires_grad_(True) for s in range(4)]
        Ts = [t[("cam_T_cam", 0, f)].clone().to(dv).requires_grad_(True) for f in (-1, 1)]
        o = {("disp", s): disps[s] for s in range(4)}
        o[("cam_T_cam", 0, -1)], o[("cam_T_cam", 0, 1)] = Ts
        if is_multi:
            o["consistency_mask"] = t["consistency_mask"].to(dv)
            o["augmentation_mask"] = torch.tensor([0.0, 1.0]).view(B, 1, 1, 1).to(dv)
            for s in range(4):
                o[("mono_depth", 0, s)] = (1.0 + 5.0 * t[("mono_disp", 0)]).to(dv)
        noises = [n.to(dv) for n in t["noise"]]
        if who == "oracle":
            O.images_pred(inputs, o, num_scales=4, height=H, width=W, is_multi=is_multi)
            losses, aux = O.dynamicdepth_compute_losses(inputs, o, (0, 1, 2, 3), is_multi=is_multi, noises=noises)
            masks = [aux[("mask", s)].reshape(B, H, W).float() for s in range(4)]
        else:
            opt = SimpleNamespace(scales=[0, 1, 2, 3], selec_reproj=True, zero_img=True, no_ssim="false", height=H,
                                  width=W, min_depth=0.1, max_depth=100.0, disable_automasking=False,
                                  disable_motion_masking=False, no_matching_augmentation="false",
                                  disparity_smoothness=1e-3)
            trainer_ops.generate_images_pred_dynamicdepth(inputs, o, opt, is_multi=is_multi)
            losses = trainer_ops.compute_losses_dynamicdepth(inputs, o, opt, is_multi=is_multi, noises=noises)
            masks = [(o[("mal_selection", s)] >> 7).float().reshape(B, H, W) for s in range(4)]
        grads = torch.autograd.grad(losses["loss"], disps + ([] if is_multi else Ts))
        res.append((losses, [m.cpu() for m in masks], [g.cpu() for g in grads], inputs[("color", 0, 0)].cpu()))
    (lo, mo, go, to_), (lk, mk, gk, tk) = res
    for k in lo:
        assert abs(float(lk[k]) - float(lo[k])) <= 1e-5 * abs(float(lo[k])), k
    if not is_multi:
        for a, b in zip(mk, mo):
            assert torch.equal(a, b)
    assert torch.equal(tk, to_)
    assert float((to_ == 0).float().mean()) > 0.02       # the case does zero part of the target
    for a, b in zip(gk, go):
        assert _gerr(a, b) < 1e-4
