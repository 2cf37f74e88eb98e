"""Generate tests/golden/*.npz from the REFERENCE's own Python (build container only).

Every array under the `ref_` prefix is an output of unmodified reference code imported
from /root/reference (manydepth.layers, manydepth.loss_utils, manydepth.trainer.Trainer
pure methods, manydepth.networks.resnet_encoder.ResnetEncoderMatching.match_features);
arrays under `in_` are the seeded synthetic inputs they were computed from, stored so the
fixtures do not depend on RNG implementation details of the machine that replays them.

    python tests/golden/make_golden.py        # rewrites the fixtures, then re-pins the oracle

The fixtures are small on purpose (32x48 images, 8x12 matching grid).
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.pin_against_reference import (forward_warp_inputs, load_reference,  # noqa: E402
                                          reference_forward_warp, reference_matcher,
                                          reference_trainer_shell, run_pin)
from mal_b200.utils.synthetic import make_cost_volume_inputs, make_photometric_inputs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _np(t):
    return t.detach().cpu().numpy()


def photometric_case(name, batch, height, width, seed, white_noise):
    ref = load_reference()
    L, U = ref.layers, ref.loss_utils
    inputs, t = make_photometric_inputs(batch, height, width, num_scales=1, seed=seed,
                                        white_noise=white_noise)
    shell = reference_trainer_shell(ref, batch, height, width)
    ssim = L.SSIM()
    out = {}
    for f in (0, -1, 1):
        out[f"in_color_{f}"] = _np(inputs[("color", f, 0)])
    out["in_K"], out["in_inv_K"] = _np(inputs[("K", 0)]), _np(inputs[("inv_K", 0)])
    for f in (-1, 1):
        out[f"in_T_{f}"] = _np(t[("cam_T_cam", 0, f)])
        out[f"in_syn_{f}"] = _np(t[("syn", f, 0)])
    out["in_mono_disp"], out["in_multi_disp"] = _np(t[("mono_disp", 0)]), _np(t[("multi_disp", 0)])
    out["in_noise_mono"], out["in_noise_main"] = _np(t["noise"][0]), _np(t["noise"][1])
    out["in_consistency_mask"] = _np(t["consistency_mask"])
    out["in_augmentation_mask"] = _np(t["augmentation_mask"])

    def with_noise(nz, fn):
        # the reference draws torch.randn(shape) in-line; feed it our stored tensor
        orig = torch.randn
        torch.randn = lambda *a, **k: nz.clone()
        try:
            return fn()
        finally:
            torch.randn = orig

    # ---- teacher (mono) pass ------------------------------------------------------
    Ts = {f: t[("cam_T_cam", 0, f)].clone().requires_grad_(True) for f in (-1, 1)}
    mono = {("disp", 0): t[("mono_disp", 0)].clone().requires_grad_(True)}
    for f in (-1, 1):
        mono[("cam_T_cam", 0, f)] = Ts[f]
        mono[("syn", f, 0)] = t[("syn", f, 0)]
    shell.generate_images_pred(inputs, mono)
    for f in (-1, 1):
        out[f"ref_mono_sample_{f}"] = _np(mono[("sample", f, 0)])
        out[f"ref_mono_color_{f}"] = _np(mono[("color", f, 0)])
    out["ref_mono_depth"] = _np(mono[("depth", 0, 0)])
    for temporal, tag in ((False, "plain"), (True, "temporal")):
        losses, mono_reproj = with_noise(
            t["noise"][0], lambda: U.compute_mono_losses(ssim, inputs, mono, temporal, True))
        out[f"ref_mono_{tag}_loss"] = _np(losses["loss"])
        out[f"ref_mono_{tag}_reproj_loss"] = _np(losses["reproj_loss/0"])
        out[f"ref_mono_{tag}_min_reproj"] = _np(mono_reproj)
        # selection indices, recomputed with the reference's own primitives
        target = inputs[("color", 0, 0)]
        cands = [U.compute_reprojection_loss(ssim, mono[("color", f, 0)], target) for f in (-1, 1)]
        if temporal:
            cands += [U.compute_reprojection_loss(ssim, mono[("syn", f, 0)], target) for f in (-1, 1)]
        rl, fidx = torch.min(torch.cat(cands, 1), dim=1, keepdim=True)
        ident = torch.min(torch.cat([U.compute_reprojection_loss(ssim, inputs[("color", f, 0)], target)
                                     for f in (-1, 1)], 1), dim=1, keepdim=True)[0]
        ident = ident + t["noise"][0] * 0.00001
        out[f"ref_mono_{tag}_frame_idx"] = _np(fidx).astype(np.uint8)
        out[f"ref_mono_{tag}_automask"] = _np(U.compute_loss_masks(rl, ident)).astype(np.uint8)
        out[f"ref_identity_min"] = _np(torch.min(torch.cat(
            [U.compute_reprojection_loss(ssim, inputs[("color", f, 0)], target) for f in (-1, 1)], 1),
            dim=1, keepdim=True)[0])
        grads = torch.autograd.grad(losses["loss"], [mono[("disp", 0)], Ts[-1], Ts[1]], retain_graph=True)
        out[f"ref_mono_{tag}_grad_disp"] = _np(grads[0])
        out[f"ref_mono_{tag}_grad_T_-1"], out[f"ref_mono_{tag}_grad_T_1"] = _np(grads[1]), _np(grads[2])
    mono_reproj = mono_reproj.detach()

    # ---- ensemble + student (multi) pass -----------------------------------------
    disp_ens = (t[("mono_disp", 0)] + t[("multi_disp", 0)]) / 2.0
    ens = shell.generate_images_pred_ensemble(inputs, t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)], disp_ens)
    out["ref_ensemble_reproj"] = _np(ens)
    multi = {("disp", 0): t[("multi_disp", 0)].clone().requires_grad_(True),
             "consistency_mask": t["consistency_mask"], "augmentation_mask": t["augmentation_mask"],
             ("mono_depth", 0, 0): mono[("depth", 0, 0)].detach()}
    for f in (-1, 1):
        multi[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
        multi[("syn", f, 0)] = t[("syn", f, 0)]
    shell.generate_images_pred(inputs, multi, is_multi=True)
    low = 1 / (mono[("depth", 0, 0)].detach()[:, 0] * (0.4 + t["consistency_mask"] * 2.0))
    multi["lowest_cost"] = low
    out["in_lowest_cost"] = _np(low)
    out["ref_matching_mask"] = _np(shell.compute_matching_mask(multi)).astype(np.uint8)
    for ens_t, has_ins, blc, tag in ((ens, False, True, "ens_blc"), (None, True, False, "noens_ins")):
        opt = SimpleNamespace(batch_size=batch, dual_distil=False, learn_ens=False, pareto=False,
                              loss_blc=blc, min_depth=0.1, max_depth=100.0)
        losses, _, loss_list = with_noise(t["noise"][1], lambda: U.compute_main_losses(
            ssim, inputs, multi, mono_reproj, ens_t, opt, None, None, has_ins))
        for k in ("loss", "distil_loss", "reproj_loss/0", "consistency_loss/0"):
            out[f"ref_main_{tag}_{k.replace('/', '_')}"] = _np(losses[k])
        out[f"ref_main_{tag}_consistency_target"] = _np(multi["consistency_target/0"])
        total = losses["loss"] + (0.25 * loss_list[1] if blc else 0)
        g, = torch.autograd.grad(total, multi[("disp", 0)], retain_graph=True)
        out[f"ref_main_{tag}_grad_disp"] = _np(g)   # d(loss + 0.25*distil)/d disp when blc
    out["meta_torch"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, {k: v.shape for k, v in out.items() if k.startswith("ref_")}.__len__(), "ref arrays")


def cost_volume_case(name, batch, height, width, channels, bins, seed):
    ref = load_reference()
    cv = make_cost_volume_inputs(batch, height, width, channels=channels, num_bins=bins, seed=seed,
                                 zero_pose_sample=batch - 1, num_lookup=2)
    cv["relative_poses"][0, 1] = 0  # one missing lookup frame on a live sample
    enc = reference_matcher(ref, height, width, cv["bins"])
    vol, miss = enc.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"],
                                   cv["K"], cv["inv_K"])
    conf = enc.compute_confidence_mask(vol * (1 - miss))
    viz = vol.clone()
    viz[viz == 0] = 100
    _, am = torch.min(viz, 1)
    out = {"in_" + k: _np(v) for k, v in cv.items()}
    out.update(ref_cost_volume=_np(vol), ref_missing=_np(miss).astype(np.uint8), ref_confidence=_np(conf),
               ref_argmin=_np(am).astype(np.int32), ref_lowest_cost=_np(enc.indices_to_disparity(am)),
               meta_torch=np.array(torch.__version__))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name)


def forward_warp_case(name):
    """dynamicdepth/rigid_warp.forward_warp (coalesce restated, see pin_against_reference)."""
    import warnings
    from oracle import mal_oracle as O
    fw = reference_forward_warp()
    img, depth, pose, K = forward_warp_inputs()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        img_w, depth_w, valid = fw(img, depth, pose, K, upscale=3)
    Ku_inv, K_inv, proj = O.forward_warp_matrices(pose, K, 3)
    np.savez_compressed(os.path.join(HERE, name), in_img=_np(img), in_depth=_np(depth), in_pose=_np(pose),
                        in_K=_np(K), in_Ku_inv=_np(Ku_inv), in_K_inv=_np(K_inv), in_proj=_np(proj),
                        ref_img_w=_np(img_w), ref_depth_w=_np(depth_w), ref_valid=_np(valid).astype(np.uint8),
                        meta_torch=np.array(torch.__version__))
    print("wrote", name)


def corr_lookup_case(name):
    """dualrefine/networks/corr.py CoordSampler.register + __call__, imported by file path."""
    import importlib.util
    from oracle.pin_against_reference import REFERENCE_ROOT, corr_case
    spec = importlib.util.spec_from_file_location(
        "ref_dualrefine_corr", os.path.join(REFERENCE_ROOT, "dualrefine", "networks", "corr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = {}
    for tag, (B, Cn, h, w, L, D, heads) in {"a": (2, 64, 12, 24, 3, 5, 1), "b": (2, 32, 12, 20, 3, 4, 2)}.items():
        fmap1, fmap2, coords = corr_case(B, Cn, h, w, L, D)
        cs = mod.CoordSampler(None)
        cs.register(fmap1, fmap2, num_levels=L)
        out[f"{tag}_in_fmap1"], out[f"{tag}_in_fmap2"], out[f"{tag}_in_coords"] = _np(fmap1), _np(fmap2), _np(coords)
        out[f"{tag}_heads"] = np.int32(heads)
        out[f"{tag}_ref_corr"] = _np(cs(coords, num_levels=L, num_head=heads))
        out[f"{tag}_ref_pyramid"] = _np(torch.cat([p.reshape(-1) for p in cs.f2_pyramid]))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name)


if __name__ == "__main__":
    torch.set_num_threads(1)
    photometric_case("photometric_smooth.npz", 2, 32, 48, seed=101, white_noise=False)
    photometric_case("photometric_noise.npz", 1, 24, 40, seed=202, white_noise=True)
    cost_volume_case("cost_volume.npz", 2, 32, 48, channels=8, bins=12, seed=303)
    forward_warp_case("forward_warp.npz")
    corr_lookup_case("corr_lookup.npz")
    ok = run_pin()
    print("oracle pinned:", ok)
    sys.exit(0 if ok else 1)
