"""CPU oracle for the MAL photometric hot path.  TEST INFRASTRUCTURE ONLY.

This module is a restatement, in plain torch-on-CPU fp32, of the arithmetic the
reference (YuejiangDong/MAL) performs on the path SURVEY.md section 8 scopes:
backproject -> project -> bilinear warp -> SSIM+L1 -> min-reprojection/automask
-> MAL temporal/distillation selection, the plane-sweep matching cost volume,
the smoothness term, the host-side loss balancers, DynamicDepth's forward warp,
the temporal-hint image synthesis and DualRefine's epipolar correlation lookup
(CoordSampler, sample_tgt).

Rules (see DESIGN.md "Oracle"):
  * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference legs may import this file.  The product package
    `mal_b200` never imports it and has no CPU fallback.
  * Every function cites the reference file:line it restates (paths relative
    to /root/reference).
  * Parity is PINNED: oracle/pin_against_reference.py imports the reference's
    own modules in the build container and asserts bitwise equality with this
    restatement on seeded inputs; tests/golden/*.npz are outputs of the
    reference itself (made by tests/golden/make_golden.py) and
    tests/test_oracle_golden.py re-checks this file against them anywhere.
  * The arithmetic below is the reference's, op for op (same torch calls in
    the same order) because the selection indices must match bit for bit.

Third-party arithmetic the reference leans on and that is restated here through
the same torch 2.11 ATen CPU kernels: grid_sampler_2d, avg_pool2d,
reflection_pad2d, upsample_bilinear2d, argmin/min.  torch_sparse.coalesce
(dynamicdepth/rigid_warp.py:577, version unpinned, absent from the tree) is
restated from its published semantics (duplicate-index max reduction).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

MANYDEPTH = 0   # Project3D normalisation x/(W-1), grid_sample align_corners=True
DUALREFINE = 1  # Project3D 2(x+0.5)/W-1, grid_sample align_corners=False


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def disp_to_depth(disp, min_depth, max_depth):
    """manydepth/layers.py:14-23."""
    lo = 1 / max_depth
    hi = 1 / min_depth
    scaled = lo + (hi - lo) * disp
    return scaled, 1 / scaled


def rot_from_axisangle(vec):
    """manydepth/layers.py:61-100 (Rodrigues, 4x4)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), device=vec.device)
    entries = {(0, 0): x * xC + ca, (0, 1): xyC - zs, (0, 2): zxC + ys,
               (1, 0): xyC + zs, (1, 1): y * yC + ca, (1, 2): yzC - xs,
               (2, 0): zxC - ys, (2, 1): yzC + xs, (2, 2): z * zC + ca}
    for (r, c), v in entries.items():
        rot[:, r, c] = torch.squeeze(v)
    rot[:, 3, 3] = 1
    return rot


def get_translation_matrix(t):
    """manydepth/layers.py:45-58."""
    T = torch.zeros(t.shape[0], 4, 4, device=t.device)
    for i in range(4):
        T[:, i, i] = 1
    T[:, :3, 3, None] = t.contiguous().view(-1, 3, 1)
    return T


def transformation_from_parameters(axisangle, translation, invert=False):
    """manydepth/layers.py:26-42."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t *= -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def pixel_grid(batch, height, width):
    """Homogeneous pixel coordinates (B,3,HW), x = column; layers.py:149-161."""
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32),
                            torch.arange(width, dtype=torch.float32), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width)], 0)
    return pix.unsqueeze(0).repeat(batch, 1, 1)


def backproject(depth, inv_K):
    """BackprojectDepth.forward, manydepth/layers.py:163-168 -> (B,4,HW)."""
    B, _, H, W = depth.shape
    rays = torch.matmul(inv_K[:, :3, :3], pixel_grid(B, H, W))
    cam = depth.view(B, 1, -1) * rays
    return torch.cat([cam, torch.ones(B, 1, H * W)], 1)


def project3d(points, K, T, height, width, convention=MANYDEPTH, eps=1e-7, return_z=False):
    """Project3D.forward: manydepth/layers.py:184-199, dualrefine/layers.py:216-226."""
    B = points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    cam = torch.matmul(P, points)
    pix = cam[:, :2, :] / (cam[:, 2, :].unsqueeze(1) + eps)
    pix = pix.view(B, 2, height, width).permute(0, 2, 3, 1)
    if convention == MANYDEPTH:
        pix[..., 0] /= width - 1
        pix[..., 1] /= height - 1
        pix = (pix - 0.5) * 2
    else:
        pix[..., 0] = 2 * (pix[..., 0] + 0.5) / width - 1
        pix[..., 1] = 2 * (pix[..., 1] + 0.5) / height - 1
    if return_z:
        return pix, cam[:, 2, :].unsqueeze(1).view(B, 1, height, width)
    return pix


def warp(img, grid, convention=MANYDEPTH, padding_mode="border"):
    """F.grid_sample call sites: manydepth/trainer.py:1122-1125 (align_corners=True),
    dualrefine/trainer.py:444-447 (align_corners=False)."""
    return F.grid_sample(img, grid, padding_mode=padding_mode,
                         align_corners=(convention == MANYDEPTH))


def upsample_disp(disp, height, width):
    """manydepth/trainer.py:1093-1094."""
    return F.interpolate(disp, [height, width], mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------
# photometric terms
# --------------------------------------------------------------------------
_C1 = 0.01 ** 2
_C2 = 0.03 ** 2


def ssim(x, y):
    """SSIM.forward, manydepth/layers.py:243-257 (3x3 box, reflection pad 1)."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    box = lambda t: F.avg_pool2d(t, 3, 1)
    mu_x, mu_y = box(x), box(y)
    sigma_x = box(x ** 2) - mu_x ** 2
    sigma_y = box(y ** 2) - mu_y ** 2
    sigma_xy = box(x * y) - mu_x * mu_y
    num = (2 * mu_x * mu_y + _C1) * (2 * sigma_xy + _C2)
    den = (mu_x ** 2 + mu_y ** 2 + _C1) * (sigma_x + sigma_y + _C2)
    return torch.clamp((1 - num / den) / 2, 0, 1)


def reprojection_loss(pred, target, no_ssim=False):
    """manydepth/loss_utils.py:46-55, manydepth/trainer.py:1211-1223."""
    l1 = torch.abs(target - pred).mean(1, True)
    if no_ssim:
        return l1
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def loss_masks(reproj, identity):
    """manydepth/loss_utils.py:27-44."""
    if identity is None:
        return torch.ones_like(reproj)
    idx = torch.argmin(torch.cat([reproj, identity], dim=1), dim=1, keepdim=True)
    return (idx == 0).float()


def smooth_loss(disp, img):
    """get_smooth_loss, manydepth/layers.py:210-223."""
    gdx = torch.abs(disp[:, :, :, :-1] - disp[:, :, :, 1:])
    gdy = torch.abs(disp[:, :, :-1, :] - disp[:, :, 1:, :])
    gix = torch.mean(torch.abs(img[:, :, :, :-1] - img[:, :, :, 1:]), 1, keepdim=True)
    giy = torch.mean(torch.abs(img[:, :, :-1, :] - img[:, :, 1:, :]), 1, keepdim=True)
    gdx = gdx * torch.exp(-gix)
    gdy = gdy * torch.exp(-giy)
    return gdx.mean() + gdy.mean()


def normalised_smooth_loss(disp, img):
    """manydepth/loss_utils.py:119-121 (mean-normalised disparity)."""
    mean_disp = disp.mean(2, True).mean(3, True)
    return smooth_loss(disp / (mean_disp + 1e-7), img)


# --------------------------------------------------------------------------
# trainer-resident glue (restated; manydepth.trainer cannot be imported)
# --------------------------------------------------------------------------
def images_pred(inputs, outputs, frame_ids=(0, -1, 1), num_scales=1, height=192, width=640,
                min_depth=0.1, max_depth=100.0, is_multi=False, convention=MANYDEPTH,
                automask=True, v1_multiscale=False):
    """Trainer.generate_images_pred, manydepth/trainer.py:1078-1158.  v1_multiscale (:1089-1095)
    keeps the disparity at its own scale and warps the images / intrinsics of that scale."""
    for scale in range(num_scales):
        ss = scale if v1_multiscale else 0
        hs, ws = (height // 2 ** scale, width // 2 ** scale) if v1_multiscale else (height, width)
        disp = outputs[("disp", scale)] if v1_multiscale else upsample_disp(outputs[("disp", scale)], height, width)
        _, depth = disp_to_depth(disp, min_depth, max_depth)
        outputs[("depth", 0, scale)] = depth
        for fid in frame_ids[1:]:
            T = outputs[("cam_T_cam", 0, fid)]
            if is_multi:
                T = T.detach()
            cam = backproject(depth, inputs[("inv_K", ss)])
            grid = project3d(cam, inputs[("K", ss)], T, hs, ws, convention)
            outputs[("sample", fid, scale)] = grid
            outputs[("color", fid, scale)] = warp(inputs[("color", fid, ss)], grid, convention)
            if automask:
                outputs[("color_identity", fid, scale)] = inputs[("color", fid, ss)]
    return outputs


def images_pred_ensemble(inputs, T_prev, T_next, disp, frame_ids=(0, -1, 1), height=192,
                         width=640, min_depth=0.1, max_depth=100.0, no_ssim=False):
    """Trainer.generate_images_pred_ensemble, manydepth/trainer.py:1172-1207."""
    disp = upsample_disp(disp, height, width)
    _, depth = disp_to_depth(disp, min_depth, max_depth)
    target = inputs[("color", 0, 0)]
    cands = []
    for T, fid in zip((T_prev, T_next), frame_ids[1:]):
        cam = backproject(depth, inputs[("inv_K", 0)])
        grid = project3d(cam, inputs[("K", 0)], T, height, width)
        cands.append(reprojection_loss(warp(inputs[("color", fid, 0)], grid), target, no_ssim))
    return torch.min(torch.cat(cands, 1), dim=1, keepdim=True)[0]


def matching_mask(outputs):
    """Trainer.compute_matching_mask, manydepth/trainer.py:1066-1076."""
    mono = outputs[("mono_depth", 0, 0)]
    matching = 1 / outputs["lowest_cost"].unsqueeze(1)
    mask = ((matching - mono) / mono) < 1.0
    mask *= ((mono - matching) / matching) < 1.0
    return mask[:, 0]


def mono_losses(inputs, outputs, temporal, has_ins, noise=None, no_ssim=False):
    """compute_mono_losses, manydepth/loss_utils.py:57-129.

    `noise` replaces the reference's in-line torch.randn(shape) draw (:105) so both
    sides of a parity test see the same tie-break bits; None draws it like the
    reference does."""
    target = inputs[("color", 0, 0)]
    cands = [reprojection_loss(outputs[("color", f, 0)], target, no_ssim) for f in (-1, 1)]
    if temporal and has_ins:
        cands += [reprojection_loss(outputs[("syn", f, 0)], target, no_ssim) for f in (-1, 1)]
    cands = torch.cat(cands, 1)
    ident = torch.cat([reprojection_loss(inputs[("color", f, 0)], target, no_ssim)
                       for f in (-1, 1)], 1)
    ident, _ = torch.min(ident, dim=1, keepdim=True)
    reproj, frame_idx = torch.min(cands, dim=1, keepdim=True)
    if noise is None:
        noise = torch.randn(ident.shape)
    ident = ident + noise * 0.00001
    mask = loss_masks(reproj, ident)
    reproj_loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
    losses = {"reproj_loss/0": reproj_loss}
    loss = reproj_loss + 1e-3 * normalised_smooth_loss(outputs[("disp", 0)],
                                                       inputs[("color", 0, 0)]) / (2 ** 0)
    losses["loss/0"] = loss
    losses["loss"] = 0 + loss
    aux = {"frame_idx": frame_idx, "automask": mask}
    return losses, torch.min(cands, dim=1, keepdim=True)[0], aux


def main_losses(inputs, outputs, mono_reproj, ensemble_reproj, *, batch_size,
                multi_has_ins=False, dual_distil=False, loss_blc=False, w_list=None,
                noise=None, no_ssim=False):
    """compute_main_losses, manydepth/loss_utils.py:131-281 (pareto / learn_ens off).

    The automask computed at :181 is discarded at :192, as in the reference; only
    the RNG draw matters to callers that share a generator."""
    target = inputs[("color", 0, 0)]
    cands = [reprojection_loss(outputs[("color", f, 0)], target, no_ssim) for f in (-1, 1)]
    if multi_has_ins:
        cands += [reprojection_loss(outputs[("syn", f, 0)], target, no_ssim) for f in (-1, 1)]
    cands = torch.cat(cands, 1)
    reproj, frame_idx = torch.min(cands, dim=1, keepdim=True)
    multi_reproj = reproj.clone()
    if noise is None:
        noise = torch.randn(reproj.shape)  # drawn then unused, loss_utils.py:178-192

    mask = torch.ones_like(reproj)
    mask = mask * outputs["consistency_mask"].unsqueeze(1)
    mask = mask * (1 - outputs["augmentation_mask"][:batch_size])
    cons_mask = (1 - mask).float()

    reproj_loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
    multi_depth = outputs[("depth", 0, 0)]
    mono_depth = outputs[("mono_depth", 0, 0)].detach()
    cons_loss = (torch.abs(multi_depth - mono_depth) * cons_mask).mean()
    cons_target = 1 / (mono_depth.detach() * cons_mask + multi_depth.detach() * (1 - cons_mask))

    losses = {"consistency_loss/0": cons_loss, "reproj_loss/0": reproj_loss}
    loss = reproj_loss + cons_loss
    loss = loss + 1e-3 * normalised_smooth_loss(outputs[("disp", 0)], inputs[("color", 0, 0)])

    if ensemble_reproj is None:
        _, sel = torch.min(torch.cat([mono_reproj, multi_reproj], 1), dim=1, keepdim=True)
        teacher = outputs[("mono_depth", 0, 0)] if dual_distil else mono_depth
        distil_depth = torch.where(sel == 0, teacher, multi_depth)
    else:
        _, sel = torch.min(torch.cat([mono_reproj, ensemble_reproj, multi_reproj], 1),
                           dim=1, keepdim=True)
        ens_depth = (mono_depth + multi_depth) / 2.0
        distil_depth = torch.where(sel == 0, mono_depth, ens_depth)
        distil_depth = torch.where(sel == 2, multi_depth, distil_depth)
    distil_loss = (torch.abs(distil_depth - multi_depth) * (1 - cons_mask)).mean()

    losses["distil_loss"] = distil_loss
    if loss_blc:
        loss_list = [loss.clone(), distil_loss]
        new_w = w_list
    else:
        loss = loss + distil_loss
        loss_list, new_w = None, None
    losses["loss/0"] = loss
    losses["loss"] = loss
    aux = {"frame_idx": frame_idx, "distil_idx": sel, "multi_reproj": multi_reproj,
           "consistency_target": cons_target}
    return losses, new_w, loss_list, aux


def trainer_compute_losses(inputs, outputs, *, num_scales=1, is_multi=False, temporal=False,
                           has_ins=False, automask=True, motion_masking=True,
                           matching_augmentation=True, batch_size=None, no_ssim=False,
                           smoothness=1e-3, noises=None, v1_multiscale=False):
    """Trainer.compute_losses, manydepth/trainer.py:1248-1475 (the non-distil path,
    one pass per scale, total / num_scales).  dynamicdepth/trainer.py:1006-1128 uses
    the same chain over opt.scales."""
    losses, total = {}, 0
    aux = {}
    for scale in range(num_scales):
        ss = scale if v1_multiscale else 0          # source_scale, trainer.py:1260-1263
        target = inputs[("color", 0, ss)]
        cands = [reprojection_loss(outputs[("color", f, scale)], target, no_ssim) for f in (-1, 1)]
        if (not is_multi) and temporal and has_ins:
            cands += [reprojection_loss(outputs[("syn", f, scale)], target, no_ssim)
                      for f in (-1, 1)]
        cands = torch.cat(cands, 1)
        ident = torch.cat([reprojection_loss(inputs[("color", f, ss)], target, no_ssim)
                           for f in (-1, 1)], 1)
        ident, _ = torch.min(ident, dim=1, keepdim=True)
        reproj, frame_idx = torch.min(cands, dim=1, keepdim=True)
        if automask:
            nz = noises[scale] if noises is not None else torch.randn(ident.shape)
            ident = ident + nz * 0.00001
        # trainer.py:1292-1311: disable_automasking only skips the tie-break noise; the identity loss is
        # still compared (compute_loss_masks always gets it in this trainer)
        mask = loss_masks(reproj, ident)
        cons_loss = 0
        if is_multi:
            mask = torch.ones_like(mask)
            if motion_masking:
                mask = mask * outputs["consistency_mask"].unsqueeze(1)
            if matching_augmentation:
                mask = mask * (1 - outputs["augmentation_mask"][:batch_size])
            cons_mask = (1 - mask).float()
        reproj_loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
        if is_multi:
            multi_depth = outputs[("depth", 0, scale)]
            mono_depth = outputs[("mono_depth", 0, scale)].detach()
            cons_loss = (torch.abs(multi_depth - mono_depth) * cons_mask).mean()
            losses[f"consistency_loss/{scale}"] = cons_loss
        losses[f"reproj_loss/{scale}"] = reproj_loss
        loss = reproj_loss + cons_loss
        loss = loss + smoothness * normalised_smooth_loss(
            outputs[("disp", scale)], inputs[("color", 0, scale)]) / (2 ** scale)
        total = total + loss
        losses[f"loss/{scale}"] = loss
        aux[("frame_idx", scale)] = frame_idx
        aux[("mask", scale)] = mask
    losses["loss"] = total / num_scales
    return losses, aux


def dualrefine_images_pred(inputs, outputs, scales=(0, 1, 2, 3), n_losses=1, height=192, width=640,
                           min_depth=0.1, max_depth=100.0, Dstar_T0_pair=False):
    """dualrefine/trainer.py generate_images_pred :395-455 (half-pixel Project3D,
    align_corners=False; per (scale, deq_iter); scale 1 is skipped).  dualrefine.trainer cannot be
    imported (networks/lib is missing), so this restatement is checked only through
    the pinned primitives it is built from."""
    for scale in scales:
        n = n_losses + 1 if scale in (0, 1, 2) else 1
        for it in range(n):
            if scale == 1:
                continue
            disp = upsample_disp(outputs[("disp", scale, it)], height, width)
            _, depth = disp_to_depth(disp, min_depth, max_depth)
            outputs[("depth", 0, scale, it)] = depth
            for fid in (-1, 1):
                if fid == 1:
                    T = outputs[("cam_T_cam", 0, fid)]
                    if it > 0:
                        T = T.detach()
                elif it > 0:
                    T = outputs[("cam_T_cam", 0, fid)].detach() if Dstar_T0_pair else outputs[("cam_T_cam", 0, fid, 1)]
                else:
                    T = outputs[("cam_T_cam", 0, fid)]
                cam = backproject(depth, inputs[("inv_K", 0)])
                grid = project3d(cam, inputs[("K", 0)], T, height, width, DUALREFINE)
                outputs[("sample", fid, scale, it)] = grid
                outputs[("color", fid, scale, it)] = warp(inputs[("color", fid, 0)], grid, DUALREFINE)
    return outputs


def dualrefine_compute_losses(inputs, outputs, scales=(0, 1, 2, 3), n_losses=1, automask=True,
                              motion_masking=True, smoothness=1e-3, noises=None, no_ssim=False,
                              avg_reprojection=False):
    """dualrefine/trainer.py compute_losses :530-697, the f_thres > 0 branch (:543-626): per
    (scale, deq_iter); for deq_iter > 0 the automask is multiplied by the consistency mask and the
    consistency term pulls towards the deq_iter-0 depth; total / len(scales)."""
    losses, total, aux = {}, 0, {}
    target = inputs[("color", 0, 0)]
    draw = 0
    for scale in scales:
        loss = 0
        n = n_losses + 1 if scale in (0, 1, 2) else 1
        for it in range(n):
            if scale == 1:
                continue
            cands = torch.cat([reprojection_loss(outputs[("color", f, scale, it)], target, no_ssim) for f in (-1, 1)], 1)
            ident = None
            if automask:
                ident = torch.cat([reprojection_loss(inputs[("color", f, 0)], target, no_ssim) for f in (-1, 1)], 1)
                if avg_reprojection:      # dualrefine/trainer.py:575-576
                    ident = ident.mean(1, keepdim=True)
                else:
                    ident, _ = torch.min(ident, dim=1, keepdim=True)
            if avg_reprojection:          # :585-586
                reproj, frame_idx = cands.mean(1, keepdim=True), torch.zeros_like(cands[:, :1]).long()
            else:
                reproj, frame_idx = torch.min(cands, dim=1, keepdim=True)
            if automask:
                nz = noises[draw] if noises is not None else torch.randn(ident.shape)
                draw += 1
                ident = ident + nz * 0.00001
            mask = loss_masks(reproj, ident)
            if it > 0:
                if motion_masking:
                    mask = mask * outputs["consistency_mask"]
                cons_mask = (1 - mask).float()
            reproj_loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
            cons = 0
            if it > 0:
                multi_depth = outputs[("depth", 0, scale, it)]
                mono_depth = outputs[("depth", 0, scale, 0)].detach()
                cons = (torch.abs(multi_depth - mono_depth) * cons_mask).mean()
                losses[f"consistency_loss/{scale}_{it}"] = cons
            losses[f"reproj_loss/{scale}"] = reproj_loss
            loss = loss + reproj_loss + cons
            loss = loss + smoothness * normalised_smooth_loss(outputs[("disp", scale, it)],
                                                              inputs[("color", 0, scale)]) / (2 ** scale)
            total = total + loss
            losses[f"loss/{scale}_{it}"] = loss
            aux[("frame_idx", scale, it)] = frame_idx
            aux[("mask", scale, it)] = mask
    losses["loss"] = total / len(scales)
    return losses, aux


def dynamicdepth_reprojection_loss(pred, target, zero_img=True, no_ssim=False):
    """dynamicdepth/trainer.py compute_reprojection_loss :958-975.  With zero_img the dark pixels
    of `pred` (DOMD warping holes) are zeroed in a copy of pred AND IN PLACE in `target`."""
    if zero_img:
        mask = (pred.sum(1) < 0.1).unsqueeze(1).repeat([1, 3, 1, 1]).detach()
        pred = pred.clone()
        pred[mask] = 0
        target[mask] = 0
    l1 = torch.abs(target - pred).mean(1, True)
    if no_ssim:
        return l1
    return 0.85 * ssim(pred, target).mean(1, True) + 0.15 * l1


def dynamicdepth_compute_losses(inputs, outputs, scales=(0, 1, 2, 3), is_multi=False, automask=True,
                                avg_reprojection=False, selec_reproj=True, zero_img=True, motion_masking=True,
                                matching_augmentation=True, no_teacher_warp=False, train_teacher_only=False,
                                smoothness=1e-3, noises=None, no_ssim=False):
    """dynamicdepth/trainer.py compute_losses :1006-1128 (feat_loss off).  NOTE: mutates
    inputs[("color", 0, 0)] in place when zero_img is set, exactly like the reference."""
    losses, total, aux = {}, 0, {}
    rl = lambda p, t: dynamicdepth_reprojection_loss(p, t, zero_img, no_ssim)
    for si, scale in enumerate(scales):
        loss = 0
        disp, color, target = outputs[("disp", scale)], inputs[("color", 0, scale)], inputs[("color", 0, 0)]
        cands = torch.cat([rl(outputs[("color", f, scale)], target) for f in (-1, 1)], 1)
        ident = None
        if automask:
            preds = [(inputs[("ori_color", f, 0)] if ((not is_multi) and no_teacher_warp and not train_teacher_only)
                      else inputs[("color", f, 0)]) for f in (-1, 1)]
            ident = torch.cat([rl(p, target) for p in preds], 1)
            ident = ident.mean(1, keepdim=True) if avg_reprojection else torch.min(ident, dim=1, keepdim=True)[0]
        reproj = cands.mean(1, keepdim=True) if avg_reprojection else torch.min(cands, dim=1, keepdim=True)[0]
        if selec_reproj:
            maskm1 = (outputs[("color", -1, scale)].sum(1) < 0.1).detach()
            maskp1 = (outputs[("color", 1, scale)].sum(1) < 0.1).detach()
            maskand = (maskm1 * maskp1).detach()
            reproj[maskm1.unsqueeze(1)] = (cands[:, 1, :, :])[maskm1]
            reproj[maskp1.unsqueeze(1)] = (cands[:, 0, :, :])[maskp1]
            reproj[maskand.unsqueeze(1)] = 0
        if automask:
            nz = noises[si] if noises is not None else torch.randn(ident.shape)
            ident = ident + nz * 0.00001
        mask = loss_masks(reproj, ident)
        if is_multi:
            mask = torch.ones_like(mask)
            if motion_masking:
                mask = mask * outputs["consistency_mask"].unsqueeze(1)
            if matching_augmentation:
                mask = mask * (1 - outputs["augmentation_mask"])
            cons_mask = (1 - mask).float()
        reproj_loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
        cons = 0
        if is_multi:
            multi_depth = outputs[("depth", 0, scale)]
            mono_depth = outputs[("mono_depth", 0, scale)].detach()
            cons = (torch.abs(multi_depth - mono_depth) * cons_mask).mean()
            losses[f"consistency_loss/{scale}"] = cons
        losses[f"reproj_loss/{scale}"] = reproj_loss
        loss = loss + reproj_loss + cons
        loss = loss + smoothness * normalised_smooth_loss(disp, color) / (2 ** scale)
        total = total + loss
        losses[f"loss/{scale}"] = loss
        aux[("mask", scale)] = mask
        aux[("reproj", scale)] = reproj
    losses["loss"] = total / len(scales)
    return losses, aux


# --------------------------------------------------------------------------
# plane-sweep matching cost volume
# --------------------------------------------------------------------------
def depth_bins(min_depth_bin, max_depth_bin, num_bins=96, binning="linear"):
    """ResnetEncoderMatching.compute_depth_bins, networks/resnet_encoder.py:121-141."""
    lo, hi = float(min_depth_bin), float(max_depth_bin)
    if binning == "linear":
        return torch.linspace(lo, hi, num_bins)
    if binning == "inverse":
        b = 1 / np.linspace(1 / hi, 1 / lo, num_bins)[::-1]
        return torch.from_numpy(np.ascontiguousarray(b)).float()
    if binning == "log":
        base, it = torch.log(torch.tensor(lo)), torch.log(torch.tensor(hi / lo))
        return torch.exp(torch.Tensor([base + it * i / num_bins for i in range(num_bins)]))
    raise NotImplementedError(binning)


def match_features(current_feats, lookup_feats, relative_poses, K, invK, bins,
                   convention=MANYDEPTH, set_missing_to_max=True):
    """ResnetEncoderMatching.match_features, networks/resnet_encoder.py:151-233
    (dualrefine/networks/resnet_encoder.py:163-245 differs in align_corners only)."""
    B, C, h, w = current_feats.shape
    nb = len(bins)
    planes = bins.view(nb, 1, 1, 1).float().expand(nb, 1, h, w).contiguous()
    volumes, masks = [], []
    for b in range(B):
        cost = torch.zeros(nb, h, w)
        counts = torch.zeros(nb, h, w)
        world = backproject(planes, invK[b:b + 1])
        for li in range(lookup_feats.shape[1]):
            pose = relative_poses[b:b + 1, li]
            if pose.sum() == 0:
                continue
            feat = lookup_feats[b:b + 1, li].repeat([nb, 1, 1, 1])
            locs = project3d(world, K[b:b + 1], pose, h, w, convention)
            warped = F.grid_sample(feat, locs, padding_mode="zeros", mode="bilinear",
                                   align_corners=(convention == MANYDEPTH))
            xv = (locs[..., 0] / 2 + 0.5) * (w - 1)
            yv = (locs[..., 1] / 2 + 0.5) * (h - 1)
            edge = ((xv >= 2.0) * (xv <= w - 2) * (yv >= 2.0) * (yv <= h - 2)).float()
            inner = torch.zeros_like(edge)
            inner[:, 2:-2, 2:-2] = 1.0
            edge = edge * inner
            diffs = torch.abs(warped - current_feats[b:b + 1]).mean(1) * edge
            cost = cost + diffs
            counts = counts + (diffs > 0).float()
        cost = cost / (counts + 1e-7)
        missing = (cost == 0).float()
        if set_missing_to_max:
            cost = cost * (1 - missing) + cost.max(0)[0].unsqueeze(0) * missing
        volumes.append(cost)
        masks.append(missing)
    return torch.stack(volumes, 0), torch.stack(masks, 0)


def occlusion_batch(lookup_images, h, w):
    """dynamicdepth/networks/resnet_encoder.py:160 (the reference hard-codes [48, 128], its
    matching resolution): black (< 0.15 summed RGB) pixels of the DOMD-processed lookup image."""
    return F.interpolate((lookup_images.sum(1).unsqueeze(1) < 0.15).float(), [h, w])


def match_features_dynamic(current_feats, lookup_feats, relative_poses, K, invK, bins, lookup_images, cv_min,
                           aug_mask, set_1, pool, pool_r, pool_th, set_missing_to_max=True):
    """DynamicDepth's match_features, dynamicdepth/networks/resnet_encoder.py:148-249: min over
    lookup frames (cv_min) and the occlusion fill of the warped features (set_1 / pool)."""
    B, C, h, w = current_feats.shape
    nb = len(bins)
    planes = bins.view(nb, 1, 1, 1).float().expand(nb, 1, h, w).contiguous()
    occ_batch = occlusion_batch(lookup_images, h, w)
    volumes, masks = [], []
    for b in range(B):
        if cv_min:
            cost = torch.ones(nb, h, w)
        else:
            cost, counts = torch.zeros(nb, h, w), torch.zeros(nb, h, w)
        world = backproject(planes, invK[b:b + 1])
        for li in range(lookup_feats.shape[1]):
            pose = relative_poses[b:b + 1, li]
            if pose.sum() == 0:
                continue
            feat = lookup_feats[b:b + 1, li].repeat([nb, 1, 1, 1])
            locs = project3d(world, K[b:b + 1], pose, h, w)
            warped = F.grid_sample(feat, locs, padding_mode="zeros", mode="bilinear", align_corners=True)
            if aug_mask[b][0][0][0] == 0 and (set_1 or pool):
                occ_mask = (occ_batch[b] > 0).unsqueeze(0).repeat([nb, C, 1, 1])
                mask = (F.grid_sample(occ_mask.float(), locs, padding_mode="zeros", mode="bilinear",
                                      align_corners=True) > pool_th).detach()
                if set_1:
                    warped[mask] = 1.0
                elif pool:
                    x = warped.clone()
                    x[mask] = 0
                    x = F.max_pool3d(x.permute(1, 0, 2, 3), pool_r * 2 + 1, stride=1, padding=pool_r).permute(1, 0, 2, 3)
                    warped[mask] = x[mask]
            xv = (locs[..., 0] / 2 + 0.5) * (w - 1)
            yv = (locs[..., 1] / 2 + 0.5) * (h - 1)
            edge = ((xv >= 2.0) * (xv <= w - 2) * (yv >= 2.0) * (yv <= h - 2)).float()
            inner = torch.zeros_like(edge)
            inner[:, 2:-2, 2:-2] = 1.0
            edge = edge * inner
            diffs = torch.abs(warped - current_feats[b:b + 1]).mean(1) * edge
            if cv_min:
                diffs[diffs == 0] = 1.0
                cost = torch.minimum(diffs, cost)
            else:
                cost = cost + diffs
                counts = counts + (diffs > 0).float()
        if cv_min:
            cost[cost == 1] = 0
        else:
            cost = cost / (counts + 1e-7)
        missing = (cost == 0).float()
        if set_missing_to_max:
            cost = cost * (1 - missing) + cost.max(0)[0].unsqueeze(0) * missing
        volumes.append(cost)
        masks.append(missing)
    return torch.stack(volumes, 0), torch.stack(masks, 0)


def confidence_mask(cost_volume, num_bins_threshold=None):
    """compute_confidence_mask, networks/resnet_encoder.py:255-262."""
    if num_bins_threshold is None:
        num_bins_threshold = cost_volume.shape[1]
    return ((cost_volume > 0).sum(1) == num_bins_threshold).float()


def lowest_cost(cost_volume, bins):
    """networks/resnet_encoder.py:309-313 + indices_to_disparity :247-253."""
    viz = cost_volume.clone()
    viz[viz == 0] = 100
    _, idx = torch.min(viz, 1)
    return 1 / bins[idx.reshape(-1)].reshape(idx.shape), idx


def cost_volume_head(current_feats, lookup_feats, relative_poses, K, invK, bins,
                     convention=MANYDEPTH):
    """ResnetEncoderMatching.forward :303-317: volume, confidence, argmin, masked volume."""
    cv, missing = match_features(current_feats, lookup_feats, relative_poses, K, invK, bins,
                                 convention)
    conf = confidence_mask(cv * (1 - missing))
    low, idx = lowest_cost(cv, bins)
    return cv * conf.unsqueeze(1), low, conf, idx, missing


# --------------------------------------------------------------------------
# DynamicDepth forward warp (z-buffered splat + inverse warp)
# --------------------------------------------------------------------------
def _pixel2cam(depth, intrinsics_inv):
    """dynamicdepth/rigid_warp.py:34-50 (set_id_grid :14-21)."""
    b, h, w = depth.shape
    i_range = torch.arange(0, h).view(1, h, 1).expand(1, h, w).type_as(depth)
    j_range = torch.arange(0, w).view(1, 1, w).expand(1, h, w).type_as(depth)
    ones = torch.ones(1, h, w).type_as(depth)
    pix = torch.stack((j_range, i_range, ones), dim=1).expand(b, 3, h, w).reshape(b, 3, -1)
    cam = (intrinsics_inv @ pix).reshape(b, 3, h, w)
    return cam * depth.unsqueeze(1)


def _euler2mat(angle):
    """dynamicdepth/rigid_warp.py:204-241."""
    B = angle.size(0)
    x, y, z = angle[:, 0], angle[:, 1], angle[:, 2]
    cosz, sinz = torch.cos(z), torch.sin(z)
    zeros = z.detach() * 0
    ones = zeros.detach() + 1
    zmat = torch.stack([cosz, -sinz, zeros, sinz, cosz, zeros, zeros, zeros, ones], dim=1).reshape(B, 3, 3)
    cosy, siny = torch.cos(y), torch.sin(y)
    ymat = torch.stack([cosy, zeros, siny, zeros, ones, zeros, -siny, zeros, cosy], dim=1).reshape(B, 3, 3)
    cosx, sinx = torch.cos(x), torch.sin(x)
    xmat = torch.stack([ones, zeros, zeros, zeros, cosx, -sinx, zeros, sinx, cosx], dim=1).reshape(B, 3, 3)
    return xmat @ ymat @ zmat


def quat2mat(quat):
    """dynamicdepth/rigid_warp.py:243-265."""
    nq = torch.cat([quat[:, :1].detach() * 0 + 1, quat], dim=1)
    nq = nq / nq.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = nq[:, 0], nq[:, 1], nq[:, 2], nq[:, 3]
    B = quat.size(0)
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz = w * x, w * y, w * z
    xy, xz, yz = x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).reshape(B, 3, 3)


def pose_vec2mat(vec, rotation_mode="euler"):
    """dynamicdepth/rigid_warp.py:268-284."""
    rot = _euler2mat(vec[:, 3:]) if rotation_mode == "euler" else quat2mat(vec[:, 3:])
    return torch.cat([rot, vec[:, :3].unsqueeze(-1)], dim=2)


def _mat2euler(R):
    """dynamicdepth/rigid_warp.py:175-200."""
    sy = torch.sqrt(R[:, 0, 0] * R[:, 0, 0] + R[:, 1, 0] * R[:, 1, 0])
    singular = (sy < 1e-6).float()
    x = torch.atan2(R[:, 2, 1], R[:, 2, 2])
    y = torch.atan2(-R[:, 2, 0], sy)
    z = torch.atan2(R[:, 1, 0], R[:, 0, 0])
    xs = torch.atan2(-R[:, 1, 2], R[:, 1, 1])
    ys = torch.atan2(-R[:, 2, 0], sy)
    zs = R[:, 1, 0] * 0
    return torch.stack([x * (1 - singular) + xs * singular, y * (1 - singular) + ys * singular,
                        z * (1 - singular) + zs * singular], dim=-1)


def forward_warp_matrices(pose, intrinsics, upscale):
    """The (B,3,3)/(B,3,4) constants forward_warp derives on the way (rigid_warp.py:562-589,
    inverse_warp :355-362): inverse of the up-scaled intrinsics, inverse intrinsics, and the
    target->source projection K @ [R(euler(inv pose)) | t(inv pose)]."""
    bs = pose.shape[0]
    intrinsic_u = torch.cat((intrinsics[:, 0:2] * upscale, intrinsics[:, 2:]), dim=1)
    aux = torch.tensor([0, 0, 0, 1]).type_as(pose).unsqueeze(0).unsqueeze(0).repeat(bs, 1, 1)
    pose_inv_mat = torch.inverse(torch.cat([pose, aux], dim=1))
    pose_inv = torch.cat([pose_inv_mat[:, :3, 3], _mat2euler(pose_inv_mat[:, :3, :3])], dim=1)
    pose_mat = torch.cat([_euler2mat(pose_inv[:, 3:]), pose_inv[:, :3].unsqueeze(-1)], dim=2)
    return intrinsic_u.inverse(), intrinsics.inverse(), intrinsics @ pose_mat


def forward_warp(img, depth, pose, intrinsics, upscale=3, matrices=None):
    """dynamicdepth/rigid_warp.forward_warp :534-597 (cam2pix_trans :513-530, inverse_warp
    :337-373, cam2pixel :54-83).  torch_sparse.coalesce(op='max') (:577; third-party, version
    unpinned, absent from the tree) is restated from its published semantics as a scatter-max of
    1/z onto the (hh+1, ww+1) grid whose last row / column collect the out-of-range points."""
    bs, _, hh, ww = depth.shape
    Ku_inv, K_inv, proj = matrices if matrices is not None else forward_warp_matrices(pose, intrinsics, upscale)
    depth_u = F.interpolate(depth, scale_factor=upscale).squeeze(1)
    cam = _pixel2cam(depth_u, Ku_inv)
    b, _, h, w = cam.shape
    rot, tr = pose[:, :, :3], pose[:, :, -1:]
    trans = rot @ cam.reshape(b, 3, -1) + tr
    X, Y, Z = trans[:, 0], trans[:, 1], trans[:, 2].clamp(min=1e-3)
    P_norm = torch.stack([X / Z, Y / Z, Z / Z], dim=1)
    pcoords = (intrinsics @ P_norm).permute(0, 2, 1)[:, :, :2].reshape(b, h, w, 2)
    depth_w, fw_val = [], []
    for coo, z in zip(pcoords, Z.reshape(b, 1, h, w)):
        idx = coo.reshape(-1, 2).permute(1, 0).long()[[1, 0]]
        val = z.reshape(-1)
        idx[0][idx[0] < 0] = hh
        idx[0][idx[0] > hh - 1] = hh
        idx[1][idx[1] < 0] = ww
        idx[1][idx[1] > ww - 1] = ww
        dense = torch.zeros((hh + 1) * (ww + 1)).scatter_reduce(0, idx[0] * (ww + 1) + idx[1], 1 / val, "amax",
                                                                include_self=False)
        dense = dense.view(hh + 1, ww + 1)[:-1, :-1]
        depth_w.append(1 / dense)
        fw_val.append(1 - (dense == 0).float())
    depth_w, fw_val = torch.stack(depth_w, 0), torch.stack(fw_val, 0)
    depth_w[fw_val == 0] = 0
    # inverse_warp(img, depth_w, pose_inv, intrinsics) with zeros padding
    cam2 = _pixel2cam(depth_w, K_inv)
    pc = proj[:, :, :3] @ cam2.reshape(bs, 3, -1) + proj[:, :, -1:]
    Xs, Ys, Zs = pc[:, 0], pc[:, 1], pc[:, 2].clamp(min=1e-3)
    grid = torch.stack([2 * (Xs / Zs) / (ww - 1) - 1, 2 * (Ys / Zs) / (hh - 1) - 1], dim=2).reshape(bs, hh, ww, 2)
    img_w = F.grid_sample(img, grid, padding_mode="zeros", align_corners=True)
    iw_val = (grid.abs().max(dim=-1)[0] <= 1).float().unsqueeze(1)
    depth_w = depth_w.unsqueeze(1)
    valid = fw_val.unsqueeze(1) * iw_val
    return img_w * valid, depth_w * valid, valid


# --------------------------------------------------------------------------
# MAL temporal hint: image synthesis from matched instance masks
# --------------------------------------------------------------------------
def fill_dynamic_obj(mask, delta_x, delta_y, source, img):
    """manydepth/dyn_utils.py:5-36 (TorchScript in the reference; same tensor ops here)."""
    N, H, W = mask.shape
    start_hl = torch.max(torch.zeros_like(delta_x), delta_x)
    end_hl = torch.min(torch.ones_like(delta_x) * H, H + delta_x)
    start_hr = torch.max(torch.zeros_like(delta_x), -delta_x)
    end_hr = torch.min(torch.ones_like(delta_x) * H, H - delta_x)
    start_wl = torch.max(torch.zeros_like(delta_y), delta_y)
    end_wl = torch.min(torch.ones_like(delta_y) * W, W + delta_y)
    start_wr = torch.max(torch.zeros_like(delta_y), -delta_y)
    end_wr = torch.min(W - delta_y, torch.ones_like(delta_y) * W)
    chn = img.shape[0]
    source_mv = torch.zeros((N, chn, H, W))
    mask_mv = torch.zeros(mask.shape, dtype=torch.bool)
    for i in range(len(mask)):
        a, b, c, d = int(start_hl[i]), int(end_hl[i]), int(start_wl[i]), int(end_wl[i])
        e, f, g, h = int(start_hr[i]), int(end_hr[i]), int(start_wr[i]), int(end_wr[i])
        if b > a and d > c:
            source_mv[i, :, a:b, c:d] = source[:, e:f, g:h]
            mask_mv[i, a:b, c:d] = mask[i, e:f, g:h]
    img_mv = mask_mv.unsqueeze(1).repeat(1, chn, 1, 1) * source_mv
    img_sum = img_mv.sum(dim=0)
    mask_or = torch.zeros((H, W), dtype=torch.bool)
    for m in mask_mv:
        mask_or = mask_or | m
    return torch.where(mask_or, img_sum, img)


def _extents(mask, grid, axis_sum, idx):
    """low/top (or right/left) of one set of masks, manydepth/dyn_utils.py:56-78."""
    inf = (mask.shape[1] + 1) * (mask.shape[2] + 1)
    s = (mask * grid).sum(dim=axis_sum)
    nz = torch.where(s == 0, 0, idx)
    hi = nz.argmax(dim=1)
    nz = torch.where(nz == 0, inf, nz)
    lo = nz.argmin(dim=1)
    return hi, lo


def generate_dynamic_instance(mask_last, mask_next, img_last, img_next, replace=False):
    """manydepth/dyn_utils.py:38-119 -> (ori_last, ori_next, (dx_last, dy_last, dx_next, dy_next))."""
    mask_or = mask_last | mask_next
    mask_or_ = torch.zeros_like(mask_or[0])
    for m in mask_or:
        mask_or_ = mask_or_ | m
    num, H, W = mask_last.shape
    x, y = torch.arange(H), torch.arange(W)
    grid_h, grid_w = torch.meshgrid(x, y, indexing="ij")
    grid_h, grid_w = grid_h.repeat(num, 1, 1), grid_w.repeat(num, 1, 1)
    low_last, top_last = _extents(mask_last, grid_h, 2, x)
    right_last, left_last = _extents(mask_last, grid_w, 1, y)
    low_next, top_next = _extents(mask_next, grid_h, 2, x)
    right_next, left_next = _extents(mask_next, grid_w, 1, y)
    sl = torch.arange(num)
    delta_x = torch.stack([low_next - low_last, top_next - top_last], dim=1)
    delta_x_sel = delta_x[sl, delta_x.abs().argmax(dim=1)]
    delta_y = torch.stack([right_next - right_last, left_next - left_last], dim=1)
    delta_y_sel = delta_y[sl, delta_y.abs().argmax(dim=1)]
    disp_x = torch.round(delta_x_sel / 2).long()
    disp_y = torch.round(delta_y_sel / 2).long()
    if replace:
        dxl = torch.where(disp_x.abs() < 3, 0, disp_x)
        dyl = torch.where(disp_y.abs() < 3, 0, disp_y)
        dxn = torch.where(disp_x.abs() < 3, 0, -disp_x)
        dyn = torch.where(disp_y.abs() < 3, 0, -disp_y)
    else:
        dxl, dyl, dxn, dyn = disp_x, disp_y, -disp_x, -disp_y
    mask = mask_last & (~mask_next)
    mask_bg = torch.zeros_like(mask[0])
    for ms in mask:
        mask_bg = mask_bg | ms
    img_bg = torch.where(mask_bg, img_next, img_last)
    mask2 = mask_next & (~mask_last)
    mask_bg2 = torch.zeros_like(mask2[0])
    for ms in mask2:
        mask_bg2 = mask_bg2 | ms
    img_bg2 = torch.where(mask_bg2, img_last, img_next)
    syn_last = fill_dynamic_obj(mask_last, dxl, dyl, img_last, img_bg)
    ori_last = torch.where(mask_or_, syn_last, img_last)
    syn_next = fill_dynamic_obj(mask_next, dxn, dyn, img_next, img_bg2)
    ori_next = torch.where(mask_or_, syn_next, img_next)
    return ori_last, ori_next, (dxl, dyl, dxn, dyn)


# --------------------------------------------------------------------------
# host-side loss balancers (fp64 numpy state, as the reference)
# --------------------------------------------------------------------------
class LossBalancing:
    """manydepth/loss_utils.py:283-345."""

    def __init__(self, num_loss, num_train_data, bs):
        self.num_loss, self.num_data, self.bs = num_loss, num_train_data, bs
        self.w_list = np.array([1. / num_loss, 1. / num_loss])
        self.scale0 = np.array([1. / num_loss, 1. / num_loss])
        self.train_scores = np.zeros((num_train_data, num_loss))
        self.initialised = False
        self.last_rebalancing_iter = 0
        self.previous_total_loss = 0
        self.previous_loss = 0

    def compute_loss(self, loss_list, index_iter):
        loss = 0
        for ib in range(self.bs):
            rec = self.bs * index_iter + ib
            if rec < self.num_data:
                for k in range(self.num_loss):
                    loss = loss + self.w_list[k] * loss_list[k]
                for k in range(self.num_loss):
                    self.train_scores[rec, k] = float(loss_list[k])
        return loss

    def update_weight(self, i, lam):
        mean = self.train_scores[self.last_rebalancing_iter * self.bs:(i + 1) * self.bs].mean(axis=0)
        total = np.sum(mean * self.w_list)
        if not self.initialised:
            for k in range(self.num_loss):
                self.w_list[k] = (total * self.scale0[k]) / mean[k]
            self.initialised = True
        else:
            prev_w = np.array(self.w_list)
            if self.previous_total_loss > 0:
                for k in range(self.num_loss):
                    adj = 1 + lam * ((total / self.previous_total_loss) *
                                     (self.previous_loss[k] / mean[k]) - 1)
                    self.w_list[k] = prev_w[k] * min(max(adj, 0.5), 2.0)
        self.previous_total_loss = np.sum(mean * self.w_list)
        self.previous_loss = mean
        return self.w_list[0], self.w_list[1]


# --------------------------------------------------------------------------
# DualRefine epipolar correlation lookup
def corr_pyramid(fmap2, num_levels):
    """dualrefine/networks/corr.py:11-22 (CoordSampler.register): fmap2 and its avg-pooled levels."""
    levels = [fmap2]
    f2 = fmap2
    for _ in range(num_levels - 1):
        f2 = F.avg_pool2d(f2, 2, stride=2)
        levels.append(f2)
    return levels


def corr_lookup(fmap1, pyramid, coords, num_head=1):
    """dualrefine/networks/corr.py:24-50 (CoordSampler.__call__; __corr__ :52-76 is num_head == 1).

    fmap1 (B,C,h,w), pyramid = corr_pyramid(fmap2, L), coords (B,2,L,D,h,w) -> (B, L*heads*D, h, w)."""
    batch, _, n1, d1, h1, w1 = coords.shape
    c = coords.permute(2, 0, 4, 5, 3, 1).reshape(n1, batch, h1 * w1, d1, 2)
    f1 = fmap1[..., None]
    out_pyramid = []
    for i in range(n1):
        f2 = pyramid[i]
        xgrid, ygrid = c[i].split([1, 1], dim=-1)
        xgrid = 2 * (xgrid + 0.5) / (w1) - 1
        ygrid = 2 * (ygrid + 0.5) / (h1) - 1
        grid = torch.cat([xgrid, ygrid], dim=-1)
        f2 = F.grid_sample(f2, grid, align_corners=False)
        f2 = f2.view(batch, -1, h1, w1, d1)
        corr = torch.abs(f1 - f2)
        corr = corr.view(batch, num_head, -1, h1, w1, d1).mean(2)
        corr = corr.permute(0, 2, 3, 1, 4).reshape(batch, h1, w1, -1)
        out_pyramid.append(corr)
    out = torch.cat(out_pyramid, dim=-1)
    return out.permute(0, 3, 1, 2).contiguous().float()


def sample_tgt(tgt_feat, p2, tgt_w):
    """dualrefine/networks/utils/utils.py:383-404 (PoseUpdate.sample_tgt): the target features at the
    projected point and its +-1 pixel neighbours -> warped features, central-difference gradients, and the
    warped confidence weight.  p2 is (B, 2, 1, 5, h, w) from depth2gradcoords (:213-231)."""
    batch, _, n1, d1, h1, w1 = p2.shape
    p2 = p2.permute(2, 0, 4, 5, 3, 1).reshape(batch, h1 * w1, d1, 2)
    xgrid, ygrid = p2.split([1, 1], dim=-1)
    xgrid = 2 * (xgrid + 0.5) / (w1) - 1
    ygrid = 2 * (ygrid + 0.5) / (h1) - 1
    grid = torch.cat([xgrid, ygrid], dim=-1)
    f = F.grid_sample(tgt_feat, grid, align_corners=False)
    f = f.view(batch, -1, h1, w1, d1)
    warped_tgt_feat = f[..., 0]
    warped_tgt_gradients = torch.stack([(f[..., 1] - f[..., 2]) / 2, (f[..., 3] - f[..., 4]) / 2], dim=-1)
    grid_0 = grid[:, :, :1]
    warped_tgt_w = F.grid_sample(tgt_w.type(grid_0.dtype), grid_0, align_corners=False).reshape(batch, 1, h1, w1)
    return warped_tgt_feat, warped_tgt_gradients, warped_tgt_w
