"""The pure (state-free) hot-path methods of the reference Trainers, as functions over `opt`.

    generate_images_pred            manydepth/trainer.py:1078-1165   (dualrefine/trainer.py:395-455,
                                                                      dynamicdepth/trainer.py:906-955)
    generate_images_pred_ensemble   manydepth/trainer.py:1172-1207
    compute_matching_mask           manydepth/trainer.py:1066-1076
    compute_losses                  manydepth/trainer.py:1248-1475   (non-distil path, opt.sclm+1 scales)
    process_batch_losses            manydepth/trainer.py:574-644     (the loss half of process_batch)

`opt` is any object with the reference's option names (height, width, min_depth, max_depth,
frame_ids, sclm, batch_size, temporal, main_temporal, distil, no_ens, loss_blc, dual_distil,
disable_automasking, disable_motion_masking, no_matching_augmentation, disparity_smoothness,
no_ssim, v1_multiscale, ensemble).  Missing names take the reference's defaults
(manydepth/options.py).

generate_images_pred does not materialise warped images by default: it leaves a WarpSpec in
outputs[("warp_spec", scale)] and the loss functions run the fused kernel.  Pass
materialize=True to also get outputs[("sample", f, s)] / ("color", f, s) like the reference.
"""
from __future__ import annotations

import torch

from . import ops, raw
from .layers import disp_to_depth
from .loss_utils import WarpSpec, _draw_noise, compute_main_losses, compute_mono_losses, identity_reprojection

_DEFAULTS = dict(height=192, width=640, min_depth=0.1, max_depth=100.0, frame_ids=[0, -1, 1], sclm=0,
                 temporal=False, main_temporal=False, distil=False, no_ens=False, loss_blc=False,
                 dual_distil=False, learn_ens=False, pareto=False, disable_automasking=False,
                 disable_motion_masking=False, no_matching_augmentation=False,
                 disparity_smoothness=1e-3, no_ssim=False, v1_multiscale=False, ensemble=False,
                 convention=raw.CONV_MANYDEPTH)


def _o(opt, name):
    return getattr(opt, name, _DEFAULTS[name])


class _Ssim:
    """Carrier for opt.no_ssim where a function wants the reference's `ssim` module argument."""

    def __init__(self, no_ssim):
        self.no_ssim = bool(no_ssim)


def generate_images_pred(inputs, outputs, opt, is_multi=False, materialize=False):
    """Per scale: up-sample disp, disp -> depth, and describe the two source-frame warps."""
    H, W = _o(opt, "height"), _o(opt, "width")
    v1 = bool(_o(opt, "v1_multiscale"))
    for scale in range(_o(opt, "sclm") + 1):
        disp = outputs[("disp", scale)]
        # trainer.py:1089-1095: v1_multiscale warps at the disparity's own scale (images, K of that scale);
        # otherwise the disparity is up-sampled to the full image.  The fused kernels read the
        # low-resolution disparity directly (mal_photo_args.depth_height); the full-resolution depth map
        # is only built for the dict.
        ss = scale if v1 else 0
        hs, ws = (disp.shape[-2], disp.shape[-1]) if v1 else (H, W)
        disp_full = disp if disp.shape[-2:] == (hs, ws) else ops.upsample_bilinear(disp, (hs, ws))
        _, depth = disp_to_depth(disp_full, _o(opt, "min_depth"), _o(opt, "max_depth"))
        outputs[("depth", 0, scale)] = depth
        Ts = []
        for frame_id in _o(opt, "frame_ids")[1:]:
            T = outputs[("cam_T_cam", 0, frame_id)]
            Ts.append(T.detach() if is_multi else T)   # "don't update posenet based on multi frame prediction"
            if not _o(opt, "disable_automasking"):
                outputs[("color_identity", frame_id, scale)] = inputs[("color", frame_id, ss)]
        outputs[("warp_spec", scale)] = WarpSpec(disp, inputs[("K", ss)], inputs[("inv_K", ss)], Ts,
                                                 _o(opt, "convention"), _o(opt, "min_depth"), _o(opt, "max_depth"),
                                                 source_scale=ss)
        if materialize:
            from .layers import BackprojectDepth, Project3D
            B = disp.shape[0]
            back, proj = BackprojectDepth(B, hs, ws), Project3D(B, hs, ws, convention=_o(opt, "convention"))
            for T, frame_id in zip(Ts, _o(opt, "frame_ids")[1:]):
                cam = back(depth, inputs[("inv_K", ss)])
                pix = proj(cam, inputs[("K", ss)], T)
                outputs[("sample", frame_id, scale)] = pix
                outputs[("color", frame_id, scale)] = ops.grid_sample(
                    inputs[("color", frame_id, ss)], pix, padding_mode="border",
                    align_corners=_o(opt, "convention") == raw.CONV_MANYDEPTH)
    return outputs


def generate_images_pred_ensemble(inputs, T_l, T_n, disp, opt):
    """min over the two source frames of the reprojection loss under `disp` (no gradient)."""
    # trainer.py:1176-1177: a low-resolution disparity is up-sampled inside the kernel
    with torch.no_grad():
        _, min_reproj, _ = ops.photo(inputs[("color", 0, 0)], [inputs[("color", f, 0)] for f in (-1, 1)],
                                     depth=disp.detach(), K=inputs[("K", 0)], inv_K=inputs[("inv_K", 0)],
                                     T=[T_l.detach(), T_n.detach()], mode=raw.PHOTO_WARP,
                                     convention=_o(opt, "convention"), depth_is_disp=True,
                                     no_ssim=_o(opt, "no_ssim"), min_depth=_o(opt, "min_depth"),
                                     max_depth=_o(opt, "max_depth"))
    return min_reproj


def compute_matching_mask(outputs):
    """Where the cost-volume depth and the teacher depth agree within a factor of two -> bool (B,H,W)."""
    return ops.matching_mask(outputs["lowest_cost"], outputs[("mono_depth", 0, 0)]) > 0


def compute_losses(inputs, outputs, opt, is_multi=False, has_ins=False, noises=None):
    """Trainer.compute_losses: the non-distil loss over opt.sclm+1 scales.  Returns (losses, [])."""
    losses, total_loss = {}, 0
    ssim = _Ssim(_o(opt, "no_ssim"))
    num_scales = _o(opt, "sclm") + 1
    idents = {}
    for scale in range(num_scales):
        spec = outputs[("warp_spec", scale)]
        ss = spec.source_scale                      # 0 unless v1_multiscale (trainer.py:1260-1263)
        target = inputs[("color", 0, ss)]
        with_syn = (not is_multi) and _o(opt, "temporal") and has_ins
        kw = dict(src=[inputs[("color", f, ss)] for f in (-1, 1)],
                  syn=[outputs[("syn", f, scale)] for f in (-1, 1)] if with_syn else None, depth=spec.disp,
                  K=spec.K, inv_K=spec.inv_K, T=spec.T, convention=spec.convention, depth_is_disp=True,
                  min_depth=spec.min_depth, max_depth=spec.max_depth, no_ssim=ssim.no_ssim)
        consistency_loss = 0
        if not is_multi:
            # trainer.py:1292-1311: the identity loss is ALWAYS computed and compared; disable_automasking
            # only drops the tie-break noise (and its randn draw)
            if ss not in idents:
                idents[ss] = identity_reprojection(ssim, inputs, ss)
            ident = idents[ss]
            if not _o(opt, "disable_automasking"):
                noise = _draw_noise(ident.shape, target.device, None if noises is None else noises[scale])
            else:
                noise = torch.zeros_like(ident)
            kw.update(identity_min=ident, noise=noise)
            sums, _, sel = ops.photo(target, **kw)
        else:
            if not _o(opt, "disable_automasking") and noises is None:
                torch.randn(target.shape[0], 1, *target.shape[-2:])   # keep the reference's RNG stream
            pm = outputs["consistency_mask"] if not _o(opt, "disable_motion_masking") else torch.ones_like(target[:, 0])
            sm = None if _o(opt, "no_matching_augmentation") else outputs["augmentation_mask"][:opt.batch_size]
            sums, multi_reproj, sel = ops.photo(target, pixel_mask=pm, sample_mask=sm, **kw)
            consistency_loss, _, _, tgt = ops.main_terms(
                outputs[("depth", 0, scale)], outputs[("mono_depth", 0, scale)].detach(), pm, sm,
                multi_reproj, None, multi_reproj)
            outputs["consistency_target/{}".format(scale)] = tgt
            losses["consistency_loss/{}".format(scale)] = consistency_loss
            if _o(opt, "ensemble"):
                multi_depth, mono_depth = outputs[("depth", 0, scale)], outputs[("mono_depth", 0, scale)].detach()
                mask = pm.unsqueeze(1) * (1 - sm if sm is not None else 1)
                ensemble_loss = (torch.abs((mono_depth + multi_depth) / 2.0 - multi_depth) * mask).mean()
                losses["ensemble_loss/{}".format(scale)] = ensemble_loss
                consistency_loss = consistency_loss + ensemble_loss
        outputs[("mal_selection", scale)] = sel
        losses["reproj_loss/{}".format(scale)] = sums[2]
        loss = sums[2] + consistency_loss
        smooth_loss = ops.smooth(outputs[("disp", scale)], inputs[("color", 0, scale)], normalise=True)
        loss = loss + _o(opt, "disparity_smoothness") * smooth_loss / (2 ** scale)
        total_loss = total_loss + loss
        losses["loss/{}".format(scale)] = loss
    losses["loss"] = total_loss / num_scales
    return losses, []


def generate_images_pred_dualrefine(inputs, outputs, opt):
    """dualrefine/trainer.py generate_images_pred :395-455: one WarpSpec per (scale, deq_iter) with
    DualRefine's convention (half-pixel Project3D, align_corners=False) and pose-detach rules."""
    H, W = _o(opt, "height"), _o(opt, "width")
    for scale in opt.scales:
        n = getattr(opt, "n_losses", 1) + 1 if scale in (0, 1, 2) else 1
        for it in range(n):
            if scale == 1:
                continue
            disp = outputs[("disp", scale, it)]
            # dualrefine/trainer.py:412-413: the fused kernels read the low-resolution disparity directly
            disp_full = disp if disp.shape[-2:] == (H, W) else ops.upsample_bilinear(disp, (H, W))
            _, depth = disp_to_depth(disp_full, _o(opt, "min_depth"), _o(opt, "max_depth"))
            outputs[("depth", 0, scale, it)] = depth
            Ts = []
            for frame_id in (-1, 1):
                if frame_id == 1:
                    T = outputs[("cam_T_cam", 0, frame_id)]
                    T = T.detach() if it > 0 else T
                elif it > 0:
                    T = (outputs[("cam_T_cam", 0, frame_id)].detach() if getattr(opt, "Dstar_T0_pair", False)
                         else outputs[("cam_T_cam", 0, frame_id, 1)])
                else:
                    T = outputs[("cam_T_cam", 0, frame_id)]
                Ts.append(T)
            outputs[("warp_spec", scale, it)] = WarpSpec(disp, inputs[("K", 0)], inputs[("inv_K", 0)], Ts,
                                                         raw.CONV_DUALREFINE, _o(opt, "min_depth"), _o(opt, "max_depth"))
    return outputs


def compute_losses_dualrefine(inputs, outputs, opt, noises=None):
    """dualrefine/trainer.py compute_losses :530-697 (f_thres > 0 branch): per (scale, deq_iter)
    fused photometric pass with automask x consistency mask, consistency towards the deq_iter-0
    depth, smoothness.  Returns the losses dict."""
    avg = bool(getattr(opt, "avg_reprojection", False))   # mean instead of min over frames (:575-586): a kernel mode
    losses, total_loss = {}, 0
    target = inputs[("color", 0, 0)]
    ssim = _Ssim(_o(opt, "no_ssim"))
    automask = not _o(opt, "disable_automasking")
    ident, draw = None, 0
    for scale in opt.scales:
        loss = 0
        n = getattr(opt, "n_losses", 1) + 1 if scale in (0, 1, 2) else 1
        for it in range(n):
            if scale == 1:
                continue
            spec = outputs[("warp_spec", scale, it)]
            kw = dict(src=[inputs[("color", f, 0)] for f in (-1, 1)], depth=spec.disp, K=spec.K, inv_K=spec.inv_K,
                      T=spec.T, convention=spec.convention, depth_is_disp=True, min_depth=spec.min_depth,
                      max_depth=spec.max_depth, no_ssim=ssim.no_ssim, avg_reprojection=avg)
            if automask:
                ident = identity_reprojection(ssim, inputs, avg_reprojection=avg) if ident is None else ident
                kw.update(identity_min=ident,
                          noise=_draw_noise(ident.shape, target.device, None if noises is None else noises[draw]))
                draw += 1
            pm = None
            if it > 0 and not _o(opt, "disable_motion_masking"):
                pm = outputs["consistency_mask"].reshape(target.shape[0], *target.shape[-2:])
            sums, min_reproj, sel = ops.photo(target, pixel_mask=pm, **kw)
            consistency_loss = 0
            if it > 0:
                # consistency_mask = 1 - (automask * consistency mask): the final per-pixel weight
                weight = (sel >> 7).float()[:, 0] * (pm if pm is not None else 1.0)
                consistency_loss, _, _, tgt = ops.main_terms(
                    outputs[("depth", 0, scale, it)], outputs[("depth", 0, scale, 0)].detach(), weight.contiguous(),
                    None, min_reproj, None, min_reproj)
                outputs["consistency_target/{}_{}".format(scale, it)] = tgt
                losses["consistency_loss/{}_{}".format(scale, it)] = consistency_loss
            outputs[("mal_selection", scale, it)] = sel
            losses["reproj_loss/{}".format(scale)] = sums[2]
            loss = loss + sums[2] + consistency_loss
            smooth_loss = ops.smooth(outputs[("disp", scale, it)], inputs[("color", 0, scale)], normalise=True)
            loss = loss + _o(opt, "disparity_smoothness") * smooth_loss / (2 ** scale)
            total_loss = total_loss + loss
            losses["loss/{}_{}".format(scale, it)] = loss
    losses["loss"] = total_loss / len(opt.scales)
    return losses


def compute_reprojection_loss_dynamicdepth(pred, target, zero_img=True, no_ssim=False):
    """dynamicdepth/trainer.py compute_reprojection_loss :958-975: with zero_img the dark pixels of
    `pred` (DOMD warping holes, RGB sum < 0.1) are zeroed in a copy of pred and IN PLACE in `target`
    (the reference mutates the caller's target image; so do we)."""
    if zero_img:
        mask = (pred.sum(1) < 0.1).unsqueeze(1).repeat([1, 3, 1, 1]).detach()
        pred = pred.clone()
        pred[mask] = 0
        target[mask] = 0
        # later calls keep zeroing `target` in place: hand the kernel (which keeps its inputs for the
        # backward) a snapshot of the image as this call saw it
        return ops.reprojection_loss_map(pred, target.detach().clone(), no_ssim=no_ssim)
    return ops.reprojection_loss_map(pred, target, no_ssim=no_ssim)


def compute_losses_dynamicdepth(inputs, outputs, opt, is_multi=False, noises=None):
    """dynamicdepth/trainer.py compute_losses :1006-1128 over opt.scales, with `selec_reproj`
    (:1058-1064), `zero_img` (:961-965), `avg_reprojection` and `no_teacher_warp`.

    zero_img makes every candidate's loss depend on the order of the calls before it (each call
    zeroes more of the shared target), so this trainer variant scores one materialised warp at a
    time (`generate_images_pred(..., materialize=True)`) instead of using the fused multi-candidate
    pass; the SSIM+L1 maps and their gradients still come from the photometric kernel (PRED mode)."""
    if str(getattr(opt, "feat_loss", "false")) == "true":
        raise NotImplementedError("feat_loss (get_feature_metric_loss) is outside the hot path")
    if ("warp_spec", list(opt.scales)[0]) in outputs and not getattr(opt, "avg_reprojection", False):
        return _compute_losses_dynamicdepth_fused(inputs, outputs, opt, is_multi, noises)
    zero_img, no_ssim = getattr(opt, "zero_img", True), str(_o(opt, "no_ssim")) in ("True", "true")
    rl = lambda p, t: compute_reprojection_loss_dynamicdepth(p, t, zero_img, no_ssim)
    avg = getattr(opt, "avg_reprojection", False)
    losses, total_loss = {}, 0
    scales = list(opt.scales)
    for si, scale in enumerate(scales):
        disp, color, target = outputs[("disp", scale)], inputs[("color", 0, scale)], inputs[("color", 0, 0)]
        cands = torch.cat([rl(outputs[("color", f, scale)], target) for f in (-1, 1)], 1)
        ident = None
        if not _o(opt, "disable_automasking"):
            use_ori = (not is_multi) and getattr(opt, "no_teacher_warp", False) and not getattr(opt, "train_teacher_only", False)
            preds = [inputs[("ori_color" if use_ori else "color", f, 0)] for f in (-1, 1)]
            ident = torch.cat([rl(p, target) for p in preds], 1)
            ident = ident.mean(1, keepdim=True) if avg else torch.min(ident, dim=1, keepdim=True)[0]
        reproj = cands.mean(1, keepdim=True) if avg else torch.min(cands, dim=1, keepdim=True)[0]
        if getattr(opt, "selec_reproj", True):
            maskm1 = (outputs[("color", -1, scale)].sum(1) < 0.1).detach()
            maskp1 = (outputs[("color", 1, scale)].sum(1) < 0.1).detach()
            maskand = (maskm1 * maskp1).detach()
            reproj = reproj.clone()
            reproj[maskm1.unsqueeze(1)] = (cands[:, 1, :, :])[maskm1]
            reproj[maskp1.unsqueeze(1)] = (cands[:, 0, :, :])[maskp1]
            reproj[maskand.unsqueeze(1)] = 0
        if ident is not None:
            ident = ident + _draw_noise(ident.shape, target.device, None if noises is None else noises[si]) * 0.00001
        mask = (~(ident < reproj)).float() if ident is not None else torch.ones_like(reproj)
        consistency_loss = 0
        if is_multi:
            mask = torch.ones_like(mask)
            if not _o(opt, "disable_motion_masking"):
                mask = mask * outputs["consistency_mask"].unsqueeze(1)
            if not str(_o(opt, "no_matching_augmentation")) == "true":
                mask = mask * (1 - outputs["augmentation_mask"])
            consistency_mask = (1 - mask).float()
        reprojection_loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
        if is_multi:
            multi_depth, mono_depth = outputs[("depth", 0, scale)], outputs[("mono_depth", 0, scale)].detach()
            consistency_loss = (torch.abs(multi_depth - mono_depth) * consistency_mask).mean()
            outputs["consistency_target/{}".format(scale)] = 1 / (mono_depth * consistency_mask +
                                                                  multi_depth.detach() * (1 - consistency_mask))
            losses["consistency_loss/{}".format(scale)] = consistency_loss
        losses["reproj_loss/{}".format(scale)] = reprojection_loss
        loss = reprojection_loss + consistency_loss
        loss = loss + _o(opt, "disparity_smoothness") * ops.smooth(disp, color, normalise=True) / (2 ** scale)
        total_loss = total_loss + loss
        losses["loss/{}".format(scale)] = loss
        outputs[("mal_mask", scale)] = mask
    losses["loss"] = total_loss / len(scales)
    return losses


def generate_images_pred_dynamicdepth(inputs, outputs, opt, is_multi=False):
    """dynamicdepth/trainer.py generate_images_pred :906-955 for the fused path: one WarpSpec per scale of
    opt.scales (the disparity of scale s is up-sampled to the full image inside the kernel), depth maps for the
    dict; `compute_losses_dynamicdepth` then runs one fused pass per scale.  Pass materialize-style warped images
    yourself (outputs[("color", f, s)]) to use the op-by-op path instead."""
    H, W = _o(opt, "height"), _o(opt, "width")
    for scale in opt.scales:
        disp = outputs[("disp", scale)]
        disp_full = disp if disp.shape[-2:] == (H, W) else ops.upsample_bilinear(disp, (H, W))
        _, depth = disp_to_depth(disp_full, _o(opt, "min_depth"), _o(opt, "max_depth"))
        outputs[("depth", 0, scale)] = depth
        Ts = [outputs[("cam_T_cam", 0, f)].detach() if is_multi else outputs[("cam_T_cam", 0, f)] for f in (-1, 1)]
        outputs[("warp_spec", scale)] = WarpSpec(disp, inputs[("K", 0)], inputs[("inv_K", 0)], Ts, raw.CONV_MANYDEPTH,
                                                 _o(opt, "min_depth"), _o(opt, "max_depth"))
    return outputs


def _compute_losses_dynamicdepth_fused(inputs, outputs, opt, is_multi, noises):
    """compute_losses_dynamicdepth with ONE fused photometric pass per scale (photo_kernel<..., DD>): the warps are
    made in the kernel, the identity candidates ride along as candidates 2 and 3, zero_img's cumulative zeroing of
    the target is a byte mask per tile element, selec_reproj a per-pixel override; the target as a scale leaves it
    is handed to the next scale and finally written back into inputs[("color", 0, 0)] like the reference's in-place
    edits (dynamicdepth/trainer.py:961-965)."""
    zero_img, no_ssim = getattr(opt, "zero_img", True), str(_o(opt, "no_ssim")) in ("True", "true")
    selec = getattr(opt, "selec_reproj", True)
    automask = not _o(opt, "disable_automasking")
    losses, total_loss = {}, 0
    scales = list(opt.scales)
    target0 = inputs[("color", 0, 0)]
    target = target0
    for si, scale in enumerate(scales):
        spec = outputs[("warp_spec", scale)]
        src = [inputs[("color", f, 0)] for f in (-1, 1)]
        ident = None
        if automask:
            use_ori = (not is_multi) and getattr(opt, "no_teacher_warp", False) and not getattr(opt, "train_teacher_only", False)
            ident = [inputs[("ori_color" if use_ori else "color", f, 0)] for f in (-1, 1)]
        noise = None
        if automask:
            noise = _draw_noise((target.shape[0], 1) + tuple(target.shape[-2:]), target.device,
                                None if noises is None else noises[si])
        pm = sm = None
        if is_multi:
            pm = outputs["consistency_mask"] if not _o(opt, "disable_motion_masking") else torch.ones_like(target[:, 0])
            if not str(_o(opt, "no_matching_augmentation")) == "true":
                sm = outputs["augmentation_mask"]
        res = ops.photo(target, src, syn=ident, depth=spec.disp, K=spec.K, inv_K=spec.inv_K, T=spec.T, noise=noise,
                        pixel_mask=pm, sample_mask=sm, mode=raw.PHOTO_WARP, convention=raw.CONV_MANYDEPTH,
                        depth_is_disp=True, no_ssim=no_ssim, min_depth=spec.min_depth, max_depth=spec.max_depth,
                        zero_img=zero_img, selec_reproj=selec, ignore_automask=is_multi, identity_in_pass=True)
        sums, _, sel = res[0], res[1], res[2]
        if zero_img:
            target = res[3]
        consistency_loss = 0
        if is_multi:
            weight = (pm if pm is not None else 1.0) * (1 - sm.reshape(-1, 1, 1) if sm is not None else 1.0)
            multi_depth, mono_depth = outputs[("depth", 0, scale)], outputs[("mono_depth", 0, scale)].detach()
            consistency_mask = (1 - weight).unsqueeze(1).float()
            consistency_loss = (torch.abs(multi_depth - mono_depth) * consistency_mask).mean()
            outputs["consistency_target/{}".format(scale)] = 1 / (mono_depth * consistency_mask +
                                                                  multi_depth.detach() * (1 - consistency_mask))
            losses["consistency_loss/{}".format(scale)] = consistency_loss
        losses["reproj_loss/{}".format(scale)] = sums[2]
        loss = sums[2] + consistency_loss
        # scale 0's smoothness reads the image the reprojection calls have just zeroed in place (:961-965, :1114)
        color = target if (zero_img and scale == 0) else inputs[("color", 0, scale)]
        loss = loss + _o(opt, "disparity_smoothness") * ops.smooth(outputs[("disp", scale)], color,
                                                                   normalise=True) / (2 ** scale)
        total_loss = total_loss + loss
        losses["loss/{}".format(scale)] = loss
        outputs[("mal_selection", scale)] = sel
    if zero_img:
        with torch.no_grad():
            target0.copy_(target)   # the reference zeroes the caller's image in place
    losses["loss"] = total_loss / len(scales)
    return losses


def process_batch_losses(inputs, mono_outputs, outputs, opt, *, has_ins=False, multi_has_ins=False,
                         loss_blc=None, index_iter=0, current_lambda_for_adjust=0.0, w_list=None,
                         noises=None, freeze_tp=False):
    """Everything Trainer.process_batch does after the networks have run (trainer.py:574-644):
    teacher warps + losses, teacher -> student hand-over, matching mask, ensemble reprojection,
    student warps + MAL losses, loss balancing.  Returns (outputs, losses)."""
    ssim = _Ssim(_o(opt, "no_ssim"))
    temporal = _o(opt, "temporal")
    generate_images_pred(inputs, mono_outputs, opt)
    has_ins = has_ins and temporal
    if _o(opt, "distil"):
        mono_losses, mono_reproj = compute_mono_losses(ssim, inputs, mono_outputs, temporal, has_ins,
                                                       noise=None if noises is None else noises[0])
    else:
        mono_losses, _ = compute_losses(inputs, mono_outputs, opt, is_multi=False, has_ins=has_ins, noises=noises)
    for key in list(mono_outputs.keys()):
        if isinstance(key, tuple) and key[0] in ("depth", "disp"):
            outputs[("mono_" + key[0],) + tuple(key[1:])] = mono_outputs[key]
    # outputs["consistency_mask"] * compute_matching_mask(outputs), trainer.py:592-593
    if outputs["lowest_cost"].shape[-2:] != outputs[("mono_depth", 0, 0)].shape[-2:]:
        # matching-resolution lowest_cost / confidence straight from the cost-volume head: the
        # nearest up-sampling of repdepth.py:331-336 happens inside the kernel
        outputs["consistency_mask"] = ops.matching_mask(outputs["lowest_cost"], outputs[("mono_depth", 0, 0)],
                                                        confidence=outputs["consistency_mask"])
    else:
        outputs["consistency_mask"] = outputs["consistency_mask"] * ops.matching_mask(
            outputs["lowest_cost"], outputs[("mono_depth", 0, 0)])
    ensemble_reproj = None
    if _o(opt, "distil") and not _o(opt, "no_ens"):
        disp_ensemble = (mono_outputs[("disp", 0)].detach() + outputs[("disp", 0)].detach()) / 2.0
        ensemble_reproj = generate_images_pred_ensemble(
            inputs, outputs[("cam_T_cam", 0, -1)].detach(), outputs[("cam_T_cam", 0, 1)].detach(), disp_ensemble, opt)
    generate_images_pred(inputs, outputs, opt, is_multi=True)
    loss_list = None
    if _o(opt, "distil"):
        losses, w_list, loss_list = compute_main_losses(
            ssim, inputs, outputs, mono_reproj, ensemble_reproj, opt, None, w_list,
            multi_has_ins and _o(opt, "main_temporal"), noise=None if noises is None else noises[1])
    else:
        losses, _ = compute_losses(inputs, outputs, opt, is_multi=True, noises=noises)
    if not freeze_tp:
        for key, val in mono_losses.items():
            losses[key] = losses[key] + val if key in losses else val
        if _o(opt, "loss_blc") and loss_list is not None:
            loss_list[0] = loss_list[0] + mono_losses["loss"]
    if _o(opt, "loss_blc") and loss_blc is not None and loss_list is not None:
        losses["loss"] = loss_blc.compute_loss(loss_list, index_iter)
        losses["w_ori"], losses["w_distil"] = loss_blc.update_weight(index_iter, current_lambda_for_adjust)
    return outputs, losses
