"""MAL loss functions with the reference's signatures (manydepth/loss_utils.py), on the fused
sm_100a kernels.

Drop-in surface:
    compute_reprojection_loss(ssim, pred, target)               loss_utils.py:46-55
    compute_loss_masks(reprojection_loss, identity_loss)        loss_utils.py:27-44
    compute_mono_losses(ssim, inputs, outputs, temporal, has_ins)          :57-129
    compute_main_losses(ssim, inputs, outputs, mono_reproj, ensemble_reproj,
                        opt, model, w_list, multi_has_ins)                 :131-281
    LossBalancing(num_loss, num_train_data, bs)                            :283-345

Two ways in:
  * fused   - `outputs` carries a ("warp_spec", scale) entry left by
              mal_b200.trainer_ops.generate_images_pred: disparity, intrinsics and poses go
              straight into ONE kernel that backprojects, projects, samples, scores SSIM+L1,
              takes the per-pixel min, the automask and the masked mean, and leaves the
              gradients for depth and pose behind (no warped image ever reaches HBM);
  * classic - `outputs` carries warped images ("color", f, scale) made by someone else: the same
              kernel scores them (PRED mode) and returns d loss / d pred to autograd.

The tie-break noise is drawn exactly like the reference does (torch.randn on the CPU generator,
then moved to the device) unless a `noise=` tensor is passed, so a seeded run selects the same
pixels.  `ssim` is accepted for signature compatibility; the kernels implement layers.SSIM.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, raw

__all__ = ["compute_reprojection_loss", "compute_loss_masks", "compute_mono_losses",
           "compute_main_losses", "LossBalancing", "WarpSpec"]


class WarpSpec:
    """What the fused kernel needs to re-create outputs[("color", f, scale)] on the fly."""

    __slots__ = ("disp", "K", "inv_K", "T", "convention", "min_depth", "max_depth", "depth_is_disp",
                 "source_scale")

    def __init__(self, disp, K, inv_K, T, convention=raw.CONV_MANYDEPTH, min_depth=0.1, max_depth=100.0,
                 depth_is_disp=True, source_scale=0):
        self.disp, self.K, self.inv_K, self.T = disp, K, inv_K, list(T)
        self.convention, self.min_depth, self.max_depth = convention, min_depth, max_depth
        self.depth_is_disp = depth_is_disp
        self.source_scale = source_scale   # v1_multiscale: images / intrinsics of the disparity's own scale


def _no_ssim(ssim):
    return bool(getattr(ssim, "no_ssim", False))


def compute_reprojection_loss(ssim, pred, target):
    """0.85 * SSIM(pred, target).mean(1) + 0.15 * |target - pred|.mean(1)  -> (B,1,H,W)."""
    return ops.reprojection_loss_map(pred, target, no_ssim=_no_ssim(ssim))


def compute_loss_masks(reprojection_loss, identity_reprojection_loss):
    """argmin([reprojection, identity], 1) == 0 as a float mask (first index wins ties)."""
    if identity_reprojection_loss is None:
        return torch.ones_like(reprojection_loss)
    return (~(identity_reprojection_loss < reprojection_loss)).float()


def _draw_noise(shape, device, noise):
    if noise is None:
        noise = torch.randn(shape)   # CPU generator, like loss_utils.py:105 / :178
    return noise.to(device, non_blocking=True)


def _candidates(inputs, outputs, scale, with_syn):
    syn = [outputs[("syn", f, scale)] for f in (-1, 1)] if with_syn else None
    spec = outputs.get(("warp_spec", scale))
    if spec is not None:
        return dict(src=[inputs[("color", f, 0)] for f in (-1, 1)], syn=syn, depth=spec.disp, K=spec.K,
                    inv_K=spec.inv_K, T=spec.T, mode=raw.PHOTO_WARP, convention=spec.convention,
                    depth_is_disp=spec.depth_is_disp, min_depth=spec.min_depth, max_depth=spec.max_depth)
    return dict(src=[outputs[("color", f, scale)] for f in (-1, 1)], syn=syn, mode=raw.PHOTO_PRED)


def identity_reprojection(ssim, inputs, source_scale=0, avg_reprojection=False):
    """min (mean with avg_reprojection) over f of compute_reprojection_loss(inputs[("color", f, 0)], target),
    loss_utils.py:92-101."""
    target = inputs[("color", 0, source_scale)]
    with torch.no_grad():
        _, ident, _ = ops.photo(target, [inputs[("color", -1, source_scale)], inputs[("color", 1, source_scale)]],
                                mode=raw.PHOTO_PRED, no_ssim=_no_ssim(ssim), avg_reprojection=avg_reprojection)
    return ident


def compute_mono_losses(ssim, inputs, outputs, temporal, has_ins, noise=None):
    """Teacher (single-frame) losses at scale 0.  Returns (losses, mono_reproj (B,1,H,W))."""
    scale = 0
    target = inputs[("color", 0, 0)]
    ident = identity_reprojection(ssim, inputs)
    noise = _draw_noise(ident.shape, target.device, noise)
    sums, min_reproj, sel = ops.photo(target, identity_min=ident, noise=noise, no_ssim=_no_ssim(ssim),
                                      **_candidates(inputs, outputs, scale, temporal and has_ins))
    reprojection_loss = sums[2]
    losses = {"reproj_loss/{}".format(scale): reprojection_loss}
    smooth_loss = ops.smooth(outputs[("disp", scale)], inputs[("color", 0, scale)], normalise=True)
    loss = reprojection_loss + 1e-3 * smooth_loss / (2 ** scale)
    losses["loss/{}".format(scale)] = loss
    losses["loss"] = 0 + loss
    outputs[("mal_selection", scale)] = sel
    return losses, min_reproj


def compute_main_losses(ssim, inputs, outputs, mono_reproj, ensemble_reproj, opt, model, w_list,
                        multi_has_ins, noise=None):
    """Student (multi-frame) losses at scale 0 with the MAL distillation selection.
    Returns (losses, new_w_list, loss_list)."""
    if getattr(opt, "pareto", False):
        raise NotImplementedError("opt.pareto needs manydepth/pareto.py, which the reference does not ship")
    if getattr(opt, "learn_ens", False):
        raise NotImplementedError("opt.learn_ens: the reference model never emits outputs['ens_disp']")
    target = inputs[("color", 0, 0)]
    B = target.shape[0]
    # the reference draws the tie-break noise and computes an automask that it then overwrites
    # with ones (loss_utils.py:178-192): keep the draw so a seeded run stays aligned
    if noise is None:
        torch.randn(B, 1, *target.shape[-2:])
    pixel_mask = outputs["consistency_mask"]
    sample_mask = outputs["augmentation_mask"][:opt.batch_size]
    sums, multi_reproj, sel = ops.photo(target, pixel_mask=pixel_mask, sample_mask=sample_mask,
                                        no_ssim=_no_ssim(ssim), **_candidates(inputs, outputs, 0, multi_has_ins))
    reprojection_loss = sums[2]
    spec = outputs.get(("warp_spec", 0))
    if spec is not None and spec.depth_is_disp and ("mono_disp", 0) in outputs:
        multi, mono, as_disp = spec.disp, outputs[("mono_disp", 0)], True
        lo, hi = spec.min_depth, spec.max_depth
    else:
        multi, mono, as_disp = outputs[("depth", 0, 0)], outputs[("mono_depth", 0, 0)], False
        lo, hi = opt.min_depth, opt.max_depth
    dual = bool(getattr(opt, "dual_distil", False)) and ensemble_reproj is None
    consistency_loss, distil_loss, distil_idx, target_depth = ops.main_terms(
        multi, mono if dual else mono.detach(), pixel_mask, sample_mask, mono_reproj.detach(),
        None if ensemble_reproj is None else ensemble_reproj.detach(), multi_reproj,
        inputs_are_disp=as_disp, dual_distil=dual, min_depth=lo, max_depth=hi)
    outputs["consistency_target/0"] = target_depth
    outputs[("mal_selection", 0)] = sel
    outputs["mal_distil_index"] = distil_idx
    losses = {"consistency_loss/0": consistency_loss, "reproj_loss/0": reprojection_loss}
    loss = reprojection_loss + consistency_loss
    smooth_loss = ops.smooth(outputs[("disp", 0)], inputs[("color", 0, 0)], normalise=True)
    loss = loss + 1e-3 * smooth_loss / (2 ** 0)
    if getattr(opt, "loss_blc", False):
        loss_list = [loss.clone(), distil_loss]
        losses["distil_loss"] = distil_loss
        new_w_list = w_list
    else:
        losses["distil_loss"] = distil_loss
        loss = loss + distil_loss
        new_w_list, loss_list = None, None
    losses["loss/0"] = loss
    losses["loss"] = loss
    return losses, new_w_list, loss_list


class LossBalancing:
    """Host-side loss re-weighting, state and arithmetic as manydepth/loss_utils.py:283-345
    (fp64 numpy).  `compute_loss` keeps the reference's quirk of adding the weighted sum once per
    batch element (the result is batch_size x the weighted loss)."""

    def __init__(self, num_loss, num_train_data, bs):
        self.num_loss = num_loss
        self.weight_initialization = True
        self.weight_initialization_done = False
        self.last_rebalancing_iter = 0
        self.previous_total_loss = 0
        self.previous_loss = 0
        self.w_list = np.array([1. / num_loss, 1. / num_loss])
        self.loss_initialize_scale = np.array([1. / num_loss, 1. / num_loss])
        self.train_scores = np.zeros((num_train_data, num_loss))
        self.train_metrics = np.zeros((num_train_data, 7))
        self.num_data = num_train_data
        self.bs = bs
        # Running column sums of train_scores[0:_rows]: numpy's mean over axis 0 adds the rows one after the
        # other, so a sum kept row by row is the same fp64 number (checked in tests/test_api.py) and
        # update_weight stays O(1) instead of re-reading every score recorded so far.  Any write that is not
        # "the next row" (a new epoch restarting at index 0, a skipped step) drops the shortcut until the next
        # update_weight rebuilds it from the array.
        self._rows = 0
        self._sum = np.zeros(num_loss)

    def _recorded(self, index_record, scores):
        if index_record == self._rows:
            self._sum = self._sum + np.asarray(scores, dtype=np.float64)
            self._rows += 1
        else:
            self._rows = -1   # out of order: the next update_weight recomputes from train_scores

    def compute_loss(self, loss_list, index_iter):
        loss = 0
        scores = None
        for index_batch in range(self.bs):
            index_record = self.bs * index_iter + index_batch
            if index_record < self.num_data:
                for k in range(self.num_loss):
                    loss = loss + self.w_list[k] * loss_list[k]
                if scores is None:   # one device->host read per step instead of bs * num_loss
                    scores = [float(v.detach()) if torch.is_tensor(v) else float(v) for v in loss_list[:self.num_loss]]
                self.train_scores[index_record, :] = scores
                self._recorded(index_record, scores)
        return loss

    def record_scores(self, index_iter, scores):
        """The bookkeeping half of compute_loss for callers that formed the weighted sum on the
        device (mal_b200.step.MalStep under a CUDA graph): same train_scores rows, no tensors."""
        for index_batch in range(self.bs):
            index_record = self.bs * index_iter + index_batch
            if index_record < self.num_data:
                self.train_scores[index_record, :] = scores
                self._recorded(index_record, scores)

    def _mean_scores(self, i):
        lo, hi = self.last_rebalancing_iter * self.bs, min((i + 1) * self.bs, self.num_data)
        if lo == 0 and hi > 0 and self._rows == hi:
            return self._sum / hi
        window = self.train_scores[lo:(i + 1) * self.bs, :]
        if lo == 0 and hi > 0:   # rebuild the running sum exactly as numpy's mean accumulates it
            self._sum, self._rows = np.add.reduce(window, axis=0), hi
        return window.mean(axis=0)

    def update_weight(self, i, current_lambda_for_adjust):
        # A loss whose recorded mean is 0 (e.g. distil_loss == 0 in the first step) makes the reference divide by
        # zero: its weight becomes inf, previous_total_loss NaN, and two updates later (inf / inf) both weights
        # are NaN for the rest of the run.  Same arithmetic here (pinned), without numpy's
        # RuntimeWarnings; mal_b200.step.MalStep warns once when such weights reach the device.
        with np.errstate(divide="ignore", invalid="ignore"):
            return self._update_weight(i, current_lambda_for_adjust)

    def _update_weight(self, i, current_lambda_for_adjust):
        mean = self._mean_scores(i)
        total_loss = np.sum(mean * self.w_list)
        if self.weight_initialization and not self.weight_initialization_done:
            for k in range(self.num_loss):
                self.w_list[k] = (total_loss * self.loss_initialize_scale[k]) / mean[k]
            self.weight_initialization_done = True
        else:
            previous = np.array(self.w_list)
            if self.previous_total_loss > 0:
                for k in range(self.num_loss):
                    adjust = 1 + current_lambda_for_adjust * (
                        (total_loss / self.previous_total_loss) * (self.previous_loss[k] / mean[k]) - 1)
                    adjust = min(max(adjust, 1.0 / 2.0), 2.0 / 1.0)
                    self.w_list[k] = previous[k] * adjust
        self.previous_total_loss = np.sum(mean * self.w_list)
        self.previous_loss = mean
        return self.w_list[0], self.w_list[1]
