"""torch custom operators (namespace ``mal_b200::``) over the C ABI, with autograd.

Each operator is registered with ``torch.library.custom_op``; the forward enqueues the fused
sm_100a kernel through ``mal_b200.raw`` and, when a gradient will be needed, the same launch
also leaves the un-normalised backward planes behind (the kernels are forward+backward fused:
see DESIGN.md).  ``register_autograd`` then only rescales those planes by the incoming
gradient - there is no second pass over the images.

There is no CPU implementation: tensors must live on a CUDA device and ``libmal_b200.so`` must
have been built (``python -m mal_b200.build``); anything else raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _capi, raw

__all__ = ["photo", "reprojection_loss_map", "smooth", "main_terms", "cost_volume", "matching_mask",
           "backproject", "project3d", "ssim", "forward_warp", "dynamic_instance", "fill_dynamic_obj",
           "grid_sample", "upsample_bilinear", "corr_pyramid", "corr_lookup"]


def _lib(t: Tensor):
    if not t.is_cuda:
        raise RuntimeError(
            "mal_b200 operators run on CUDA tensors only (sm_100a); got a tensor on "
            f"{t.device}.  There is no CPU fallback.")
    return _capi.lib()


_warned = set()


def _warn_once(msg):
    if msg not in _warned:
        _warned.add(msg)
        import warnings
        warnings.warn(msg, stacklevel=3)


def _empty(ref: Tensor) -> Tensor:
    return torch.empty(0, device=ref.device, dtype=torch.float32)


def _opt(t: Tensor) -> Optional[Tensor]:
    return t if t.numel() else None


# --------------------------------------------------------------------------------------------
# photometric loss
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::photo", mutates_args=())
def _photo_op(target: Tensor, src0: Tensor, src1: Optional[Tensor], syn0: Optional[Tensor],
              syn1: Optional[Tensor], depth: Optional[Tensor], K: Optional[Tensor],
              inv_K: Optional[Tensor], T0: Optional[Tensor], T1: Optional[Tensor],
              identity_min: Optional[Tensor], noise: Optional[Tensor], pixel_mask: Optional[Tensor],
              sample_mask: Optional[Tensor], mode: int, convention: int, depth_is_disp: bool,
              no_ssim: bool, min_depth: float, max_depth: float, eps: float,
              need_grad: bool, need_grad_syn: bool, avg_reprojection: bool = False, zero_img: bool = False,
              selec_reproj: bool = False, ignore_automask: bool = False, identity_in_pass: bool = False) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    h = _lib(target)
    src = [src0] if src1 is None else [src0, src1]
    out = raw.photo(h, target=target, src=src, syn=None if syn0 is None else [syn0, syn1],
                    depth=depth, K=K, inv_K=inv_K, T=None if T0 is None else [T0, T1],
                    identity_min=identity_min, noise=noise, pixel_mask=pixel_mask,
                    sample_mask=sample_mask, mode=mode, convention=convention,
                    depth_is_disp=depth_is_disp, no_ssim=no_ssim, with_grad=need_grad,
                    min_depth=min_depth, max_depth=max_depth, eps=eps,
                    want_grad_syn=need_grad and need_grad_syn and syn0 is not None, avg_reprojection=avg_reprojection,
                    zero_img=zero_img, selec_reproj=selec_reproj, ignore_automask=ignore_automask, want_target_out=zero_img,
                    identity_in_pass=identity_in_pass)
    gp = out.get("grad_pred", [None, None])
    gs = out.get("grad_syn", [None, None])
    pick = lambda v: v if v is not None else _empty(target)   # a fresh tensor each: outputs may not alias
    return (out["sums"], out["min_reproj"], out["selection"], pick(out.get("grad_depth")),
            pick(out.get("grad_P")), pick(gp[0]), pick(gp[1]), pick(gs[0]), pick(gs[1]), pick(out.get("target_out")))


@_photo_op.register_fake
def _(target, src0, src1, syn0, syn1, depth, K, inv_K, T0, T1, identity_min, noise, pixel_mask,
      sample_mask, mode, convention, depth_is_disp, no_ssim, min_depth, max_depth, eps, need_grad, need_grad_syn,
      avg_reprojection=False, zero_img=False, selec_reproj=False, ignore_automask=False, identity_in_pass=False):
    B, _, H, W = target.shape
    f = lambda *s: target.new_empty(s)
    warp = mode == raw.PHOTO_WARP
    return (f(4), f(B, 1, H, W), target.new_empty((B, 1, H, W), dtype=torch.uint8),
            f(B, 1, H, W) if need_grad and warp else f(0), f(B, 2, 12) if need_grad and warp else f(0),
            f(B, 3, H, W) if need_grad and not warp else f(0),
            f(B, 3, H, W) if need_grad and not warp and src1 is not None else f(0),
            f(B, 3, H, W) if need_grad and need_grad_syn and syn0 is not None else f(0),
            f(B, 3, H, W) if need_grad and need_grad_syn and syn0 is not None else f(0),
            f(B, 3, H, W) if zero_img else f(0))


def _photo_setup(ctx, inputs, output):
    (target, src0, src1, syn0, syn1, depth, K, inv_K, T0, T1, identity_min, noise, pixel_mask,
     sample_mask, mode, *_rest) = inputs
    sums, _, _, g_depth, g_P, g_p0, g_p1, g_s0, g_s1, _t = output
    ctx.mode = mode
    ctx.need_grad = inputs[21]
    ctx.depth_size = None if depth is None else tuple(depth.shape[-2:])
    ctx.save_for_backward(sums, g_depth, g_P, g_p0, g_p1, K if K is not None else sums, g_s0, g_s1)


def _photo_backward(ctx, g_sums, *_unused):
    sums, g_depth, g_P, g_p0, g_p1, K, g_s0, g_s1 = ctx.saved_tensors
    if not ctx.need_grad:
        raise RuntimeError("mal_b200::photo was run with need_grad=False but a gradient is requested")
    # sums = [S, W, S / (W + 1e-7), 0]; the planes hold d S / d input
    coef = g_sums[0] + g_sums[2] / (sums[1] + 1e-7)
    grads = [None] * 28
    if g_s0.numel():   # temporal-hint candidates, either mode
        grads[3], grads[4] = coef * g_s0, coef * g_s1
    if ctx.mode == raw.PHOTO_WARP:
        grads[5] = coef * g_depth
        if ctx.depth_size != tuple(g_depth.shape[-2:]):   # the kernel up-sampled a low-resolution disparity
            grads[5] = _upsample_bwd_op(grads[5], ctx.depth_size[0], ctx.depth_size[1])
        gP = (coef * g_P).view(-1, 2, 3, 4)
        Kt = K[:, :3, :].transpose(1, 2)      # P = (K @ T)[:3]  =>  dT = K[:3]^T @ dP
        grads[8] = Kt @ gP[:, 0]
        grads[9] = Kt @ gP[:, 1]
    else:
        grads[1] = coef * g_p0
        if g_p1.numel():
            grads[2] = coef * g_p1
    return tuple(grads)


_photo_op.register_autograd(_photo_backward, setup_context=_photo_setup)


def photo(target, src, *, syn=None, depth=None, K=None, inv_K=None, T=None, identity_min=None,
          noise=None, pixel_mask=None, sample_mask=None, mode=raw.PHOTO_WARP,
          convention=raw.CONV_MANYDEPTH, depth_is_disp=True, no_ssim=False, min_depth=0.1,
          max_depth=100.0, eps=1e-7, avg_reprojection=False, zero_img=False, selec_reproj=False,
          ignore_automask=False, identity_in_pass=False):
    """Fused photometric loss.  Returns ``(sums, min_reproj, selection)`` (+ the target image as the pass leaves
    it with ``zero_img``).

    ``zero_img`` / ``selec_reproj`` / ``ignore_automask``: DynamicDepth's compute_losses as one pass per scale
    (dynamicdepth/trainer.py:958-975, :1006-1128): WARP mode with the identity candidates of the automask given as
    ``syn`` and ``noise`` for its tie-break (no ``identity_min``).

    ``avg_reprojection`` (opt.avg_reprojection of the DualRefine / DynamicDepth trainers): the mean instead of
    the min over the two candidates.

    ``sums = [sum(w*reproj), sum(w), sum(w*reproj)/(sum(w)+1e-7), 0]`` is differentiable with
    respect to ``depth`` and ``T`` (WARP mode) or the predictions in ``src`` (PRED mode), and in either mode
    to the temporal-hint candidates ``syn`` when they require grad (autograd then carries that gradient
    into the warped images ``image_synthesis`` copied them from, as in the reference);
    ``min_reproj`` (B,1,H,W) and ``selection`` (uint8: arg-min candidate | automask << 7) are not.
    """
    src = list(src)
    syn_grad = torch.is_grad_enabled() and bool(syn) and any(t.requires_grad for t in syn)
    diff = ([depth, *(T or [])] if mode == raw.PHOTO_WARP else src) + (list(syn) if syn_grad else [])
    need_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in diff)
    sample_mask = None if sample_mask is None else sample_mask.reshape(-1)
    out = _photo_op(target, src[0], src[1] if len(src) > 1 else None,
                    syn[0] if syn else None, syn[1] if syn else None, depth, K, inv_K,
                    T[0] if T else None, T[1] if T else None, identity_min, noise, pixel_mask,
                    sample_mask, mode, convention, depth_is_disp, no_ssim, float(min_depth),
                    float(max_depth), float(eps), need_grad, syn_grad, bool(avg_reprojection), bool(zero_img),
                    bool(selec_reproj), bool(ignore_automask), bool(identity_in_pass))
    if zero_img:
        return out[0], out[1], out[2], out[9]
    return out[0], out[1], out[2]


@torch.library.custom_op("mal_b200::reprojection_loss_map", mutates_args=())
def _reproj_map_op(pred: Tensor, target: Tensor, no_ssim: bool) -> Tensor:
    out = raw.photo(_lib(pred), target=target, src=[pred], mode=raw.PHOTO_PRED, no_ssim=no_ssim,
                    want_selection=False)
    return out["min_reproj"]


@_reproj_map_op.register_fake
def _(pred, target, no_ssim):
    return pred.new_empty((pred.shape[0], 1, pred.shape[2], pred.shape[3]))


@torch.library.custom_op("mal_b200::reprojection_loss_map_backward", mutates_args=())
def _reproj_map_bwd_op(pred: Tensor, target: Tensor, grad_map: Tensor, no_ssim: bool) -> Tensor:
    # S = sum_p w_p * loss_p with w := the incoming gradient, so dS/dpred is the VJP
    out = raw.photo(_lib(pred), target=target, src=[pred], mode=raw.PHOTO_PRED, no_ssim=no_ssim,
                    pixel_mask=grad_map[:, 0], with_grad=True, want_selection=False, want_min_reproj=False)
    return out["grad_pred"][0]


@_reproj_map_bwd_op.register_fake
def _(pred, target, grad_map, no_ssim):
    return pred.new_empty(pred.shape)


def _reproj_map_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.no_ssim = inputs[2]


def _reproj_map_backward(ctx, g):
    pred, target = ctx.saved_tensors
    return _reproj_map_bwd_op(pred, target, g.contiguous(), ctx.no_ssim), None, None


_reproj_map_op.register_autograd(_reproj_map_backward, setup_context=_reproj_map_setup)


def reprojection_loss_map(pred, target, no_ssim=False):
    """compute_reprojection_loss as a (B,1,H,W) map, differentiable w.r.t. `pred` (the target is
    data in every reference call site and receives no gradient)."""
    return _reproj_map_op(pred, target.detach(), bool(no_ssim))


# --------------------------------------------------------------------------------------------
# smoothness
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::smooth", mutates_args=())
def _smooth_op(disp: Tensor, img: Tensor, normalise: bool, need_grad: bool) -> Tuple[Tensor, Tensor]:
    out = raw.smooth(_lib(disp), disp=disp, img=img, normalise=normalise, with_grad=need_grad)
    return out["loss"], out["grad_disp"] if need_grad else _empty(disp)


@_smooth_op.register_fake
def _(disp, img, normalise, need_grad):
    return disp.new_empty(1), disp.new_empty(disp.shape if need_grad else (0,))


def _smooth_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _smooth_backward(ctx, g_loss, _g):
    (g,) = ctx.saved_tensors
    if not g.numel():
        raise RuntimeError("mal_b200::smooth was run with need_grad=False but a gradient is requested")
    return g_loss * g, None, None, None


_smooth_op.register_autograd(_smooth_backward, setup_context=_smooth_setup)


def smooth(disp, img, normalise=False):
    """Edge-aware smoothness (layers.get_smooth_loss), optionally on mean-normalised disparity."""
    need_grad = torch.is_grad_enabled() and disp.requires_grad
    return _smooth_op(disp, img, bool(normalise), need_grad)[0][0]


# --------------------------------------------------------------------------------------------
# MAL student terms
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::main_terms", mutates_args=())
def _main_terms_op(multi: Tensor, mono: Tensor, pixel_mask: Tensor, sample_mask: Optional[Tensor],
                   mono_reproj: Tensor, ens_reproj: Optional[Tensor], multi_reproj: Tensor,
                   inputs_are_disp: bool, dual_distil: bool, min_depth: float, max_depth: float,
                   need_grad: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    out = raw.main_terms(_lib(multi), multi=multi, mono=mono, pixel_mask=pixel_mask,
                         sample_mask=sample_mask, mono_reproj=mono_reproj, multi_reproj=multi_reproj,
                         ens_reproj=ens_reproj, inputs_are_disp=inputs_are_disp,
                         dual_distil=dual_distil, with_grad=need_grad, min_depth=min_depth,
                         max_depth=max_depth)
    pick = lambda v: v if v is not None else _empty(multi)
    return (out["sums"], out["distil_index"], out["consistency_target"], pick(out["grad_cons"]),
            pick(out["grad_distil"]), pick(out["grad_distil_mono"]))


@_main_terms_op.register_fake
def _(multi, mono, pixel_mask, sample_mask, mono_reproj, ens_reproj, multi_reproj, inputs_are_disp,
      dual_distil, min_depth, max_depth, need_grad):
    f = lambda s: multi.new_empty(s)
    g = multi.shape if need_grad else (0,)
    return (f(2), multi.new_empty(multi.shape, dtype=torch.uint8), f(multi.shape), f(g), f(g),
            f(multi.shape if need_grad and dual_distil else (0,)))


def _main_terms_setup(ctx, inputs, output):
    ctx.save_for_backward(output[3], output[4], output[5])


def _main_terms_backward(ctx, g_sums, *_unused):
    g_cons, g_distil, g_mono = ctx.saved_tensors
    if not g_cons.numel():
        raise RuntimeError("mal_b200::main_terms was run with need_grad=False but a gradient is requested")
    grads = [None] * 12
    grads[0] = g_sums[0] * g_cons + g_sums[1] * g_distil
    if g_mono.numel():
        grads[1] = g_sums[1] * g_mono
    return tuple(grads)


_main_terms_op.register_autograd(_main_terms_backward, setup_context=_main_terms_setup)


def main_terms(multi, mono, pixel_mask, sample_mask, mono_reproj, ens_reproj, multi_reproj, *,
               inputs_are_disp=False, dual_distil=False, min_depth=0.1, max_depth=100.0):
    """Consistency + distillation terms.  Returns ``(consistency_loss, distil_loss, distil_index,
    consistency_target)``; the two losses are differentiable w.r.t. ``multi`` (and ``mono`` under
    ``dual_distil``)."""
    need_grad = torch.is_grad_enabled() and (multi.requires_grad or (dual_distil and mono.requires_grad))
    sample_mask = None if sample_mask is None else sample_mask.reshape(-1)
    sums, idx, target, *_ = _main_terms_op(multi, mono, pixel_mask, sample_mask, mono_reproj, ens_reproj,
                                           multi_reproj, bool(inputs_are_disp), bool(dual_distil),
                                           float(min_depth), float(max_depth), need_grad)
    return sums[0], sums[1], idx, target


# --------------------------------------------------------------------------------------------
# cost volume, matching mask (no gradient in the reference: built under torch.no_grad())
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::cost_volume", mutates_args=())
def _cost_volume_op(current: Tensor, lookup: Tensor, poses: Tensor, K: Tensor, inv_K: Tensor,
                    bins: Tensor, convention: int, set_missing_to_max: bool, apply_confidence: bool,
                    num_bins_threshold: int, eps: float, cv_min: bool, occ: Optional[Tensor], occ_mode: int,
                    pool_radius: int, pool_th: float,
                    aug_mask: Optional[Tensor]) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    out = raw.cost_volume(_lib(current), current=current, lookup=lookup, poses=poses, K=K, inv_K=inv_K,
                          bins=bins, convention=convention, set_missing_to_max=set_missing_to_max,
                          apply_confidence=apply_confidence, num_bins_threshold=num_bins_threshold, eps=eps,
                          cv_min=cv_min, occ=occ, occ_mode=occ_mode, pool_radius=pool_radius, pool_th=pool_th,
                          aug_mask=aug_mask)
    return out["cost_volume"], out["missing_mask"], out["confidence"], out["argmin"], out["lowest_cost"]


@_cost_volume_op.register_fake
def _(current, lookup, poses, K, inv_K, bins, convention, set_missing_to_max, apply_confidence,
      num_bins_threshold, eps, cv_min, occ, occ_mode, pool_radius, pool_th, aug_mask):
    B, _, h, w = current.shape
    nb = bins.shape[0]
    f = lambda *s: current.new_empty(s)
    return f(B, nb, h, w), f(B, nb, h, w), f(B, h, w), current.new_empty((B, h, w), dtype=torch.int32), f(B, h, w)


def cost_volume(current, lookup, poses, K, inv_K, bins, *, convention=raw.CONV_MANYDEPTH,
                set_missing_to_max=True, apply_confidence=False, num_bins_threshold=0, eps=1e-7,
                cv_min=False, occ=None, occ_mode=raw.OCC_NONE, pool_radius=1, pool_th=0.7, aug_mask=None):
    """Plane-sweep cost volume + head.  Returns ``(cost_volume, missing_mask, confidence, argmin,
    lowest_cost)``; inputs are detached (the reference builds the volume under no_grad).
    ``cv_min`` / ``occ`` / ``occ_mode`` select DynamicDepth's variant."""
    with torch.no_grad():
        return _cost_volume_op(current.detach(), lookup.detach(), poses.detach(), K, inv_K, bins, convention,
                               bool(set_missing_to_max), bool(apply_confidence), int(num_bins_threshold),
                               float(eps), bool(cv_min), occ, int(occ_mode), int(pool_radius), float(pool_th),
                               None if aug_mask is None else aug_mask.reshape(-1).float())


@torch.library.custom_op("mal_b200::matching_mask", mutates_args=())
def _matching_mask_op(lowest_cost: Tensor, confidence: Optional[Tensor], mono: Tensor, mono_is_disp: bool,
                      min_depth: float, max_depth: float) -> Tensor:
    return raw.matching_mask(_lib(mono), lowest_cost=lowest_cost, confidence=confidence, mono=mono,
                             mono_is_disp=mono_is_disp, min_depth=min_depth, max_depth=max_depth)


@_matching_mask_op.register_fake
def _(lowest_cost, confidence, mono, mono_is_disp, min_depth, max_depth):
    B, _, H, W = mono.shape
    return mono.new_empty((B, H, W))


def matching_mask(lowest_cost, mono, confidence=None, *, mono_is_disp=False, min_depth=0.1, max_depth=100.0):
    """compute_matching_mask (x nearest-upsampled confidence) as a float (B,H,W) mask."""
    with torch.no_grad():
        return _matching_mask_op(lowest_cost.detach(), confidence, mono.detach(), bool(mono_is_disp),
                                 float(min_depth), float(max_depth))


# --------------------------------------------------------------------------------------------
# stand-alone layers: BackprojectDepth, Project3D, SSIM
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::backproject", mutates_args=())
def _backproject_op(depth: Tensor, inv_K: Tensor) -> Tensor:
    return raw.backproject(_lib(depth), depth, inv_K)


@_backproject_op.register_fake
def _(depth, inv_K):
    B, _, H, W = depth.shape
    return depth.new_empty((B, 4, H * W))


@torch.library.custom_op("mal_b200::backproject_backward", mutates_args=())
def _backproject_bwd_op(grad_out: Tensor, inv_K: Tensor, height: int, width: int) -> Tensor:
    return raw.backproject_backward(_lib(grad_out), grad_out, inv_K, height, width)


@_backproject_bwd_op.register_fake
def _(grad_out, inv_K, height, width):
    return grad_out.new_empty((grad_out.shape[0], 1, height, width))


def _backproject_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[1])
    ctx.hw = inputs[0].shape[-2:]


def _backproject_backward(ctx, g):
    (inv_K,) = ctx.saved_tensors
    return _backproject_bwd_op(g.contiguous(), inv_K, ctx.hw[0], ctx.hw[1]), None


_backproject_op.register_autograd(_backproject_backward, setup_context=_backproject_setup)


def backproject(depth, inv_K):
    """BackprojectDepth.forward: (B,1,H,W), (B,4,4) -> (B,4,H*W); differentiable w.r.t. depth."""
    return _backproject_op(depth, inv_K)


@torch.library.custom_op("mal_b200::project3d", mutates_args=())
def _project3d_op(points: Tensor, K: Tensor, T: Tensor, height: int, width: int, convention: int, eps: float,
                  want_z: bool) -> Tuple[Tensor, Tensor]:
    pix, z = raw.project3d(_lib(points), points, K, T, height, width, convention, eps, want_z)
    return pix, z if z is not None else _empty(points)


@_project3d_op.register_fake
def _(points, K, T, height, width, convention, eps, want_z):
    B = points.shape[0]
    return points.new_empty((B, height, width, 2)), points.new_empty((B, 1, height, width) if want_z else (0,))


@torch.library.custom_op("mal_b200::project3d_backward", mutates_args=())
def _project3d_bwd_op(points: Tensor, K: Tensor, T: Tensor, grad_pix: Tensor, grad_z: Optional[Tensor],
                      height: int, width: int, convention: int, eps: float) -> Tuple[Tensor, Tensor]:
    return raw.project3d_backward(_lib(points), points, K, T, grad_pix, grad_z, height, width, convention, eps)


@_project3d_bwd_op.register_fake
def _(points, K, T, grad_pix, grad_z, height, width, convention, eps):
    return points.new_empty(points.shape), points.new_empty((points.shape[0], 12))


def _project3d_setup(ctx, inputs, output):
    points, K, T, height, width, convention, eps, want_z = inputs
    ctx.save_for_backward(points, K, T)
    ctx.args = (height, width, convention, eps, want_z)


def _project3d_backward(ctx, g_pix, g_z):
    points, K, T = ctx.saved_tensors
    height, width, convention, eps, want_z = ctx.args
    g_points, g_P = _project3d_bwd_op(points, K, T, g_pix.contiguous(),
                                      g_z.contiguous() if want_z and g_z is not None else None, height, width,
                                      convention, eps)
    gP = g_P.view(-1, 3, 4)
    # P = (K @ T)[:3]:  dT = K[:3]^T @ dP,  dK[:3] = dP @ T^T
    g_T = K[:, :3, :].transpose(1, 2) @ gP
    g_K = torch.zeros_like(K)
    g_K[:, :3, :] = gP @ T.transpose(1, 2)
    return g_points, g_K, g_T, None, None, None, None, None


_project3d_op.register_autograd(_project3d_backward, setup_context=_project3d_setup)


def project3d(points, K, T, height, width, convention=raw.CONV_MANYDEPTH, eps=1e-7, want_z=False):
    """Project3D.forward -> (B,H,W,2) sampling grid [and z]; differentiable w.r.t. points, K, T."""
    pix, z = _project3d_op(points, K, T, int(height), int(width), int(convention), float(eps), bool(want_z))
    return (pix, z) if want_z else pix


@torch.library.custom_op("mal_b200::ssim", mutates_args=())
def _ssim_op(x: Tensor, y: Tensor) -> Tensor:
    return raw.ssim(_lib(x), x, y)


@_ssim_op.register_fake
def _(x, y):
    return x.new_empty(x.shape)


@torch.library.custom_op("mal_b200::ssim_backward", mutates_args=())
def _ssim_bwd_op(x: Tensor, y: Tensor, grad_out: Tensor) -> Tuple[Tensor, Tensor]:
    return raw.ssim_backward(_lib(x), x, y, grad_out, True)


@_ssim_bwd_op.register_fake
def _(x, y, grad_out):
    return x.new_empty(x.shape), x.new_empty(x.shape)


def _ssim_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _ssim_backward(ctx, g):
    x, y = ctx.saved_tensors
    return _ssim_bwd_op(x, y, g.contiguous())


_ssim_op.register_autograd(_ssim_backward, setup_context=_ssim_setup)


def ssim(x, y):
    """SSIM.forward: clamp((1 - SSIM(x, y)) / 2, 0, 1), differentiable w.r.t. both images."""
    return _ssim_op(x, y)


# --------------------------------------------------------------------------------------------
# DynamicDepth forward warp (no gradient in the reference)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::forward_warp", mutates_args=())
def _forward_warp_op(img: Tensor, depth: Tensor, pose: Tensor, K: Tensor, Ku_inv: Tensor, K_inv: Tensor,
                     proj: Tensor, upscale: int) -> Tuple[Tensor, Tensor, Tensor]:
    return raw.forward_warp(_lib(img), img=img, depth=depth, pose=pose, K=K, Ku_inv=Ku_inv, K_inv=K_inv, proj=proj,
                            upscale=upscale)


@_forward_warp_op.register_fake
def _(img, depth, pose, K, Ku_inv, K_inv, proj, upscale):
    return img.new_empty(img.shape), depth.new_empty(depth.shape), depth.new_empty(depth.shape)


def forward_warp(img, depth, pose, K, Ku_inv, K_inv, proj, upscale=3):
    """Z-buffered forward splat + inverse warp -> (img_w * valid, depth_w * valid, valid)."""
    with torch.no_grad():
        c = lambda t: t.detach().contiguous()
        return _forward_warp_op(c(img), c(depth), c(pose), c(K), c(Ku_inv), c(K_inv), c(proj), int(upscale))


# --------------------------------------------------------------------------------------------
# MAL temporal hint (integer / byte work, no gradient)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::dynamic_instance", mutates_args=())
def _dynamic_instance_op(mask_last: Tensor, mask_next: Tensor, img_last: Tensor, img_next: Tensor,
                         replace: bool) -> Tuple[Tensor, Tensor, Tensor]:
    a, b, d = raw.dynamic_instance(_lib(img_last), mask_last=mask_last, mask_next=mask_next, img_last=img_last,
                                   img_next=img_next, replace=replace)
    return a, b, d.clone()


@_dynamic_instance_op.register_fake
def _(mask_last, mask_next, img_last, img_next, replace):
    return (img_last.new_empty(img_last.shape), img_next.new_empty(img_next.shape),
            img_last.new_empty((4, mask_last.shape[0]), dtype=torch.int32))


@torch.library.custom_op("mal_b200::dynamic_instance_backward", mutates_args=())
def _dynamic_instance_bwd_op(mask_last: Tensor, mask_next: Tensor, deltas: Tensor, g_last: Tensor,
                             g_next: Tensor) -> Tuple[Tensor, Tensor]:
    return raw.dynamic_instance_backward(_lib(g_last), mask_last=mask_last, mask_next=mask_next, deltas=deltas,
                                         grad_ori_last=g_last, grad_ori_next=g_next)


@_dynamic_instance_bwd_op.register_fake
def _(mask_last, mask_next, deltas, g_last, g_next):
    return g_last.new_empty(g_last.shape), g_next.new_empty(g_next.shape)


def _dynamic_instance_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], output[2])


def _dynamic_instance_backward(ctx, g_last, g_next, _g_delta):
    mask_last, mask_next, deltas = ctx.saved_tensors
    gl, gn = _dynamic_instance_bwd_op(mask_last, mask_next, deltas, g_last.contiguous(), g_next.contiguous())
    return None, None, gl, gn, None


_dynamic_instance_op.register_autograd(_dynamic_instance_backward, setup_context=_dynamic_instance_setup)


def dynamic_instance(mask_last, mask_next, img_last, img_next, replace=False):
    """-> (ori_last, ori_next, deltas (4,N) int32 = dx_last, dy_last, dx_next, dy_next).  The two
    images are differentiable w.r.t. img_last / img_next (pure copies and selections)."""
    return _dynamic_instance_op(mask_last.contiguous(), mask_next.contiguous(), img_last.contiguous(),
                                img_next.contiguous(), bool(replace))


@torch.library.custom_op("mal_b200::fill_dynamic_obj", mutates_args=())
def _fill_dynamic_obj_op(mask: Tensor, delta_x: Tensor, delta_y: Tensor, source: Tensor, img: Tensor) -> Tensor:
    return raw.fill_dynamic_obj(_lib(img), mask=mask, delta_x=delta_x, delta_y=delta_y, source=source, img=img)


@_fill_dynamic_obj_op.register_fake
def _(mask, delta_x, delta_y, source, img):
    return img.new_empty(img.shape)


def fill_dynamic_obj(mask, delta_x, delta_y, source, img):
    with torch.no_grad():
        return _fill_dynamic_obj_op(mask.contiguous(), delta_x, delta_y, source.detach().contiguous(),
                                    img.detach().contiguous())


# --------------------------------------------------------------------------------------------
# grid_sample with the reference's (CPU) rounding
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::grid_sample", mutates_args=())
def _grid_sample_op(img: Tensor, grid: Tensor, align_corners: bool, border: bool) -> Tensor:
    return raw.grid_sample(_lib(img), img, grid, align_corners, border)


@_grid_sample_op.register_fake
def _(img, grid, align_corners, border):
    return img.new_empty((img.shape[0], img.shape[1], grid.shape[1], grid.shape[2]))


@torch.library.custom_op("mal_b200::grid_sample_backward", mutates_args=())
def _grid_sample_bwd_op(img: Tensor, grid: Tensor, grad_out: Tensor, align_corners: bool, border: bool) -> Tensor:
    return raw.grid_sample_backward(_lib(img), img, grid, grad_out, align_corners, border)


@_grid_sample_bwd_op.register_fake
def _(img, grid, grad_out, align_corners, border):
    return grid.new_empty(grid.shape)


def _grid_sample_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])
    ctx.flags = (inputs[2], inputs[3])


def _grid_sample_backward(ctx, g):
    img, grid = ctx.saved_tensors
    return None, _grid_sample_bwd_op(img, grid, g.contiguous(), ctx.flags[0], ctx.flags[1]), None, None


_grid_sample_op.register_autograd(_grid_sample_backward, setup_context=_grid_sample_setup)


def grid_sample(img, grid, padding_mode="border", align_corners=True):
    """F.grid_sample(img, grid, mode="bilinear", ...) bit-identical to torch's CPU kernel (which the
    reference's golden outputs come from); differentiable w.r.t. `grid` (the image is data)."""
    if padding_mode not in ("border", "zeros"):
        raise NotImplementedError("padding_mode must be 'border' or 'zeros'")
    return _grid_sample_op(img.detach().contiguous(), grid.contiguous(), bool(align_corners),
                           padding_mode == "border")


# --------------------------------------------------------------------------------------------
# bilinear up-sampling with the reference's (CPU) rounding
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::upsample_bilinear", mutates_args=())
def _upsample_op(x: Tensor, out_height: int, out_width: int) -> Tensor:
    return raw.upsample_bilinear(_lib(x), x, (out_height, out_width))


@_upsample_op.register_fake
def _(x, out_height, out_width):
    return x.new_empty(tuple(x.shape[:-2]) + (out_height, out_width))


@torch.library.custom_op("mal_b200::upsample_bilinear_backward", mutates_args=())
def _upsample_bwd_op(grad_out: Tensor, in_height: int, in_width: int) -> Tensor:
    return raw.upsample_bilinear_backward(_lib(grad_out), grad_out, (in_height, in_width))


@_upsample_bwd_op.register_fake
def _(grad_out, in_height, in_width):
    return grad_out.new_empty(tuple(grad_out.shape[:-2]) + (in_height, in_width))


def _upsample_setup(ctx, inputs, output):
    ctx.in_size = tuple(inputs[0].shape[-2:])


def _upsample_backward(ctx, g):
    return _upsample_bwd_op(g.contiguous(), ctx.in_size[0], ctx.in_size[1]), None, None


_upsample_op.register_autograd(_upsample_backward, setup_context=_upsample_setup)


def upsample_bilinear(x, size):
    """F.interpolate(x, size, mode="bilinear", align_corners=False) bit-identical to torch's CPU kernel
    (torch's CUDA kernel rounds differently, which can flip a min-reprojection selection at scales > 0);
    differentiable (deterministic gather adjoint)."""
    return _upsample_op(x.contiguous(), int(size[0]), int(size[1]))


# --------------------------------------------------------------------------------------------
# DualRefine epipolar correlation lookup (dualrefine/networks/corr.py)
# --------------------------------------------------------------------------------------------
@torch.library.custom_op("mal_b200::corr_pyramid", mutates_args=())
def _corr_pyramid_op(fmap2: Tensor, num_levels: int) -> Tensor:
    return raw.corr_pyramid(_lib(fmap2), fmap2, num_levels)


@_corr_pyramid_op.register_fake
def _(fmap2, num_levels):
    B, C, h, w = fmap2.shape
    n = 0
    for _ in range(num_levels):
        n += B * C * h * w
        h, w = h // 2, w // 2
    return fmap2.new_empty((n,))


def _corr_pyramid_setup(ctx, inputs, output):
    ctx.shape = tuple(inputs[0].shape)
    ctx.levels = inputs[1]


def _corr_pyramid_backward(ctx, g):
    # adjoint of the chain of 2x2 average pools: level l+1's gradient is spread over its 2x2 parents / 4
    B, C, h, w = ctx.shape
    parts = raw.pyramid_levels(g, B, C, h, w, ctx.levels)
    total = None
    for gl in reversed(parts):
        if total is not None:
            up = torch.zeros_like(gl)
            hh, ww = total.shape[-2:]
            up[..., :2 * hh, :2 * ww] = total.repeat_interleave(2, -2).repeat_interleave(2, -1) * 0.25
            total = gl + up
        else:
            total = gl.clone()
    return total, None


_corr_pyramid_op.register_autograd(_corr_pyramid_backward, setup_context=_corr_pyramid_setup)


@torch.library.custom_op("mal_b200::corr_lookup", mutates_args=())
def _corr_lookup_op(fmap1: Tensor, pyramid: Tensor, coords: Tensor, num_head: int) -> Tensor:
    return raw.corr_lookup(_lib(fmap1), fmap1, pyramid, coords, num_head)


@_corr_lookup_op.register_fake
def _(fmap1, pyramid, coords, num_head):
    B, _, L, D, h, w = coords.shape
    return fmap1.new_empty((B, L * num_head * D, h, w))


@torch.library.custom_op("mal_b200::corr_lookup_backward", mutates_args=())
def _corr_lookup_bwd_op(fmap1: Tensor, pyramid: Tensor, coords: Tensor, grad_out: Tensor, num_head: int,
                        want_coords: bool, want_fmap1: bool, want_pyramid: bool) -> Tuple[Tensor, Tensor, Tensor]:
    gc, g1, gp = raw.corr_lookup_backward(_lib(fmap1), fmap1, pyramid, coords, grad_out, num_head, want_coords,
                                          want_fmap1, want_pyramid)
    pick = lambda v: v if v is not None else _empty(fmap1)
    return pick(gc), pick(g1), pick(gp)


@_corr_lookup_bwd_op.register_fake
def _(fmap1, pyramid, coords, grad_out, num_head, want_coords, want_fmap1, want_pyramid):
    e = lambda t, w: t.new_empty(t.shape) if w else t.new_empty((0,))
    return e(coords, want_coords), e(fmap1, want_fmap1), e(pyramid, want_pyramid)


def _corr_lookup_setup(ctx, inputs, output):
    fmap1, pyramid, coords, num_head = inputs
    ctx.save_for_backward(fmap1, pyramid, coords)
    ctx.num_head = num_head


def _corr_lookup_backward(ctx, g):
    fmap1, pyramid, coords = ctx.saved_tensors
    need = ctx.needs_input_grad
    gc, g1, gp = _corr_lookup_bwd_op(fmap1, pyramid, coords, g.contiguous(), ctx.num_head, need[2], need[0], need[1])
    return (g1 if need[0] else None, gp if need[1] else None, gc if need[2] else None, None)


_corr_lookup_op.register_autograd(_corr_lookup_backward, setup_context=_corr_lookup_setup)


def corr_pyramid(fmap2, num_levels):
    """fmap2 and its (num_levels - 1) 2x2 average-pooled levels as one flat differentiable buffer
    (CoordSampler.register, dualrefine/networks/corr.py:11-22)."""
    return _corr_pyramid_op(fmap2.contiguous(), int(num_levels))


def corr_lookup(fmap1, pyramid, coords, num_head=1):
    """CoordSampler.__call__ (dualrefine/networks/corr.py:24-50): mean_c |fmap1 - grid_sample(level l, coords)|
    for every level and epipolar candidate -> (B, L*heads*D, h, w); differentiable w.r.t. all three."""
    return _corr_lookup_op(fmap1.contiguous(), pyramid, coords.contiguous(), int(num_head))
