"""Thin tensor -> C-ABI marshalling for libmal_b200 (no autograd here).

Each function validates dtype / layout / shapes, allocates the outputs and workspaces with
torch (the caller owns every buffer the library touches) and enqueues the kernels on the
current CUDA stream.  `handle` is the ctypes library; the product passes `_capi.lib()`.
(The CPU-only tests pass the host-emulated twin built from the same sources, with CPU
tensors; nothing in this package ever does.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _capi

CONV_MANYDEPTH, CONV_DUALREFINE = 0, 1
# kernels launched through this module since the caller last reset it (bench.py's gpu_launches)
LAUNCHES = [0]
PHOTO_WARP, PHOTO_PRED = 0, 1
OCC_NONE, OCC_SET_1, OCC_POOL = 0, 1, 2


def _f32(t, name, shape=None):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if not t.is_contiguous():
        t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(t):
    if t.is_cuda:
        if t.device.index != torch.cuda.current_device():
            # kernels are launched into the calling thread's current CUDA context
            raise RuntimeError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                               "call torch.cuda.set_device() (one process per GPU) first")
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    return C.c_void_p(0)


def _same_device(ts):
    dev = None
    for t in ts:
        if t is None:
            continue
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def photo(handle, *, target, src, syn=None, depth=None, depth_b=None, K=None, inv_K=None, T=None,
          identity_min=None, noise=None, pixel_mask=None, sample_mask=None,
          mode=PHOTO_WARP, convention=CONV_MANYDEPTH, depth_is_disp=True, no_ssim=False,
          with_grad=False, min_depth=0.1, max_depth=100.0, eps=1e-7,
          want_min_reproj=True, want_selection=True, want_weight=False, want_grad_syn=False, finalize=True,
          avg_reprojection=False, split_min=False, zero_img=False, selec_reproj=False, ignore_automask=False,
          want_target_out=False, identity_in_pass=False, warped=None):
    """mal_photo_forward.  Returns a dict of output tensors (see include/mal_b200.h).

    finalize=False leaves `sums` / `grad_P` unreduced until photo_finalize(handle, out) is called (on any
    stream ordered after this call) - a scheduler uses it to keep the tiny reduction kernel off the critical
    path between two heavy kernels - or for good, when only the per-pixel maps are wanted."""
    B, C3, H, W = target.shape
    if C3 != 3:
        raise ValueError("target must be (B,3,H,W)")
    img = (B, 3, H, W)
    target = _f32(target, "target", img)
    src = [_f32(s, f"src[{i}]", img) for i, s in enumerate(src)]
    if len(src) == 1:
        src = [src[0], None]
    syn = [_f32(s, f"syn[{i}]", img) for i, s in enumerate(syn)] if syn is not None else [None, None]
    # WARP mode: the caller's materialised warped sources are staged instead of re-warped (gradients unchanged)
    warped = [_f32(s, f"warped[{i}]", img) for i, s in enumerate(warped)] if warped is not None else [None, None]
    plane = (B, 1, H, W)
    # a low-resolution disparity is up-sampled inside the kernel (trainer.py:1093-1094 fused away)
    dplane, dh, dw = plane, 0, 0
    if depth is not None and tuple(depth.shape[-2:]) != (H, W):
        dh, dw = int(depth.shape[-2]), int(depth.shape[-1])
        dplane = (B, 1, dh, dw)
    depth, depth_b = _f32(depth, "depth", dplane), _f32(depth_b, "depth_b", dplane)
    K, inv_K = _f32(K, "K", (B, 4, 4)), _f32(inv_K, "inv_K", (B, 4, 4))
    T = [_f32(t, f"T[{i}]", (B, 4, 4)) for i, t in enumerate(T)] if T is not None else [None, None]
    identity_min, noise = _f32(identity_min, "identity_min", plane), _f32(noise, "noise", plane)
    pixel_mask = _f32(pixel_mask, "pixel_mask", (B, H, W))
    if sample_mask is not None:
        sample_mask = _f32(sample_mask.reshape(-1), "sample_mask", (B,))
    dev = _same_device([target, *src, *syn, depth, K, inv_K, *T, identity_min, noise, pixel_mask, sample_mask])
    new = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)

    out = {}
    out["sums"] = new((4,))
    out["min_reproj"] = new(plane) if want_min_reproj else None
    out["selection"] = new(plane, torch.uint8) if want_selection else None
    out["weight"] = new(plane) if want_weight else None
    out["min_reproj_b"] = new(plane) if split_min else None   # second min, over the `syn` candidates
    out["target_out"] = new(img) if want_target_out else None
    if with_grad and mode == PHOTO_WARP:
        out["grad_depth"], out["grad_P"] = new(plane), new((B, 2, 12))
    if with_grad and mode == PHOTO_PRED:
        out["grad_pred"] = [new(img), new(img) if src[1] is not None else None]
    if with_grad and want_grad_syn and syn[0] is not None:   # either mode
        out["grad_syn"] = [new(img), new(img)]
    partials = new((handle.mal_photo_partials_floats(B, H, W),))

    a = _capi.PhotoArgs()
    a.batch, a.height, a.width = B, H, W
    a.mode, a.convention = mode, convention
    a.depth_is_disp, a.no_ssim, a.with_grad = int(depth_is_disp), int(no_ssim), int(with_grad)
    a.min_depth, a.max_depth, a.eps = float(min_depth), float(max_depth), float(eps)
    a.target = _ptr(target)
    for i in range(2):
        a.src[i], a.syn[i], a.T[i] = _ptr(src[i]), _ptr(syn[i]), _ptr(T[i])
        a.grad_pred[i] = _ptr(out["grad_pred"][i]) if "grad_pred" in out else None
        a.grad_syn[i] = _ptr(out["grad_syn"][i]) if "grad_syn" in out else None
    a.depth, a.depth_b, a.K, a.inv_K = _ptr(depth), _ptr(depth_b), _ptr(K), _ptr(inv_K)
    a.depth_height, a.depth_width = dh, dw
    a.identity_min, a.noise = _ptr(identity_min), _ptr(noise)
    a.pixel_mask, a.sample_mask = _ptr(pixel_mask), _ptr(sample_mask)
    a.min_reproj, a.selection, a.weight = _ptr(out["min_reproj"]), _ptr(out["selection"]), _ptr(out["weight"])
    a.grad_depth, a.grad_P = _ptr(out.get("grad_depth")), _ptr(out.get("grad_P"))
    a.partials, a.sums = _ptr(partials), _ptr(out["sums"])
    a.skip_finalize = 0 if finalize else 1
    a.avg_reprojection = int(bool(avg_reprojection))
    a.min_reproj_b = _ptr(out["min_reproj_b"])
    a.zero_img, a.selec_reproj, a.ignore_automask = int(bool(zero_img)), int(bool(selec_reproj)), int(bool(ignore_automask))
    a.identity_in_pass = int(bool(identity_in_pass))
    a.warped[0], a.warped[1] = _ptr(warped[0]), _ptr(warped[1])
    a.target_out = _ptr(out["target_out"])
    _capi.check(handle.mal_photo_forward(C.byref(a), _stream(target)), handle)
    LAUNCHES[0] += 2 if finalize else 1   # photo_kernel (+ photo_finalize_kernel)
    out["_keepalive"] = (partials,)
    out["_args"] = a
    return out


def photo_finalize(handle, out):
    """mal_photo_finalize for a photo(..., finalize=False) result, on the current stream."""
    _capi.check(handle.mal_photo_finalize(C.byref(out["_args"]), _stream(out["sums"])), handle)
    LAUNCHES[0] += 1
    return out


def cost_volume(handle, *, current, lookup, poses, K, inv_K, bins, convention=CONV_MANYDEPTH,
                set_missing_to_max=True, apply_confidence=False, num_bins_threshold=0, eps=1e-7,
                want_missing=True, want_head=True, cv_min=False, occ=None, occ_mode=OCC_NONE, pool_radius=1,
                pool_th=0.7, aug_mask=None):
    """mal_cost_volume_forward.  `want_head` adds confidence / argmin / lowest_cost."""
    B, Cn, h, w = current.shape
    F_ = lookup.shape[1]
    nb = bins.shape[0]
    current = _f32(current, "current", (B, Cn, h, w))
    lookup = _f32(lookup, "lookup", (B, F_, Cn, h, w))
    poses = _f32(poses, "poses", (B, F_, 4, 4))
    K, inv_K = _f32(K, "K", (B, 4, 4)), _f32(inv_K, "inv_K", (B, 4, 4))
    bins = _f32(bins, "bins", (nb,))
    occ = _f32(occ, "occ", (B, h, w))
    if aug_mask is not None:
        aug_mask = _f32(aug_mask.reshape(-1).float(), "aug_mask", (B,))
    dev = _same_device([current, lookup, poses, K, inv_K, bins, occ, aug_mask])
    new = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)
    out = {"cost_volume": new((B, nb, h, w))}
    out["missing_mask"] = new((B, nb, h, w)) if want_missing else None
    out["confidence"] = new((B, h, w)) if want_head else None
    out["argmin"] = new((B, h, w), torch.int32) if want_head else None
    out["lowest_cost"] = new((B, h, w)) if want_head else None
    packed = new((handle.mal_cost_volume_workspace_floats(B, Cn, h, w, F_),))
    a = _capi.CostVolumeArgs()
    a.batch, a.channels, a.height, a.width, a.num_lookup, a.num_bins = B, Cn, h, w, F_, nb
    a.convention, a.set_missing_to_max = convention, int(set_missing_to_max)
    a.apply_confidence, a.num_bins_threshold, a.eps = int(apply_confidence), int(num_bins_threshold), float(eps)
    a.current, a.lookup, a.poses = _ptr(current), _ptr(lookup), _ptr(poses)
    a.K, a.inv_K, a.bins = _ptr(K), _ptr(inv_K), _ptr(bins)
    a.cost_volume, a.missing_mask = _ptr(out["cost_volume"]), _ptr(out["missing_mask"])
    a.confidence, a.argmin, a.lowest_cost = _ptr(out["confidence"]), _ptr(out["argmin"]), _ptr(out["lowest_cost"])
    a.packed = _ptr(packed)
    a.cv_min, a.occ_mode, a.pool_radius, a.pool_th = int(bool(cv_min)), int(occ_mode), int(pool_radius), float(pool_th)
    a.occ, a.aug_mask = _ptr(occ), _ptr(aug_mask)
    # the pool fill projects every (lookup frame, bin, pixel) once into a descriptor volume (12 B per sample) that
    # the pool windows read their neighbours from
    dyn = bool(cv_min) or (occ is not None and occ_mode != OCC_NONE)
    quad = (Cn + 15) // 16 <= 4 and os.environ.get("MAL_CV_KERNEL", "")[:1] != "l"
    desc = None
    if occ is not None and occ_mode == OCC_POOL:
        desc = new((handle.mal_cost_volume_desc_floats(B, Cn, F_, nb, h, w),))
    a.desc = _ptr(desc)
    _capi.check(handle.mal_cost_volume_forward(C.byref(a), _stream(current)), handle)
    # the four-lanes-per-pixel sweep (C <= 64, no DynamicDepth extras) reads the current features in place:
    # lookup pack + sweep; the general kernel packs both operands first
    if quad:
        # cv_pack (lookup), cv_sweep_quad (+ cv_pack (current), cv_project, cv_interior, cv_pack_cm, cv_sample, cv_pool)
        LAUNCHES[0] += 2 + (6 if desc is not None else 0)
    else:
        LAUNCHES[0] += 3 + (5 if desc is not None else 0)   # + cv_project, cv_interior, cv_pack_cm, cv_sample, cv_pool
    out["_keepalive"] = (packed, desc)
    return out


def smooth(handle, *, disp, img, normalise=True, with_grad=False, disp_b=None, defer_fix=False):
    """mal_smooth_forward -> {"loss": (1,), "grad_disp": (B,1,h,w)} (+ "loss_b", "grad_disp_b" with `disp_b`: a second
    disparity scored against the same image in the same launch; + "stats" (B,2,2)).  defer_fix leaves the
    gradient planes un-chained through the mean-normalisation for mal_step_combine(smooth_stats=...)."""
    B, _, h, w = disp.shape
    disp, img = _f32(disp, "disp", (B, 1, h, w)), _f32(img, "img", (B, 3, h, w))
    disp_b = _f32(disp_b, "disp_b", (B, 1, h, w))
    dev = _same_device([disp, img, disp_b])
    new = lambda shape: torch.empty(shape, dtype=torch.float32, device=dev)
    dual = disp_b is not None
    out = {"loss": new((1,)), "grad_disp": new((B, 1, h, w)) if with_grad else None,
           "loss_b": new((1,)) if dual else None, "grad_disp_b": new((B, 1, h, w)) if (with_grad and dual) else None,
           "stats": new((B, 2, 2))}
    ws = new((handle.mal_smooth_workspace_floats(B, h, w),))
    a = _capi.SmoothArgs()
    a.batch, a.height, a.width, a.normalise, a.with_grad = B, h, w, int(normalise), int(with_grad)
    a.disp, a.img, a.grad_disp, a.workspace, a.loss = _ptr(disp), _ptr(img), _ptr(out["grad_disp"]), _ptr(ws), _ptr(out["loss"])
    a.disp_b, a.grad_disp_b, a.loss_b = _ptr(disp_b), _ptr(out["grad_disp_b"]), _ptr(out["loss_b"])
    a.defer_fix, a.stats = int(bool(defer_fix) and normalise and with_grad), _ptr(out["stats"])
    _capi.check(handle.mal_smooth_forward(C.byref(a), _stream(disp)), handle)
    fix = normalise and with_grad and not defer_fix
    LAUNCHES[0] += 1 + (int(fix) * (2 if dual else 1))   # smooth_kernel [+ smooth_fix_kernel per term]
    out["_keepalive"] = (ws,)
    return out


def main_terms(handle, *, multi, mono, pixel_mask, sample_mask=None, mono_reproj, multi_reproj,
               ens_reproj=None, inputs_are_disp=False, dual_distil=False, with_grad=False,
               min_depth=0.1, max_depth=100.0, want_index=True, want_target=True):
    """mal_main_terms_forward."""
    B, _, H, W = multi.shape
    plane = (B, 1, H, W)
    multi, mono = _f32(multi, "multi", plane), _f32(mono, "mono", plane)
    pixel_mask = _f32(pixel_mask, "pixel_mask", (B, H, W))
    if sample_mask is not None:
        sample_mask = _f32(sample_mask.reshape(-1), "sample_mask", (B,))
    mono_reproj, multi_reproj = _f32(mono_reproj, "mono_reproj", plane), _f32(multi_reproj, "multi_reproj", plane)
    ens_reproj = _f32(ens_reproj, "ens_reproj", plane)
    dev = _same_device([multi, mono, pixel_mask, sample_mask, mono_reproj, multi_reproj, ens_reproj])
    new = lambda shape, dt=torch.float32: torch.empty(shape, dtype=dt, device=dev)
    out = {"sums": new((2,))}
    out["distil_index"] = new(plane, torch.uint8) if want_index else None
    out["consistency_target"] = new(plane) if want_target else None
    out["grad_cons"] = new(plane) if with_grad else None
    out["grad_distil"] = new(plane) if with_grad else None
    out["grad_distil_mono"] = new(plane) if (with_grad and dual_distil) else None
    partials = new((handle.mal_main_terms_partials_floats(B, H, W),))
    a = _capi.MainTermsArgs()
    a.batch, a.height, a.width = B, H, W
    a.inputs_are_disp, a.dual_distil, a.with_grad = int(inputs_are_disp), int(dual_distil), int(with_grad)
    a.min_depth, a.max_depth = float(min_depth), float(max_depth)
    a.multi, a.mono, a.pixel_mask, a.sample_mask = _ptr(multi), _ptr(mono), _ptr(pixel_mask), _ptr(sample_mask)
    a.mono_reproj, a.ens_reproj, a.multi_reproj = _ptr(mono_reproj), _ptr(ens_reproj), _ptr(multi_reproj)
    a.distil_index, a.consistency_target = _ptr(out["distil_index"]), _ptr(out["consistency_target"])
    a.grad_cons, a.grad_distil, a.grad_distil_mono = _ptr(out["grad_cons"]), _ptr(out["grad_distil"]), _ptr(out["grad_distil_mono"])
    a.partials, a.sums = _ptr(partials), _ptr(out["sums"])
    _capi.check(handle.mal_main_terms_forward(C.byref(a), _stream(multi)), handle)
    LAUNCHES[0] += 1   # main_terms_kernel (its last CTA forms the two means)
    out["_keepalive"] = (partials,)
    return out


def matching_mask(handle, *, lowest_cost, mono, confidence=None, height=None, width=None,
                  mono_is_disp=False, min_depth=0.1, max_depth=100.0):
    """mal_matching_mask -> (B,H,W) float mask (x nearest-upsampled confidence when given)."""
    B, h, w = lowest_cost.shape
    H, W = mono.shape[-2:]
    lowest_cost = _f32(lowest_cost, "lowest_cost", (B, h, w))
    confidence = _f32(confidence, "confidence", (B, h, w))
    mono = _f32(mono, "mono", (B, 1, H, W))
    dev = _same_device([lowest_cost, confidence, mono])
    out = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    a = _capi.MatchingMaskArgs()
    a.batch, a.height, a.width, a.low_height, a.low_width = B, H, W, h, w
    a.mono_is_disp, a.min_depth, a.max_depth = int(mono_is_disp), float(min_depth), float(max_depth)
    a.lowest_cost, a.confidence, a.mono, a.out_mask = _ptr(lowest_cost), _ptr(confidence), _ptr(mono), _ptr(out)
    _capi.check(handle.mal_matching_mask(C.byref(a), _stream(mono)), handle)
    LAUNCHES[0] += 1
    return out


def _vp(t):
    return None if t is None else t.data_ptr()


def backproject(handle, depth, inv_K):
    B, _, H, W = depth.shape
    depth, inv_K = _f32(depth, "depth", (B, 1, H, W)), _f32(inv_K, "inv_K", (B, 4, 4))
    _same_device([depth, inv_K])
    out = torch.empty((B, 4, H * W), dtype=torch.float32, device=depth.device)
    _capi.check(handle.mal_backproject(_vp(depth), _vp(inv_K), B, H, W, _vp(out), _stream(depth)), handle)
    return out


def backproject_backward(handle, grad_out, inv_K, height, width):
    B = grad_out.shape[0]
    grad_out, inv_K = _f32(grad_out, "grad_out", (B, 4, height * width)), _f32(inv_K, "inv_K", (B, 4, 4))
    g = torch.empty((B, 1, height, width), dtype=torch.float32, device=grad_out.device)
    _capi.check(handle.mal_backproject_backward(_vp(grad_out), _vp(inv_K), B, height, width, _vp(g),
                                                _stream(grad_out)), handle)
    return g


def project3d(handle, points, K, T, height, width, convention=CONV_MANYDEPTH, eps=1e-7, want_z=False):
    B = points.shape[0]
    points = _f32(points, "points", (B, 4, height * width))
    K, T = _f32(K, "K", (B, 4, 4)), _f32(T, "T", (B, 4, 4))
    _same_device([points, K, T])
    pix = torch.empty((B, height, width, 2), dtype=torch.float32, device=points.device)
    z = torch.empty((B, 1, height, width), dtype=torch.float32, device=points.device) if want_z else None
    _capi.check(handle.mal_project3d(_vp(points), _vp(K), _vp(T), B, height, width, convention, float(eps),
                                     _vp(pix), _vp(z), _stream(points)), handle)
    return pix, z


def project3d_backward(handle, points, K, T, grad_pix, grad_z, height, width, convention=CONV_MANYDEPTH,
                       eps=1e-7):
    B = points.shape[0]
    points = _f32(points, "points", (B, 4, height * width))
    K, T = _f32(K, "K", (B, 4, 4)), _f32(T, "T", (B, 4, 4))
    grad_pix = _f32(grad_pix, "grad_pix", (B, height, width, 2))
    grad_z = _f32(grad_z, "grad_z", (B, 1, height, width))
    g_points = torch.empty_like(points)
    g_P = torch.empty((B, 12), dtype=torch.float32, device=points.device)
    partials = torch.empty((handle.mal_project3d_partials_floats(B, height, width),), dtype=torch.float32,
                           device=points.device)
    _capi.check(handle.mal_project3d_backward(_vp(points), _vp(K), _vp(T), _vp(grad_pix), _vp(grad_z), B, height,
                                              width, convention, float(eps), _vp(g_points), _vp(g_P),
                                              _vp(partials), _stream(points)), handle)
    return g_points, g_P


def ssim(handle, x, y):
    if x.shape != y.shape or x.dim() != 4:
        raise ValueError("ssim: x and y must be (B,C,H,W) of the same shape")
    B, Cn, H, W = x.shape
    x, y = _f32(x, "x"), _f32(y, "y")
    _same_device([x, y])
    out = torch.empty_like(x)
    _capi.check(handle.mal_ssim(_vp(x), _vp(y), B * Cn, H, W, _vp(out), _stream(x)), handle)
    return out


def ssim_backward(handle, x, y, grad_out, want_grad_y=True):
    B, Cn, H, W = x.shape
    x, y, grad_out = _f32(x, "x"), _f32(y, "y"), _f32(grad_out, "grad_out", x.shape)
    gx = torch.empty_like(x)
    gy = torch.empty_like(x) if want_grad_y else None
    ws = torch.empty((4 * x.numel(),), dtype=torch.float32, device=x.device)
    _capi.check(handle.mal_ssim_backward(_vp(x), _vp(y), _vp(grad_out), B * Cn, H, W, _vp(gx), _vp(gy), _vp(ws),
                                         _stream(x)), handle)
    return gx, gy


def step_combine(handle, *, batch, height, width, weights, sums_teacher, sums_student, smooth_teacher,
                 smooth_student, main_sums, K, gd_teacher, gs_teacher, gP_teacher, gd_student, gs_student,
                 g_cons, g_distil, g_distil_mono=None, smoothness=1e-3, smooth_stats=None):
    """mal_step_combine -> {"scalars": (8,), "grad_disp_teacher", "grad_disp_student", "grad_T": [2 x (B,4,4)]}."""
    dev = gd_teacher.device
    plane = (batch, 1, height, width)
    new = lambda shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = {"scalars": new((8,)), "grad_disp_teacher": new(plane), "grad_disp_student": new(plane),
           "grad_T": [new((batch, 4, 4)), new((batch, 4, 4))]}
    a = _capi.StepCombineArgs()
    a.batch, a.height, a.width, a.smoothness = batch, height, width, float(smoothness)
    a.weights = _ptr(_f32(weights, "weights", (2,)))
    a.sums_teacher, a.sums_student = _ptr(_f32(sums_teacher, "sums", (4,))), _ptr(_f32(sums_student, "sums", (4,)))
    a.smooth_teacher, a.smooth_student = _ptr(_f32(smooth_teacher, "smooth", (1,))), _ptr(_f32(smooth_student, "smooth", (1,)))
    a.main_sums, a.K = _ptr(_f32(main_sums, "main_sums", (2,))), _ptr(_f32(K, "K", (batch, 4, 4)))
    a.gd_teacher, a.gs_teacher = _ptr(_f32(gd_teacher, "gd_teacher", plane)), _ptr(_f32(gs_teacher, "gs_teacher", plane))
    a.gP_teacher = _ptr(_f32(gP_teacher, "gP_teacher", (batch, 2, 12)))
    a.gd_student, a.gs_student = _ptr(_f32(gd_student, "gd_student", plane)), _ptr(_f32(gs_student, "gs_student", plane))
    a.g_cons, a.g_distil = _ptr(_f32(g_cons, "g_cons", plane)), _ptr(_f32(g_distil, "g_distil", plane))
    a.g_distil_mono = _ptr(_f32(g_distil_mono, "g_distil_mono", plane))
    a.smooth_stats = _ptr(_f32(smooth_stats, "smooth_stats", (batch, 2, 2)))
    a.scalars = _ptr(out["scalars"])
    a.grad_disp_teacher, a.grad_disp_student = _ptr(out["grad_disp_teacher"]), _ptr(out["grad_disp_student"])
    a.grad_T[0], a.grad_T[1] = _ptr(out["grad_T"][0]), _ptr(out["grad_T"][1])
    _capi.check(handle.mal_step_combine(C.byref(a), _stream(gd_teacher)), handle)
    LAUNCHES[0] += 1
    return out


def forward_warp(handle, *, img, depth, pose, K, Ku_inv, K_inv, proj, upscale=3):
    """mal_forward_warp -> (img_w * valid, depth_w * valid, valid)."""
    B, Cn, H, W = img.shape
    img, depth = _f32(img, "img", (B, Cn, H, W)), _f32(depth, "depth", (B, 1, H, W))
    pose, proj = _f32(pose, "pose", (B, 3, 4)), _f32(proj, "proj", (B, 3, 4))
    K, Ku_inv, K_inv = _f32(K, "K", (B, 3, 3)), _f32(Ku_inv, "Ku_inv", (B, 3, 3)), _f32(K_inv, "K_inv", (B, 3, 3))
    dev = _same_device([img, depth, pose, K, Ku_inv, K_inv, proj])
    img_w = torch.empty_like(img)
    depth_w, valid = torch.empty_like(depth), torch.empty_like(depth)
    zbuf = torch.empty((B, H, W), dtype=torch.int32, device=dev)
    a = _capi.ForwardWarpArgs()
    a.batch, a.channels, a.height, a.width, a.upscale = B, Cn, H, W, int(upscale)
    a.img, a.depth, a.pose, a.K = _ptr(img), _ptr(depth), _ptr(pose), _ptr(K)
    a.Ku_inv, a.K_inv, a.proj = _ptr(Ku_inv), _ptr(K_inv), _ptr(proj)
    a.img_w, a.depth_w, a.valid, a.zbuf = _ptr(img_w), _ptr(depth_w), _ptr(valid), _ptr(zbuf)
    _capi.check(handle.mal_forward_warp(C.byref(a), _stream(img)), handle)
    LAUNCHES[0] += 2
    return img_w, depth_w, valid


def _mask_u8(m, name, shape):
    if m.dtype not in (torch.bool, torch.uint8):
        raise TypeError(f"{name}: expected a bool mask, got {m.dtype}")
    if tuple(m.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(m.shape)}")
    return m.contiguous().view(torch.uint8)


def dynamic_instance(handle, *, mask_last, mask_next, img_last, img_next, replace=False):
    """mal_dynamic_instance -> (ori_last, ori_next, deltas (4,N) int32)."""
    N, H, W = mask_last.shape
    Cn = img_last.shape[0]
    ml, mn = _mask_u8(mask_last, "mask_last", (N, H, W)), _mask_u8(mask_next, "mask_next", (N, H, W))
    il, inx = _f32(img_last, "img_last", (Cn, H, W)), _f32(img_next, "img_next", (Cn, H, W))
    dev = _same_device([ml, mn, il, inx])
    ol, on = torch.empty_like(il), torch.empty_like(inx)
    ws = torch.empty((12 * N,), dtype=torch.int32, device=dev)
    a = _capi.DynamicInstanceArgs()
    a.num, a.channels, a.height, a.width, a.replace = N, Cn, H, W, int(bool(replace))
    a.mask_last, a.mask_next, a.img_last, a.img_next = _ptr(ml), _ptr(mn), _ptr(il), _ptr(inx)
    a.ori_last, a.ori_next, a.workspace = _ptr(ol), _ptr(on), _ptr(ws)
    _capi.check(handle.mal_dynamic_instance(C.byref(a), _stream(il)), handle)
    LAUNCHES[0] += 3
    return ol, on, ws[8 * N:].view(4, N)


def fill_dynamic_obj(handle, *, mask, delta_x, delta_y, source, img):
    """mal_fill_dynamic_obj -> (C,H,W)."""
    N, H, W = mask.shape
    Cn = img.shape[0]
    m = _mask_u8(mask, "mask", (N, H, W))
    src, im = _f32(source, "source", (Cn, H, W)), _f32(img, "img", (Cn, H, W))
    dx = delta_x.to(torch.int32).contiguous()
    dy = delta_y.to(torch.int32).contiguous()
    _same_device([m, src, im, dx, dy])
    out = torch.empty_like(im)
    _capi.check(handle.mal_fill_dynamic_obj(_vp(m), _vp(dx), _vp(dy), _vp(src), _vp(im), N, Cn, H, W, _vp(out),
                                            _stream(im)), handle)
    LAUNCHES[0] += 1
    return out


def dynamic_instance_backward(handle, *, mask_last, mask_next, deltas, grad_ori_last, grad_ori_next):
    """mal_dynamic_instance_backward -> (grad_img_last, grad_img_next)."""
    N, H, W = mask_last.shape
    Cn = grad_ori_last.shape[0]
    ml, mn = _mask_u8(mask_last, "mask_last", (N, H, W)), _mask_u8(mask_next, "mask_next", (N, H, W))
    gol, gon = _f32(grad_ori_last, "grad_ori_last", (Cn, H, W)), _f32(grad_ori_next, "grad_ori_next", (Cn, H, W))
    d = deltas.to(torch.int32).contiguous()
    dev = _same_device([ml, mn, gol, gon, d])
    gl, gn = torch.empty_like(gol), torch.empty_like(gon)
    flags = torch.empty((H, W), dtype=torch.uint8, device=dev)
    a = _capi.DynamicInstanceArgs()
    a.num, a.channels, a.height, a.width, a.replace = N, Cn, H, W, 0
    a.mask_last, a.mask_next = _ptr(ml), _ptr(mn)
    _capi.check(handle.mal_dynamic_instance_backward(C.byref(a), _vp(d), _vp(gol), _vp(gon), _vp(gl), _vp(gn),
                                                     _vp(flags), _stream(gol)), handle)
    LAUNCHES[0] += 2
    return gl, gn


# ---- temporal hint inside the step (csrc/temporal.cu) ---------------------------------------------------------
def _temporal_args(B, H, W, convention, depth_is_disp, min_depth, max_depth, eps, replace=False):
    a = _capi.TemporalArgs()
    a.batch, a.height, a.width, a.convention = B, H, W, convention
    a.depth_is_disp, a.replace = int(depth_is_disp), int(bool(replace))
    a.min_depth, a.max_depth, a.eps = float(min_depth), float(max_depth), float(eps)
    return a


def _temporal_geom(a, src, depth, K, inv_K, T, B, H, W):
    img, keep = (B, 3, H, W), []
    for i in range(2):
        s_, t_ = _f32(src[i], f"src[{i}]", img), _f32(T[i], f"T[{i}]", (B, 4, 4))
        a.src[i], a.T[i] = _ptr(s_), _ptr(t_)
        keep += [s_, t_]
    depth, K, inv_K = _f32(depth, "depth", (B, 1, H, W)), _f32(K, "K", (B, 4, 4)), _f32(inv_K, "inv_K", (B, 4, 4))
    a.depth, a.K, a.inv_K = _ptr(depth), _ptr(K), _ptr(inv_K)
    return keep + [depth, K, inv_K]


def temporal_warp(handle, *, src, depth, K, inv_K, T, convention=CONV_MANYDEPTH, depth_is_disp=True, min_depth=0.1,
                  max_depth=100.0, eps=1e-7):
    """mal_temporal_warp -> [warped(-1), warped(+1)] (B,3,H,W): outputs[("color", f, 0)] of generate_images_pred."""
    B, _, H, W = src[0].shape
    a = _temporal_args(B, H, W, convention, depth_is_disp, min_depth, max_depth, eps)
    keep = _temporal_geom(a, src, depth, K, inv_K, T, B, H, W)
    out = [torch.empty((B, 3, H, W), dtype=torch.float32, device=src[0].device) for _ in range(2)]
    a.warped[0], a.warped[1] = _ptr(out[0]), _ptr(out[1])
    _capi.check(handle.mal_temporal_warp(C.byref(a), _stream(src[0])), handle)
    LAUNCHES[0] += 1
    del keep
    return out


def temporal_pack_masks(handle, *, masks_last, masks_next, counts):
    """(B,N,H,W) bool / uint8 instance masks (N <= 32) -> two (B,H,W) int32 planes, bit n = instance n."""
    B, N, H, W = masks_last.shape
    ml, mn = _mask_u8(masks_last, "masks_last", (B, N, H, W)), _mask_u8(masks_next, "masks_next", (B, N, H, W))
    cnt = counts.to(torch.int32).contiguous()
    _same_device([ml, mn, cnt])
    pl = torch.empty((B, H, W), dtype=torch.int32, device=ml.device)
    pn = torch.empty_like(pl)
    _capi.check(handle.mal_temporal_pack_masks(_vp(ml), _vp(mn), _vp(cnt), B, N, H, W, _vp(pl), _vp(pn), _stream(ml)), handle)
    LAUNCHES[0] += 1
    return pl, pn


def _packed(t, name, shape):
    if t.dtype != torch.int32 or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous int32 tensor of shape {tuple(shape)}")
    return t


def temporal_synthesis(handle, *, warped, packed_last, packed_next, counts, replace=False):
    """mal_temporal_synthesis -> {"syn": [2 x (B,3,H,W)], "deltas": (B,2,32) int32}."""
    B, _, H, W = warped[0].shape
    a = _temporal_args(B, H, W, CONV_MANYDEPTH, True, 0.1, 100.0, 1e-7, replace)
    w = [_f32(warped[i], f"warped[{i}]", (B, 3, H, W)) for i in range(2)]
    pl, pn = _packed(packed_last, "packed_last", (B, H, W)), _packed(packed_next, "packed_next", (B, H, W))
    cnt = counts.to(torch.int32).contiguous()
    dev = _same_device([*w, pl, pn, cnt])
    syn = [torch.empty_like(w[0]), torch.empty_like(w[1])]
    ext = torch.empty((B, 256), dtype=torch.int32, device=dev)
    deltas = torch.zeros((B, 2, 32), dtype=torch.int32, device=dev)
    for i in range(2):
        a.warped[i], a.syn[i] = _ptr(w[i]), _ptr(syn[i])
    a.packed_last, a.packed_next, a.counts, a.ext, a.deltas = _ptr(pl), _ptr(pn), _ptr(cnt), _ptr(ext), _ptr(deltas)
    _capi.check(handle.mal_temporal_synthesis(C.byref(a), _stream(w[0])), handle)
    LAUNCHES[0] += 2
    return {"syn": syn, "deltas": deltas, "_keepalive": (ext, cnt, w)}


def temporal_backward(handle, *, grad_syn, packed_last, packed_next, counts, deltas, want_grad_warped=True,
                      src=None, depth=None, K=None, inv_K=None, T=None, grad_depth=None, grad_P=None,
                      convention=CONV_MANYDEPTH, depth_is_disp=True, min_depth=0.1, max_depth=100.0, eps=1e-7):
    """mal_temporal_backward: d loss / d syn -> {"grad_warped": [2 x (B,3,H,W)]} and / or, when `grad_depth` and
    `grad_P` are given (the photometric pass's own planes), the chain into depth and pose accumulated onto them."""
    B, _, H, W = grad_syn[0].shape
    a = _temporal_args(B, H, W, convention, depth_is_disp, min_depth, max_depth, eps)
    g = [_f32(grad_syn[i], f"grad_syn[{i}]", (B, 3, H, W)) for i in range(2)]
    pl, pn = _packed(packed_last, "packed_last", (B, H, W)), _packed(packed_next, "packed_next", (B, H, W))
    cnt = counts.to(torch.int32).contiguous()
    dev = _same_device([*g, pl, pn, cnt, deltas])
    keep = [cnt]
    out = {}
    if src is not None:
        keep += _temporal_geom(a, src, depth, K, inv_K, T, B, H, W)
    elif K is not None:
        K, inv_K = _f32(K, "K", (B, 4, 4)), _f32(inv_K, "inv_K", (B, 4, 4))
        a.K, a.inv_K = _ptr(K), _ptr(inv_K)
        for i in range(2):
            t_ = _f32(T[i], f"T[{i}]", (B, 4, 4))
            a.T[i] = _ptr(t_)
            keep.append(t_)
        keep += [K, inv_K]
    if want_grad_warped:
        out["grad_warped"] = [torch.empty_like(g[0]), torch.empty_like(g[1])]
        a.grad_warped[0], a.grad_warped[1] = _ptr(out["grad_warped"][0]), _ptr(out["grad_warped"][1])
    if grad_depth is not None:
        parts = torch.empty((handle.mal_temporal_partials_floats(B, H, W),), dtype=torch.float32, device=dev)
        a.grad_depth, a.grad_P, a.partials = _ptr(_f32(grad_depth, "grad_depth", (B, 1, H, W))), _ptr(_f32(grad_P, "grad_P", (B, 2, 12))), _ptr(parts)
        keep.append(parts)
    for i in range(2):
        a.grad_syn[i] = _ptr(g[i])
    a.packed_last, a.packed_next, a.counts, a.deltas = _ptr(pl), _ptr(pn), _ptr(cnt), _ptr(deltas)
    _capi.check(handle.mal_temporal_backward(C.byref(a), _stream(g[0])), handle)
    LAUNCHES[0] += 2 if grad_depth is not None else 1
    out["_keepalive"] = keep
    return out


def grid_sample(handle, img, grid, align_corners=True, border=True):
    B, Cn, H, W = img.shape
    Ho, Wo = grid.shape[1:3]
    img, grid = _f32(img, "img"), _f32(grid, "grid", (B, Ho, Wo, 2))
    _same_device([img, grid])
    out = torch.empty((B, Cn, Ho, Wo), dtype=torch.float32, device=img.device)
    _capi.check(handle.mal_grid_sample(_vp(img), _vp(grid), B, Cn, H, W, Ho, Wo, int(align_corners), int(border),
                                       _vp(out), _stream(img)), handle)
    return out


def grid_sample_backward(handle, img, grid, grad_out, align_corners=True, border=True):
    B, Cn, H, W = img.shape
    Ho, Wo = grid.shape[1:3]
    img, grid = _f32(img, "img"), _f32(grid, "grid", (B, Ho, Wo, 2))
    grad_out = _f32(grad_out, "grad_out", (B, Cn, Ho, Wo))
    g = torch.empty_like(grid)
    _capi.check(handle.mal_grid_sample_backward(_vp(img), _vp(grid), _vp(grad_out), B, Cn, H, W, Ho, Wo,
                                                int(align_corners), int(border), _vp(g), _stream(img)), handle)
    return g


def upsample_bilinear(handle, x, size):
    """mal_upsample_bilinear: F.interpolate(x, size, mode="bilinear", align_corners=False), CPU-kernel rounding."""
    x = _f32(x, "x", None)
    oh, ow = int(size[0]), int(size[1])
    out = torch.empty(tuple(x.shape[:-2]) + (oh, ow), dtype=torch.float32, device=x.device)
    planes = x.numel() // (x.shape[-2] * x.shape[-1])
    _capi.check(handle.mal_upsample_bilinear(_vp(x), planes, x.shape[-2], x.shape[-1], oh, ow, _vp(out), _stream(x)),
                handle)
    LAUNCHES[0] += 1
    return out


def upsample_bilinear_backward(handle, grad_out, in_size):
    """mal_upsample_bilinear_backward: the adjoint of upsample_bilinear (deterministic gather)."""
    g = _f32(grad_out, "grad_out", None)
    ih, iw = int(in_size[0]), int(in_size[1])
    gin = torch.empty(tuple(g.shape[:-2]) + (ih, iw), dtype=torch.float32, device=g.device)
    planes = g.numel() // (g.shape[-2] * g.shape[-1])
    _capi.check(handle.mal_upsample_bilinear_backward(_vp(g), planes, ih, iw, g.shape[-2], g.shape[-1], _vp(gin),
                                                      _stream(g)), handle)
    LAUNCHES[0] += 1
    return gin


def corr_pyramid(handle, fmap2, num_levels):
    """mal_corr_pyramid: fmap2 and its 2x2 average-pooled levels in one flat buffer (corr.py:11-22)."""
    fmap2 = _f32(fmap2, "fmap2", None)
    B, Cn, h, w = fmap2.shape
    out = torch.empty((handle.mal_corr_pyramid_floats(B, Cn, h, w, num_levels),), dtype=torch.float32,
                      device=fmap2.device)
    _capi.check(handle.mal_corr_pyramid(_vp(fmap2), B, Cn, h, w, num_levels, _vp(out), _stream(fmap2)), handle)
    LAUNCHES[0] += num_levels - 1
    return out


def pyramid_levels(pyramid, B, Cn, h, w, num_levels):
    """The levels of a flat pyramid buffer as (B, C, h>>l, w>>l) tensors.  The buffer stores every level
    channel-quad interleaved ([B][C/4][h][w][4], csrc/corr.cu); this undoes the interleave (a copy)."""
    out, off = [], 0
    for _ in range(num_levels):
        n = B * Cn * h * w
        out.append(pyramid[off:off + n].view(B, Cn // 4, h, w, 4).permute(0, 1, 4, 2, 3).reshape(B, Cn, h, w))
        off += n
        h, w = h // 2, w // 2
    return out


def _corr_args(fmap1, pyramid, coords, num_head):
    B, Cn, h, w = fmap1.shape
    _, two, L, D, h1, w1 = coords.shape
    if two != 2 or (h1, w1) != (h, w) or coords.shape[0] != B:
        raise ValueError(f"coords must be (B,2,L,D,{h},{w}), got {tuple(coords.shape)}")
    a = _capi.CorrArgs()
    a.batch, a.channels, a.height, a.width = B, Cn, h, w
    a.num_levels, a.num_samples, a.num_head = L, D, num_head
    a.fmap1, a.pyramid, a.coords = _ptr(fmap1), _ptr(pyramid), _ptr(coords)
    return a, (B, L * num_head * D, h, w)


def corr_lookup(handle, fmap1, pyramid, coords, num_head=1):
    """mal_corr_lookup -> (B, L*heads*D, h, w)."""
    fmap1, pyramid, coords = _f32(fmap1, "fmap1"), _f32(pyramid, "pyramid"), _f32(coords, "coords")
    a, oshape = _corr_args(fmap1, pyramid, coords, num_head)
    out = torch.empty(oshape, dtype=torch.float32, device=fmap1.device)
    a.out = _ptr(out)
    _capi.check(handle.mal_corr_lookup(C.byref(a), _stream(fmap1)), handle)
    LAUNCHES[0] += 1
    return out


def corr_lookup_backward(handle, fmap1, pyramid, coords, grad_out, num_head=1, want_coords=True, want_fmap1=True,
                         want_pyramid=True):
    """mal_corr_lookup_backward -> (grad_coords, grad_fmap1, grad_pyramid) (None where not wanted)."""
    fmap1, pyramid, coords = _f32(fmap1, "fmap1"), _f32(pyramid, "pyramid"), _f32(coords, "coords")
    a, oshape = _corr_args(fmap1, pyramid, coords, num_head)
    grad_out = _f32(grad_out, "grad_out", oshape)
    gc = torch.empty_like(coords) if want_coords else None
    g1 = torch.zeros_like(fmap1) if want_fmap1 else None
    gp = torch.zeros_like(pyramid) if want_pyramid else None
    a.grad_out, a.grad_coords, a.grad_fmap1, a.grad_pyramid = _ptr(grad_out), _ptr(gc), _ptr(g1), _ptr(gp)
    # d/d fmap1 through a channel-quad interleaved workspace (128-bit reductions) and an un-packing kernel
    ws = torch.empty_like(fmap1) if want_fmap1 else None
    a.workspace = _ptr(ws)
    _capi.check(handle.mal_corr_lookup_backward(C.byref(a), _stream(fmap1)), handle)
    LAUNCHES[0] += 2 if want_fmap1 else 1
    return gc, g1, gp
