"""Thin tensor -> C-ABI marshalling for libmal_b200 (no autograd here).

Each function validates dtype / layout / shapes, allocates the outputs and workspaces with
torch (the caller owns every buffer the library touches) and enqueues the kernels on the
current CUDA stream.  `handle` is the ctypes library; the product passes `_capi.lib()`.
(The CPU-only tests pass the host-emulated twin built from the same sources, with CPU
tensors; nothing in this package ever does.)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _capi

CONV_MANYDEPTH, CONV_DUALREFINE = 0, 1
PHOTO_WARP, PHOTO_PRED = 0, 1


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


def photo(handle, *, target, src, syn=None, depth=None, K=None, inv_K=None, T=None,
          identity_min=None, noise=None, pixel_mask=None, sample_mask=None,
          mode=PHOTO_WARP, convention=CONV_MANYDEPTH, depth_is_disp=True, no_ssim=False,
          with_grad=False, min_depth=0.1, max_depth=100.0, eps=1e-7,
          want_min_reproj=True, want_selection=True, want_weight=False):
    """mal_photo_forward.  Returns a dict of output tensors (see include/mal_b200.h)."""
    B, C3, H, W = target.shape
    if C3 != 3:
        raise ValueError("target must be (B,3,H,W)")
    img = (B, 3, H, W)
    target = _f32(target, "target", img)
    src = [_f32(s, f"src[{i}]", img) for i, s in enumerate(src)]
    syn = [_f32(s, f"syn[{i}]", img) for i, s in enumerate(syn)] if syn is not None else [None, None]
    plane = (B, 1, H, W)
    depth = _f32(depth, "depth", plane)
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
    if with_grad and mode == PHOTO_WARP:
        out["grad_depth"], out["grad_P"] = new(plane), new((B, 2, 12))
    if with_grad and mode == PHOTO_PRED:
        out["grad_pred"] = [new(img), new(img)]
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
    a.depth, a.K, a.inv_K = _ptr(depth), _ptr(K), _ptr(inv_K)
    a.identity_min, a.noise = _ptr(identity_min), _ptr(noise)
    a.pixel_mask, a.sample_mask = _ptr(pixel_mask), _ptr(sample_mask)
    a.min_reproj, a.selection, a.weight = _ptr(out["min_reproj"]), _ptr(out["selection"]), _ptr(out["weight"])
    a.grad_depth, a.grad_P = _ptr(out.get("grad_depth")), _ptr(out.get("grad_P"))
    a.partials, a.sums = _ptr(partials), _ptr(out["sums"])
    _capi.check(handle.mal_photo_forward(C.byref(a), _stream(target)), handle)
    out["_keepalive"] = (partials,)
    return out
