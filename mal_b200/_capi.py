"""ctypes binding of the C ABI declared in include/mal_b200.h.

`lib()` loads `libmal_b200.so` (built in-tree by mal_b200/build.py for sm_100a) and fails
loudly when it is missing: there is no CPU fallback.  `bind(handle)` attaches the prototypes
to any handle exporting the same ABI.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAL_B200_LIB") or os.path.join(_PKG, "libmal_b200.so")   # override: A/B tuning runs only

c_float_p = C.c_void_p  # raw device addresses are passed as integers


class PhotoArgs(C.Structure):
    """struct mal_photo_args (include/mal_b200.h)."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("mode", C.c_int32), ("convention", C.c_int32), ("depth_is_disp", C.c_int32),
        ("no_ssim", C.c_int32), ("with_grad", C.c_int32),
        ("min_depth", C.c_double), ("max_depth", C.c_double), ("eps", C.c_float),
        ("target", C.c_void_p), ("src", C.c_void_p * 2), ("syn", C.c_void_p * 2),
        ("depth", C.c_void_p), ("K", C.c_void_p), ("inv_K", C.c_void_p), ("T", C.c_void_p * 2),
        ("identity_min", C.c_void_p), ("noise", C.c_void_p), ("pixel_mask", C.c_void_p),
        ("sample_mask", C.c_void_p),
        ("min_reproj", C.c_void_p), ("selection", C.c_void_p), ("weight", C.c_void_p),
        ("grad_depth", C.c_void_p), ("grad_pred", C.c_void_p * 2), ("partials", C.c_void_p),
        ("sums", C.c_void_p), ("grad_P", C.c_void_p), ("depth_b", C.c_void_p),
        ("grad_syn", C.c_void_p * 2),
        ("depth_height", C.c_int32), ("depth_width", C.c_int32), ("min_reproj_b", C.c_void_p),
        ("zero_img", C.c_int32), ("selec_reproj", C.c_int32), ("ignore_automask", C.c_int32), ("identity_in_pass", C.c_int32),
        ("target_out", C.c_void_p),
        ("warped", C.c_void_p * 2),
        ("avg_reprojection", C.c_int32),
        ("skip_finalize", C.c_int32),
    ]


class StepCombineArgs(C.Structure):
    """struct mal_step_combine_args."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("smoothness", C.c_float),
        ("weights", C.c_void_p), ("sums_teacher", C.c_void_p), ("sums_student", C.c_void_p),
        ("smooth_teacher", C.c_void_p), ("smooth_student", C.c_void_p), ("main_sums", C.c_void_p),
        ("K", C.c_void_p), ("gd_teacher", C.c_void_p), ("gs_teacher", C.c_void_p), ("gP_teacher", C.c_void_p),
        ("gd_student", C.c_void_p), ("gs_student", C.c_void_p), ("g_cons", C.c_void_p), ("g_distil", C.c_void_p),
        ("g_distil_mono", C.c_void_p), ("scalars", C.c_void_p), ("grad_disp_teacher", C.c_void_p),
        ("grad_disp_student", C.c_void_p), ("grad_T", C.c_void_p * 2), ("smooth_stats", C.c_void_p),
    ]


class ForwardWarpArgs(C.Structure):
    """struct mal_forward_warp_args."""
    _fields_ = [(n, C.c_int32) for n in ("batch", "channels", "height", "width", "upscale")] + \
               [(n, C.c_void_p) for n in ("img", "depth", "pose", "K", "Ku_inv", "K_inv", "proj", "img_w",
                                          "depth_w", "valid", "zbuf")]


class DynamicInstanceArgs(C.Structure):
    """struct mal_dynamic_instance_args."""
    _fields_ = [(n, C.c_int32) for n in ("num", "channels", "height", "width", "replace")] + \
               [(n, C.c_void_p) for n in ("mask_last", "mask_next", "img_last", "img_next", "ori_last", "ori_next",
                                          "workspace")]


class CostVolumeArgs(C.Structure):
    """struct mal_cost_volume_args."""
    _fields_ = [
        ("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("num_lookup", C.c_int32), ("num_bins", C.c_int32), ("convention", C.c_int32),
        ("set_missing_to_max", C.c_int32), ("apply_confidence", C.c_int32),
        ("num_bins_threshold", C.c_int32), ("eps", C.c_float),
        ("current", C.c_void_p), ("lookup", C.c_void_p), ("poses", C.c_void_p), ("K", C.c_void_p),
        ("inv_K", C.c_void_p), ("bins", C.c_void_p),
        ("cost_volume", C.c_void_p), ("missing_mask", C.c_void_p), ("confidence", C.c_void_p),
        ("argmin", C.c_void_p), ("lowest_cost", C.c_void_p), ("packed", C.c_void_p),
        ("cv_min", C.c_int32), ("occ_mode", C.c_int32), ("pool_radius", C.c_int32), ("pool_th", C.c_float),
        ("occ", C.c_void_p), ("aug_mask", C.c_void_p), ("desc", C.c_void_p),
    ]


class CorrArgs(C.Structure):
    """struct mal_corr_args."""
    _fields_ = [(n, C.c_int32) for n in ("batch", "channels", "height", "width", "num_levels", "num_samples",
                                          "num_head")] + \
               [(n, C.c_void_p) for n in ("fmap1", "pyramid", "coords", "out", "grad_out", "grad_coords",
                                          "grad_fmap1", "grad_pyramid", "workspace")]


class SmoothArgs(C.Structure):
    """struct mal_smooth_args."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("normalise", C.c_int32), ("with_grad", C.c_int32),
        ("disp", C.c_void_p), ("img", C.c_void_p), ("grad_disp", C.c_void_p),
        ("workspace", C.c_void_p), ("loss", C.c_void_p),
        ("disp_b", C.c_void_p), ("grad_disp_b", C.c_void_p), ("loss_b", C.c_void_p),
        ("defer_fix", C.c_int32), ("stats", C.c_void_p),
    ]


class MainTermsArgs(C.Structure):
    """struct mal_main_terms_args."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("inputs_are_disp", C.c_int32), ("dual_distil", C.c_int32), ("with_grad", C.c_int32),
        ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("multi", C.c_void_p), ("mono", C.c_void_p), ("pixel_mask", C.c_void_p),
        ("sample_mask", C.c_void_p), ("mono_reproj", C.c_void_p), ("ens_reproj", C.c_void_p),
        ("multi_reproj", C.c_void_p),
        ("distil_index", C.c_void_p), ("consistency_target", C.c_void_p), ("grad_cons", C.c_void_p),
        ("grad_distil", C.c_void_p), ("grad_distil_mono", C.c_void_p), ("partials", C.c_void_p),
        ("sums", C.c_void_p),
    ]


class MatchingMaskArgs(C.Structure):
    """struct mal_matching_mask_args."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("low_height", C.c_int32), ("low_width", C.c_int32), ("mono_is_disp", C.c_int32),
        ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("lowest_cost", C.c_void_p), ("confidence", C.c_void_p), ("mono", C.c_void_p),
        ("out_mask", C.c_void_p),
    ]


class TemporalArgs(C.Structure):
    """struct mal_temporal_args."""
    _fields_ = [
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("convention", C.c_int32),
        ("depth_is_disp", C.c_int32), ("replace", C.c_int32),
        ("min_depth", C.c_double), ("max_depth", C.c_double), ("eps", C.c_float),
        ("src", C.c_void_p * 2), ("depth", C.c_void_p), ("K", C.c_void_p), ("inv_K", C.c_void_p), ("T", C.c_void_p * 2),
        ("warped", C.c_void_p * 2), ("packed_last", C.c_void_p), ("packed_next", C.c_void_p), ("counts", C.c_void_p),
        ("syn", C.c_void_p * 2), ("ext", C.c_void_p), ("deltas", C.c_void_p), ("grad_syn", C.c_void_p * 2),
        ("grad_warped", C.c_void_p * 2), ("grad_depth", C.c_void_p), ("partials", C.c_void_p), ("grad_P", C.c_void_p),
    ]


EXPORTS = {
    # name: (restype, argtypes)
    "mal_abi_version": (C.c_int, []),
    "mal_last_error": (C.c_char_p, []),
    "mal_check_device": (C.c_int, [C.c_int]),
    "mal_photo_partials_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "mal_photo_finalize": (C.c_int, [C.POINTER(PhotoArgs), C.c_void_p]),
    "mal_photo_forward": (C.c_int, [C.POINTER(PhotoArgs), C.c_void_p]),
    "mal_cost_volume_workspace_floats": (C.c_size_t, [C.c_int] * 5),
    "mal_cost_volume_desc_floats": (C.c_size_t, [C.c_int] * 6),
    "mal_cost_volume_forward": (C.c_int, [C.POINTER(CostVolumeArgs), C.c_void_p]),
    "mal_smooth_workspace_floats": (C.c_size_t, [C.c_int] * 3),
    "mal_smooth_forward": (C.c_int, [C.POINTER(SmoothArgs), C.c_void_p]),
    "mal_main_terms_partials_floats": (C.c_size_t, [C.c_int] * 3),
    "mal_main_terms_forward": (C.c_int, [C.POINTER(MainTermsArgs), C.c_void_p]),
    "mal_matching_mask": (C.c_int, [C.POINTER(MatchingMaskArgs), C.c_void_p]),
    "mal_step_combine": (C.c_int, [C.POINTER(StepCombineArgs), C.c_void_p]),
    "mal_forward_warp": (C.c_int, [C.POINTER(ForwardWarpArgs), C.c_void_p]),
    "mal_dynamic_instance": (C.c_int, [C.POINTER(DynamicInstanceArgs), C.c_void_p]),
    "mal_dynamic_instance_backward": (C.c_int, [C.POINTER(DynamicInstanceArgs)] + [C.c_void_p] * 7),
    "mal_fill_dynamic_obj": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p, C.c_void_p]),
    "mal_temporal_partials_floats": (C.c_size_t, [C.c_int] * 3),
    "mal_temporal_warp": (C.c_int, [C.POINTER(TemporalArgs), C.c_void_p]),
    "mal_temporal_pack_masks": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p] * 3),
    "mal_temporal_synthesis": (C.c_int, [C.POINTER(TemporalArgs), C.c_void_p]),
    "mal_temporal_backward": (C.c_int, [C.POINTER(TemporalArgs), C.c_void_p]),
    "mal_backproject": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mal_backproject_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mal_project3d_partials_floats": (C.c_size_t, [C.c_int] * 3),
    "mal_project3d": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mal_project3d_backward": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_float] + [C.c_void_p] * 4),
    "mal_grid_sample": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p, C.c_void_p]),
    "mal_grid_sample_backward": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_void_p, C.c_void_p]),
    "mal_corr_pyramid_floats": (C.c_size_t, [C.c_int] * 5),
    "mal_corr_pyramid": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "mal_corr_lookup": (C.c_int, [C.POINTER(CorrArgs), C.c_void_p]),
    "mal_corr_lookup_backward": (C.c_int, [C.POINTER(CorrArgs), C.c_void_p]),
    "mal_upsample_bilinear": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "mal_upsample_bilinear_backward": (C.c_int, [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "mal_ssim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mal_ssim_backward": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 4),
}

ABI_VERSION = 5


def bind(handle):
    for name, (res, args) in EXPORTS.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    v = handle.mal_abi_version()
    if v != ABI_VERSION:
        raise RuntimeError(f"libmal_b200 ABI version {v}, binding expects {ABI_VERSION}")
    return handle


_lib = None


def lib():
    """The CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m mal_b200.build` "
                "(mal_b200 has no CPU fallback)")
        _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def check(rc, handle=None):
    if rc != 0:
        h = handle or lib()
        raise RuntimeError(f"libmal_b200 error {rc}: {h.mal_last_error().decode()}")
