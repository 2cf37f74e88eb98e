"""The reference's geometry / photometric layers (manydepth/layers.py, dualrefine/layers.py,
dynamicdepth/layers.py) with identical names, constructor arguments and return shapes, on the
sm_100a kernels.

    disp_to_depth(disp, min_depth, max_depth)             layers.py:14-23
    transformation_from_parameters / rot_from_axisangle / get_translation_matrix   :26-100
    BackprojectDepth(batch_size, height, width)           :138-168
    Project3D(batch_size, height, width, dc=False, eps=1e-7)    :171-199
          (convention=CONV_DUALREFINE gives dualrefine/layers.py:216-226)
    SSIM()                                                :226-257
    get_smooth_loss(disp, img)                            :210-223
    compute_depth_errors(gt, pred)                        (evaluation metric, plain torch)

The modules hold no buffers: the pixel grid the reference stores as nn.Parameters is generated
inside the kernels.  All of them require CUDA tensors (no CPU fallback).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops, raw
from .pose import get_translation_matrix, rot_from_axisangle, transformation_from_parameters  # noqa: F401

CONV_MANYDEPTH, CONV_DUALREFINE = raw.CONV_MANYDEPTH, raw.CONV_DUALREFINE


def disp_to_depth(disp, min_depth, max_depth):
    """Sigmoid output -> (scaled disparity, depth).  Tiny element-wise producer, kept in torch so
    autograd links it to the decoder; the fused kernels take the disparity directly."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled_disp = min_disp + (max_disp - min_disp) * disp
    depth = 1 / scaled_disp
    return scaled_disp, depth


class BackprojectDepth(nn.Module):
    """Depth image -> homogeneous camera points (B,4,H*W)."""

    def __init__(self, batch_size, height, width):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width

    def forward(self, depth, inv_K):
        depth = depth.reshape(-1, 1, self.height, self.width)
        if inv_K.shape[0] != depth.shape[0]:   # the cost-volume call site broadcasts one inv_K over the bins
            inv_K = inv_K.expand(depth.shape[0], 4, 4)
        return ops.backproject(depth, inv_K.contiguous())


class Project3D(nn.Module):
    """Camera points -> normalised sampling grid (B,H,W,2) for F.grid_sample."""

    def __init__(self, batch_size, height, width, dc=False, eps=1e-7, convention=CONV_MANYDEPTH):
        super().__init__()
        self.batch_size, self.height, self.width = batch_size, height, width
        self.dc, self.eps, self.convention = dc, eps, convention

    def forward(self, points, K, T):
        B = points.shape[0]
        K = K.expand(B, 4, 4).contiguous() if K.shape[0] != B else K
        T = T.expand(B, 4, 4).contiguous() if T.shape[0] != B else T
        return ops.project3d(points, K, T, self.height, self.width, self.convention, self.eps, want_z=self.dc)


class SSIM(nn.Module):
    """clamp((1 - SSIM(x, y)) / 2, 0, 1) with a 3x3 box window and reflection padding."""

    def __init__(self, no_ssim=False):
        super().__init__()
        self.no_ssim = no_ssim   # lets loss_utils honour opt.no_ssim through the `ssim` argument
        self.C1, self.C2 = 0.01 ** 2, 0.03 ** 2

    def forward(self, x, y):
        return ops.ssim(x, y)


def grid_sample(img, grid, padding_mode="border", align_corners=True):
    """Bilinear F.grid_sample with the reference's CPU rounding (differentiable w.r.t. the grid)."""
    return ops.grid_sample(img, grid, padding_mode=padding_mode, align_corners=align_corners)


def get_smooth_loss(disp, img):
    """Edge-aware smoothness of a disparity image."""
    return ops.smooth(disp, img, normalise=False)


def compute_depth_errors(gt, pred):
    """Evaluation metrics between predicted and ground-truth depths (manydepth/layers.py:260-283)."""
    thresh = torch.max((gt / pred), (pred / gt))
    a1 = (thresh < 1.25).float().mean()
    a2 = (thresh < 1.25 ** 2).float().mean()
    a3 = (thresh < 1.25 ** 3).float().mean()
    rmse = torch.sqrt(((gt - pred) ** 2).mean())
    rmse_log = torch.sqrt(((torch.log(gt) - torch.log(pred)) ** 2).mean())
    abs_rel = torch.mean(torch.abs(gt - pred) / gt)
    sq_rel = torch.mean((gt - pred) ** 2 / gt)
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3
