"""Pose-parameter helpers.

`transformation_from_parameters`, `rot_from_axisangle` and `get_translation_matrix`
keep the reference's names, argument meaning and element arithmetic
(manydepth/layers.py:26-100).  They are (B,1,3)-sized host-side tensor algebra
whose gradients flow to the pose network, so they stay in torch; the per-pixel
work that consumes the 4x4 result is in the CUDA kernels.
"""
from __future__ import annotations

import torch


def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """Axis-angle (B,1,3) -> 4x4 rotation (manydepth/layers.py:61-100)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero, one = torch.zeros_like(ca), torch.ones_like(ca)
    rows = [
        [x * xC + ca, xyC - zs, zxC + ys, zero],
        [xyC + zs, y * yC + ca, yzC - xs, zero],
        [zxC - ys, yzC + xs, z * zC + ca, zero],
        [zero, zero, zero, one],
    ]
    return torch.cat([torch.cat(r, dim=2) for r in rows], dim=1)


def get_translation_matrix(translation_vector: torch.Tensor) -> torch.Tensor:
    """Translation (B,1,3) or (B,3) -> 4x4 (manydepth/layers.py:45-58)."""
    t = translation_vector.contiguous().view(-1, 3, 1)
    eye = torch.eye(4, device=t.device, dtype=t.dtype).unsqueeze(0).repeat(t.shape[0], 1, 1)
    top = torch.cat([eye[:, :3, :3], t], dim=2)
    return torch.cat([top, eye[:, 3:, :]], dim=1)


def transformation_from_parameters(axisangle, translation, invert=False):
    """(axisangle, translation) -> 4x4 camera transform (manydepth/layers.py:26-42)."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)
