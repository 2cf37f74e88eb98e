"""dynamicdepth/rigid_warp.py's public entry on the sm_100a kernels.

    forward_warp(img, depth, pose, intrinsics, upscale=None, rotation_mode='euler',
                 padding_mode='zeros') -> (img_w * valid, depth_w * valid, valid)      :534-597

plus the small pose helpers it relies on (mat2euler :175-200, euler2mat :204-241, pose_vec2mat
:269-283), kept in torch because they act on a handful of scalars per sample.  Only
`forward_warp` is called by the DynamicDepth trainer (dynamicdepth/trainer.py:502,517,525); the
per-pixel work (pixel2cam, cam2pix_trans, the torch_sparse.coalesce(op='max') z-buffer, inverse_warp,
cam2pixel) is two kernels (csrc/warp.cu) instead of a Python loop over sparse tensors.
"""
from __future__ import annotations

import torch

from . import ops


def mat2euler(R):
    """Rotation matrices (B,3,3) -> euler angles (B,3)."""
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


def euler2mat(angle):
    """Euler angles (B,3) -> rotation matrices (B,3,3), R = X @ Y @ Z."""
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
    """The first three coefficients of a rotation quaternion (B,3), the scalar part fixed to 1 before
    normalisation -> rotation matrices (B,3,3)  (dynamicdepth/rigid_warp.py:243-265)."""
    q = torch.cat([quat[:, :1].detach() * 0 + 1, quat], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    rows = [w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
            2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
            2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2]
    return torch.stack(rows, dim=1).reshape(quat.size(0), 3, 3)


def pose_vec2mat(vec, rotation_mode="euler"):
    """6-DoF (tx,ty,tz,rx,ry,rz) (B,6) -> (B,3,4)  (dynamicdepth/rigid_warp.py:268-284)."""
    if rotation_mode == "euler":
        rot = euler2mat(vec[:, 3:])
    elif rotation_mode == "quat":
        rot = quat2mat(vec[:, 3:])
    else:
        raise ValueError("rotation_mode must be 'euler' or 'quat', got %r" % (rotation_mode,))
    return torch.cat([rot, vec[:, :3].unsqueeze(-1)], dim=2)


def forward_warp_matrices(pose, intrinsics, upscale):
    """(Ku_inv, K_inv, proj): the per-sample constants of forward_warp (:562-564, :585-589,
    inverse_warp :355-362), derived with the same torch calls as the reference."""
    bs = pose.shape[0]
    intrinsic_u = torch.cat((intrinsics[:, 0:2] * upscale, intrinsics[:, 2:]), dim=1)
    aux = torch.tensor([0, 0, 0, 1]).type_as(pose).unsqueeze(0).unsqueeze(0).repeat(bs, 1, 1)
    pose_mat_inv = torch.inverse(torch.cat([pose, aux], dim=1))
    pose_inv = torch.cat([pose_mat_inv[:, :3, 3], mat2euler(pose_mat_inv[:, :3, :3])], dim=1)
    return intrinsic_u.inverse(), intrinsics.inverse(), intrinsics @ pose_vec2mat(pose_inv)


def forward_warp(img, depth, pose, intrinsics, upscale=None, rotation_mode="euler", padding_mode="zeros",
                 matrices=None):
    """Warp `img` (B,C,H,W) with its own `depth` (B,1,H,W) into the view reached by `pose`
    (B,3,4), z-buffering collisions.  Returns (img_w * valid, depth_w * valid, valid).
    `rotation_mode` and `padding_mode` are accepted and unused, exactly as in the reference (its forward_warp goes
    through pose_vec2mat's default and never samples with a padding mode: rigid_warp.py:534-597)."""
    if upscale is None or int(upscale) != upscale:
        raise ValueError("upscale must be an integer (the reference passes upscale=3)")
    with torch.no_grad():
        Ku_inv, K_inv, proj = matrices if matrices is not None else forward_warp_matrices(pose, intrinsics, int(upscale))
        return ops.forward_warp(img, depth, pose, intrinsics, Ku_inv, K_inv, proj, int(upscale))
