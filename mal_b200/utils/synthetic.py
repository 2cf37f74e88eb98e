"""Seeded synthetic KITTI / Cityscapes-shaped inputs for the MAL hot path.

Shapes, intrinsics and dict keys follow the reference data pipeline
(manydepth/datasets/mono_dataset.py:181-190 for the K scaling,
manydepth/datasets/kitti_dataset.py:26-29 for the normalised KITTI K) and the
recipe in SURVEY.md section 8(d).  Everything is generated on the CPU with an
explicit torch.Generator so a parity test can hand identical bits to the CUDA
path and to the oracle; `to(device)` moves a whole bundle.

This module is plain tensor plumbing: it does not touch the CUDA library and
does not import the oracle.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

KITTI_K = ((0.58, 0.0, 0.5, 0.0), (0.0, 1.92, 0.5, 0.0), (0.0, 0.0, 1.0, 0.0), (0.0, 0.0, 0.0, 1.0))
# Cityscapes: fx = fy = 2262.52 on a 2048x1024 sensor, bottom quarter cropped (reference
# manydepth/datasets/cityscapes_preprocessed_dataset.py uses 1024x384 crops).
CITYSCAPES_K = ((1.104, 0.0, 0.535, 0.0), (0.0, 2.212, 0.501, 0.0), (0.0, 0.0, 1.0, 0.0),
                (0.0, 0.0, 0.0, 1.0))


def _smooth_field(gen, shape, k=9):
    """Low-pass random field in [0,1]: k x k box blur of uniform noise."""
    b, c, h, w = shape
    noise = torch.rand(b, c, h + k - 1, w + k - 1, generator=gen)
    return F.avg_pool2d(noise, k, 1)


def _rodrigues(axisangle, translation, invert):
    from ..pose import transformation_from_parameters
    return transformation_from_parameters(axisangle, translation, invert=invert)


def intrinsics(batch, height, width, scale=0, normalised=KITTI_K):
    """K and pinv(K) at a pyramid level, (B,4,4) each (mono_dataset.py:181-190)."""
    K = torch.tensor(normalised, dtype=torch.float32).clone()
    K[0, :] *= width // (2 ** scale)
    K[1, :] *= height // (2 ** scale)
    inv_K = torch.linalg.pinv(K)
    return K.unsqueeze(0).repeat(batch, 1, 1), inv_K.unsqueeze(0).repeat(batch, 1, 1)


def make_photometric_inputs(batch=2, height=192, width=640, num_scales=1, seed=1234,
                            white_noise=False, with_syn=True, normalised_K=KITTI_K,
                            contrast=4.0, translation_scale=1.0):
    """One batch of hot-path inputs.

    Returns (inputs, tensors): `inputs` uses the reference's dict keys
    (("color",f,s), ("K",s), ("inv_K",s)); `tensors` holds the network-side
    quantities a trainer would produce (disp pyramids, poses, masks, tie-break noise).
    """
    gen = torch.Generator().manual_seed(seed)
    inputs, t = {}, {}
    for f in (0, -1, 1):
        if white_noise:
            img = torch.rand(batch, 3, height, width, generator=gen)
        else:
            # stretch the blurred field around 0.5 so SSIM/L1 are not degenerate
            img = ((_smooth_field(gen, (batch, 3, height, width)) - 0.5) * contrast + 0.5).clamp(0, 1)
        inputs[("color", f, 0)] = img
        inputs[("color_aug", f, 0)] = img
    for s in range(1, max(num_scales, 4)):
        inputs[("color", 0, s)] = F.interpolate(inputs[("color", 0, 0)], scale_factor=1 / 2 ** s,
                                                mode="area")
    for s in range(4):
        inputs[("K", s)], inputs[("inv_K", s)] = intrinsics(batch, height, width, s, normalised_K)

    for name in ("mono", "multi"):
        for s in range(num_scales):
            h, w = height // 2 ** s, width // 2 ** s
            field = _smooth_field(gen, (batch, 1, h, w), 5)
            t[(name + "_disp", s)] = torch.sigmoid((field - 0.5) * 12 + torch.randn(
                batch, 1, h, w, generator=gen) * 0.05)
    for f in (-1, 1):
        aa = torch.randn(batch, 1, 3, generator=gen) * 0.01
        tr = torch.randn(batch, 1, 3, generator=gen) * 0.05
        tr[..., 2] += 0.1
        tr = tr * translation_scale
        t[("axisangle", f)], t[("translation", f)] = aa, tr
        t[("cam_T_cam", 0, f)] = _rodrigues(aa, tr, invert=(f < 0))
        if with_syn:
            if white_noise:
                t[("syn", f, 0)] = torch.rand(batch, 3, height, width, generator=gen)
            else:
                t[("syn", f, 0)] = (inputs[("color", 0, 0)] + 0.08 * (
                    _smooth_field(gen, (batch, 3, height, width), 5) - 0.5)).clamp(0, 1)
    t["noise"] = [torch.randn(batch, 1, height, width, generator=gen) for _ in range(max(num_scales, 2))]
    cm = (torch.rand(batch, 1, height // 4, width // 4, generator=gen) < 0.8).float()
    t["consistency_mask"] = F.interpolate(cm, [height, width], mode="nearest")[:, 0]
    t["augmentation_mask"] = (torch.rand(batch, 1, 1, 1, generator=gen) < 0.5).float()
    return inputs, t


def make_cost_volume_inputs(batch=2, height=192, width=640, channels=64, num_lookup=1,
                            num_bins=96, seed=4321, zero_pose_sample=None,
                            normalised_K=KITTI_K, min_bin=0.1, max_bin=20.0, translation_scale=1.0):
    """Matching-branch inputs at 1/4 resolution (networks/resnet_encoder.py:264-305)."""
    gen = torch.Generator().manual_seed(seed)
    h, w = height // 4, width // 4
    cur = _smooth_field(gen, (batch, channels, h, w), 3) * 2
    look = torch.stack([(cur + 0.3 * torch.rand(batch, channels, h, w, generator=gen)).clamp_min(0)
                        for _ in range(num_lookup)], 1)
    poses = []
    for _ in range(num_lookup):
        aa = torch.randn(batch, 1, 3, generator=gen) * 0.01
        tr = torch.randn(batch, 1, 3, generator=gen) * 0.05
        tr[..., 2] += 0.1
        poses.append(_rodrigues(aa, tr * translation_scale, invert=True))
    poses = torch.stack(poses, 1)
    if zero_pose_sample is not None:
        poses[zero_pose_sample] = 0
    K, inv_K = intrinsics(batch, height, width, 2, normalised_K)
    bins = torch.linspace(min_bin, max_bin, num_bins)
    return {"current_feats": cur, "lookup_feats": look, "relative_poses": poses, "K": K,
            "inv_K": inv_K, "bins": bins}


def make_instance_masks(num=6, height=192, width=640, seed=99, max_shift=12, empty=None, max_ry=None, max_rx=None):
    """Synthetic Mask2Former-shaped matched instance masks: `num` random
    rectangles / ellipses in the "last" frame and the same shapes shifted by up to `max_shift`
    pixels in the "next" frame; instance `empty` (if given) is absent from the next frame.
    Half-extents are drawn from [3, max_ry) x [3, max_rx) (default height // 4, width // 6)."""
    gen = torch.Generator().manual_seed(seed)
    ys = torch.arange(height).view(-1, 1).float()
    xs = torch.arange(width).view(1, -1).float()
    last = torch.zeros(num, height, width, dtype=torch.bool)
    nxt = torch.zeros(num, height, width, dtype=torch.bool)
    for n in range(num):
        cy = float(torch.randint(0, height, (1,), generator=gen))
        cx = float(torch.randint(0, width, (1,), generator=gen))
        ry = float(torch.randint(3, max(4, max_ry or height // 4), (1,), generator=gen))
        rx = float(torch.randint(3, max(4, max_rx or width // 6), (1,), generator=gen))
        dy = float(torch.randint(-max_shift, max_shift + 1, (1,), generator=gen))
        dx = float(torch.randint(-max_shift, max_shift + 1, (1,), generator=gen))
        if n % 2 == 0:
            shape = lambda oy, ox: ((ys - cy - oy).abs() <= ry) & ((xs - cx - ox).abs() <= rx)
        else:
            shape = lambda oy, ox: ((ys - cy - oy) / ry) ** 2 + ((xs - cx - ox) / rx) ** 2 <= 1.0
        last[n] = shape(0.0, 0.0)
        nxt[n] = shape(dy, dx)
    if empty is not None:
        nxt[empty] = False
    return last, nxt


def warp_by_depth(feat, depth, K, inv_K, T):
    """Sample `feat` (B,C,h,w) where the pixels of a view with `depth` (B,1,h,w) land after the
    rigid motion T: plain-torch pinhole geometry, used only to make synthetic feature pairs
    that are photo-consistent under a known depth."""
    B, _, h, w = depth.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(h * w)], 0).unsqueeze(0)
    cam = depth.view(B, 1, -1) * (inv_K[:, :3, :3] @ pix)
    cam = torch.cat([cam, torch.ones(B, 1, h * w)], 1)
    p = (K @ T)[:, :3, :] @ cam
    xy = p[:, :2] / (p[:, 2:3] + 1e-7)
    gx = (xy[:, 0] / (w - 1) - 0.5) * 2
    gy = (xy[:, 1] / (h - 1) - 0.5) * 2
    grid = torch.stack([gx, gy], -1).view(B, h, w, 2)
    return F.grid_sample(feat, grid, padding_mode="border", align_corners=True)


def to_device(obj, device):
    """Recursively move a nested dict/list of tensors."""
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_device(v, device) for v in obj)
    return obj
