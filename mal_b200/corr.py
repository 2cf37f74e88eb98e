"""DualRefine's feature sampler with the reference's interface (dualrefine/networks/corr.py).

    sampler = CoordSampler(args)
    sampler.register(fmap1, fmap2, num_levels=3)          # corr.py:11-22
    corr = sampler(coords, num_levels=3, num_head=1)      # corr.py:24-50, inside every DEQ iteration
    corr0 = sampler.__corr__(coords0)                     # corr.py:52-76

`coords` is (B, 2, L, D, h, w) as produced by depth2epipolarcoords (dualrefine/networks/utils/utils.py:177-211).
The result is (B, L*heads*D, h, w) float32, differentiable with respect to the coordinates and both feature
maps.  One fused kernel (mal_b200/csrc/corr.cu) replaces the per-level grid_sample / abs / mean chain and its
(B, C, h, w, D) intermediates.
"""
from __future__ import annotations

import torch

from . import ops


class CoordSampler:
    def __init__(self, args=None):
        self.args = args
        self.num_levels = 0
        self.fmap1 = None
        self._pyramid = None
        self._shape = None

    def register(self, fmap1, fmap2, num_levels=4):
        self.num_levels = num_levels
        self.fmap1 = fmap1
        self._shape = tuple(fmap2.shape)
        self._pyramid = ops.corr_pyramid(fmap2, num_levels)

    @property
    def f2_pyramid(self):
        """The levels as (B, C, h>>l, w>>l) views, like the reference's list."""
        from .raw import pyramid_levels
        B, C, h, w = self._shape
        return pyramid_levels(self._pyramid, B, C, h, w, self.num_levels)

    def _lookup(self, coords, num_levels, num_head):
        if self._pyramid is None:
            raise RuntimeError("CoordSampler.register must be called first")
        if coords.shape[2] != num_levels or num_levels > self.num_levels:
            raise ValueError(f"coords carry {coords.shape[2]} levels, num_levels={num_levels}, "
                             f"registered {self.num_levels}")
        return ops.corr_lookup(self.fmap1, self._pyramid, coords, num_head)

    def __call__(self, coords, num_levels=1, num_head=1):
        return self._lookup(coords, num_levels, num_head)

    def __corr__(self, coords, num_levels=1, num_head=1):
        return self._lookup(coords, num_levels, 1)

    def _update_fmap1(self, fmap1):
        self.fmap1 = fmap1


def _normalised_grid(points, height, width):
    """(..., 2) pixel coordinates (x, y) -> grid_sample coordinates for align_corners=False.  The divisors are
    device tensors: torch's CUDA kernels turn division by a host scalar into a multiplication by its
    reciprocal, which is not the IEEE division the reference's CPU path performs."""
    size = torch.tensor([float(width), float(height)], device=points.device, dtype=points.dtype)
    return 2 * (points + 0.5) / size - 1


def sample_tgt(tgt_feat, p2, tgt_w):
    """PoseUpdate.sample_tgt (dualrefine/networks/utils/utils.py:383-404): the target features at the projected
    point and at its four +-1 pixel neighbours (p2: (B,2,1,5,h,w) from depth2gradcoords :213-231) ->

        warped_tgt_feat       (B,C,h,w)    the sample at the point itself
        warped_tgt_gradients  (B,C,h,w,2)  central differences (x, y) of the neighbour samples
        warped_tgt_w          (B,1,h,w)    the confidence map `tgt_w` sampled at the point

    (the reference keeps the third in self.warped_tgt_w).  The sampler is mal_b200's grid_sample kernel with
    ATen's CPU rounding (zeros padding, align_corners=False); gradients reach p2."""
    B, _, _, K, h, w = p2.shape
    pts = p2[:, :, 0].permute(0, 3, 4, 2, 1).reshape(B, h * w, K, 2)          # (B, pixels, 5 probes, xy)
    grid = _normalised_grid(pts, h, w)
    probes = ops.grid_sample(tgt_feat, grid, padding_mode="zeros", align_corners=False).view(B, -1, h, w, K)
    centre, x_plus, x_minus, y_plus, y_minus = probes.unbind(-1)
    gradients = torch.stack([(x_plus - x_minus) / 2, (y_plus - y_minus) / 2], dim=-1)
    weight = ops.grid_sample(tgt_w.to(grid.dtype), grid[:, :, :1], padding_mode="zeros", align_corners=False)
    return centre, gradients, weight.reshape(B, 1, h, w)
