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
