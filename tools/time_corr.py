"""Device timing of the DualRefine correlation lookup at its training shape (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mal_b200 import _capi, raw


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    B, C, h, w, L, D = 12, 64, 48, 160, 3, 17
    dev = torch.device("cuda:0")
    hnd = _capi.lib()
    g = torch.Generator().manual_seed(0)
    f1, f2 = torch.rand(B, C, h, w, generator=g).to(dev), torch.rand(B, C, h, w, generator=g).to(dev)
    ys, xs = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    # epipolar candidates: a line through each pixel, spacing growing with the level
    dx = torch.linspace(-8, 8, D)[None, None, None, :, None, None] * torch.tensor([1.0, 2.0, 4.0])[None, None, :, None, None, None] * 0.4
    coords = (torch.stack([xs, ys])[None, :, None, None] + dx * torch.tensor([1.0, 0.15])[None, :, None, None, None, None]).repeat(B, 1, 1, 1, 1, 1).to(dev)
    pyr = raw.corr_pyramid(hnd, f2, L)
    go = torch.randn(B, L * D, h, w, device=dev)
    print("pyramid          %8.1f us" % timeit(lambda: raw.corr_pyramid(hnd, f2, L)))
    print("lookup forward   %8.1f us" % timeit(lambda: raw.corr_lookup(hnd, f1, pyr, coords)))
    print("lookup backward  %8.1f us" % timeit(lambda: raw.corr_lookup_backward(hnd, f1, pyr, coords, go)))
    print("backward, coords only %8.1f us" % timeit(lambda: raw.corr_lookup_backward(hnd, f1, pyr, coords, go, want_fmap1=False, want_pyramid=False)))
    print("backward, coords + fmap1   %8.1f us" % timeit(lambda: raw.corr_lookup_backward(hnd, f1, pyr, coords, go, want_pyramid=False)))
    print("backward, coords + pyramid %8.1f us" % timeit(lambda: raw.corr_lookup_backward(hnd, f1, pyr, coords, go, want_fmap1=False)))


if __name__ == "__main__":
    main()
