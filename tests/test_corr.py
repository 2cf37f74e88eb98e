"""DualRefine epipolar correlation lookup (dualrefine/networks/corr.py) against the
reference's golden outputs, the oracle on seeded inputs (bit-exact forward), and the oracle's autograd
(gradients within 1e-4 of the gradient's max magnitude)."""
import numpy as np
import pytest
import torch

from mal_b200 import raw
from mal_b200.corr import CoordSampler, sample_tgt
from oracle import mal_oracle as O
from tests.backends import BACKENDS, handle_and_device
from tests.helpers import load_npz

GRAD_RTOL = 1e-4


def _case(B, C, h, w, L, D, seed=5, spread=2.5):
    g = torch.Generator().manual_seed(seed)
    fmap1, fmap2 = torch.rand(B, C, h, w, generator=g), torch.rand(B, C, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    coords = torch.stack([xs, ys])[None, :, None, None] + spread * torch.randn(B, 2, L, D, h, w, generator=g)
    return fmap1, fmap2, coords


def test_oracle_reproduces_the_reference_golden():
    g = load_npz("corr_lookup.npz")
    for tag in ("a", "b"):
        f1, f2, c = (torch.from_numpy(g[f"{tag}_in_{k}"]) for k in ("fmap1", "fmap2", "coords"))
        pyr = O.corr_pyramid(f2, c.shape[2])
        assert np.array_equal(torch.cat([p.reshape(-1) for p in pyr]).numpy(), g[f"{tag}_ref_pyramid"])
        assert np.array_equal(O.corr_lookup(f1, pyr, c, int(g[f"{tag}_heads"])).numpy(), g[f"{tag}_ref_corr"])


@pytest.mark.parametrize("backend", BACKENDS)
def test_kernel_against_reference_golden(backend):
    h, dev = handle_and_device(backend)
    g = load_npz("corr_lookup.npz")
    for tag in ("a", "b"):
        f1, f2, c = (torch.from_numpy(g[f"{tag}_in_{k}"]).to(dev) for k in ("fmap1", "fmap2", "coords"))
        pyr = raw.corr_pyramid(h, f2, c.shape[2])
        levels = raw.pyramid_levels(pyr, *f2.shape, c.shape[2])
        assert np.array_equal(torch.cat([p.reshape(-1) for p in levels]).cpu().numpy(), g[f"{tag}_ref_pyramid"])
        out = raw.corr_lookup(h, f1, pyr, c, int(g[f"{tag}_heads"]))
        assert np.array_equal(out.cpu().numpy(), g[f"{tag}_ref_corr"])


# (B, C, h, w, L, D, heads): with and without ATen's reduction tail (h*w*D % 32), odd map sizes (pyramid floor),
# DualRefine's own shape class (64 channels, 3 levels, 17 candidates)
CASES = [(1, 64, 8, 16, 3, 17, 1), (2, 64, 9, 13, 2, 5, 1), (1, 48, 10, 12, 3, 3, 2), (2, 16, 6, 10, 1, 7, 4)]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("case", CASES)
def test_forward_is_bit_exact_against_the_oracle(backend, case):
    h, dev = handle_and_device(backend)
    B, C, hh, ww, L, D, heads = case
    f1, f2, c = _case(B, C, hh, ww, L, D, seed=sum(case))
    want = O.corr_lookup(f1, O.corr_pyramid(f2, L), c, heads)
    pyr = raw.corr_pyramid(h, f2.to(dev), L)
    got = raw.corr_lookup(h, f1.to(dev), pyr, c.to(dev), heads)
    assert torch.equal(got.cpu(), want)


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


@pytest.mark.parametrize("heads", [1, 2])
def test_gradients_match_the_oracle_autograd(op_device, heads):
    dev = op_device
    B, C, hh, ww, L, D = 2, 32, 10, 14, 3, 5
    f1, f2, c = _case(B, C, hh, ww, L, D, seed=17 + heads)
    go = torch.randn(B, L * heads * D, hh, ww, generator=torch.Generator().manual_seed(2))
    leaves = [t.clone().requires_grad_(True) for t in (f1, f2, c)]
    want_out = O.corr_lookup(leaves[0], O.corr_pyramid(leaves[1], L), leaves[2], heads)
    want = torch.autograd.grad(want_out, leaves, go)

    d_leaves = [t.clone().to(dev).requires_grad_(True) for t in (f1, f2, c)]
    sampler = CoordSampler()
    sampler.register(d_leaves[0], d_leaves[1], num_levels=L)
    out = sampler(d_leaves[2], num_levels=L, num_head=heads)
    assert torch.equal(out.detach().cpu(), want_out.detach())
    got = torch.autograd.grad(out, d_leaves, go.to(dev))
    for name, a, b in zip(("fmap1", "fmap2", "coords"), got, want):
        assert _rel(a.cpu(), b) < GRAD_RTOL, name


def test_sampler_interface(op_device):
    dev = op_device
    f1, f2, c = _case(1, 16, 8, 12, 2, 3)
    s = CoordSampler(None)
    s.register(f1.to(dev), f2.to(dev), num_levels=3)
    assert [tuple(p.shape) for p in s.f2_pyramid] == [(1, 16, 8, 12), (1, 16, 4, 6), (1, 16, 2, 3)]
    out = s.__corr__(c.to(dev), num_levels=2)
    assert tuple(out.shape) == (1, 6, 8, 12) and out.dtype == torch.float32
    s._update_fmap1(f2.to(dev))
    assert not torch.equal(s(c.to(dev), num_levels=2), out)
    with pytest.raises(ValueError):
        s(c.to(dev), num_levels=3)


def test_sample_tgt_matches_the_oracle(op_device):
    """PoseUpdate.sample_tgt (dualrefine/networks/utils/utils.py:383-404): bit-exact values, gradient to p2."""
    dev = op_device
    B, C, hh, ww = 2, 16, 10, 14
    g = torch.Generator().manual_seed(23)
    feat, wmap = torch.rand(B, C, hh, ww, generator=g), torch.rand(B, 1, hh, ww, generator=g)
    ys, xs = torch.meshgrid(torch.arange(hh).float(), torch.arange(ww).float(), indexing="ij")
    c1 = torch.stack([xs, ys])[None, :, None, None] + 1.5 * torch.randn(B, 2, 1, 1, hh, ww, generator=g)
    delta = torch.tensor([[0., 1., -1., 0., 0.], [0., 0., 0., 1., -1.]]).reshape(1, 2, 1, 5, 1, 1)   # utils.py:220-226
    p2 = (c1 + delta).requires_grad_(True)
    want = O.sample_tgt(feat, p2, wmap)
    gw = torch.autograd.grad(want[0].sum() + (want[1] ** 2).sum() + want[2].sum(), p2)[0]
    p2d = p2.detach().clone().to(dev).requires_grad_(True)
    got = sample_tgt(feat.to(dev), p2d, wmap.to(dev))
    for a, b in zip(got, want):
        assert torch.equal(a.detach().cpu(), b.detach())
    gg = torch.autograd.grad(got[0].sum() + (got[1] ** 2).sum() + got[2].sum(), p2d)[0]
    assert _rel(gg.cpu(), gw) < GRAD_RTOL
