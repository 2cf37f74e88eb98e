"""Smoothness, MAL student terms and matching mask kernels against the oracle / golden outputs.

Bars: distillation arg-min and matching mask bit-exact; losses 1e-5 relative; gradients 1e-4
relative to the gradient's max magnitude.
"""
import numpy as np
import pytest
import torch

from mal_b200 import raw
from mal_b200.utils.synthetic import make_photometric_inputs
from oracle import mal_oracle as O
from tests.backends import BACKENDS, handle_and_device
from tests.helpers import photometric_golden

LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4


def _gerr(a, b):
    s = float(b.abs().max())
    return float((a - b).abs().max()) / (s if s > 0 else 1.0)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(2, 32, 48), (1, 19, 77), (3, 8, 33)])
@pytest.mark.parametrize("normalise", [True, False])
def test_smooth_loss_and_gradient(backend, shape, normalise):
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t = make_photometric_inputs(B, H, W, seed=31)
    disp = t[("mono_disp", 0)].clone().requires_grad_(True)
    img = inputs[("color", 0, 0)]
    want = O.normalised_smooth_loss(disp, img) if normalise else O.smooth_loss(disp, img)
    g, = torch.autograd.grad(want, disp)
    out = raw.smooth(h, disp=disp.detach().to(dev), img=img.to(dev), normalise=normalise, with_grad=True)
    assert abs(float(out["loss"]) - float(want)) <= LOSS_RTOL * abs(float(want))
    assert _gerr(out["grad_disp"].cpu(), g) < GRAD_RTOL
    out2 = raw.smooth(h, disp=disp.detach().to(dev), img=img.to(dev), normalise=normalise, with_grad=False)
    assert torch.equal(out2["loss"], out["loss"])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(2, 32, 48), (1, 19, 77)])   # 48: 128-bit path; 77: scalar path
def test_smooth_two_disparities_one_launch(backend, shape):
    """Teacher and student disparity against the same image in one launch (the edge weights are shared), and
    the deferred chain through the mean-normalisation: (g - L_b / HW) * s_b  ==  the single-term results."""
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t = make_photometric_inputs(B, H, W, seed=33)
    img = inputs[("color", 0, 0)].to(dev)
    da, db = t[("mono_disp", 0)].to(dev), t[("multi_disp", 0)].to(dev)
    one_a = raw.smooth(h, disp=da, img=img, normalise=True, with_grad=True)
    one_b = raw.smooth(h, disp=db, img=img, normalise=True, with_grad=True)
    two = raw.smooth(h, disp=da, img=img, normalise=True, with_grad=True, disp_b=db)
    assert torch.equal(two["loss"], one_a["loss"]) and torch.equal(two["loss_b"], one_b["loss"])
    assert torch.equal(two["grad_disp"], one_a["grad_disp"]) and torch.equal(two["grad_disp_b"], one_b["grad_disp"])
    lazy = raw.smooth(h, disp=da, img=img, normalise=True, with_grad=True, disp_b=db, defer_fix=True)
    st = lazy["stats"]
    for g, want, term in ((lazy["grad_disp"], one_a["grad_disp"], 0), (lazy["grad_disp_b"], one_b["grad_disp"], 1)):
        fixed = (g - st[:, term, 0].view(B, 1, 1, 1) / (H * W)) * st[:, term, 1].view(B, 1, 1, 1)
        assert _gerr(fixed.cpu(), want.cpu()) < 1e-6
    # against the oracle as well
    for d, loss in ((t[("mono_disp", 0)], two["loss"]), (t[("multi_disp", 0)], two["loss_b"])):
        want = O.normalised_smooth_loss(d, inputs[("color", 0, 0)])
        assert abs(float(loss) - float(want)) <= LOSS_RTOL * abs(float(want))


def _main_case(B, H, W, seed):
    inputs, t = make_photometric_inputs(B, H, W, seed=seed)
    gen = torch.Generator().manual_seed(seed + 1)
    planes = [torch.rand(B, 1, H, W, generator=gen) * 0.2 for _ in range(3)]
    # exact ties exercise the first-index rule
    planes[2][:, :, ::3, ::5] = planes[0][:, :, ::3, ::5]
    planes[1][:, :, 1::4, ::2] = planes[0][:, :, 1::4, ::2]
    return inputs, t, planes


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(2, 32, 48), (1, 17, 23)])
@pytest.mark.parametrize("use_ens,dual", [(True, False), (False, False), (False, True)])
@pytest.mark.parametrize("as_disp", [False, True])
def test_main_terms(backend, shape, use_ens, dual, as_disp):
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t, (r_mono, r_ens, r_multi) = _main_case(B, H, W, 41)
    multi_disp = t[("multi_disp", 0)].clone().requires_grad_(True)
    mono_disp = t[("mono_disp", 0)].clone().requires_grad_(True)
    multi_depth, mono_depth = O.disp_to_depth(multi_disp, 0.1, 100.0)[1], O.disp_to_depth(mono_disp, 0.1, 100.0)[1]
    multi_leaf = multi_disp if as_disp else multi_depth.detach().clone().requires_grad_(True)
    mono_leaf = mono_disp if as_disp else mono_depth.detach().clone().requires_grad_(True)
    md = multi_depth if as_disp else multi_leaf
    mo = mono_depth if as_disp else mono_leaf
    # oracle restatement of loss_utils.py:192-254
    mask = torch.ones(B, 1, H, W) * t["consistency_mask"].unsqueeze(1) * (1 - t["augmentation_mask"])
    cm = (1 - mask).float()
    cons = (torch.abs(md - mo.detach()) * cm).mean()
    target = 1 / (mo.detach() * cm + md.detach() * (1 - cm))
    if use_ens:
        _, sel = torch.min(torch.cat([r_mono, r_ens, r_multi], 1), dim=1, keepdim=True)
        dd = torch.where(sel == 0, mo.detach(), (mo.detach() + md) / 2.0)
        dd = torch.where(sel == 2, md, dd)
    else:
        _, sel = torch.min(torch.cat([r_mono, r_multi], 1), dim=1, keepdim=True)
        dd = torch.where(sel == 0, mo if dual else mo.detach(), md)
    distil = (torch.abs(dd - md) * (1 - cm)).mean()
    leaves = [multi_leaf] + ([mono_leaf] if dual else [])
    g_cons = torch.autograd.grad(cons, multi_leaf, retain_graph=True)[0]
    g_dist = torch.autograd.grad(distil, leaves, retain_graph=True, allow_unused=True)

    d = lambda x: x.detach().to(dev)
    out = raw.main_terms(h, multi=d(multi_leaf), mono=d(mono_leaf), pixel_mask=d(t["consistency_mask"]),
                         sample_mask=d(t["augmentation_mask"]), mono_reproj=d(r_mono), multi_reproj=d(r_multi),
                         ens_reproj=d(r_ens) if use_ens else None, inputs_are_disp=as_disp, dual_distil=dual,
                         with_grad=True)
    assert np.array_equal(out["distil_index"].cpu().numpy(), sel.numpy().astype(np.uint8))
    assert abs(float(out["sums"][0]) - float(cons)) <= LOSS_RTOL * abs(float(cons))
    assert abs(float(out["sums"][1]) - float(distil)) <= LOSS_RTOL * abs(float(distil))
    assert torch.equal(out["consistency_target"].cpu(), target)
    assert _gerr(out["grad_cons"].cpu(), g_cons) < GRAD_RTOL
    assert _gerr(out["grad_distil"].cpu(), g_dist[0]) < GRAD_RTOL
    if dual:
        assert _gerr(out["grad_distil_mono"].cpu(), g_dist[1]) < GRAD_RTOL


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ["photometric_smooth.npz", "photometric_noise.npz"])
def test_matching_mask_against_reference_golden(backend, name):
    h, dev = handle_and_device(backend)
    inputs, t, ref = photometric_golden(name)
    mono_depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    got = raw.matching_mask(h, lowest_cost=t["lowest_cost"].to(dev), mono=mono_depth.to(dev))
    assert np.array_equal(got.cpu().numpy().astype(np.uint8), ref["matching_mask"])
    got2 = raw.matching_mask(h, lowest_cost=t["lowest_cost"].to(dev), mono=t[("mono_disp", 0)].to(dev),
                             mono_is_disp=True)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("backend", BACKENDS)
def test_matching_mask_fused_nearest_upsampling(backend):
    """repdepth.py:331-336 nearest up-sampling + trainer.py:592-593 product, in one kernel."""
    import torch.nn.functional as F
    h, dev = handle_and_device(backend)
    B, H, W = 2, 48, 80
    inputs, t = make_photometric_inputs(B, H, W, seed=51)
    gen = torch.Generator().manual_seed(52)
    mono_depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    low = 1 / (F.avg_pool2d(mono_depth, 4)[:, 0] * (0.3 + 2.5 * torch.rand(B, H // 4, W // 4, generator=gen)))
    conf = (torch.rand(B, H // 4, W // 4, generator=gen) < 0.7).float()
    outputs = {("mono_depth", 0, 0): mono_depth,
               "lowest_cost": F.interpolate(low.unsqueeze(1), [H, W], mode="nearest")[:, 0]}
    want = F.interpolate(conf.unsqueeze(1), [H, W], mode="nearest")[:, 0] * O.matching_mask(outputs)
    got = raw.matching_mask(h, lowest_cost=low.to(dev), confidence=conf.to(dev), mono=mono_depth.to(dev))
    assert torch.equal(got.cpu(), want)
