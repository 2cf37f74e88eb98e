"""Full-size (BASELINE.json configs[1]: batch 12, 192x640, 96 bins x 64 channels) checks on the GPU
through properties that do not need the oracle to finish in seconds:

  * batch sharding: a batch of 12 equals 12 batches of 1 (per-pixel outputs bit-exact; the masked
    sums add up) - the property the data-parallel split relies on;
  * the fused WARP kernel equals the PRED kernel fed with materialised warps (two code paths);
  * cost-volume self-consistency: arg-min / lowest_cost / confidence / missing fill agree with the
    volume they were derived from;
  * linearity of the masked sums in the per-pixel weights;
  * one oracle spot check on a single sample at full resolution.
"""
import numpy as np
import pytest
import torch

from mal_b200 import _capi, layers, raw, step as S
from mal_b200.utils.synthetic import to_device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    opt = S.default_opt(12)
    b = to_device(S.synthetic_batch(opt, seed=2024), torch.device("cuda:0"))
    return opt, b, _capi.lib()


def _teacher(h, b, sl=slice(None), **kw):
    tgt, src = b["color_0"][sl], [b["color_-1"][sl], b["color_1"][sl]]
    ident = raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False)["min_reproj"]
    return raw.photo(h, target=tgt, src=src, syn=[b["syn_-1"][sl], b["syn_1"][sl]], depth=b["mono_disp"][sl].detach(),
                     K=b["K"][sl], inv_K=b["inv_K"][sl], T=[b["T_-1"][sl].detach(), b["T_1"][sl].detach()],
                     identity_min=ident, noise=b["noise_mono"][sl], with_grad=True, **kw)


def test_photo_batch_sharding(full):
    opt, b, h = full
    whole = _teacher(h, b)
    S_sum, W_sum = 0.0, 0.0
    for i in range(opt.batch_size):
        one = _teacher(h, b, slice(i, i + 1))
        for k in ("min_reproj", "selection", "grad_depth"):
            assert torch.equal(one[k], whole[k][i:i + 1]), (k, i)
        assert torch.allclose(one["grad_P"], whole["grad_P"][i:i + 1], rtol=1e-5, atol=1e-7)
        S_sum += float(one["sums"][0])
        W_sum += float(one["sums"][1])
    assert abs(S_sum - float(whole["sums"][0])) <= 1e-5 * abs(S_sum)
    assert W_sum == float(whole["sums"][1])


def test_fused_warp_equals_materialised_warp(full):
    opt, b, h = full
    H, W, B = opt.height, opt.width, opt.batch_size
    depth = layers.disp_to_depth(b["multi_disp"].detach(), opt.min_depth, opt.max_depth)[1]
    cam = raw.backproject(h, depth, b["inv_K"])
    preds = []
    for f in (-1, 1):
        grid, _ = raw.project3d(h, cam, b["K"], b["T_%d" % f].detach(), H, W)
        preds.append(raw.grid_sample(h, b["color_%d" % f], grid, align_corners=True, border=True))
    mask = (b["noise_main"][:, 0] > 0).float()
    fused = raw.photo(h, target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["multi_disp"].detach(),
                      K=b["K"], inv_K=b["inv_K"], T=[b["T_-1"].detach(), b["T_1"].detach()], pixel_mask=mask)
    classic = raw.photo(h, target=b["color_0"], src=preds, mode=raw.PHOTO_PRED, pixel_mask=mask)
    # two code paths (in-kernel warp vs stand-alone backproject / project3d / grid_sample kernels): bit-identical
    assert torch.equal(fused["min_reproj"], classic["min_reproj"])
    assert torch.equal(fused["selection"], classic["selection"])
    assert torch.equal(fused["sums"], classic["sums"])


def test_sums_are_linear_in_the_pixel_weights(full):
    opt, b, h = full
    kw = dict(target=b["color_0"], src=[b["color_-1"], b["color_1"]], depth=b["multi_disp"].detach(), K=b["K"],
              inv_K=b["inv_K"], T=[b["T_-1"].detach(), b["T_1"].detach()])
    m1 = (b["noise_main"][:, 0] > 0.3).float()
    m2 = (b["noise_mono"][:, 0] < -0.2).float()
    s1, s2 = raw.photo(h, pixel_mask=m1, **kw)["sums"], raw.photo(h, pixel_mask=m2, **kw)["sums"]
    s12 = raw.photo(h, pixel_mask=m1 + 2 * m2, **kw)["sums"]
    assert abs(float(s12[0]) - float(s1[0] + 2 * s2[0])) <= 2e-5 * float(s12[0])
    assert abs(float(s12[1]) - float(s1[1] + 2 * s2[1])) <= 1e-6 * float(s12[1])


def test_cost_volume_self_consistency_and_sharding(full):
    opt, b, h = full
    kw = lambda sl: dict(current=b["current_feats"][sl], lookup=b["lookup_feats"][sl], poses=b["relative_poses"][sl],
                         K=b["K2"][sl], inv_K=b["inv_K2"][sl], bins=b["bins"])
    out = raw.cost_volume(h, **kw(slice(None)))
    cv, miss, conf, idx, low = (out[k] for k in ("cost_volume", "missing_mask", "confidence", "argmin", "lowest_cost"))
    assert cv.shape == (12, 96, 48, 160)
    viz = torch.where(cv == 0, torch.full_like(cv, 100.0), cv)
    mn, am = viz.min(1)
    assert torch.equal(am.int(), idx)                                     # first-index arg-min
    assert torch.equal(low, 1 / b["bins"][idx.long()])
    assert torch.equal(conf, ((cv * (1 - miss)) > 0).sum(1).eq(96).float())
    mx = (cv * (1 - miss)).max(1, keepdim=True)[0]
    assert torch.equal(torch.where(miss > 0, mx.expand_as(cv), cv), cv)   # missing entries hold the per-pixel max
    assert 0.3 < float(conf.mean()) < 0.95
    for i in (0, 7, 11):
        one = raw.cost_volume(h, **kw(slice(i, i + 1)))
        for k in ("cost_volume", "missing_mask", "confidence", "argmin", "lowest_cost"):
            assert torch.equal(one[k], out[k][i:i + 1]), (k, i)
    masked = raw.cost_volume(h, apply_confidence=True, want_missing=False, **kw(slice(None)))["cost_volume"]
    assert torch.equal(masked, cv * conf.unsqueeze(1))


def test_full_resolution_sample_against_oracle(full):
    """One sample at 192x640 / 96 bins x 64 channels through the whole step oracle (a few seconds)."""
    from oracle.step_oracle import oracle_step
    opt1 = S.default_opt(1)
    b1 = S.synthetic_batch(opt1, seed=77)
    total, loss_list, grads, aux = oracle_step(b1, opt1)
    scalars, g, outputs = S.fused_step(_capi.lib(), to_device(b1, torch.device("cuda:0")), opt1,
                                       torch.tensor([0.5, 0.5], device="cuda:0"))
    assert abs(float(scalars[0]) - float(total)) <= 1e-5 * abs(float(total))
    assert torch.equal(outputs["cost_volume"].cpu(), aux["cv"])
    assert torch.equal(outputs["consistency_mask"].cpu(), aux["mask"])
    assert np.array_equal(outputs["mal_distil_index"].cpu().numpy(), aux["distil_idx"].numpy().astype(np.uint8))
    for a, bb in zip(g, grads):
        assert float((a.cpu() - bb).abs().max()) <= 1e-4 * float(bb.abs().max())


def test_low_resolution_disparity_full_size(full):
    """SURVEY.md 8 f.3 at 192x640: the teacher pass reading the scale-1..3 disparity in place equals the same
    pass on F.interpolate(disp) computed on the CPU (selection, min-reprojection and gradient planes bit for
    bit), and the gradient handed to the low resolution is the adjoint of the full-resolution one."""
    import torch.nn.functional as F
    opt, b, h = full
    for s in (1, 2, 3):
        lo = F.avg_pool2d(b["mono_disp"].detach(), 2 ** s)
        up_cpu = F.interpolate(lo.cpu(), [opt.height, opt.width], mode="bilinear", align_corners=False)
        up = raw.upsample_bilinear(h, lo, (opt.height, opt.width))
        assert torch.equal(up.cpu(), up_cpu)
        bb = dict(b)
        bb["mono_disp"] = up
        want = _teacher(h, bb)
        bb["mono_disp"] = lo
        got = _teacher(h, bb)
        for k in ("min_reproj", "selection", "grad_depth", "grad_P", "sums"):
            assert torch.equal(got[k], want[k]), (s, k)
        g_lo = raw.upsample_bilinear_backward(h, got["grad_depth"], lo.shape[-2:])
        lo_leaf = lo.cpu().clone().requires_grad_(True)
        (ref,) = torch.autograd.grad(F.interpolate(lo_leaf, [opt.height, opt.width], mode="bilinear", align_corners=False),
                                     lo_leaf, got["grad_depth"].cpu())
        assert float((g_lo.cpu() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_correlation_lookup_full_size_against_the_oracle():
    """SURVEY.md 8 f.2 at DualRefine's training shape (64 channels, 48x160, 3 levels x 17 candidates, two
    samples): bit-exact forward against the CPU oracle; gradients within 1e-4."""
    from oracle import mal_oracle as O
    h = _capi.lib()
    dev = torch.device("cuda:0")
    B, C, hh, ww, L, D = 2, 64, 48, 160, 3, 17
    g = torch.Generator().manual_seed(99)
    f1, f2 = torch.rand(B, C, hh, ww, generator=g), torch.rand(B, C, hh, ww, generator=g)
    ys, xs = torch.meshgrid(torch.arange(hh).float(), torch.arange(ww).float(), indexing="ij")
    dx = torch.linspace(-8, 8, D)[None, None, None, :, None, None] * torch.tensor([1.0, 2.0, 4.0])[None, None, :, None, None, None]
    coords = torch.stack([xs, ys])[None, :, None, None] + 0.4 * dx * torch.tensor([1.0, 0.2])[None, :, None, None, None, None] \
        + 0.3 * torch.randn(B, 2, L, D, hh, ww, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (f1, f2, coords)]
    want = O.corr_lookup(leaves[0], O.corr_pyramid(leaves[1], L), leaves[2])
    go = torch.randn(want.shape, generator=g)
    want_g = torch.autograd.grad(want, leaves, go)
    pyr = raw.corr_pyramid(h, f2.to(dev), L)
    got = raw.corr_lookup(h, f1.to(dev), pyr, coords.to(dev))
    assert torch.equal(got.cpu(), want.detach())
    gc, g1, gp = raw.corr_lookup_backward(h, f1.to(dev), pyr, coords.to(dev), go.to(dev))
    rel = lambda a, b_: float((a - b_).abs().max()) / float(b_.abs().max())
    assert rel(gc.cpu(), want_g[2]) < 1e-4
    assert rel(g1.cpu(), want_g[0]) < 1e-4
    # gradient with respect to fmap2 = adjoint of the pooling chain applied to the pyramid gradient
    from mal_b200 import ops
    f2d = f2.to(dev).requires_grad_(True)
    out = ops.corr_lookup(f1.to(dev), ops.corr_pyramid(f2d, L), coords.to(dev))
    (g2,) = torch.autograd.grad(out, f2d, go.to(dev))
    assert rel(g2.cpu(), want_g[1]) < 1e-4


@pytest.mark.parametrize("mode", ["masks", "main_temporal"])
def test_whole_step_at_batch_12_against_the_oracle(mode):
    """configs[1] at its full size (batch 12, 192x640, 96 bins x 64 channels) through MalStep's captured graph
    against oracle_step (a few seconds of CPU): losses 1e-5, gradients 1e-4, cost volume / matching mask /
    distillation indices bit-exact.  "masks": the temporal hint synthesised inside the step from packed instance
    masks; "main_temporal": ready-made syn images that the STUDENT pass uses as well (multi_has_ins, a 4-candidate
    gradient pass with masks, compute_main_losses :146-176)."""
    from oracle.step_oracle import oracle_step
    opt = S.default_opt(12, main_temporal=(mode == "main_temporal"))
    b = S.synthetic_batch(opt, seed=77, with_masks=(mode == "masks"))
    kw = dict(multi_has_ins=True) if mode == "main_temporal" else {}
    total, loss_list, grads, aux = oracle_step(b, opt, (0.5, 0.5), **kw)
    d = to_device(b, torch.device("cuda:0"))
    with torch.no_grad():
        scalars, g, outputs = S.fused_step(_capi.lib(), d, opt, torch.full((2,), 0.5, device="cuda:0"), **kw)
    torch.cuda.synchronize()
    assert abs(float(scalars[0]) - float(total)) <= 1e-5 * abs(float(total))
    for a, w in zip((scalars[1], scalars[2]), loss_list):
        assert abs(float(a) - float(w)) <= 1e-5 * abs(float(w))
    assert torch.equal(outputs["cost_volume"].cpu(), aux["cv"])
    assert torch.equal(outputs["consistency_mask"].cpu(), aux["mask"])
    assert np.array_equal(outputs["mal_distil_index"].cpu().numpy(), aux["distil_idx"].numpy().astype(np.uint8))
    for a, w, name in zip(g, grads, S.LEAVES):
        scale = float(w.abs().max())
        assert float((a.cpu().reshape(w.shape) - w).abs().max()) <= 1e-4 * scale, name
