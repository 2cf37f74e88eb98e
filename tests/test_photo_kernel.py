"""Fused photometric kernel (mal_photo_forward) against the reference's golden outputs and
against the oracle on seeded inputs.

Bars (BASELINE.json north_star): selection indices and automask bit-exact; per-pixel
min-reprojection bit-exact (it feeds the distillation argmin); loss within 1e-5 relative;
gradients within 1e-4 relative (of the gradient's max magnitude).
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


def _grad_err(a, b):
    scale = float(b.abs().max())
    if scale == 0.0:
        return float(a.abs().max())
    return float((a - b).abs().max()) / scale


def _run_mono(h, dev, inputs, t, temporal, with_grad=True):
    d = lambda x: x.to(dev)
    tgt = d(inputs[("color", 0, 0)])
    src = [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
    ident = raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False)["min_reproj"]
    out = raw.photo(h, target=tgt, src=src,
                    syn=[d(t[("syn", -1, 0)]), d(t[("syn", 1, 0)])] if temporal else None,
                    depth=d(t[("mono_disp", 0)]), K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]),
                    T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])],
                    identity_min=ident, noise=d(t["noise"][0]), with_grad=with_grad, want_weight=True)
    return ident, out


def _oracle_mono(inputs, t, temporal):
    H, W = inputs[("color", 0, 0)].shape[-2:]
    Ts = {f: t[("cam_T_cam", 0, f)].clone().requires_grad_(True) for f in (-1, 1)}
    o = {("disp", 0): t[("mono_disp", 0)].clone().requires_grad_(True)}
    for f in (-1, 1):
        o[("cam_T_cam", 0, f)] = Ts[f]
        o[("syn", f, 0)] = t[("syn", f, 0)]
    O.images_pred(inputs, o, height=H, width=W)
    losses, mono_reproj, aux = O.mono_losses(inputs, o, temporal, True, noise=t["noise"][0])
    g = torch.autograd.grad(losses["reproj_loss/0"], [o[("disp", 0)], Ts[-1], Ts[1]])
    return losses, mono_reproj, aux, g


def _check_grads(out, inputs, g):
    scale = 1.0 / (out["sums"][1].cpu() + 1e-7)
    assert _grad_err(out["grad_depth"].cpu() * scale, g[0]) < GRAD_RTOL
    Kt = inputs[("K", 0)][:, :3, :].transpose(1, 2)
    gP = out["grad_P"].cpu() * scale
    for i in range(2):
        assert _grad_err(Kt @ gP[:, i].view(-1, 3, 4), g[1 + i]) < GRAD_RTOL


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("name", ["photometric_smooth.npz", "photometric_noise.npz"])
@pytest.mark.parametrize("temporal,tag", [(False, "plain"), (True, "temporal")])
def test_mono_pass_against_reference_golden(backend, name, temporal, tag):
    h, dev = handle_and_device(backend)
    inputs, t, ref = photometric_golden(name)
    ident, out = _run_mono(h, dev, inputs, t, temporal)
    assert np.array_equal(ident.cpu().numpy(), ref["identity_min"])
    sel = out["selection"].cpu().numpy()
    assert np.array_equal(sel & 0x7F, ref[f"mono_{tag}_frame_idx"])          # bit-exact argmin
    assert np.array_equal(sel >> 7, ref[f"mono_{tag}_automask"])              # bit-exact automask
    assert np.array_equal(out["min_reproj"].cpu().numpy(), ref[f"mono_{tag}_min_reproj"])
    assert np.array_equal(out["weight"].cpu().numpy(), ref[f"mono_{tag}_automask"].astype(np.float32))
    got, want = float(out["sums"][2]), float(ref[f"mono_{tag}_reproj_loss"])
    assert abs(got - want) <= LOSS_RTOL * abs(want)
    _, _, _, g = _oracle_mono(inputs, t, temporal)
    _check_grads(out, inputs, g)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape,seed,white", [((1, 40, 72), 5, False), ((2, 35, 50), 6, True), ((1, 16, 32), 7, False)])
def test_mono_pass_against_oracle_ragged(backend, shape, seed, white):
    """Sizes that are not tile multiples (tile 32x16) and a one-tile image."""
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t = make_photometric_inputs(B, H, W, seed=seed, white_noise=white)
    for temporal in (False, True):
        ident, out = _run_mono(h, dev, inputs, t, temporal)
        losses, mono_reproj, aux, g = _oracle_mono(inputs, t, temporal)
        sel = out["selection"].cpu().numpy()
        assert np.array_equal(sel & 0x7F, aux["frame_idx"].numpy().astype(np.uint8))
        assert np.array_equal(sel >> 7, aux["automask"].numpy().astype(np.uint8))
        assert torch.equal(out["min_reproj"].cpu(), mono_reproj)
        want = float(losses["reproj_loss/0"])
        assert abs(float(out["sums"][2]) - want) <= LOSS_RTOL * abs(want)
        _check_grads(out, inputs, g)


@pytest.mark.parametrize("backend", BACKENDS)
def test_no_grad_variant_matches_grad_variant(backend):
    h, dev = handle_and_device(backend)
    inputs, t = make_photometric_inputs(1, 48, 64, seed=9)
    _, a = _run_mono(h, dev, inputs, t, True, with_grad=True)
    _, b = _run_mono(h, dev, inputs, t, True, with_grad=False)
    assert torch.equal(a["min_reproj"], b["min_reproj"]) and torch.equal(a["selection"], b["selection"])
    assert torch.equal(a["sums"], b["sums"])


@pytest.mark.parametrize("backend", BACKENDS)
def test_pred_mode_gradients(backend):
    """PRED mode: loss from already-warped images, d/d pred against autograd on the oracle."""
    h, dev = handle_and_device(backend)
    inputs, t = make_photometric_inputs(2, 33, 47, seed=10)
    tgt = inputs[("color", 0, 0)]
    preds = [t[("syn", f, 0)].clone().requires_grad_(True) for f in (-1, 1)]
    cands = torch.cat([O.reprojection_loss(p, tgt) for p in preds], 1)
    reproj, idx = torch.min(cands, 1, keepdim=True)
    mask = t["consistency_mask"].unsqueeze(1)[:, :, :33, :47] * (1 - t["augmentation_mask"])
    loss = (reproj * mask).sum() / (mask.sum() + 1e-7)
    g = torch.autograd.grad(loss, preds)
    out = raw.photo(h, target=tgt.to(dev), src=[p.detach().to(dev) for p in preds], mode=raw.PHOTO_PRED,
                    pixel_mask=mask[:, 0].contiguous().to(dev), with_grad=True)
    assert np.array_equal(out["selection"].cpu().numpy() & 0x7F, idx.numpy().astype(np.uint8))
    assert abs(float(out["sums"][2]) - float(loss)) <= LOSS_RTOL * abs(float(loss))
    scale = 1.0 / (out["sums"][1].cpu() + 1e-7)
    for i in range(2):
        assert _grad_err(out["grad_pred"][i].cpu() * scale, g[i]) < GRAD_RTOL


def test_argument_errors_are_reported():
    from tests.emu.emu_lib import emu
    h = emu()
    inputs, t = make_photometric_inputs(1, 16, 32, seed=1)
    with pytest.raises(RuntimeError, match="WARP mode needs"):
        raw.photo(h, target=inputs[("color", 0, 0)], src=[inputs[("color", -1, 0)], inputs[("color", 1, 0)]])
    with pytest.raises(ValueError):
        raw.photo(h, target=inputs[("color", 0, 0)], src=[inputs[("color", -1, 0)][:, :, :8], inputs[("color", 1, 0)]],
                  mode=raw.PHOTO_PRED)
    # the options added in round 2 refuse the combinations they do not implement
    tgt, src = inputs[("color", 0, 0)], [inputs[("color", -1, 0)], inputs[("color", 1, 0)]]
    geom = dict(depth=t[("mono_disp", 0)], K=inputs[("K", 0)], inv_K=inputs[("inv_K", 0)],
                T=[t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)]])
    with pytest.raises(RuntimeError, match="WARP-mode input"):
        raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, warped=src)
    with pytest.raises(RuntimeError, match="plain WARP pass"):
        raw.photo(h, target=tgt, src=src, warped=src, zero_img=True, want_target_out=True, **geom)
    ident = raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False)["min_reproj"]
    with pytest.raises(RuntimeError, match="DynamicDepth"):
        raw.photo(h, target=tgt, src=src, identity_min=ident, noise=t["noise"][0], selec_reproj=True, **geom)
    with pytest.raises(RuntimeError, match="needs `noise`"):
        raw.photo(h, target=tgt, src=src, syn=src, identity_in_pass=True, **geom)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape,K_name", [((1, 48, 128), "CITYSCAPES_K"), ((2, 24, 80), "KITTI_K")])
def test_dualrefine_convention_and_cityscapes_intrinsics(backend, shape, K_name):
    """BASELINE configs 3 and 4: Cityscapes-shaped intrinsics (192x512 aspect) and DualRefine's
    half-pixel Project3D + align_corners=False sampling (dualrefine/layers.py:224-225,
    dualrefine/trainer.py:444-447), forward selections bit-exact and gradients vs autograd."""
    from mal_b200.utils import synthetic
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t = make_photometric_inputs(B, H, W, seed=23, normalised_K=getattr(synthetic, K_name),
                                        translation_scale=0.3)
    d = lambda x: x.to(dev)
    for conv, oconv in ((raw.CONV_DUALREFINE, O.DUALREFINE), (raw.CONV_MANYDEPTH, O.MANYDEPTH)):
        Ts = {f: t[("cam_T_cam", 0, f)].clone().requires_grad_(True) for f in (-1, 1)}
        o = {("disp", 0): t[("mono_disp", 0)].clone().requires_grad_(True)}
        for f in (-1, 1):
            o[("cam_T_cam", 0, f)] = Ts[f]
        O.images_pred(inputs, o, height=H, width=W, convention=oconv)
        losses, mono_reproj, aux = O.mono_losses(inputs, o, False, False, noise=t["noise"][0])
        g = torch.autograd.grad(losses["reproj_loss/0"], [o[("disp", 0)], Ts[-1], Ts[1]])
        tgt = d(inputs[("color", 0, 0)])
        src = [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
        ident = raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False)["min_reproj"]
        out = raw.photo(h, target=tgt, src=src, depth=d(t[("mono_disp", 0)]), K=d(inputs[("K", 0)]),
                        inv_K=d(inputs[("inv_K", 0)]), T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])],
                        identity_min=ident, noise=d(t["noise"][0]), with_grad=True, convention=conv)
        sel = out["selection"].cpu().numpy()
        assert np.array_equal(sel & 0x7F, aux["frame_idx"].numpy().astype(np.uint8))
        assert np.array_equal(sel >> 7, aux["automask"].numpy().astype(np.uint8))
        assert torch.equal(out["min_reproj"].cpu(), mono_reproj)
        want = float(losses["reproj_loss/0"].detach())
        assert abs(float(out["sums"][2]) - want) <= LOSS_RTOL * abs(want)
        _check_grads(out, inputs, g)


@pytest.mark.parametrize("backend", BACKENDS)
def test_ensemble_disparity_average_in_kernel(backend):
    """depth_b: the kernel averages two disparities like (a.detach() + b.detach()) / 2.0
    (manydepth/trainer.py:598) before warping."""
    h, dev = handle_and_device(backend)
    inputs, t = make_photometric_inputs(2, 32, 48, seed=29, translation_scale=0.3)
    d = lambda x: x.to(dev)
    want = O.images_pred_ensemble(inputs, t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)],
                                  (t[("mono_disp", 0)] + t[("multi_disp", 0)]) / 2.0, height=32, width=48)
    out = raw.photo(h, target=d(inputs[("color", 0, 0)]), src=[d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])],
                    depth=d(t[("mono_disp", 0)]), depth_b=d(t[("multi_disp", 0)]), K=d(inputs[("K", 0)]),
                    inv_K=d(inputs[("inv_K", 0)]), T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])],
                    want_selection=False)
    assert torch.equal(out["min_reproj"].cpu(), want)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(2, 48, 96), (1, 37, 50)])
def test_ready_made_warps_give_the_same_bits(backend, shape):
    """mal_photo_args.warped: a WARP-mode pass that stages the caller's materialised warps (mal_temporal_warp:
    trainer.py:1122-1125 makes them for image_synthesis) instead of re-warping - every output identical,
    gradients included; ragged sizes take the plain loader, border tiles the reflected halo."""
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t = make_photometric_inputs(B, H, W, seed=321)
    d = lambda x: x.to(dev)
    tgt, src = d(inputs[("color", 0, 0)]), [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
    geom = dict(K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]), T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])])
    warped = raw.temporal_warp(h, src=src, depth=d(t[("mono_disp", 0)]), **geom)
    ident = raw.photo(h, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False)["min_reproj"]
    for syn, grad_syn in ((None, False), ([d(t[("syn", -1, 0)]), d(t[("syn", 1, 0)])], True)):
        kw = dict(target=tgt, src=src, syn=syn, depth=d(t[("mono_disp", 0)]), identity_min=ident, noise=d(t["noise"][0]),
                  with_grad=True, want_weight=True, want_grad_syn=grad_syn, **geom)
        want = raw.photo(h, **kw)
        got = raw.photo(h, warped=warped, **kw)
        for k in ("sums", "min_reproj", "selection", "weight", "grad_depth", "grad_P"):
            assert torch.equal(got[k], want[k]), k
        if grad_syn:
            for a, b in zip(got["grad_syn"], want["grad_syn"]):
                assert torch.equal(a, b)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("automask", [True, False])
def test_dynamicdepth_mode_at_the_c_level(backend, automask):
    """photo_kernel<..., DD> through the raw binding (zero_img + selec_reproj, the identity candidates given as `syn`):
    masked mean, the zeroed target and d/d disparity against the oracle's compute_losses of one scale
    (dynamicdepth/trainer.py:958-975, :1006-1128)."""
    h, dev = handle_and_device(backend)
    inputs, t = make_photometric_inputs(2, 40, 72, num_scales=1, seed=19, translation_scale=0.3)
    inputs[("color", -1, 0)][:, :, 4:14, 6:30] = 0.0
    inputs[("color", 1, 0)][:, :, 18:30, 30:60] = 0.0
    disp = t[("mono_disp", 0)].clone().requires_grad_(True)
    o = {("disp", 0): disp, ("cam_T_cam", 0, -1): t[("cam_T_cam", 0, -1)], ("cam_T_cam", 0, 1): t[("cam_T_cam", 0, 1)]}
    O.images_pred(inputs, o, num_scales=1, height=40, width=72)
    ref_inputs = {k: v.clone() for k, v in inputs.items()}
    want, aux = O.dynamicdepth_compute_losses(ref_inputs, o, (0,), noises=t["noise"][:1], automask=automask)
    want_g, = torch.autograd.grad(want["reproj_loss/0"], disp)
    d = lambda x: x.to(dev)
    ident = [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
    out = raw.photo(h, target=d(inputs[("color", 0, 0)]), src=ident, syn=ident if automask else None,
                    depth=d(t[("mono_disp", 0)]), K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]),
                    T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])], noise=d(t["noise"][0]) if automask else None,
                    zero_img=True, selec_reproj=True, identity_in_pass=True, want_target_out=True, with_grad=True)
    w = float(want["reproj_loss/0"])
    assert abs(float(out["sums"][2]) - w) <= LOSS_RTOL * abs(w)
    assert torch.equal(out["target_out"].cpu(), ref_inputs[("color", 0, 0)])
    if automask:
        assert torch.equal((out["selection"].cpu() >> 7).float(), aux[("mask", 0)].float())
    assert _grad_err(out["grad_depth"].cpu() / (out["sums"][1].cpu() + 1e-7), want_g) < GRAD_RTOL
