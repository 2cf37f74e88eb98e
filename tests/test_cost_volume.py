"""Plane-sweep cost volume (mal_cost_volume_forward) against the reference's golden output and
against the oracle on seeded inputs.

Bars: missing mask, confidence mask and arg-min indices bit-exact; cost volume and lowest_cost
bit-exact on the machine that made the golden (the kernel follows torch's CPU summation tree),
checked here to 1e-6 relative so another host BLAS cannot fail the CPU suite.
"""
import numpy as np
import pytest
import torch

from mal_b200 import raw
from mal_b200.utils.synthetic import make_cost_volume_inputs
from oracle import mal_oracle as O
from tests.backends import BACKENDS, handle_and_device
from tests.helpers import load_npz, rel_err


def _run(h, dev, cv, **kw):
    d = lambda x: x.to(dev)
    return raw.cost_volume(h, current=d(cv["current_feats"]), lookup=d(cv["lookup_feats"]),
                           poses=d(cv["relative_poses"]), K=d(cv["K"]), inv_K=d(cv["inv_K"]), bins=d(cv["bins"]), **kw)


@pytest.mark.parametrize("backend", BACKENDS)
def test_cost_volume_against_reference_golden(backend):
    h, dev = handle_and_device(backend)
    g = load_npz("cost_volume.npz")
    cv = {k[3:]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith("in_")}
    out = _run(h, dev, cv)
    assert np.array_equal(out["missing_mask"].cpu().numpy().astype(np.uint8), g["ref_missing"])
    assert np.array_equal(out["cost_volume"].cpu().numpy(), g["ref_cost_volume"])
    assert np.array_equal(out["confidence"].cpu().numpy(), g["ref_confidence"])
    assert np.array_equal(out["argmin"].cpu().numpy(), g["ref_argmin"])
    assert np.array_equal(out["lowest_cost"].cpu().numpy(), g["ref_lowest_cost"])


CASES = [
    # B, H, W, C, bins, lookups, seed, zero-pose sample
    (1, 96, 160, 64, 96, 1, 11, None),     # the training shape's channel/bin counts
    (2, 64, 100, 24, 20, 2, 12, 1),        # ragged: 25-pixel rows, channels not a multiple of 16
    (1, 48, 72, 80, 37, 1, 13, None),      # 5 chunks, bins not a multiple of the group size
]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("B,H,W,C,bins,F,seed,zero", CASES)
def test_cost_volume_against_oracle(backend, B, H, W, C, bins, F, seed, zero):
    h, dev = handle_and_device(backend)
    cv = make_cost_volume_inputs(B, H, W, channels=C, num_lookup=F, num_bins=bins, seed=seed, zero_pose_sample=zero,
                                 max_bin=10.0)
    vol, miss = O.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"], cv["inv_K"],
                                 cv["bins"])
    conf = O.confidence_mask(vol * (1 - miss))
    low, idx = O.lowest_cost(vol, cv["bins"])
    out = _run(h, dev, cv)
    assert torch.equal(out["missing_mask"].cpu(), miss)
    assert rel_err(out["cost_volume"].cpu(), vol) < 1e-6
    assert torch.equal(out["cost_volume"].cpu(), vol)
    assert torch.equal(out["confidence"].cpu(), conf)
    assert torch.equal(out["argmin"].cpu().long(), idx)
    assert torch.equal(out["lowest_cost"].cpu(), low)
    # forward(): cost_volume *= confidence_mask.unsqueeze(1)
    out2 = _run(h, dev, cv, apply_confidence=True, want_missing=False, want_head=False)
    assert torch.equal(out2["cost_volume"].cpu(), vol * conf.unsqueeze(1))


@pytest.mark.parametrize("backend", BACKENDS)
def test_cost_volume_dualrefine_convention(backend):
    """dualrefine/networks/resnet_encoder.py:202: half-pixel projection + align_corners=False."""
    h, dev = handle_and_device(backend)
    cv = make_cost_volume_inputs(1, 64, 96, channels=32, num_lookup=1, num_bins=24, seed=21, max_bin=10.0)
    vol, miss = O.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"], cv["inv_K"],
                                 cv["bins"], convention=O.DUALREFINE)
    out = _run(h, dev, cv, convention=raw.CONV_DUALREFINE)
    assert torch.equal(out["missing_mask"].cpu(), miss)
    assert torch.equal(out["cost_volume"].cpu(), vol)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("cv_min,set_1,pool,speckle", [(True, False, True, False), (False, True, False, False),
                                                       (True, False, False, False), (False, False, True, False),
                                                       (True, False, True, True)])
def test_dynamicdepth_cost_volume_variant(backend, cv_min, set_1, pool, speckle, monkeypatch):
    """dynamicdepth/networks/resnet_encoder.py:148-249: min over lookup frames and the occlusion
    fill (set_1 / 3-D max-pool) of the warped features, one sample with augmentation on (no fill).
    `speckle`: isolated occluded pixels everywhere, so that most samples sit next to one - more than the pool
    fill's vector cache holds (the overflow is warped in place)."""
    h, dev = handle_and_device(backend)
    B, H, W, C, nb = 2, 64, 96, 32, 20
    cv = make_cost_volume_inputs(B, H, W, channels=C, num_lookup=2, num_bins=nb, seed=91, min_bin=0.5, max_bin=6.0,
                                 translation_scale=0.5)
    gen = torch.Generator().manual_seed(92)
    look_img = torch.rand(B, 3, H, W, generator=gen)
    look_img[:, :, 20:44, 24:64] = 0.0
    look_img[:, :, 4:16, 70:90] = 0.01
    if speckle:
        look_img[:, :, 0::8, 0::8] = 0.0       # F.interpolate(nearest) keeps every second of these
    aug = torch.zeros(B, 1, 1, 1)
    aug[1] = 1
    want_vol, want_miss = O.match_features_dynamic(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"],
                                                   cv["K"], cv["inv_K"], cv["bins"], look_img, cv_min, aug, set_1, pool,
                                                   1, 0.7)
    occ = (O.occlusion_batch(look_img, H // 4, W // 4)[:, 0] > 0).float()
    assert 0.02 < float(occ.mean()) < 0.5
    mode = raw.OCC_SET_1 if set_1 else (raw.OCC_POOL if pool else raw.OCC_NONE)
    out = _run(h, dev, cv, cv_min=cv_min, occ=occ.to(dev), occ_mode=mode, pool_radius=1, pool_th=0.7,
               aug_mask=aug.to(dev))
    assert torch.equal(out["missing_mask"].cpu(), want_miss)
    assert torch.equal(out["cost_volume"].cpu(), want_vol)
    # the same through the one-pixel-per-lane kernel (what C > 64 takes)
    monkeypatch.setenv("MAL_CV_KERNEL", "lane")
    out2 = _run(h, dev, cv, cv_min=cv_min, occ=occ.to(dev), occ_mode=mode, pool_radius=1, pool_th=0.7,
                aug_mask=aug.to(dev))
    monkeypatch.delenv("MAL_CV_KERNEL")
    assert torch.equal(out2["cost_volume"].cpu(), want_vol) and torch.equal(out2["missing_mask"].cpu(), want_miss)
    if mode != raw.OCC_NONE:   # the fill really changed something, and only on the un-augmented sample
        plain, _ = O.match_features_dynamic(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"],
                                            cv["inv_K"], cv["bins"], look_img, cv_min, aug, False, False, 1, 0.7)
        assert not torch.equal(plain[0], want_vol[0]) and torch.equal(plain[1], want_vol[1])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("C", [32, 80])   # 32: four-lanes-per-pixel sweep; 80 (5 chunks): one-pixel-per-lane sweep
def test_cost_volume_nan_and_inf_features(backend, C):
    """A NaN / +Inf in the current features poisons every plane of that pixel (the reference multiplies the
    channel mean by the edge mask, NaN * 0, and the max-fill spreads it).  torch.min then returns the first
    NaN; the arg-min must follow and never leave [0, bins) (it indexes the depth bins for lowest_cost)."""
    h, dev = handle_and_device(backend)
    cv = make_cost_volume_inputs(1, 64, 96, channels=C, num_lookup=1, num_bins=24, seed=31, max_bin=10.0)
    cv["current_feats"][0, 3, 7, 9] = float("nan")
    cv["current_feats"][0, 5, 9, 14] = float("inf")
    vol, miss = O.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"], cv["inv_K"],
                                 cv["bins"])
    low, idx = O.lowest_cost(vol, cv["bins"])
    out = _run(h, dev, cv)
    am = out["argmin"].cpu().long()
    assert int(am.min()) >= 0 and int(am.max()) < 24
    assert np.array_equal(out["cost_volume"].cpu().numpy(), vol.numpy(), equal_nan=True)
    assert torch.equal(out["missing_mask"].cpu(), miss)
    assert torch.equal(am, idx)
    assert torch.equal(out["lowest_cost"].cpu(), low)
    # both pixels end up all-NaN in the reference: a masked plane is NaN * 0 (Inf * 0) and max-filling spreads it
    assert bool(torch.isnan(vol[0, :, 7, 9]).all()) and bool(torch.isnan(vol[0, :, 9, 14]).all())


def test_pool_fill_needs_its_workspace():
    """MAL_CV_OCC_POOL without the descriptor workspace is an argument error at the C ABI (no silent slow path)."""
    import ctypes as C
    from mal_b200 import _capi
    from tests.emu.emu_lib import emu
    h = emu()
    a = _capi.CostVolumeArgs()
    a.batch, a.channels, a.height, a.width, a.num_lookup, a.num_bins = 1, 16, 8, 8, 1, 4
    buf = torch.zeros(4096)
    for f in ("current", "lookup", "poses", "K", "inv_K", "bins", "cost_volume", "packed", "occ"):
        setattr(a, f, buf.data_ptr())
    a.occ_mode, a.pool_radius = raw.OCC_POOL, 1
    assert h.mal_cost_volume_forward(C.byref(a), None) == 1          # MAL_ERR_ARGUMENT
    assert b"desc workspace" in h.mal_last_error()


@pytest.mark.parametrize("backend", BACKENDS)
def test_dynamicdepth_pool_fill_radius_2(backend):
    """--cv_pool_radius 2 (5x5x5 windows): the generic-radius pre-passes and pool kernel against the oracle's
    max_pool3d (dynamicdepth/networks/resnet_encoder.py:196-202)."""
    h, dev = handle_and_device(backend)
    B, H, W, C, nb = 1, 64, 96, 32, 12
    cv = make_cost_volume_inputs(B, H, W, channels=C, num_lookup=2, num_bins=nb, seed=93, min_bin=0.5, max_bin=6.0,
                                 translation_scale=0.5)
    look_img = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(94))
    look_img[:, :, 16:44, 20:60] = 0.0
    aug = torch.zeros(B, 1, 1, 1)
    want_vol, want_miss = O.match_features_dynamic(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"],
                                                   cv["K"], cv["inv_K"], cv["bins"], look_img, True, aug, False, True,
                                                   2, 0.7)
    occ = (O.occlusion_batch(look_img, H // 4, W // 4)[:, 0] > 0).float()
    out = _run(h, dev, cv, cv_min=True, occ=occ.to(dev), occ_mode=raw.OCC_POOL, pool_radius=2, pool_th=0.7,
               aug_mask=aug.to(dev))
    assert torch.equal(out["missing_mask"].cpu(), want_miss)
    assert torch.equal(out["cost_volume"].cpu(), want_vol)
