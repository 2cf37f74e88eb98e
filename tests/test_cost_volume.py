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
