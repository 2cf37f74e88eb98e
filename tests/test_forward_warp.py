"""DynamicDepth forward_warp (mal_forward_warp) against the reference's golden output and the
oracle.  Bars: validity mask and splatted depth bit-exact (the z-buffer is an exact max), warped
image bit-exact given the same per-sample matrices; through the public wrapper (matrices derived
on the device) within 1e-5."""
import numpy as np
import pytest
import torch

from mal_b200 import raw, rigid_warp
from mal_b200.utils.synthetic import CITYSCAPES_K, KITTI_K, make_photometric_inputs
from oracle import mal_oracle as O
from tests.backends import BACKENDS, handle_and_device
from tests.helpers import load_npz


@pytest.mark.parametrize("backend", BACKENDS)
def test_forward_warp_against_reference_golden(backend):
    h, dev = handle_and_device(backend)
    g = load_npz("forward_warp.npz")
    t = lambda k: torch.from_numpy(g[k].copy()).to(dev)
    img_w, depth_w, valid = raw.forward_warp(h, img=t("in_img"), depth=t("in_depth"), pose=t("in_pose"), K=t("in_K"),
                                             Ku_inv=t("in_Ku_inv"), K_inv=t("in_K_inv"), proj=t("in_proj"), upscale=3)
    assert np.array_equal(valid.cpu().numpy().astype(np.uint8), g["ref_valid"])
    assert 0.2 < g["ref_valid"].mean() < 0.95
    assert np.array_equal(depth_w.cpu().numpy(), g["ref_depth_w"])
    assert np.array_equal(img_w.cpu().numpy(), g["ref_img_w"])


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape,C,upscale,Kn,ts", [((2, 32, 80), 3, 3, CITYSCAPES_K, 0.3), ((1, 19, 37), 1, 2, KITTI_K, 1.0),
                                                   ((1, 16, 24), 4, 1, KITTI_K, 0.1)])
def test_forward_warp_against_oracle(backend, shape, C, upscale, Kn, ts):
    h, dev = handle_and_device(backend)
    B, H, W = shape
    inputs, t = make_photometric_inputs(B, H, W, seed=41, normalised_K=Kn, translation_scale=ts)
    gen = torch.Generator().manual_seed(42)
    img = torch.rand(B, C, H, W, generator=gen)
    depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    pose = t[("cam_T_cam", 0, 1)][:, :3, :].contiguous()
    K = inputs[("K", 0)][:, :3, :3].contiguous()
    mats = O.forward_warp_matrices(pose, K, upscale)
    want = O.forward_warp(img, depth, pose, K, upscale, matrices=mats)
    d = lambda x: x.to(dev)
    got = raw.forward_warp(h, img=d(img), depth=d(depth), pose=d(pose), K=d(K), Ku_inv=d(mats[0]), K_inv=d(mats[1]),
                           proj=d(mats[2]), upscale=upscale)
    for a, b, name in zip(got, want, ("img_w", "depth_w", "valid")):
        assert torch.equal(a.cpu(), b), name


def test_forward_warp_public_wrapper(op_device):
    """rigid_warp.forward_warp with the reference's signature; the 3x3 / 3x4 constants are derived
    on the operator's device, so trigonometric ulps may move a coordinate: compare with tolerance."""
    dev = op_device
    inputs, t = make_photometric_inputs(2, 32, 80, seed=43, normalised_K=CITYSCAPES_K, translation_scale=0.3)
    img = inputs[("color", 0, 0)]
    depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    pose = t[("cam_T_cam", 0, -1)][:, :3, :].contiguous()
    K = inputs[("K", 0)][:, :3, :3].contiguous()
    want = O.forward_warp(img, depth, pose, K, 3)
    got = rigid_warp.forward_warp(img.to(dev), depth.to(dev), pose.to(dev), K.to(dev), upscale=3, rotation_mode="euler",
                                  padding_mode="zeros")
    mism = (got[2].cpu() != want[2]).float().mean()
    assert float(mism) < 2e-3
    same = (got[2].cpu() == want[2])
    assert float(((got[0].cpu() - want[0]).abs() * same).max()) < 1e-4
    assert float(((got[1].cpu() - want[1]).abs() * same).max()) < 1e-4
    with pytest.raises(ValueError):
        rigid_warp.forward_warp(img.to(dev), depth.to(dev), pose.to(dev), K.to(dev))
