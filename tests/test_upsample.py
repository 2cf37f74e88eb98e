"""Bilinear up-sampling of the low-resolution disparities: the stand-alone kernel
against torch's CPU F.interpolate (bit-exact), its adjoint against autograd, and the photometric kernel
reading a low-resolution disparity directly against the same kernel fed the up-sampled plane.

Reference call sites: manydepth/trainer.py:1093-1097, :1176-1177; dualrefine/trainer.py:412-413;
dynamicdepth/trainer.py:915-916.
"""
import pytest
import torch
import torch.nn.functional as F

from mal_b200 import ops, raw
from mal_b200.utils.synthetic import make_photometric_inputs
from tests.backends import BACKENDS, handle_and_device

# (in_h, in_w) -> (out_h, out_w): scales 1..3 of KITTI 192x640 and CityScapes 192x512, one non power of two ratio
SHAPES = [((96, 320), (192, 640)), ((48, 160), (192, 640)), ((24, 80), (192, 640)), ((24, 64), (192, 512)),
          ((50, 100), (192, 640)), ((192, 640), (192, 640))]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", SHAPES)
def test_upsample_forward_is_bit_exact_against_torch_cpu(backend, shape):
    h, dev = handle_and_device(backend)
    (ih, iw), (oh, ow) = shape
    x = torch.rand(2, 1, ih, iw, generator=torch.Generator().manual_seed(ih * 1000 + iw))
    want = F.interpolate(x, [oh, ow], mode="bilinear", align_corners=False)
    got = raw.upsample_bilinear(h, x.to(dev), (oh, ow)).cpu()
    assert torch.equal(got, want)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", SHAPES[:5])
def test_upsample_backward_is_the_adjoint(backend, shape):
    h, dev = handle_and_device(backend)
    (ih, iw), (oh, ow) = shape
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 1, ih, iw, generator=g, requires_grad=True)
    go = torch.randn(2, 1, oh, ow, generator=g)
    (want,) = torch.autograd.grad(F.interpolate(x, [oh, ow], mode="bilinear", align_corners=False), x, go)
    got = raw.upsample_bilinear_backward(h, go.to(dev), (ih, iw)).cpu()
    assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_upsample_op_autograd(op_device):
    dev = op_device
    x = torch.rand(1, 1, 12, 40, generator=torch.Generator().manual_seed(3))   # out H + W > 128: see mal_math.cuh
    xd = x.clone().to(dev).requires_grad_(True)
    y = ops.upsample_bilinear(xd, (48, 160))
    w = torch.randn(1, 1, 48, 160, generator=torch.Generator().manual_seed(4))
    (y * w.to(dev)).sum().backward()
    xr = x.clone().requires_grad_(True)
    (F.interpolate(xr, [48, 160], mode="bilinear", align_corners=False) * w).sum().backward()
    assert torch.equal(y.detach().cpu(), F.interpolate(x, [48, 160], mode="bilinear", align_corners=False))
    assert float((xd.grad.cpu() - xr.grad).abs().max()) <= 1e-5 * float(xr.grad.abs().max())


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("factor", [2, 4, 8])
def test_photo_kernel_reads_low_resolution_disparity(backend, factor):
    """Fused path == the same kernel on F.interpolate(disp): every output bit for bit, and the
    low-resolution gradient == adjoint(full-resolution gradient)."""
    h, dev = handle_and_device(backend)
    B, H, W = 2, 64, 96
    inputs, t = make_photometric_inputs(B, H, W, seed=300 + factor)
    d = lambda x: x.to(dev)
    lo = F.avg_pool2d(t[("mono_disp", 0)], factor)
    lo_b = F.avg_pool2d(t[("multi_disp", 0)], factor)
    tgt, src = d(inputs[("color", 0, 0)]), [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
    common = dict(target=tgt, src=src, K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]),
                  T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])], want_weight=True)
    up = lambda x: F.interpolate(x, [H, W], mode="bilinear", align_corners=False)
    full = raw.photo(h, depth=d(up(lo)), with_grad=True, **common)
    fused = raw.photo(h, depth=d(lo), with_grad=True, **common)
    for k in ("min_reproj", "selection", "weight", "grad_depth", "grad_P", "sums"):
        assert torch.equal(fused[k].cpu(), full[k].cpu()), k
    # the ensemble form: (disp_a + disp_b) / 2 of two low-resolution planes, no gradient
    full = raw.photo(h, depth=d(up(lo)), depth_b=d(up(lo_b)), **common)
    fused = raw.photo(h, depth=d(lo), depth_b=d(lo_b), **common)
    assert torch.equal(fused["min_reproj"].cpu(), full["min_reproj"].cpu())


def test_photo_op_gradient_reaches_the_low_resolution_disparity(op_device):
    dev = op_device
    B, H, W = 1, 32, 64
    inputs, t = make_photometric_inputs(B, H, W, seed=41)
    d = lambda x: x.to(dev)
    lo = F.avg_pool2d(t[("mono_disp", 0)], 4)
    args = dict(K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]),
                T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])])
    tgt, src = d(inputs[("color", 0, 0)]), [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
    lo_a = d(lo).requires_grad_(True)
    sums, _, _ = ops.photo(tgt, src, depth=lo_a, **args)
    sums[2].backward()
    lo_b = d(lo).requires_grad_(True)
    sums_b, _, _ = ops.photo(tgt, src, depth=ops.upsample_bilinear(lo_b, (H, W)), **args)
    sums_b[2].backward()
    assert torch.equal(sums.detach().cpu(), sums_b.detach().cpu())
    assert lo_a.grad.shape == lo.shape
    assert float((lo_a.grad - lo_b.grad).abs().max()) <= 1e-6 * float(lo_b.grad.abs().max())
