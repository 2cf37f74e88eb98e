"""MAL's temporal hint inside the step (csrc/temporal.cu): warped-image materialisation, packed-mask
synthesis for a whole batch, and the backward from d loss / d syn into the warped images, the disparity and the
poses - against the oracle (manydepth/trainer.py:1078-1165, dyn_utils.py:38-170 restated in oracle/mal_oracle.py)
and its autograd.  Bars: images bit-exact; gradients 1e-4 of their max."""
import numpy as np
import pytest
import torch

from mal_b200 import raw
from mal_b200.utils.synthetic import make_instance_masks, make_photometric_inputs
from oracle import mal_oracle as O
from tests.backends import BACKENDS, handle_and_device

GRAD_RTOL = 1e-4


def _gerr(a, b):
    s = float(b.abs().max())
    return float((a - b).abs().max()) / (s if s > 0 else 1.0)


def _case(B, H, W, nmax, seed):
    inputs, t = make_photometric_inputs(B, H, W, seed=seed, translation_scale=0.3)
    counts = torch.tensor([(nmax - 2 * b) % (nmax + 1) for b in range(B)], dtype=torch.int32)   # includes 0 and nmax
    counts[0] = nmax
    ml = torch.zeros(B, nmax, H, W, dtype=torch.bool)
    mn = torch.zeros_like(ml)
    for b in range(B):
        n = int(counts[b])
        if n:
            last, nxt = make_instance_masks(n, H, W, seed=seed + 7 * b, max_shift=6, empty=(1 if n > 2 else None))
            ml[b, :n], mn[b, :n] = last, nxt
    return inputs, t, ml, mn, counts


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("shape", [(3, 32, 64, 5), (2, 24, 40, 32)])
def test_temporal_pipeline_matches_oracle(backend, shape):
    h, dev = handle_and_device(backend)
    B, H, W, nmax = shape
    inputs, t, ml, mn, counts = _case(B, H, W, nmax, 61)
    src = [inputs[("color", f, 0)] for f in (-1, 1)]
    Ts = [t[("cam_T_cam", 0, f)] for f in (-1, 1)]
    disp = t[("mono_disp", 0)]

    # ---- oracle: generate_images_pred -> generate_dynamic_instance per sample -> a loss on syn -> autograd
    disp_o = disp.clone().requires_grad_(True)
    T_o = [x.clone().requires_grad_(True) for x in Ts]
    outs = {("disp", 0): disp_o, ("cam_T_cam", 0, -1): T_o[0], ("cam_T_cam", 0, 1): T_o[1]}
    O.images_pred(inputs, outs, height=H, width=W)
    warped_o = [outs[("color", -1, 0)], outs[("color", 1, 0)]]
    syn_o = [w.clone() for w in warped_o]
    deltas_o = torch.zeros(B, 2, 32, dtype=torch.int32)
    for b in range(B):
        n = int(counts[b])
        if n == 0:
            continue
        ol, on, (dxl, dyl, _, _) = O.generate_dynamic_instance(ml[b, :n], mn[b, :n], warped_o[0][b], warped_o[1][b])
        syn_o[0] = torch.cat([syn_o[0][:b], ol[None], syn_o[0][b + 1:]])
        syn_o[1] = torch.cat([syn_o[1][:b], on[None], syn_o[1][b + 1:]])
        deltas_o[b, 0, :n], deltas_o[b, 1, :n] = dxl.int(), dyl.int()
    gen = torch.Generator().manual_seed(5)
    g_syn = [torch.randn(B, 3, H, W, generator=gen) for _ in range(2)]
    loss = (syn_o[0] * g_syn[0]).sum() + (syn_o[1] * g_syn[1]).sum()
    want_gw = torch.autograd.grad(loss, warped_o, retain_graph=True)
    want_gd, want_gT0, want_gT1 = torch.autograd.grad(loss, [disp_o, T_o[0], T_o[1]])

    # ---- kernels
    d = lambda x: x.to(dev)
    geom = dict(src=[d(s) for s in src], depth=d(disp), K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]),
                T=[d(x) for x in Ts])
    warped = raw.temporal_warp(h, **geom)
    for a, b_ in zip(warped, warped_o):
        assert torch.equal(a.cpu(), b_.detach())
    pl, pn = raw.temporal_pack_masks(h, masks_last=d(ml), masks_next=d(mn), counts=d(counts))
    bits = (ml.long() << torch.arange(nmax).view(1, -1, 1, 1)).sum(1)
    assert torch.equal(pl.cpu().long() & 0xFFFFFFFF, bits & 0xFFFFFFFF)
    syn = raw.temporal_synthesis(h, warped=warped, packed_last=pl, packed_next=pn, counts=d(counts))
    assert torch.equal(syn["deltas"].cpu(), deltas_o)
    for a, b_ in zip(syn["syn"], syn_o):
        assert torch.equal(a.cpu(), b_.detach())
    assert not torch.equal(syn["syn"][0].cpu(), warped_o[0].detach())   # the synthesis really changed something

    grad_depth = torch.zeros(B, 1, H, W, device=dev)
    grad_P = torch.zeros(B, 2, 12, device=dev)
    back = raw.temporal_backward(h, grad_syn=[d(g) for g in g_syn], packed_last=pl, packed_next=pn, counts=d(counts),
                                 deltas=syn["deltas"], grad_depth=grad_depth, grad_P=grad_P, **geom)
    for a, b_ in zip(back["grad_warped"], want_gw):
        assert _gerr(a.cpu(), b_) < 1e-6
    assert _gerr(grad_depth.cpu(), want_gd) < GRAD_RTOL
    # d/dT = K[:3,:]^T @ dP (mal_step_combine does this product on the device)
    K = inputs[("K", 0)]
    gT = torch.einsum("bkr,bfkc->bfrc", K[:, :3, :], grad_P.cpu().view(B, 2, 3, 4))
    assert _gerr(gT[:, 0], want_gT0) < GRAD_RTOL and _gerr(gT[:, 1], want_gT1) < GRAD_RTOL
