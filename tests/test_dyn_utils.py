"""MAL temporal hint (mal_dynamic_instance / mal_fill_dynamic_obj) against the oracle restatement
of manydepth/dyn_utils.py (pinned bitwise against the reference's TorchScript by
oracle/pin_against_reference.py).  Integer / byte work: everything must be bit-exact."""
import pytest
import torch

from mal_b200 import dyn_utils, raw
from mal_b200.utils.synthetic import make_instance_masks, make_photometric_inputs
from oracle import mal_oracle as O
from tests.backends import BACKENDS, handle_and_device

CASES = [(6, 48, 96, 1, None, False), (5, 40, 64, 2, 1, True), (17, 32, 48, 3, None, False),
         (3, 24, 40, 4, 0, False), (1, 16, 16, 5, None, False), (40, 33, 57, 6, 7, True)]


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("N,H,W,seed,empty,replace", CASES)
def test_dynamic_instance_bit_exact(backend, N, H, W, seed, empty, replace):
    h, dev = handle_and_device(backend)
    ml, mn = make_instance_masks(N, H, W, seed=seed, max_shift=9, empty=empty)
    g = torch.Generator().manual_seed(seed)
    il, inx = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
    want_l, want_n, deltas = O.generate_dynamic_instance(ml, mn, il, inx, replace)
    got_l, got_n, d = raw.dynamic_instance(h, mask_last=ml.to(dev), mask_next=mn.to(dev), img_last=il.to(dev),
                                           img_next=inx.to(dev), replace=replace)
    assert torch.equal(d.cpu().long(), torch.stack(deltas, 0))
    assert torch.equal(got_l.cpu(), want_l) and torch.equal(got_n.cpu(), want_n)
    assert not torch.equal(want_l, il)          # the case moves something
    dx, dy = torch.randint(-5, 6, (N,), generator=g), torch.randint(-W, W + 1, (N,), generator=g)
    got = raw.fill_dynamic_obj(h, mask=ml.to(dev), delta_x=dx.to(dev), delta_y=dy.to(dev), source=il.to(dev),
                               img=inx.to(dev))
    assert torch.equal(got.cpu(), O.fill_dynamic_obj(ml, dx, dy, il, inx))


class _Instances:
    """Minimal detectron2-Instances look-alike: .scores, .pred_masks, len(), boolean indexing."""

    def __init__(self, scores, masks):
        self.scores, self.pred_masks = scores, masks

    def __len__(self):
        return len(self.scores)

    def __getitem__(self, idx):
        return _Instances(self.scores[idx], self.pred_masks[idx])


def test_image_synthesis_orchestration(op_device):
    """dyn_utils.image_synthesis with a synthetic segmenter / matcher (Mask2Former-shaped output):
    samples without confident instances keep their warped images, the others get the composition."""
    dev = op_device
    B, H, W = 3, 32, 64
    inputs, t = make_photometric_inputs(B, H, W, seed=3)
    inputs = {k: v.to(dev) for k, v in inputs.items()}
    outputs = {("color", -1, 0): inputs[("color", -1, 0)].clone(), ("color", 1, 0): inputs[("color", 1, 0)].clone()}
    ml, mn = make_instance_masks(4, H, W, seed=8, max_shift=6)
    ml, mn = ml.to(dev), mn.to(dev)

    def ins_model(batch):
        out = []
        for i, item in enumerate(batch):
            if len(batch) == B:          # segmentation of the target frames: sample 0 has nothing confident
                scores = torch.tensor([0.2, 0.1] if i != 1 else [0.95, 0.97], device=dev)
                out.append({"instances": _Instances(scores, ml[:2])})
            else:                        # (last, next) pair of one sample
                out.append({"instances": _Instances(torch.ones(4, device=dev), ml if i == 0 else mn)})
        return out

    matcher = lambda a, b, cur: (torch.arange(4, device=dev), torch.arange(4, device=dev))
    has = dyn_utils.image_synthesis(inputs, outputs, 0, 0.9, ins_model, matcher)
    assert has
    for i in (0, 2):
        assert torch.equal(outputs[("syn", -1, 0)][i], outputs[("color", -1, 0)][i])
    want_l, want_n, _ = O.generate_dynamic_instance(ml.cpu(), mn.cpu(), outputs[("color", -1, 0)][1].cpu(),
                                                    outputs[("color", 1, 0)][1].cpu(), False)
    assert torch.equal(outputs[("syn", -1, 0)][1].cpu(), want_l)
    assert torch.equal(outputs[("syn", 1, 0)][1].cpu(), want_n)
    none = dyn_utils.image_synthesis(inputs, {k: v for k, v in outputs.items() if k[0] == "color"}, 0, 0.99,
                                     ins_model, matcher)
    assert none is False


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("N,H,W,seed,empty,replace", CASES[:4])
def test_dynamic_instance_backward(backend, N, H, W, seed, empty, replace):
    """The composition is copies and selections: its backward is exact (autograd on the oracle)."""
    h, dev = handle_and_device(backend)
    ml, mn = make_instance_masks(N, H, W, seed=seed, max_shift=9, empty=empty)
    g = torch.Generator().manual_seed(seed)
    il = torch.rand(3, H, W, generator=g).requires_grad_(True)
    inx = torch.rand(3, H, W, generator=g).requires_grad_(True)
    gl, gn = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
    want_l, want_n, deltas = O.generate_dynamic_instance(ml, mn, il, inx, replace)
    want = torch.autograd.grad((want_l * gl).sum() + (want_n * gn).sum(), [il, inx])
    got = raw.dynamic_instance_backward(h, mask_last=ml.to(dev), mask_next=mn.to(dev),
                                        deltas=torch.stack(deltas, 0).to(dev), grad_ori_last=gl.to(dev),
                                        grad_ori_next=gn.to(dev))
    for a, b in zip(got, want):
        assert torch.allclose(a.cpu(), b, rtol=1e-6, atol=1e-7)


def test_temporal_hint_gradient_reaches_the_warped_images(op_device):
    """Classic path: warped images -> image synthesis -> compute_mono_losses (PRED mode, 4 candidates):
    d loss / d warped image matches autograd on the oracle, including pixels whose min is a syn candidate."""
    from mal_b200 import loss_utils, layers
    dev = op_device
    B, H, W = 1, 32, 64
    inputs, t = make_photometric_inputs(B, H, W, seed=13, translation_scale=0.3)
    ml, mn = make_instance_masks(5, H, W, seed=14, max_shift=6)
    # moving objects: painted at their "last" position in the warped frame -1, at their "next"
    # position in frame +1 and half-way in the target, so the synthesised candidates win there
    base = {f: (inputs[("color", 0, 0)] + 0.05 * (inputs[("color", f, 0)] - 0.5)).clamp(0, 1) for f in (-1, 1)}
    inputs[("color", 0, 0)] = inputs[("color", 0, 0)].clone()
    _, _, d = O.generate_dynamic_instance(ml, mn, base[-1][0], base[1][0], False)
    gen = torch.Generator().manual_seed(15)
    for n in range(ml.shape[0]):
        col = torch.rand(3, 1, generator=gen)
        base[-1][0][:, ml[n]] = col
        base[1][0][:, mn[n]] = col
        mid = torch.roll(ml[n], shifts=(int(d[0][n]), int(d[1][n])), dims=(0, 1))
        inputs[("color", 0, 0)][0][:, mid] = col
    res = []
    for mode, dv in (("oracle", torch.device("cpu")), ("ours", dev)):
        warped = {f: base[f].clone().to(dv).requires_grad_(True) for f in (-1, 1)}
        inp = {k: v.to(dv) for k, v in inputs.items()}
        if mode == "oracle":
            sl, sn, _ = O.generate_dynamic_instance(ml, mn, warped[-1][0], warped[1][0], False)
            out = {("color", -1, 0): warped[-1], ("color", 1, 0): warped[1], ("syn", -1, 0): sl.unsqueeze(0),
                   ("syn", 1, 0): sn.unsqueeze(0), ("disp", 0): t[("mono_disp", 0)]}
            losses, mono_reproj, aux = O.mono_losses(inp, out, True, True, noise=t["noise"][0])
            frame_idx = aux["frame_idx"]
        else:
            sl, sn = dyn_utils.generate_dynamic_instance(None, None, ml.to(dv), mn.to(dv), warped[-1][0], warped[1][0], False)
            out = {("color", -1, 0): warped[-1], ("color", 1, 0): warped[1], ("syn", -1, 0): sl.unsqueeze(0),
                   ("syn", 1, 0): sn.unsqueeze(0), ("disp", 0): t[("mono_disp", 0)].to(dv)}
            losses, mono_reproj = loss_utils.compute_mono_losses(layers.SSIM(), inp, out, True, True,
                                                                 noise=t["noise"][0].to(dv))
            frame_idx = out[("mal_selection", 0)] & 0x7F
        grads = torch.autograd.grad(losses["reproj_loss/0"], [warped[-1], warped[1]])
        res.append((float(losses["reproj_loss/0"].detach()), frame_idx.cpu(), [g.cpu() for g in grads]))
    (l0, idx0, g0), (l1, idx1, g1) = res
    assert abs(l0 - l1) <= 1e-5 * abs(l0)
    assert torch.equal(idx0.to(torch.uint8), idx1.to(torch.uint8))
    assert float((idx0 >= 2).float().mean()) > 0.02          # some pixels do select a temporal-hint candidate
    for a, b in zip(g1, g0):
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max())


def test_fused_warp_path_returns_the_temporal_hint_gradient(op_device):
    """Fused WARP mode (no materialised warps inside the loss kernel): the gradient with respect to the
    temporal-hint candidates equals autograd on the oracle, so autograd can carry it into the warped images
    image_synthesis copied them from - the reference's behaviour (manydepth/loss_utils.py:84-88)."""
    from mal_b200 import ops
    dev = op_device
    B, H, W = 1, 32, 64
    inputs, t = make_photometric_inputs(B, H, W, seed=21, translation_scale=0.3)
    # candidates that win on a good share of the pixels: the target, slightly perturbed
    gen = torch.Generator().manual_seed(22)
    syn = {f: (inputs[("color", 0, 0)] + 0.02 * torch.randn(B, 3, H, W, generator=gen)).clamp(0, 1) for f in (-1, 1)}
    # oracle: syn as leaves
    leaves = {f: syn[f].clone().requires_grad_(True) for f in (-1, 1)}
    disp = t[("mono_disp", 0)].clone().requires_grad_(True)
    out = {("disp", 0): disp, ("syn", -1, 0): leaves[-1], ("syn", 1, 0): leaves[1]}
    for f in (-1, 1):
        out[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
    O.images_pred(inputs, out, height=H, width=W)
    losses, _, aux = O.mono_losses(inputs, out, True, True, noise=t["noise"][0])
    want = torch.autograd.grad(losses["reproj_loss/0"], [leaves[-1], leaves[1], disp])
    assert float((aux["frame_idx"] >= 2).float().mean()) > 0.1
    # ours: one fused call
    d = lambda x: x.to(dev)
    tgt, src = d(inputs[("color", 0, 0)]), [d(inputs[("color", -1, 0)]), d(inputs[("color", 1, 0)])]
    ident = ops.photo(tgt, src, mode=raw.PHOTO_PRED)[1]
    syn_d = [d(syn[f]).requires_grad_(True) for f in (-1, 1)]
    disp_d = d(t[("mono_disp", 0)]).requires_grad_(True)
    sums, _, sel = ops.photo(tgt, src, syn=syn_d, depth=disp_d, K=d(inputs[("K", 0)]), inv_K=d(inputs[("inv_K", 0)]),
                             T=[d(t[("cam_T_cam", 0, -1)]), d(t[("cam_T_cam", 0, 1)])], identity_min=ident.detach(),
                             noise=d(t["noise"][0]))
    assert abs(float(sums[2]) - float(losses["reproj_loss/0"])) <= 1e-5 * abs(float(losses["reproj_loss/0"]))
    assert torch.equal((sel.cpu() & 0x7F)[:, 0].long(), aux["frame_idx"].reshape(B, H, W).long())
    got = torch.autograd.grad(sums[2], [syn_d[0], syn_d[1], disp_d])
    for name, a, b in zip(("syn -1", "syn +1", "disp"), got, want):
        assert float((a.cpu() - b).abs().max()) <= 1e-4 * float(b.abs().max()), name
