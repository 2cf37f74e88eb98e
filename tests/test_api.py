"""The reference-shaped Python surface (mal_b200.layers / loss_utils / trainer_ops / matching)
against the oracle: same dict keys in, same losses / indices / gradients out."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mal_b200 import layers, loss_utils, matching, ops, trainer_ops
from mal_b200.utils.synthetic import make_cost_volume_inputs, make_photometric_inputs, to_device
from oracle import mal_oracle as O

LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4


def _gerr(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    s = float(b.abs().max())
    return float((a - b).abs().max()) / (s if s > 0 else 1.0)


def _close(a, b, rtol=LOSS_RTOL):
    a, b = float(a), float(b)
    return abs(a - b) <= rtol * abs(b)


def _opt(B, H, W, **kw):
    base = dict(height=H, width=W, batch_size=B, min_depth=0.1, max_depth=100.0, frame_ids=[0, -1, 1], sclm=0,
                temporal=True, main_temporal=False, distil=True, no_ens=False, loss_blc=True, dual_distil=False)
    base.update(kw)
    return SimpleNamespace(**base)


def _leaves(t, dev):
    leaves = {"mono_disp": t[("mono_disp", 0)].clone().to(dev).requires_grad_(True),
              "multi_disp": t[("multi_disp", 0)].clone().to(dev).requires_grad_(True)}
    for f in (-1, 1):
        leaves[f"T{f}"] = t[("cam_T_cam", 0, f)].clone().to(dev).requires_grad_(True)
    return leaves


def _step_dicts(inputs, t, leaves, dev, lowest_cost):
    inputs_d = to_device(inputs, dev)
    mono = {("disp", 0): leaves["mono_disp"]}
    multi = {("disp", 0): leaves["multi_disp"], "consistency_mask": t["consistency_mask"].to(dev),
             "augmentation_mask": t["augmentation_mask"].to(dev), "lowest_cost": lowest_cost.to(dev)}
    for f in (-1, 1):
        mono[("cam_T_cam", 0, f)] = multi[("cam_T_cam", 0, f)] = leaves[f"T{f}"]
        mono[("syn", f, 0)] = multi[("syn", f, 0)] = t[("syn", f, 0)].to(dev)
    return inputs_d, mono, multi


def _oracle_step(inputs, t, lowest_cost, opt, has_ins, multi_has_ins):
    """process_batch's loss half through the oracle (manydepth/trainer.py:574-644)."""
    H, W, B = opt.height, opt.width, opt.batch_size
    L = _leaves(t, "cpu")
    _, mono, multi = _step_dicts(inputs, t, L, "cpu", lowest_cost)
    O.images_pred(inputs, mono, height=H, width=W)
    mono_losses, mono_reproj, _ = O.mono_losses(inputs, mono, opt.temporal, has_ins, noise=t["noise"][0])
    multi[("mono_depth", 0, 0)] = mono[("depth", 0, 0)]
    multi["consistency_mask"] = multi["consistency_mask"] * O.matching_mask(multi)
    ens = None
    if not opt.no_ens:
        ens = O.images_pred_ensemble(inputs, L["T-1"].detach(), L["T1"].detach(),
                                     (L["mono_disp"].detach() + L["multi_disp"].detach()) / 2.0, height=H, width=W)
    O.images_pred(inputs, multi, height=H, width=W, is_multi=True)
    losses, _, loss_list, aux = O.main_losses(inputs, multi, mono_reproj, ens, batch_size=B,
                                              multi_has_ins=multi_has_ins, dual_distil=opt.dual_distil,
                                              loss_blc=opt.loss_blc, noise=t["noise"][1])
    for k, v in mono_losses.items():
        losses[k] = losses[k] + v
    if opt.loss_blc:
        loss_list[0] = loss_list[0] + mono_losses["loss"]
        total = 0.5 * loss_list[0] + 0.5 * loss_list[1]
    else:
        total = losses["loss"]
    grads = torch.autograd.grad(total, [L["mono_disp"], L["multi_disp"], L["T-1"], L["T1"]])
    return losses, total, grads, aux, multi


@pytest.mark.parametrize("no_ens,has_ins,multi_has_ins,loss_blc,dual", [
    (False, True, False, True, False),     # --temporal --distil --loss_blc (the README command)
    (True, False, True, False, True),      # no ensemble, main_temporal, dual distillation
])
def test_mal_step_matches_oracle(op_device, no_ens, has_ins, multi_has_ins, loss_blc, dual):
    dev = op_device
    B, H, W = 2, 32, 64
    inputs, t = make_photometric_inputs(B, H, W, seed=77)
    opt = _opt(B, H, W, no_ens=no_ens, loss_blc=loss_blc, dual_distil=dual, main_temporal=multi_has_ins)
    mono_depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    lowest = 1 / (mono_depth[:, 0] * (1.0 + 0.8 * torch.tanh(t["noise"][0][:, 0])))   # ratio in (0.2, 1.8): mixed mask
    want, want_total, want_g, aux, o_multi = _oracle_step(inputs, t, lowest, opt, has_ins, multi_has_ins)

    L = _leaves(t, dev)
    inputs_d, mono, multi = _step_dicts(inputs, t, L, dev, lowest)
    blc = loss_utils.LossBalancing(2, 100, B) if loss_blc else None
    outputs, losses = trainer_ops.process_batch_losses(
        inputs_d, mono, multi, opt, has_ins=has_ins, multi_has_ins=multi_has_ins, loss_blc=blc,
        noises=[n.to(dev) for n in t["noise"]])
    for k in ("reproj_loss/0", "consistency_loss/0", "distil_loss", "loss/0"):
        assert _close(losses[k], want[k]), k
    assert torch.equal(outputs["consistency_mask"].cpu(), o_multi["consistency_mask"])
    assert np.array_equal(outputs["mal_distil_index"].cpu().numpy(), aux["distil_idx"].numpy().astype(np.uint8))
    assert np.array_equal(outputs[("mal_selection", 0)].cpu().numpy() & 0x7F, aux["frame_idx"].numpy().astype(np.uint8))
    assert torch.equal(outputs["consistency_target/0"].cpu(), aux["consistency_target"])
    if loss_blc:
        # LossBalancing.compute_loss returns batch_size x the weighted sum (loss_utils.py:305-318)
        assert _close(losses["loss"], B * float(want_total))
        total = losses["loss"] / B
    else:
        total = losses["loss"]
        assert _close(total, want_total)
    g = torch.autograd.grad(total, [L["mono_disp"], L["multi_disp"], L["T-1"], L["T1"]])
    for a, b, name in zip(g, want_g, ("mono_disp", "multi_disp", "T-1", "T1")):
        assert _gerr(a, b) < GRAD_RTOL, name


def test_compute_losses_multiscale_matches_oracle(op_device):
    """Trainer.compute_losses (non-distil) over 3 scales, teacher and student passes."""
    dev = op_device
    B, H, W, S = 1, 32, 64, 3
    inputs, t = make_photometric_inputs(B, H, W, num_scales=S, seed=88)
    opt = _opt(B, H, W, sclm=S - 1, distil=False, temporal=True)
    for is_multi in (False, True):
        name = "multi" if is_multi else "mono"
        disps_c = [t[(name + "_disp", s)].clone().requires_grad_(True) for s in range(S)]
        Tc = {f: t[("cam_T_cam", 0, f)].clone().requires_grad_(True) for f in (-1, 1)}
        o = {("disp", s): disps_c[s] for s in range(S)}
        o.update({("cam_T_cam", 0, f): Tc[f] for f in (-1, 1)})
        o.update({"consistency_mask": t["consistency_mask"], "augmentation_mask": t["augmentation_mask"]})
        for s in range(S):
            for f in (-1, 1):
                o[("syn", f, s)] = t[("syn", f, 0)]
        O.images_pred(inputs, o, num_scales=S, height=H, width=W, is_multi=is_multi)
        if is_multi:
            for s in range(S):
                o[("mono_depth", 0, s)] = O.disp_to_depth(F.interpolate(t[("mono_disp", s)], [H, W], mode="bilinear",
                                                                         align_corners=False), 0.1, 100.0)[1]
        want, aux = O.trainer_compute_losses(inputs, o, num_scales=S, is_multi=is_multi, temporal=True, has_ins=True,
                                             batch_size=B, noises=t["noise"][:S] + t["noise"][:1])
        leaves = disps_c + ([] if is_multi else [Tc[-1], Tc[1]])
        want_g = torch.autograd.grad(want["loss"], leaves)

        inputs_d = to_device(inputs, dev)
        disps = [t[(name + "_disp", s)].clone().to(dev).requires_grad_(True) for s in range(S)]
        Td = {f: t[("cam_T_cam", 0, f)].clone().to(dev).requires_grad_(True) for f in (-1, 1)}
        od = {("disp", s): disps[s] for s in range(S)}
        od.update({("cam_T_cam", 0, f): Td[f] for f in (-1, 1)})
        od.update({"consistency_mask": t["consistency_mask"].to(dev), "augmentation_mask": t["augmentation_mask"].to(dev)})
        for s in range(S):
            for f in (-1, 1):
                od[("syn", f, s)] = t[("syn", f, 0)].to(dev)
            if is_multi:
                od[("mono_depth", 0, s)] = o[("mono_depth", 0, s)].to(dev)
        trainer_ops.generate_images_pred(inputs_d, od, opt, is_multi=is_multi)
        noises = [n.to(dev) for n in (t["noise"][:S] + t["noise"][:1])]
        got, _ = trainer_ops.compute_losses(inputs_d, od, opt, is_multi=is_multi, has_ins=True, noises=noises)
        for k in want:
            assert _close(got[k], want[k]), (name, k)
        for s in range(S):
            assert np.array_equal(od[("mal_selection", s)].cpu().numpy() & 0x7F,
                                  aux[("frame_idx", s)].numpy().astype(np.uint8))
        g = torch.autograd.grad(got["loss"], disps + ([] if is_multi else [Td[-1], Td[1]]))
        for a, b in zip(g, want_g):
            assert _gerr(a, b) < GRAD_RTOL, name


@pytest.mark.parametrize("convention", [layers.CONV_MANYDEPTH, layers.CONV_DUALREFINE])
def test_layer_classes_forward_backward(op_device, convention):
    """BackprojectDepth -> Project3D -> grid_sample -> SSIM / compute_reprojection_loss, used one by
    one like the reference trainers do, against the oracle and its autograd."""
    dev = op_device
    B, H, W = 2, 24, 40
    inputs, t = make_photometric_inputs(B, H, W, seed=99)
    K, iK = inputs[("K", 0)], inputs[("inv_K", 0)]
    src, tgt = inputs[("color", -1, 0)], inputs[("color", 0, 0)]

    def chain(mod, dev_, depth, T, Kt):
        if mod is O:
            cam = O.backproject(depth, iK)
            grid, z = O.project3d(cam, Kt, T, H, W, convention, return_z=True)
            warped = O.warp(src, grid, convention)
            return cam, grid, z, O.ssim(warped, tgt), O.reprojection_loss(warped, tgt)
        back = layers.BackprojectDepth(B, H, W)
        proj = layers.Project3D(B, H, W, dc=True, convention=convention)
        cam = back(depth, iK.to(dev_))
        grid, z = proj(cam, Kt, T)
        warped = F.grid_sample(src.to(dev_), grid, padding_mode="border", align_corners=convention == layers.CONV_MANYDEPTH)
        return cam, grid, z, layers.SSIM()(warped, tgt.to(dev_)), loss_utils.compute_reprojection_loss(
            layers.SSIM(), warped, tgt.to(dev_))

    depth0 = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    outs, grads = [], []
    for mod, dv in ((O, "cpu"), (layers, dev)):
        depth = depth0.clone().to(dv).requires_grad_(True)
        T = t[("cam_T_cam", 0, -1)].clone().to(dv).requires_grad_(True)
        Kt = K.clone().to(dv).requires_grad_(True)
        cam, grid, z, s, r = chain(mod, dv, depth, T, Kt)
        total = (s * torch.linspace(0.5, 1.5, W, device=dv)).sum() + (r ** 2).sum() * 3 + z.mean()
        outs.append((cam, grid, z, s, r))
        grads.append(torch.autograd.grad(total, [depth, T, Kt]))
    for a, b, name in zip(outs[1], outs[0], ("cam", "grid", "z", "ssim", "reproj")):
        assert torch.equal(a.detach().cpu(), b.detach()), name      # forward is bit-exact
    for a, b, name in zip(grads[1], grads[0], ("depth", "T", "K")):
        assert _gerr(a, b) < GRAD_RTOL, name


def test_ssim_gradient_both_sides(op_device):
    dev = op_device
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 17, 23, generator=gen)
    y = (x + 0.2 * torch.rand(2, 3, 17, 23, generator=gen)).clamp(0, 1)
    w = torch.rand(2, 3, 17, 23, generator=gen)
    res = []
    for fn, dv in ((O.ssim, "cpu"), (layers.SSIM(), dev)):
        a, b = x.clone().to(dv).requires_grad_(True), y.clone().to(dv).requires_grad_(True)
        out = fn(a, b)
        res.append((out, torch.autograd.grad((out * w.to(dv)).sum(), [a, b])))
    assert torch.equal(res[1][0].detach().cpu(), res[0][0].detach())
    for a, b in zip(res[1][1], res[0][1]):
        assert _gerr(a, b) < GRAD_RTOL


def test_smooth_and_masks(op_device):
    dev = op_device
    inputs, t = make_photometric_inputs(2, 16, 40, seed=3)
    disp = t[("mono_disp", 0)].clone().to(dev).requires_grad_(True)
    got = layers.get_smooth_loss(disp, inputs[("color", 0, 0)].to(dev))
    want_d = t[("mono_disp", 0)].clone().requires_grad_(True)
    want = O.smooth_loss(want_d, inputs[("color", 0, 0)])
    assert _close(got, want)
    assert _gerr(torch.autograd.grad(got, disp)[0], torch.autograd.grad(want, want_d)[0]) < GRAD_RTOL
    a, b = torch.rand(2, 1, 8, 8), torch.rand(2, 1, 8, 8)
    b[0, 0, 0] = a[0, 0, 0]   # ties -> index 0 -> mask 1
    assert torch.equal(loss_utils.compute_loss_masks(a, b), O.loss_masks(a, b))
    assert torch.equal(loss_utils.compute_loss_masks(a, None), torch.ones_like(a))


def test_matcher_methods(op_device):
    dev = op_device
    cv = make_cost_volume_inputs(2, 48, 64, channels=16, num_lookup=1, num_bins=24, seed=17, max_bin=10.0)
    m = matching.CostVolumeMatcher(num_depth_bins=24, min_depth_bin=0.1, max_depth_bin=10.0)
    assert torch.equal(m.depth_bins, cv["bins"])
    d = to_device(cv, dev)
    vol, miss = m.match_features(d["current_feats"], d["lookup_feats"], d["relative_poses"], d["K"], d["inv_K"])
    want_vol, want_miss = O.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"],
                                           cv["inv_K"], cv["bins"])
    assert torch.equal(vol.cpu(), want_vol) and torch.equal(miss.cpu(), want_miss)
    conf = m.compute_confidence_mask(vol * (1 - miss))
    assert torch.equal(conf.cpu(), O.confidence_mask(want_vol * (1 - want_miss)))
    cvm, low, conf2 = m.matching_head(d["current_feats"], d["lookup_feats"], d["relative_poses"], d["K"], d["inv_K"])
    want_cvm, want_low, want_conf, want_idx, _ = O.cost_volume_head(
        cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"], cv["inv_K"], cv["bins"])
    assert torch.equal(cvm.cpu(), want_cvm) and torch.equal(low.cpu(), want_low) and torch.equal(conf2.cpu(), want_conf)
    assert torch.equal(m.indices_to_disparity(want_idx.to(dev)).cpu(), want_low)
    for binning in ("inverse", "log"):
        mm = matching.CostVolumeMatcher(num_depth_bins=24, depth_binning=binning)
        assert torch.equal(mm.compute_depth_bins(0.3, 7.0),
                           O.depth_bins(0.3, 7.0, 24, binning))


def test_operators_refuse_cpu_tensors():
    """No CPU fallback: without the test hook, a CPU tensor is an error, not a slow path."""
    inputs, t = make_photometric_inputs(1, 16, 32, seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA tensors only"):
        ops.smooth(t[("mono_disp", 0)], inputs[("color", 0, 0)])


def test_dynamicdepth_matcher(op_device):
    """matching.DynamicCostVolumeMatcher.match_features with the reference's DynamicDepth signature."""
    dev = op_device
    B, H, W = 1, 64, 96
    cv = make_cost_volume_inputs(B, H, W, channels=16, num_lookup=1, num_bins=12, seed=19, min_bin=0.5, max_bin=6.0,
                                 translation_scale=0.5)
    look_img = torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(3))
    look_img[:, :, 16:48, 20:70] = 0.0
    aug = torch.zeros(B, 1, 1, 1)
    m = matching.DynamicCostVolumeMatcher(num_depth_bins=12, min_depth_bin=0.5, max_depth_bin=6.0)
    d = to_device(cv, dev)
    for cv_min, set_1, pool in ((True, False, True), (False, True, False)):
        want = O.match_features_dynamic(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"],
                                        cv["inv_K"], cv["bins"], look_img, cv_min, aug, set_1, pool, 1, 0.7)
        got = m.match_features(d["current_feats"], d["lookup_feats"], d["relative_poses"], d["K"], d["inv_K"],
                               look_img.to(dev), cv_min, aug.to(dev), set_1, pool, 1, 0.7)
        assert torch.equal(got[0].cpu(), want[0]) and torch.equal(got[1].cpu(), want[1])


@pytest.mark.parametrize("avg", [False, True])
def test_dualrefine_losses_match_oracle(op_device, avg):
    """BASELINE config 4: DualRefine's per-(scale, deq_iter) losses with half-pixel sampling
    (dualrefine/trainer.py:395-455, :530-626) - selections bit-exact, losses and gradients; with
    opt.avg_reprojection the mean instead of the min over the two frames (:575-586)."""
    dev = op_device
    B, H, W = 1, 32, 64
    scales = [0, 1, 2, 3]
    inputs, t = make_photometric_inputs(B, H, W, num_scales=4, seed=61, translation_scale=0.3)
    opt = SimpleNamespace(height=H, width=W, scales=scales, n_losses=1, min_depth=0.1, max_depth=100.0,
                          disparity_smoothness=1e-3, avg_reprojection=avg)
    keys = [(s, it) for s in scales if s != 1 for it in range(2 if s in (0, 1, 2) else 1)]
    noises = [torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(70 + i)) for i in range(len(keys))]
    cmask = t["consistency_mask"].unsqueeze(1)

    def build(dv):
        d = lambda x: x.to(dv)
        leaves = {(s, it): t[("mono_disp" if it == 0 else "multi_disp", s)].clone().to(dv).requires_grad_(True)
                  for s, it in keys}
        Ts = {k: t[("cam_T_cam", 0, f)].clone().to(dv).requires_grad_(True) for k, f in
              (((0, -1), -1), ((0, 1), 1), ((0, -1, 1), -1))}
        o = {("disp", s, it): leaves[(s, it)] for s, it in keys}
        o[("cam_T_cam", 0, -1)], o[("cam_T_cam", 0, 1)], o[("cam_T_cam", 0, -1, 1)] = Ts[(0, -1)], Ts[(0, 1)], Ts[(0, -1, 1)]
        o["consistency_mask"] = d(cmask)
        return to_device(inputs, dv), o, list(leaves.values()) + list(Ts.values())

    inp_c, o_c, leaves_c = build("cpu")
    O.dualrefine_images_pred(inp_c, o_c, scales, 1, H, W)
    want, aux = O.dualrefine_compute_losses(inp_c, o_c, scales, 1, noises=noises, avg_reprojection=avg)
    want_g = torch.autograd.grad(want["loss"], leaves_c)

    inp_d, o_d, leaves_d = build(dev)
    trainer_ops.generate_images_pred_dualrefine(inp_d, o_d, opt)
    got = trainer_ops.compute_losses_dualrefine(inp_d, o_d, opt, noises=[n.to(dev) for n in noises])
    for k in want:
        assert _close(got[k], want[k]), k
    for s, it in keys:
        sel = o_d[("mal_selection", s, it)].cpu().numpy()
        assert np.array_equal(sel & 0x7F, aux[("frame_idx", s, it)].numpy().astype(np.uint8)), (s, it)
        assert np.array_equal(sel >> 7, aux[("mask", s, it)].numpy().astype(np.uint8)) or it > 0   # bit-exact automask
    g = torch.autograd.grad(got["loss"], leaves_d)
    for a, b in zip(g, want_g):
        assert _gerr(a, b) < GRAD_RTOL


@pytest.mark.parametrize("is_multi,selec,zero", [(False, True, True), (True, True, True), (False, False, False)])
def test_dynamicdepth_losses_match_oracle(op_device, is_multi, selec, zero):
    """BASELINE config 5: DynamicDepth's 4-scale losses with `selec_reproj` and `zero_img`
    (dynamicdepth/trainer.py:958-975, :1006-1128), incl. the in-place zeroing of the target image;
    the oracle side is pinned bitwise against the reference's Trainer.compute_losses."""
    from oracle.pin_against_reference import dynamicdepth_loss_case
    dev = op_device
    H, W = 32, 64
    name = "multi" if is_multi else "mono"
    res = []
    for who, dv in (("oracle", torch.device("cpu")), ("ours", dev)):
        inputs, t, outs = dynamicdepth_loss_case(1, H, W)
        inputs = {k: v.to(dv) for k, v in inputs.items()}
        o = {k: (v.to(dv) if torch.is_tensor(v) else v) for k, v in outs[name].items()}
        preds = []
        for s in range(4):
            for f in (-1, 1):
                o[("color", f, s)] = o[("color", f, s)].detach().clone().requires_grad_(True)
                preds.append(o[("color", f, s)])
        disps = [o[("disp", s)].detach().clone().requires_grad_(True) for s in range(4)]
        for s in range(4):
            o[("disp", s)] = disps[s]
        if is_multi:
            for s in range(4):
                o[("depth", 0, s)] = o[("depth", 0, s)].detach().clone().requires_grad_(True)
                preds.append(o[("depth", 0, s)])
        noises = [n.to(dv) for n in (t["noise"] + t["noise"])]
        if who == "oracle":
            losses, aux = O.dynamicdepth_compute_losses(inputs, o, (0, 1, 2, 3), is_multi=is_multi, selec_reproj=selec,
                                                        zero_img=zero, noises=noises)
            masks = [aux[("mask", s)] for s in range(4)]
        else:
            opt = SimpleNamespace(scales=[0, 1, 2, 3], selec_reproj=selec, zero_img=zero, no_ssim="false",
                                  no_matching_augmentation="false", disparity_smoothness=1e-3)
            losses = trainer_ops.compute_losses_dynamicdepth(inputs, o, opt, is_multi=is_multi, noises=noises)
            masks = [o[("mal_mask", s)] for s in range(4)]
        grads = torch.autograd.grad(losses["loss"], preds + disps)
        res.append((losses, [m.cpu() for m in masks], [g.cpu() for g in grads], inputs[("color", 0, 0)].cpu()))
    (lo, mo, go, to_), (lk, mk, gk, tk) = res
    for k in lo:
        assert _close(lk[k], lo[k]), k
    for a, b in zip(mk, mo):
        assert torch.equal(a, b)
    assert torch.equal(tk, to_)                      # the target image was mutated identically
    for a, b in zip(gk, go):
        assert _gerr(a, b) < GRAD_RTOL


@pytest.mark.parametrize("is_multi,selec,zero,automask", [(False, True, True, True), (True, True, True, True),
                                                          (False, False, False, True), (False, True, False, True),
                                                          (False, False, True, True), (False, True, True, False)])
def test_dynamicdepth_fused_losses_match_oracle(op_device, is_multi, selec, zero, automask):
    """BASELINE config 5 on the fused path: generate_images_pred_dynamicdepth leaves a WarpSpec per scale and
    compute_losses_dynamicdepth runs ONE photometric pass per scale (warps made in the kernel, identity candidates
    riding along, zero_img's cumulative zeroing of the target, selec_reproj).  Black regions in the source frames
    give the warps DOMD-style holes (dynamicdepth/trainer.py:906-975, :1006-1128).  Gradients w.r.t. the
    disparities and poses."""
    dev = op_device
    B, H, W = 2, 32, 64
    res = []
    for who, dv in (("oracle", torch.device("cpu")), ("ours", dev)):
        inputs, t = make_photometric_inputs(B, H, W, num_scales=4, seed=77, translation_scale=0.3)
        inputs[("color", -1, 0)][:, :, 4:14, 6:30] = 0.0
        inputs[("color", 1, 0)][:, :, 8:20, 20:44] = 0.0
        inputs[("color", 1, 0)][1, :, 2:9, 3:20] = 0.0
        inputs = {k: v.to(dv) for k, v in inputs.items()}
        name = "multi" if is_multi else "mono"
        disps = [t[(name + "_disp", s)].clone().to(dv).requires_grad_(True) for s in range(4)]
        Ts = [t[("cam_T_cam", 0, f)].clone().to(dv).requires_grad_(True) for f in (-1, 1)]
        o = {("disp", s): disps[s] for s in range(4)}
        o[("cam_T_cam", 0, -1)], o[("cam_T_cam", 0, 1)] = Ts
        if is_multi:
            o["consistency_mask"] = t["consistency_mask"].to(dv)
            o["augmentation_mask"] = torch.tensor([0.0, 1.0]).view(B, 1, 1, 1).to(dv)   # one live, one augmented sample
            for s in range(4):
                o[("mono_depth", 0, s)] = (1.0 + 5.0 * t[("mono_disp", 0)]).to(dv)
        noises = [n.to(dv) for n in t["noise"]]
        if who == "oracle":
            O.images_pred(inputs, o, num_scales=4, height=H, width=W, is_multi=is_multi)
            losses, aux = O.dynamicdepth_compute_losses(inputs, o, (0, 1, 2, 3), is_multi=is_multi, selec_reproj=selec,
                                                        zero_img=zero, noises=noises, automask=automask)
            masks = [aux[("mask", s)].reshape(B, H, W) for s in range(4)]
        else:
            opt = SimpleNamespace(scales=[0, 1, 2, 3], selec_reproj=selec, zero_img=zero, no_ssim="false", height=H,
                                  width=W, min_depth=0.1, max_depth=100.0, disable_automasking=not automask,
                                  disable_motion_masking=False, no_matching_augmentation="false",
                                  disparity_smoothness=1e-3)
            trainer_ops.generate_images_pred_dynamicdepth(inputs, o, opt, is_multi=is_multi)
            losses = trainer_ops.compute_losses_dynamicdepth(inputs, o, opt, is_multi=is_multi, noises=noises)
            masks = [(o[("mal_selection", s)] >> 7).float().reshape(B, H, W) for s in range(4)]
        leaves = disps + ([] if is_multi else Ts)     # is_multi detaches the poses
        grads = torch.autograd.grad(losses["loss"], leaves)
        res.append((losses, [m.cpu() for m in masks], [g.cpu() for g in grads], inputs[("color", 0, 0)].cpu()))
    (lo, mo, go, to_), (lk, mk, gk, tk) = res
    for k in lo:
        assert _close(lk[k], lo[k]), k
    if not is_multi and automask:
        for a, b in zip(mk, mo):
            assert torch.equal(a, b.float())         # bit-exact automask
    assert torch.equal(tk, to_)                      # the target image was mutated identically
    for a, b in zip(gk, go):
        assert _gerr(a, b) < GRAD_RTOL


@pytest.mark.parametrize("padding,align", [("border", True), ("border", False), ("zeros", True), ("zeros", False)])
def test_grid_sample_matches_torch_cpu(op_device, padding, align):
    """layers.grid_sample: forward bit-identical to torch's CPU F.grid_sample, backward w.r.t. the grid."""
    dev = op_device
    gen = torch.Generator().manual_seed(11)
    img = torch.rand(2, 3, 24, 40, generator=gen)
    grid = (torch.rand(2, 20, 36, 2, generator=gen) * 2.4 - 1.2)
    w = torch.rand(2, 3, 20, 36, generator=gen)
    g_c = grid.clone().requires_grad_(True)
    want = F.grid_sample(img, g_c, padding_mode=padding, align_corners=align)
    want_g, = torch.autograd.grad((want * w).sum(), g_c)
    g_d = grid.clone().to(dev).requires_grad_(True)
    got = layers.grid_sample(img.to(dev), g_d, padding_mode=padding, align_corners=align)
    assert torch.equal(got.detach().cpu(), want.detach())
    got_g, = torch.autograd.grad((got * w.to(dev)).sum(), g_d)
    assert _gerr(got_g, want_g) < GRAD_RTOL


def test_loss_balancing_running_mean_is_the_reference_mean():
    """LossBalancing.update_weight keeps column sums instead of re-reading every recorded score (the
    reference's mean runs over all of them, every step): same fp64 weights over a long sequential run, over
    an epoch restart (rows overwritten from index 0) and past the end of the score table."""
    rng = np.random.default_rng(3)
    bs, n_data = 4, 4 * 260
    ours, ref = loss_utils.LossBalancing(2, n_data, bs), O.LossBalancing(2, n_data, bs)
    for epoch in range(2):
        for it in range(300):           # 300 * 4 rows > 1040: the last steps fall off the table
            scores = [0.4 + 0.1 * rng.random(), 0.01 + 0.01 * rng.random()]
            ours.record_scores(it, scores)
            ref.compute_loss(scores, it)
            assert np.array_equal(np.array(ours.update_weight(it, 3.0)), np.array(ref.update_weight(it, 3.0))), (epoch, it)
    assert np.array_equal(ours.train_scores, ref.train_scores)


@pytest.mark.parametrize("v1_multiscale,disable_automasking", [(True, False), (False, True), (True, True)])
def test_compute_losses_v1_multiscale_and_no_automask(op_device, v1_multiscale, disable_automasking):
    """Trainer.compute_losses teacher pass with the two options that change what is compared:
    v1_multiscale (images / intrinsics of the disparity's own scale, manydepth/trainer.py:1089-1095,
    :1260-1263) and disable_automasking (the identity loss is still compared, only the tie-break
    noise is dropped, :1292-1311)."""
    dev = op_device
    B, H, W, S = 1, 32, 64, 2
    inputs, t = make_photometric_inputs(B, H, W, num_scales=S, seed=123)
    for s in range(1, S):
        for f in (-1, 1):
            inputs[("color", f, s)] = F.interpolate(inputs[("color", f, 0)], scale_factor=1 / 2 ** s, mode="area")
    opt = _opt(B, H, W, sclm=S - 1, distil=False, temporal=False, v1_multiscale=v1_multiscale,
               disable_automasking=disable_automasking)
    noises = [torch.randn(B, 1, H // 2 ** s if v1_multiscale else H, W // 2 ** s if v1_multiscale else W,
                          generator=torch.Generator().manual_seed(5 + s)) for s in range(S)]
    disps_c = [t[("mono_disp", s)].clone().requires_grad_(True) for s in range(S)]
    Tc = {f: t[("cam_T_cam", 0, f)].clone().requires_grad_(True) for f in (-1, 1)}
    o = {("disp", s): disps_c[s] for s in range(S)}
    o.update({("cam_T_cam", 0, f): Tc[f] for f in (-1, 1)})
    O.images_pred(inputs, o, num_scales=S, height=H, width=W, v1_multiscale=v1_multiscale)
    want, aux = O.trainer_compute_losses(inputs, o, num_scales=S, batch_size=B, noises=noises,
                                         automask=not disable_automasking, v1_multiscale=v1_multiscale)
    want_g = torch.autograd.grad(want["loss"], disps_c + [Tc[-1], Tc[1]])

    inputs_d = to_device(inputs, dev)
    disps = [t[("mono_disp", s)].clone().to(dev).requires_grad_(True) for s in range(S)]
    Td = {f: t[("cam_T_cam", 0, f)].clone().to(dev).requires_grad_(True) for f in (-1, 1)}
    od = {("disp", s): disps[s] for s in range(S)}
    od.update({("cam_T_cam", 0, f): Td[f] for f in (-1, 1)})
    trainer_ops.generate_images_pred(inputs_d, od, opt)
    got, _ = trainer_ops.compute_losses(inputs_d, od, opt, noises=[n.to(dev) for n in noises])
    for k in want:
        assert _close(got[k], want[k]), k
    for s in range(S):
        sel = od[("mal_selection", s)].cpu().numpy()
        assert np.array_equal(sel & 0x7F, aux[("frame_idx", s)].numpy().astype(np.uint8))
        assert np.array_equal(sel >> 7, aux[("mask", s)].numpy().astype(np.uint8))      # bit-exact automask
    g = torch.autograd.grad(got["loss"], disps + [Td[-1], Td[1]])
    for a, b in zip(g, want_g):
        assert _gerr(a, b) < GRAD_RTOL
