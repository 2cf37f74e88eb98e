"""CPU oracle for one whole MAL hot-path step.  TEST INFRASTRUCTURE ONLY (see mal_oracle.py).

Composition, in the reference's order, of the functions Trainer.process_batch runs between the
networks and the loss (manydepth/trainer.py:555-644 with --temporal --distil --loss_blc):
ResnetEncoderMatching.forward's matching head (networks/resnet_encoder.py:292-317), the nearest
up-sampling in RepDepth.forward (networks/repdepth.py:331-336), generate_images_pred,
compute_mono_losses, compute_matching_mask, generate_images_pred_ensemble, compute_main_losses and
the LossBalancing weighting.  Used by tests/test_step.py as the checker and by bench.py's
cpu_baseline / --impl reference legs as the timed CPU path.
"""
import torch
import torch.nn.functional as F

from . import mal_oracle as O

LEAVES = ("mono_disp", "multi_disp", "T_-1", "T_1")   # the tensors the networks would have produced


def oracle_step(b, opt, weights=(0.5, 0.5), multi_has_ins=False):
    """manydepth/trainer.py:555-644 (--temporal --distil --loss_blc) through oracle/mal_oracle.py."""
    H, W, B = opt.height, opt.width, opt.batch_size
    leaves = {k: b[k].clone().requires_grad_(True) for k in LEAVES}
    inputs = {("color", 0, 0): b["color_0"], ("color", -1, 0): b["color_-1"], ("color", 1, 0): b["color_1"],
              ("K", 0): b["K"], ("inv_K", 0): b["inv_K"]}
    cv, low, conf, idx, _ = O.cost_volume_head(b["current_feats"], b["lookup_feats"], b["relative_poses"], b["K2"],
                                               b["inv_K2"], b["bins"])
    mono = {("disp", 0): leaves["mono_disp"]}
    multi = {("disp", 0): leaves["multi_disp"], "augmentation_mask": b["augmentation_mask"],
             "lowest_cost": F.interpolate(low.unsqueeze(1), [H, W], mode="nearest")[:, 0],
             "consistency_mask": F.interpolate(conf.unsqueeze(1), [H, W], mode="nearest")[:, 0]}
    for f in (-1, 1):
        mono[("cam_T_cam", 0, f)] = multi[("cam_T_cam", 0, f)] = leaves["T_%d" % f]
        if "syn_%d" % f in b:
            mono[("syn", f, 0)] = multi[("syn", f, 0)] = b["syn_%d" % f]
    O.images_pred(inputs, mono, height=H, width=W)
    if "masks_last" in b:
        # trainer.py:1161-1162 image_synthesis on the materialised warps (dyn_utils.py:121-170): per sample with
        # matched instances, generate_dynamic_instance; the copies keep autograd history to the warped images
        syn = [mono[("color", -1, 0)], mono[("color", 1, 0)]]
        for s_ in range(B):
            n = int(b["mask_counts"][s_])
            if n == 0:
                continue
            ml = ((b["masks_last"][s_].long().unsqueeze(0) >> torch.arange(n).view(-1, 1, 1)) & 1).bool()
            mn = ((b["masks_next"][s_].long().unsqueeze(0) >> torch.arange(n).view(-1, 1, 1)) & 1).bool()
            ol, on, _ = O.generate_dynamic_instance(ml, mn, mono[("color", -1, 0)][s_], mono[("color", 1, 0)][s_])
            syn[0] = torch.cat([syn[0][:s_], ol[None], syn[0][s_ + 1:]])
            syn[1] = torch.cat([syn[1][:s_], on[None], syn[1][s_ + 1:]])
        mono[("syn", -1, 0)], mono[("syn", 1, 0)] = syn
    mono_losses, mono_reproj, _ = O.mono_losses(inputs, mono, True, True, noise=b["noise_mono"])
    multi[("mono_depth", 0, 0)] = mono[("depth", 0, 0)]
    multi["consistency_mask"] = multi["consistency_mask"] * O.matching_mask(multi)
    ens = O.images_pred_ensemble(inputs, leaves["T_-1"].detach(), leaves["T_1"].detach(),
                                 (leaves["mono_disp"].detach() + leaves["multi_disp"].detach()) / 2.0, height=H, width=W)
    O.images_pred(inputs, multi, height=H, width=W, is_multi=True)
    if "masks_last" in b and multi_has_ins and getattr(opt, "main_temporal", False):
        # trainer.py:1164-1165: the multi pass synthesises its own hint from its own warps
        sfx = "_multi" if "masks_last_multi" in b else ""
        syn = [multi[("color", -1, 0)], multi[("color", 1, 0)]]
        for s_ in range(B):
            n = int(b["mask_counts" + sfx][s_])
            if n == 0:
                continue
            ml = ((b["masks_last" + sfx][s_].long().unsqueeze(0) >> torch.arange(n).view(-1, 1, 1)) & 1).bool()
            mn = ((b["masks_next" + sfx][s_].long().unsqueeze(0) >> torch.arange(n).view(-1, 1, 1)) & 1).bool()
            ol, on, _ = O.generate_dynamic_instance(ml, mn, multi[("color", -1, 0)][s_], multi[("color", 1, 0)][s_])
            syn[0] = torch.cat([syn[0][:s_], ol[None], syn[0][s_ + 1:]])
            syn[1] = torch.cat([syn[1][:s_], on[None], syn[1][s_ + 1:]])
        multi[("syn", -1, 0)], multi[("syn", 1, 0)] = syn
    losses, _, loss_list, aux = O.main_losses(inputs, multi, mono_reproj, ens, batch_size=B, loss_blc=True,
                                              multi_has_ins=bool(multi_has_ins and getattr(opt, "main_temporal", False)),
                                              noise=b["noise_main"])
    loss_list[0] = loss_list[0] + mono_losses["loss"]
    total = B * (weights[0] * loss_list[0] + weights[1] * loss_list[1])
    grads = torch.autograd.grad(total, [leaves[k] for k in LEAVES])
    return total.detach(), [l.detach() for l in loss_list], grads, {"cv": cv, "mask": multi["consistency_mask"],
                                                                     "distil_idx": aux["distil_idx"]}


