"""One MAL hot-path step: everything ManyDepth+MAL's Trainer.process_batch does between the
conv networks and `loss.backward()` reaching them (manydepth/trainer.py:555-644 with
`--temporal --distil --loss_blc`, README.md:19-25), for one batch:

    cost-volume head (ResnetEncoderMatching.forward :292-317)
 -> teacher warps + compute_mono_losses            (loss_utils.py:57-129)
 -> teacher->student hand-over, matching mask       (trainer.py:584-593)
 -> ensemble reprojection                           (trainer.py:1172-1207)
 -> student warps + compute_main_losses             (loss_utils.py:131-281)
 -> LossBalancing weighting                         (loss_utils.py:303-345)
 -> backward to the disparities and poses the networks produced.

`MalStep` owns static device buffers for the batch and (by default) captures the whole step -
forward and backward, 20-odd launches - into one CUDA graph, so a training loop pays one launch
per step instead of Python dispatch per kernel.  The LossBalancing state stays on the host in
fp64 exactly like the reference: the two loss scalars come back after each replay and the new
weights go up before the next one.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

from . import loss_utils, ops, raw, trainer_ops
from .utils.synthetic import make_cost_volume_inputs, make_photometric_inputs, warp_by_depth

INPUT_KEYS = ("color_0", "color_-1", "color_1", "syn_-1", "syn_1", "mono_disp", "multi_disp", "T_-1", "T_1",
              "K", "inv_K", "K2", "inv_K2", "current_feats", "lookup_feats", "relative_poses",
              "augmentation_mask", "noise_mono", "noise_main", "bins")
LEAVES = ("mono_disp", "multi_disp", "T_-1", "T_1")
# --temporal as the reference runs it (trainer.py:1161-1162): instead of ready-made syn images the batch carries
# the matched instance masks of the two warped frames, packed one bit per instance (mal_temporal_pack_masks), and
# the step materialises the warps, synthesises the temporal-hint images and back-propagates through them
MASK_KEYS = ("masks_last", "masks_next", "mask_counts")
# --main_temporal (trainer.py:1164-1165): the matched masks of the MULTI pass's warps, optional (the teacher's masks
# are reused without them)
MASK_KEYS_MULTI = tuple(k + "_multi" for k in MASK_KEYS)
SYN_KEYS = ("syn_-1", "syn_1")


# what the reference's data loader / CPU generator hands to the GPU every step (images, intrinsics, the CPU-drawn
# tie-break noise of loss_utils.py:105-106 / :178, the segmenter-shaped masks); everything else in a batch is born
# on the device in the reference (network outputs) and only travels in the all-from-host end-to-end measurement
HOST_BORN = ("color_0", "color_-1", "color_1", "K", "inv_K", "K2", "inv_K2", "noise_mono", "noise_main", "bins") + \
    MASK_KEYS + MASK_KEYS_MULTI


def batch_keys(batch):
    """The input keys a batch carries, host-born ones first (so they are one contiguous range of a staged batch):
    INPUT_KEYS, with the packed instance masks in place of the syn images when the temporal hint is synthesised
    inside the step."""
    keys = INPUT_KEYS
    if "masks_last" in batch:
        keys = tuple(k for k in INPUT_KEYS if k not in SYN_KEYS) + MASK_KEYS
        if "masks_last_multi" in batch:
            keys = keys + MASK_KEYS_MULTI
    return tuple(k for k in keys if k in HOST_BORN) + tuple(k for k in keys if k not in HOST_BORN)


def default_opt(batch, height=192, width=640, **kw):
    """The README's ManyDepth+MAL flags: --temporal --distil --loss_blc."""
    o = dict(height=height, width=width, batch_size=batch, min_depth=0.1, max_depth=100.0, frame_ids=[0, -1, 1],
             sclm=0, temporal=True, main_temporal=False, distil=True, no_ens=False, loss_blc=True,
             dual_distil=False, num_depth_bins=96, min_depth_bin=0.1, max_depth_bin=20.0, matching_channels=64)
    o.update(kw)
    return SimpleNamespace(**o)


def synthetic_masks(opt, seed=1234, max_instances=12, max_shift=8):
    """Synthetic Mask2Former-shaped matched instance masks for a batch (0-N random
    rectangles / ellipses per frame, shifted between "last" and "next"), packed one bit per instance:
    (masks_last, masks_next) int32 (B,H,W) and the per-sample instance counts.  One sample gets no instance."""
    from .utils.synthetic import make_instance_masks
    B, H, W = opt.batch_size, opt.height, opt.width
    gen = torch.Generator().manual_seed(seed + 99)
    counts = torch.randint(1, max_instances + 1, (B,), generator=gen, dtype=torch.int32)
    if B > 1:
        counts[B - 1] = 0
    pl, pn = torch.zeros(B, H, W, dtype=torch.int64), torch.zeros(B, H, W, dtype=torch.int64)
    for b in range(B):
        n = int(counts[b])
        if n == 0:
            continue
        # object-sized instances (a car at this resolution is a few dozen pixels across): half-extents up to
        # H/8 x W/10, so the masks of a 12-instance frame cover of the order of 15% of the image
        last, nxt = make_instance_masks(n, H, W, seed=seed + 31 * b, max_shift=max_shift, max_ry=max(6, H // 8),
                                        max_rx=max(6, W // 10))
        w = (1 << torch.arange(n, dtype=torch.int64)).view(-1, 1, 1)
        pl[b], pn[b] = (last.long() * w).sum(0), (nxt.long() * w).sum(0)
    to_i32 = lambda t: torch.where(t >= 2 ** 31, t - 2 ** 32, t).to(torch.int32)
    return to_i32(pl), to_i32(pn), counts


def unpack_masks(packed, count):
    """(H,W) int32 bit planes of one sample -> (count,H,W) bool."""
    bits = torch.arange(count, dtype=torch.int64).view(-1, 1, 1)
    return ((packed.long().unsqueeze(0) >> bits) & 1).bool()


def synthetic_batch(opt, seed=1234, normalised_K=None, translation_scale=0.2, adaptive_bins=True, with_masks=False):
    """Seeded synthetic KITTI-shaped host tensors for one step, keyed by INPUT_KEYS.

    The camera baseline is a fifth of the nearest depth and (adaptive_bins) the depth hypotheses
    span the teacher's depth range the way the reference's DepthBins tracker sets them
    (manydepth/trainer.py:75-103, 634), so that the confidence and matching masks cover a
    realistic share of the image instead of being empty."""
    kw = {} if normalised_K is None else {"normalised_K": normalised_K}
    inputs, t = make_photometric_inputs(opt.batch_size, opt.height, opt.width, seed=seed,
                                        translation_scale=translation_scale, **kw)
    lo, hi = opt.min_depth_bin, opt.max_depth_bin
    if adaptive_bins:
        scaled = 1.0 / opt.max_depth + (1.0 / opt.min_depth - 1.0 / opt.max_depth) * t[("mono_disp", 0)]
        depth = 1.0 / scaled
        lo, hi = 0.9 * float(depth.min()), 1.1 * float(depth.max())
    cv = make_cost_volume_inputs(opt.batch_size, opt.height, opt.width, channels=opt.matching_channels,
                                 num_lookup=1, num_bins=opt.num_depth_bins, seed=seed + 1, min_bin=lo, max_bin=hi,
                                 translation_scale=translation_scale, **kw)
    cv["relative_poses"][:, 0] = t[("cam_T_cam", 0, -1)]   # the lookup frame is frame -1
    if adaptive_bins:
        # photo-consistent features: the current frame's features are the lookup frame's, seen
        # through the teacher depth (plus noise), so the sweep's arg-min lands near that depth
        low_depth = torch.nn.functional.avg_pool2d(depth, 4)
        gen = torch.Generator().manual_seed(seed + 2)
        look = cv["lookup_feats"][:, 0]
        cur = warp_by_depth(look, low_depth, cv["K"], cv["inv_K"], cv["relative_poses"][:, 0])
        cv["current_feats"] = (cur + 0.05 * torch.rand(cur.shape, generator=gen)).contiguous()
    b = {"color_0": inputs[("color", 0, 0)], "color_-1": inputs[("color", -1, 0)], "color_1": inputs[("color", 1, 0)],
         "syn_-1": t[("syn", -1, 0)], "syn_1": t[("syn", 1, 0)], "mono_disp": t[("mono_disp", 0)],
         "multi_disp": t[("multi_disp", 0)], "T_-1": t[("cam_T_cam", 0, -1)], "T_1": t[("cam_T_cam", 0, 1)],
         "K": inputs[("K", 0)], "inv_K": inputs[("inv_K", 0)], "K2": cv["K"], "inv_K2": cv["inv_K"],
         "current_feats": cv["current_feats"], "lookup_feats": cv["lookup_feats"],
         "relative_poses": cv["relative_poses"], "augmentation_mask": t["augmentation_mask"],
         "noise_mono": t["noise"][0], "noise_main": t["noise"][1], "bins": cv["bins"]}
    if with_masks:
        for k in SYN_KEYS:
            del b[k]
        b["masks_last"], b["masks_next"], b["mask_counts"] = synthetic_masks(opt, seed=seed)
    return {k: v.contiguous() for k, v in b.items()}


def step_losses(b, opt, leaves, weights=None, has_ins=True, multi_has_ins=False):
    """The step on a dict of device tensors `b` (INPUT_KEYS) with `leaves` standing in for the
    network outputs.  Returns (total, loss_list, losses, outputs)."""
    if "masks_last" in b:
        raise ValueError("step_losses (op by op) takes ready-made syn images; the in-step synthesis is part of fused_step")
    inputs = {("color", 0, 0): b["color_0"], ("color", -1, 0): b["color_-1"], ("color", 1, 0): b["color_1"],
              ("K", 0): b["K"], ("inv_K", 0): b["inv_K"]}
    cv, lowest_cost, confidence, *_ = _head(b, opt)
    mono = {("disp", 0): leaves["mono_disp"]}
    outputs = {("disp", 0): leaves["multi_disp"], "lowest_cost": lowest_cost, "consistency_mask": confidence,
               "augmentation_mask": b["augmentation_mask"], "cost_volume": cv}
    for f in (-1, 1):
        mono[("cam_T_cam", 0, f)] = outputs[("cam_T_cam", 0, f)] = leaves["T_%d" % f]
        mono[("syn", f, 0)] = outputs[("syn", f, 0)] = b["syn_%d" % f]
    outputs, losses = trainer_ops.process_batch_losses(
        inputs, mono, outputs, opt, has_ins=has_ins, multi_has_ins=multi_has_ins, loss_blc=None,
        noises=[b["noise_mono"], b["noise_main"]])
    if opt.loss_blc:
        # LossBalancing.compute_loss: sum_k w_k * L_k, added once per batch element (loss_utils.py:305-318)
        loss_list = [losses["loss"], losses["distil_loss"]]
        w = weights if weights is not None else torch.full((2,), 0.5, device=b["color_0"].device)
        total = opt.batch_size * (w[0] * loss_list[0] + w[1] * loss_list[1])
    else:
        loss_list = [losses["loss"]]
        total = losses["loss"]
    return total, loss_list, losses, outputs


def _head(b, opt):
    return _head_op(b["current_feats"], b["lookup_feats"], b["relative_poses"], b["K2"], b["inv_K2"], b["bins"])


def _head_op(cur, look, poses, K, inv_K, bins):
    cv, _, conf, _, low = ops.cost_volume(cur, look, poses, K, inv_K, bins, apply_confidence=True)
    return cv, low, conf


def fused_step(handle, b, opt, weights=None, has_ins=True, multi_has_ins=False, side_streams=None, head=None):
    """fused_step_main + fused_step_tail in one go (see there)."""
    ctx = fused_step_main(handle, b, opt, has_ins=has_ins, multi_has_ins=multi_has_ins, side_streams=side_streams,
                          head=head)
    return fused_step_tail(handle, b, opt, weights, ctx)


def fused_step_main(handle, b, opt, has_ins=True, multi_has_ins=False, side_streams=None, head=None):
    """Everything of the fused step that does not depend on the loss-balancing weights (every kernel but the
    last); returns the intermediates fused_step_tail needs.

    The same step as `step_losses` + backward, as 20 launches of libmal_b200 and nothing else:
    no autograd graph, no intermediate depth maps, no one-element torch kernels.  The scalar tail
    and the gradient hand-over are `mal_step_combine` (csrc/step.cu).  Needs opt.distil.

    `side_streams` (two CUDA streams) lets the three independent chains of the step - the cost-volume
    head, the teacher + ensemble passes, and the two smoothness terms - run as parallel branches
    (of the captured graph), so the small kernels and the last waves of the big ones overlap.

    Returns (scalars, grads, outputs): scalars = [total, loss_list[0], loss_list[1], reproj_loss/0
    of the teacher, of the student, consistency, teacher loss, student loss]; grads follow LEAVES."""
    if not opt.distil:
        raise ValueError("fused_step implements the --distil step; use step_losses otherwise")
    B, H, W = opt.batch_size, opt.height, opt.width
    lo, hi = opt.min_depth, opt.max_depth
    tgt, src = b["color_0"], [b["color_-1"], b["color_1"]]
    in_step_syn = "masks_last" in b
    syn = None if in_step_syn else [b["syn_-1"], b["syn_1"]]
    T = [b["T_-1"], b["T_1"]]
    mono, multi = b["mono_disp"].detach(), b["multi_disp"].detach()
    geom = dict(K=b["K"], inv_K=b["inv_K"], T=[t.detach() for t in T], min_depth=lo, max_depth=hi)

    main = torch.cuda.current_stream(tgt.device) if tgt.is_cuda else None
    branch = _Branches(main, side_streams)

    with branch(0):   # cost-volume head and the matching mask that only depends on it and the teacher disparity
        if head is None:   # (a caller whose student network already consumed the cost volume hands it in)
            head = raw.cost_volume(handle, current=b["current_feats"], lookup=b["lookup_feats"],
                                   poses=b["relative_poses"], K=b["K2"], inv_K=b["inv_K2"], bins=b["bins"],
                                   apply_confidence=True, want_missing=False)
        mask = raw.matching_mask(handle, lowest_cost=head["lowest_cost"], confidence=head["confidence"], mono=mono,
                                 mono_is_disp=True, min_depth=lo, max_depth=hi)
    with branch(1):   # smoothness of both disparities: one launch, the chain through the normalisation is
        # applied by mal_step_combine
        sm = raw.smooth(handle, disp=mono, img=tgt, normalise=True, with_grad=True, disp_b=multi, defer_fix=True)
        sm_t = {"loss": sm["loss"], "grad_disp": sm["grad_disp"], "_k": sm}
        sm_s = {"loss": sm["loss_b"], "grad_disp": sm["grad_disp_b"], "stats": sm["stats"]}
    # main chain: identity -> teacher -> ensemble.  The per-pass reductions (photo_finalize_kernel: 12 CTAs,
    # ~6 us) are not needed before step_combine, so each is forked onto the smoothness branch, where it runs
    # in the tail of the next heavy kernel instead of between two of them.
    def later(out):
        with branch(1):
            raw.photo_finalize(handle, out)
        return out

    # (the identity and ensemble passes only feed their per-pixel maps forward: their sums are never read)
    if not opt.no_ens:
        # the ensemble reprojection (trainer.py:1172-1207: both warps under the averaged disparity) and the identity
        # reprojection of the automask (loss_utils.py:92-101: the un-warped sources) are two 2-candidate mins against
        # the same target: ONE forward-only pass stages the target and its window moments once for both
        aux = raw.photo(handle, target=tgt, src=src, syn=src, depth=mono, depth_b=multi, want_selection=False,
                        finalize=False, split_min=True, **geom)
        ens, ident = aux["min_reproj"], aux["min_reproj_b"]
    else:
        ens = None
        ident = raw.photo(handle, target=tgt, src=src, mode=raw.PHOTO_PRED, want_selection=False,
                          finalize=False)["min_reproj"]
    use_syn = opt.temporal and has_ins
    hint = None
    warped = None
    if use_syn and in_step_syn:
        # trainer.py:1122-1125 + :1161-1162: materialise the two warps, synthesise the temporal-hint images
        warped = raw.temporal_warp(handle, src=src, depth=mono, **geom)
        hint = raw.temporal_synthesis(handle, warped=warped, packed_last=b["masks_last"], packed_next=b["masks_next"],
                                      counts=b["mask_counts"])
        syn = hint["syn"]
    # the teacher pass scores the warps just materialised for the temporal hint instead of re-warping (same bits)
    teacher = raw.photo(handle, target=tgt, src=src, syn=syn if use_syn else None, depth=mono, identity_min=ident,
                        noise=b["noise_mono"], with_grad=True, finalize=False, want_grad_syn=hint is not None,
                        warped=warped, **geom)
    with branch(1):
        raw.photo_finalize(handle, teacher)
        if hint is not None:
            # d loss / d syn -> warped images -> onto the teacher pass's d/d disparity plane and d/d(K@T) sums
            # (on the side branch: it overlaps the ensemble and student passes)
            hint["back"] = raw.temporal_backward(handle, grad_syn=teacher["grad_syn"], packed_last=b["masks_last"],
                                                 packed_next=b["masks_next"], counts=b["mask_counts"],
                                                 deltas=hint["deltas"], want_grad_warped=False, src=src, depth=mono,
                                                 grad_depth=teacher["grad_depth"], grad_P=teacher["grad_P"], **geom)
    branch.join(0)
    sample_mask = b["augmentation_mask"].reshape(-1)[:B]
    use_syn_s = bool(opt.main_temporal and multi_has_ins)
    hint_s, warped_s, syn_s = None, None, syn
    if use_syn_s and in_step_syn:
        # trainer.py:1164-1165: the multi pass synthesises its own temporal hint from ITS warps (the multi disparity).
        # The matched masks of those warps come as masks_*_multi; without them the teacher's masks are reused.
        sfx = "_multi" if "masks_last_multi" in b else ""
        warped_s = raw.temporal_warp(handle, src=src, depth=multi, **geom)
        hint_s = raw.temporal_synthesis(handle, warped=warped_s, packed_last=b["masks_last" + sfx],
                                        packed_next=b["masks_next" + sfx], counts=b["mask_counts" + sfx])
        syn_s = hint_s["syn"]
    student = raw.photo(handle, target=tgt, src=src, syn=syn_s if use_syn_s else None, depth=multi, pixel_mask=mask,
                        sample_mask=sample_mask, with_grad=True, finalize=False, want_grad_syn=hint_s is not None,
                        warped=warped_s, **geom)
    with branch(1):
        raw.photo_finalize(handle, student)
        if hint_s is not None:
            sfx = "_multi" if "masks_last_multi" in b else ""
            hint_s["back"] = raw.temporal_backward(handle, grad_syn=student["grad_syn"], packed_last=b["masks_last" + sfx],
                                                   packed_next=b["masks_next" + sfx], counts=b["mask_counts" + sfx],
                                                   deltas=hint_s["deltas"], want_grad_warped=False, src=src, depth=multi,
                                                   grad_depth=student["grad_depth"], grad_P=student["grad_P"], **geom)
    dual = bool(opt.dual_distil) and ens is None
    mt = raw.main_terms(handle, multi=multi, mono=mono, pixel_mask=mask, sample_mask=sample_mask,
                        mono_reproj=teacher["min_reproj"], ens_reproj=ens, multi_reproj=student["min_reproj"],
                        inputs_are_disp=True, dual_distil=dual, with_grad=True, min_depth=lo, max_depth=hi)
    branch.join(1)
    return dict(head=head, mask=mask, teacher=teacher, student=student, ens=ens, sm_t=sm_t, sm_s=sm_s, mt=mt,
                ident=ident, hint=hint, hint_s=hint_s)


def fused_step_tail(handle, b, opt, weights, ctx):
    """mal_step_combine: the only kernel that reads the LossBalancing weights.  MalStep replays it as its own
    graph so that the host-side weight update of step i-1 overlaps the heavy kernels of step i."""
    B, H, W = opt.batch_size, opt.height, opt.width
    head, mask, teacher, student, ens = ctx["head"], ctx["mask"], ctx["teacher"], ctx["student"], ctx["ens"]
    sm_t, sm_s, mt, ident = ctx["sm_t"], ctx["sm_s"], ctx["mt"], ctx["ident"]
    comb = raw.step_combine(handle, batch=B, height=H, width=W, weights=weights if opt.loss_blc else None,
                            sums_teacher=teacher["sums"], sums_student=student["sums"], smooth_teacher=sm_t["loss"],
                            smooth_student=sm_s["loss"], main_sums=mt["sums"], K=b["K"],
                            gd_teacher=teacher["grad_depth"], gs_teacher=sm_t["grad_disp"],
                            gP_teacher=teacher["grad_P"], gd_student=student["grad_depth"],
                            gs_student=sm_s["grad_disp"], g_cons=mt["grad_cons"], g_distil=mt["grad_distil"],
                            g_distil_mono=mt["grad_distil_mono"], smooth_stats=sm_s["stats"])
    outputs = {"cost_volume": head["cost_volume"], "lowest_cost": head["lowest_cost"],
               "confidence_mask": head["confidence"], "consistency_mask": mask,
               "mal_distil_index": mt["distil_index"], "consistency_target/0": mt["consistency_target"],
               ("mal_selection", 0): student["selection"], "mono_reproj": teacher["min_reproj"],
               "multi_reproj": student["min_reproj"], "ensemble_reproj": ens,
               "_keepalive": (head, teacher, student, sm_t, sm_s, mt, comb, ident, ctx.get("hint"), ctx.get("hint_s"))}
    if ctx.get("hint") is not None:
        outputs[("syn", -1, 0)], outputs[("syn", 1, 0)] = ctx["hint"]["syn"]
    grads = (comb["grad_disp_teacher"], comb["grad_disp_student"], comb["grad_T"][0], comb["grad_T"][1])
    return comb["scalars"], grads, outputs


class _Branches:
    """Fork / join of side streams off the current stream (a no-op without side streams, e.g. on
    the CPU twin used by the tests).  Works inside CUDA-graph capture: forks and joins are events."""

    def __init__(self, main, side):
        self.main, self.side = main, side if (side and main is not None) else None

    def __call__(self, i):
        return _Branch(self, i)

    def join(self, i):
        if self.side is not None:
            self.main.wait_stream(self.side[i])


class _Branch:
    def __init__(self, owner, i):
        self.o, self.i, self.ctx = owner, i, None

    def __enter__(self):
        if self.o.side is not None:
            st = self.o.side[self.i]
            st.wait_stream(self.o.main)
            self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


class StagedBatch(dict):
    """A batch whose tensors are views of one flat (pinned) buffer `flat` (MalStep.staging)."""
    flat = None


def _flat_layout(batch, align=256):
    off, items, host_born = 0, {}, 0
    for k in batch_keys(batch):
        t = batch[k]
        items[k] = (off, tuple(t.shape), t.dtype)
        off += (t.numel() * t.element_size() + align - 1) // align * align
        if k in HOST_BORN:
            host_born = off
    return {"items": items, "bytes": off, "host_born_bytes": host_born}


def _flat_views(flat, layout):
    out = {}
    for k, (off, shape, dtype) in layout["items"].items():
        n = 1
        for d in shape:
            n *= d
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        out[k] = flat[off:off + nbytes].view(dtype).view(shape)
    return out


class MalStep:
    """Static-buffer, graph-captured MAL step for one GPU.

    `slots` independent sets of static input buffers (each with its own captured graph) let a
    loader fill slot i+1 while slot i is being replayed, and let a benchmark rotate over more
    input bytes than the L2 holds."""

    def __init__(self, opt, device="cuda:0", use_graph=True, slots=1, num_train_data=1 << 20,
                 lambda_for_adjust=0.0, fused=True, branches=True, has_ins=True, multi_has_ins=False):
        self.opt, self.device, self.use_graph = opt, torch.device(device), use_graph
        # Trainer.has_ins / multi_has_ins (image_synthesis found matched instances): fixed for a captured step
        self.has_ins, self.multi_has_ins = has_ins, multi_has_ins
        self.fused = fused and opt.distil   # libmal_b200-only schedule; False: op-by-op through autograd
        self.slots = [dict(buf=None, graph=None, static=None) for _ in range(slots)]
        self.weights = torch.full((2,), 0.5, device=self.device)
        self.blc = loss_utils.LossBalancing(2, num_train_data, opt.batch_size) if opt.loss_blc else None
        self.lambda_for_adjust = lambda_for_adjust
        self.index_iter = 0
        self._w_host = torch.empty(2, dtype=torch.float32).pin_memory()
        self._scalars_host = torch.empty(8, dtype=torch.float32).pin_memory()
        self.launches_per_step = None
        self._layout = None
        self._pending = None   # (event, iteration) of the step whose LossBalancing update is still due
        self._warned_weights = False
        self.copy_stream = torch.cuda.Stream(self.device)
        # parallel branches only inside a captured graph: there every buffer is static, so tensors
        # produced on one stream and consumed on another need no allocator bookkeeping
        self.side_streams = ([torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
                             if (branches and use_graph) else None)

    # -- buffers -------------------------------------------------------------------------------
    def load_async(self, batch, slot=0, host_born_only=False):
        """Enqueue the host->device copy of a (pinned) batch into `slot` on the copy stream, so it
        overlaps the step running on another slot.  The copy waits for the last step that read the
        slot; the next `__call__(slot)` waits for the copy.  host_born_only copies just the tensors the
        reference's loader delivers (HOST_BORN); the network outputs of the slot stay as they are."""
        sl = self.slots[slot]
        if sl["buf"] is None:
            return self.load(batch, slot)
        keys = [k for k in batch_keys(batch) if (k in HOST_BORN or not host_born_only)]
        cur = torch.cuda.current_stream(self.device)
        if sl.get("done") is not None:
            self.copy_stream.wait_event(sl["done"])
        else:
            self.copy_stream.wait_stream(cur)
        with torch.cuda.stream(self.copy_stream), torch.no_grad():
            if isinstance(batch, StagedBatch) and batch.flat.numel() == sl["flat"].numel():
                n = self._layout["host_born_bytes"] if host_born_only else sl["flat"].numel()
                sl["flat"][:n].copy_(batch.flat[:n], non_blocking=True)
            else:
                for k in keys:
                    sl["buf"][k].copy_(batch[k], non_blocking=True)
            sl["ready"] = torch.cuda.Event()
            sl["ready"].record(self.copy_stream)
        return sum(batch[k].numel() * batch[k].element_size() for k in keys)

    def load(self, batch, slot=0, non_blocking=True):
        """Copy one batch (host or device tensors keyed by INPUT_KEYS) into a slot's static buffers.
        Returns the bytes copied."""
        sl = self.slots[slot]
        if sl["buf"] is None:
            # every input of the slot lives in ONE device allocation (256-byte aligned views), so that a
            # batch staged with `staging()` goes up as a single host->device copy
            self._layout = self._layout or _flat_layout(batch)
            sl["flat"] = torch.empty(self._layout["bytes"], dtype=torch.uint8, device=self.device)
            sl["buf"] = _flat_views(sl["flat"], self._layout)
            for k in LEAVES:
                sl["buf"][k].requires_grad_(True)
        with torch.no_grad():
            if isinstance(batch, StagedBatch) and batch.flat.numel() == sl["flat"].numel():
                sl["flat"].copy_(batch.flat, non_blocking=non_blocking)
            else:
                for k in batch_keys(batch):
                    sl["buf"][k].copy_(batch[k], non_blocking=non_blocking)
        return sum(batch[k].numel() * batch[k].element_size() for k in batch_keys(batch))

    def staging(self, like):
        """A pinned host batch with the slot layout: fill its tensors (same keys / shapes as `like`) and hand
        it to load() / load_async(); it then travels as one contiguous copy instead of one per tensor."""
        self._layout = self._layout or _flat_layout(like)
        flat = torch.empty(self._layout["bytes"], dtype=torch.uint8).pin_memory()
        staged = StagedBatch(_flat_views(flat, self._layout))
        staged.flat = flat
        for k in batch_keys(like):
            staged[k].copy_(like[k])
        return staged

    # -- one step ------------------------------------------------------------------------------
    def _run(self, buf):
        if self.fused:
            return self._run_tail(buf, self._run_main(buf))
        leaves = {k: buf[k] for k in LEAVES}
        total, loss_list, losses, outputs = step_losses(buf, self.opt, leaves, self.weights, has_ins=self.has_ins,
                                                        multi_has_ins=self.multi_has_ins)
        grads = torch.autograd.grad(total, [leaves[k] for k in LEAVES])
        scalars = torch.stack([total.detach(), loss_list[0].detach(), loss_list[-1].detach(),
                               losses["reproj_loss/0"].detach()])
        return scalars, grads, outputs

    def _run_main(self, buf):
        from . import _capi
        with torch.no_grad():
            return fused_step_main(_capi.lib(), buf, self.opt, has_ins=self.has_ins, multi_has_ins=self.multi_has_ins,
                                   side_streams=self.side_streams)

    def _run_tail(self, buf, ctx):
        from . import _capi
        with torch.no_grad():
            return fused_step_tail(_capi.lib(), buf, self.opt, self.weights, ctx)

    def _capture(self, sl):
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            branches, self.side_streams = self.side_streams, None   # fork only under capture (static buffers)
            try:
                for _ in range(2):
                    self._run(sl["buf"])
            finally:
                self.side_streams = branches
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        raw.LAUNCHES[0] = 0
        sl["graph"] = torch.cuda.CUDAGraph()
        if self.fused:
            # two graphs sharing one memory pool: everything up to the weights, then the kernel that reads them
            with torch.cuda.graph(sl["graph"]):
                ctx = self._run_main(sl["buf"])
            sl["graph_tail"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(sl["graph_tail"], pool=sl["graph"].pool()):
                sl["static"] = self._run_tail(sl["buf"], ctx)
            sl["ctx"] = ctx
        else:
            with torch.cuda.graph(sl["graph"]):
                sl["static"] = self._run(sl["buf"])
        self.launches_per_step = raw.LAUNCHES[0]

    def finish(self):
        """Apply the LossBalancing update of the last step (it is otherwise applied while the next step's
        heavy kernels run).  Call before reading `blc` / `weights` on the host."""
        if self._pending is None:
            return
        ev, it = self._pending
        self._pending = None
        ev.synchronize()
        self.blc.record_scores(it, [float(self._scalars_host[1]), float(self._scalars_host[2])])
        w0, w1 = self.blc.update_weight(it, self.lambda_for_adjust)
        if not (np.isfinite(w0) and np.isfinite(w1)) and not self._warned_weights:
            # the reference's LossBalancing does this too when a recorded loss mean is 0 (loss_utils.py:326);
            # its next backward is then NaN as well - keep its arithmetic, but say so
            import warnings
            warnings.warn("LossBalancing produced non-finite weights (%r, %r) at iteration %d: a loss term was "
                          "exactly 0, as in the reference the following steps' gradients are not finite" % (w0, w1, it))
            self._warned_weights = True
        self._w_host[0], self._w_host[1] = float(w0), float(w1)
        self.weights.copy_(self._w_host, non_blocking=True)

    def __call__(self, slot=0, sync_weights=True):
        """Run the step on the batch loaded in `slot`.  Returns (scalars, grads, outputs): scalars is
        a device tensor [total, loss, distil_loss, reproj_loss/0]; grads follow LEAVES."""
        sl = self.slots[slot]
        cur = torch.cuda.current_stream(self.device)
        if sl.get("ready") is not None:
            cur.wait_event(sl["ready"])
            sl["ready"] = None
        # LossBalancing (host fp64, like the reference) needs two scalars of step i-1 before step i's weights are
        # final - but only the last kernel of a step reads the weights.  So: replay the heavy part of step i,
        # finish step i-1's host update while it runs, upload the weights, replay the tail.
        if self.use_graph:
            if sl["graph"] is None:
                self._capture(sl)
            if not self.fused:
                self.finish()          # the op-by-op graph reads the weights throughout
            sl["graph"].replay()
            if self.fused:
                self.finish()
                sl["graph_tail"].replay()
            res = sl["static"]
        else:
            raw.LAUNCHES[0] = 0
            if self.fused:
                ctx = self._run_main(sl["buf"])
                self.finish()
                res = self._run_tail(sl["buf"], ctx)
            else:
                self.finish()
                res = self._run(sl["buf"])
            self.launches_per_step = raw.LAUNCHES[0]
        sl["done"] = torch.cuda.Event()
        sl["done"].record(cur)
        if self.blc is not None and sync_weights:
            self._scalars_host[:res[0].numel()].copy_(res[0], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            self._pending = (ev, self.index_iter)
        self.index_iter += 1
        return res
