"""MAL's temporal hint (manydepth/dyn_utils.py) on the sm_100a kernels.

    fill_dynamic_obj(mask, delta_x, delta_y, source, img)                                   :5-36
    generate_dynamic_instance(grid_h, grid_w, mask_last, mask_next, img_last, img_next, replace)   :38-119
    image_synthesis(inputs, outputs, scale, thres, ins_model, matcher)                      :121-170
    generate_instances(images, ins_model)                                                   :172-188

The per-pixel / per-instance work (mask extents, displacement, background swap, shifted copies,
composition) is three kernels (csrc/dynsyn.cu) instead of TorchScript loops over instances with
(N,3,H,W) temporaries.  `image_synthesis` keeps the reference's orchestration: the instance
segmenter (`ins_model`, Mask2Former in the reference) and the Hungarian `matcher` are the
caller's - they are out of scope here - and any callable with the same
interface works (tests use synthetic Mask2Former-shaped masks).

Like the reference's copies, the synthesised images keep autograd history to the warped source
images (mal_dynamic_instance_backward); both loss paths (PRED mode and the fused WARP mode) return
d loss / d syn when the candidates require grad.
"""
from __future__ import annotations

import torch

from . import ops


def fill_dynamic_obj(mask, delta_x, delta_y, source, img):
    """Paste `source` under each instance mask shifted by (delta_x along H, delta_y along W) over `img`."""
    return ops.fill_dynamic_obj(mask, delta_x, delta_y, source, img)


def generate_dynamic_instance(grid_h, grid_w, mask_last, mask_next, img_last, img_next, replace: bool):
    """(N,H,W) matched masks of frames -1/+1 + the two warped images (3,H,W) -> the two synthesised
    images.  `grid_h` / `grid_w` are accepted for signature compatibility (the kernels index directly)."""
    ori_last, ori_next, _ = ops.dynamic_instance(mask_last, mask_next, img_last, img_next, bool(replace))
    return ori_last, ori_next


def generate_instances(images, ins_model):
    """RGB [0,1] batch -> the segmenter's instance predictions (detectron2 input convention)."""
    height, width = images.shape[-2:]
    images = images[:, [2, 1, 0], :, :] * 255
    batch = [{"image": img, "height": height, "width": width} for img in images]
    with torch.no_grad():
        return ins_model(batch)


def image_synthesis(inputs, outputs, scale, thres, ins_model, matcher):
    """Fill outputs[("syn", -1/+1, scale)] from the warped images; returns has_ins."""
    bs = inputs[("color", 0, 0)].shape[0]
    instances = generate_instances(inputs[("color", 0, 0)], ins_model)
    syn_last = outputs[("color", -1, scale)].clone()
    syn_next = outputs[("color", 1, scale)].clone()
    has_ins = False
    for b in range(bs):
        cur = instances[b]["instances"]
        instances_cur = cur[cur.scores > thres]
        if len(instances_cur) == 0:
            continue
        img_last = outputs[("color", -1, scale)][b].clone()
        img_next = outputs[("color", 1, scale)][b].clone()
        both = generate_instances(torch.stack([img_last, img_next], 0).detach(), ins_model)
        ins_last, ins_next = both[0]["instances"], both[1]["instances"]
        slice_last, slice_next = matcher(ins_last, ins_next, instances_cur)
        if len(slice_last) + len(slice_next) == 0:
            continue
        has_ins = True
        mask_last = ins_last.pred_masks[slice_last].bool()
        mask_next = ins_next.pred_masks[slice_next].bool()
        tmp_last, tmp_next = generate_dynamic_instance(None, None, mask_last, mask_next, img_last, img_next, replace=False)
        syn_last[b], syn_next[b] = tmp_last, tmp_next
    if has_ins:
        outputs[("syn", -1, scale)] = syn_last
        outputs[("syn", 1, scale)] = syn_next
    return has_ins
