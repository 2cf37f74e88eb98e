"""MAL temporal hint (mal_dynamic_instance / mal_fill_dynamic_obj) against the oracle restatement
of manydepth/dyn_utils.py (pinned bitwise against the reference's TorchScript by
oracle/pin_against_reference.py).  Integer / byte work: everything must be bit-exact."""
from types import SimpleNamespace

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
