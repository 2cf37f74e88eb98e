"""mal_b200.step: the whole MAL hot-path step (cost-volume head -> teacher -> matching mask ->
ensemble -> student -> balancing -> backward) against the oracle's composition of the same
reference functions, and the CUDA-graph replay against the eager run."""
import numpy as np
import pytest
import torch

from mal_b200 import step as S
from mal_b200.utils.synthetic import to_device
from oracle import mal_oracle as O
from oracle.step_oracle import oracle_step
from tests.backends import BACKENDS, handle_and_device

LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4


def _gerr(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    s = float(b.abs().max())
    return float((a - b).abs().max()) / (s if s > 0 else 1.0)


def _check(got_total, got_list, got_grads, outputs, want):
    total, loss_list, grads, aux = want
    assert float(loss_list[1]) > 0 and 0.02 < float(aux["mask"].mean()) < 0.98   # the case is not degenerate
    assert abs(float(got_total) - float(total)) <= LOSS_RTOL * abs(float(total))
    for a, b in zip(got_list, loss_list):
        assert abs(float(a) - float(b)) <= LOSS_RTOL * abs(float(b))
    assert torch.equal(outputs["cost_volume"].cpu(), aux["cv"])
    assert torch.equal(outputs["consistency_mask"].cpu(), aux["mask"])
    assert np.array_equal(outputs["mal_distil_index"].cpu().numpy(), aux["distil_idx"].numpy().astype(np.uint8))
    for a, b, k in zip(got_grads, grads, S.LEAVES):
        assert _gerr(a, b) < GRAD_RTOL, k


def test_step_losses_match_oracle(op_device):
    opt = S.default_opt(2, 32, 64, num_depth_bins=16, matching_channels=16)
    b = S.synthetic_batch(opt, seed=5)
    want = oracle_step(b, opt)
    d = to_device(b, op_device)
    leaves = {k: d[k].clone().requires_grad_(True) for k in S.LEAVES}
    total, loss_list, losses, outputs = S.step_losses(d, opt, leaves)
    grads = torch.autograd.grad(total, [leaves[k] for k in S.LEAVES])
    _check(total, loss_list, grads, outputs, want)


@pytest.mark.parametrize("backend", BACKENDS)
def test_fused_step_matches_oracle(backend):
    """The libmal_b200-only schedule (24 launches, mal_step_combine tail) against the oracle."""
    h, dev = handle_and_device(backend)
    opt = S.default_opt(2, 32, 64, num_depth_bins=16, matching_channels=16)
    b = S.synthetic_batch(opt, seed=5)
    for w in ((0.5, 0.5), (0.3, 1.7)):
        want = oracle_step(b, opt, w)
        d = to_device(b, dev)
        scalars, grads, outputs = S.fused_step(h, d, opt, torch.tensor(w, device=dev))
        _check(scalars[0], [scalars[1], scalars[2]], grads, outputs, want)
    # without loss balancing the total is the plain sum (loss_utils.py:271-275)
    opt2 = S.default_opt(2, 32, 64, num_depth_bins=16, matching_channels=16, loss_blc=False)
    scalars, _, _ = S.fused_step(h, to_device(b, dev), opt2)
    assert abs(float(scalars[0]) - float(scalars[1] + scalars[2])) < 1e-6


@pytest.mark.parametrize("backend", BACKENDS)
def test_fused_step_synthesises_the_temporal_hint_in_step(backend):
    """--temporal as the reference runs it (manydepth/trainer.py:1161-1162): the batch carries packed instance
    masks instead of syn images; the step materialises the warps, runs the synthesis and back-propagates
    d loss / d syn into the disparity and the poses.  Against the oracle's image_synthesis + autograd."""
    h, dev = handle_and_device(backend)
    opt = S.default_opt(3, 32, 64, num_depth_bins=16, matching_channels=16)
    b = S.synthetic_batch(opt, seed=5, with_masks=True)
    assert "syn_-1" not in b and int(b["mask_counts"].max()) > 0 and int(b["mask_counts"].min()) == 0
    want = oracle_step(b, opt, (0.4, 0.9))
    d = to_device(b, dev)
    scalars, grads, outputs = S.fused_step(h, d, opt, torch.tensor((0.4, 0.9), device=dev))
    _check(scalars[0], [scalars[1], scalars[2]], grads, outputs, want)
    # the temporal-hint candidates are really selected somewhere, and the gradient differs from the one a run with
    # the same images treated as data would give
    sel = outputs["_keepalive"][1]["selection"].cpu() & 0x7F
    assert int((sel >= 2).sum()) > 0
    b2 = {k: v for k, v in b.items() if k not in S.MASK_KEYS}
    b2["syn_-1"], b2["syn_1"] = outputs[("syn", -1, 0)].cpu(), outputs[("syn", 1, 0)].cpu()
    _, grads2, _ = S.fused_step(h, to_device(b2, dev), opt, torch.tensor((0.4, 0.9), device=dev))
    assert not torch.equal(grads[0].cpu(), grads2[0].cpu())


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("own_masks", [False, True])
def test_fused_step_main_temporal_synthesises_the_students_hint(backend, own_masks):
    """--main_temporal with instance masks (manydepth/trainer.py:1164-1165): the multi pass synthesises its own
    temporal hint from ITS warps (the multi disparity) and back-propagates through the copies; `own_masks` hands it
    the matched masks of those warps (masks_*_multi), otherwise the teacher's are reused."""
    h, dev = handle_and_device(backend)
    opt = S.default_opt(3, 32, 64, num_depth_bins=16, matching_channels=16)
    opt.main_temporal = True
    b = S.synthetic_batch(opt, seed=5, with_masks=True)
    if own_masks:
        other = S.synthetic_batch(opt, seed=6, with_masks=True)
        for k in S.MASK_KEYS:
            b[k + "_multi"] = other[k]
    want = oracle_step(b, opt, (0.4, 0.9), multi_has_ins=True)
    scalars, grads, outputs = S.fused_step(h, to_device(b, dev), opt, torch.tensor((0.4, 0.9), device=dev), multi_has_ins=True)
    _check(scalars[0], [scalars[1], scalars[2]], grads, outputs, want)
    sel = outputs[("mal_selection", 0)].cpu() & 0x7F
    assert int((sel >= 2).sum()) > 0          # the student really selects its hint candidates somewhere
    plain = oracle_step(b, opt, (0.4, 0.9), multi_has_ins=False)
    assert not torch.equal(plain[2][1], want[2][1])   # ... and that changes the gradient of the multi disparity


@pytest.mark.gpu
@pytest.mark.parametrize("use_graph,fused,with_masks", [(False, False, False), (True, False, False), (False, True, False),
                                                         (True, True, False), (False, True, True), (True, True, True)])
def test_malstep_graph_and_eager(use_graph, fused, with_masks):
    opt = S.default_opt(2, 64, 96, num_depth_bins=32, matching_channels=32)
    b = S.synthetic_batch(opt, seed=11, with_masks=with_masks)
    want = oracle_step(b, opt)
    st = S.MalStep(opt, use_graph=use_graph, fused=fused)
    st.load(b)
    for it in range(3):   # replays must reproduce the first run (weights start at 0.5, lambda 0 keeps them)
        scalars, grads, outputs = st(0, sync_weights=False)
        torch.cuda.synchronize()
        _check(scalars[0], [scalars[1], scalars[2]], grads, outputs, want)
    assert st.launches_per_step and 8 <= st.launches_per_step <= 17


@pytest.mark.gpu
def test_malstep_main_temporal_with_the_multi_pass_masks():
    """The captured step with --main_temporal and instance masks for both passes (masks_*_multi travel with the
    staged batch): three replays against the oracle."""
    opt = S.default_opt(2, 64, 96, num_depth_bins=32, matching_channels=32)
    opt.main_temporal = True
    b = S.synthetic_batch(opt, seed=11, with_masks=True)
    other = S.synthetic_batch(opt, seed=12, with_masks=True)
    for k in S.MASK_KEYS:
        b[k + "_multi"] = other[k]
    want = oracle_step(b, opt, multi_has_ins=True)
    st = S.MalStep(opt, use_graph=True, fused=True, multi_has_ins=True)
    st.load(b)
    for it in range(3):
        scalars, grads, outputs = st(0, sync_weights=False)
        torch.cuda.synchronize()
        _check(scalars[0], [scalars[1], scalars[2]], grads, outputs, want)


@pytest.mark.gpu
def test_malstep_loss_balancing_follows_reference():
    """Host-side LossBalancing driven from the graph replays == the oracle's LossBalancing driven
    by the oracle's losses (weights change every step, loss_utils.py:320-345)."""
    opt = S.default_opt(2, 64, 96, num_depth_bins=32, matching_channels=32)
    b = S.synthetic_batch(opt, seed=11)
    st = S.MalStep(opt, use_graph=True, num_train_data=64, lambda_for_adjust=3.0)
    ref = O.LossBalancing(2, 64, opt.batch_size)
    w = (0.5, 0.5)
    for it in range(3):
        b["noise_mono"] = torch.randn(b["noise_mono"].shape, generator=torch.Generator().manual_seed(100 + it))
        st.load(b)
        scalars, _, _ = st()
        total, loss_list, _, _ = oracle_step(b, opt, w)
        assert abs(float(scalars[0]) - float(total)) <= 2e-5 * abs(float(total))
        ref.compute_loss(loss_list, it)
        w = ref.update_weight(it, 3.0)
        st.finish()   # the host-side update of a step is otherwise applied while the next step's kernels run
        assert np.allclose(st.blc.w_list, ref.w_list, rtol=1e-4)


@pytest.mark.gpu
def test_pipelined_loss_balancing_equals_the_synchronous_sequence():
    """The host-side weight update of step i-1 is applied while step i's heavy kernels run (only the last
    kernel reads the weights): totals, weights and gradients of every step equal those of a run that
    finishes each update before starting the next step."""
    opt = S.default_opt(2, 64, 96, num_depth_bins=32, matching_channels=32)
    batches = []
    for i in range(2):
        b = S.synthetic_batch(opt, seed=11)
        b["noise_mono"] = torch.randn(b["noise_mono"].shape, generator=torch.Generator().manual_seed(200 + i))
        b["mono_disp"] = (b["mono_disp"] * (1.0 + 0.05 * i)).clamp(0, 1)
        batches.append(b)
    runs = []
    for pipelined in (False, True):
        st = S.MalStep(opt, use_graph=True, num_train_data=64, lambda_for_adjust=3.0, slots=2)
        for i in range(2):
            st.load(batches[i], slot=i)
        seq = []
        for it in range(5):
            scalars, grads, _ = st(it % 2)
            if not pipelined:
                st.finish()
            torch.cuda.synchronize()
            seq.append((scalars.clone().cpu(), [g.clone().cpu() for g in grads]))
        st.finish()
        runs.append((seq, np.array(st.blc.w_list)))
    (a, wa), (b, wb) = runs
    assert np.all(np.isfinite(wa)) and np.array_equal(wa, wb)
    assert len({float(s[0][0]) for s in a}) > 1          # the weights did change the total from step to step
    for (sa, ga), (sb, gb) in zip(a, b):
        assert torch.equal(sa, sb)
        for x, y in zip(ga, gb):
            assert torch.equal(x, y)
