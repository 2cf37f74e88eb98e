"""Host-side pieces of the boundary that carry no kernel: the pose helpers and
MultiLossManager (a17), against the pinned oracle / a numpy restatement of the reference's arithmetic.
oracle/pin_against_reference.py pins the same two against the reference's own modules."""
import numpy as np
import torch

from mal_b200 import pose
from mal_b200.multilossmanager import MultiLossManager
from oracle import mal_oracle as O


def test_pose_helpers_match_the_pinned_oracle():
    """manydepth/layers.py:26-100: bitwise, both `invert` branches, incl. a zero rotation."""
    g = torch.Generator().manual_seed(17)
    aa = torch.randn(5, 1, 3, generator=g) * 0.02
    aa[0] = 0.0                                            # angle 0: axis = vec / 1e-7
    tr = torch.randn(5, 1, 3, generator=g) * 0.1
    assert torch.equal(pose.rot_from_axisangle(aa), O.rot_from_axisangle(aa))
    assert torch.equal(pose.get_translation_matrix(tr), O.get_translation_matrix(tr))
    for inv in (False, True):
        assert torch.equal(pose.transformation_from_parameters(aa, tr, inv),
                           O.transformation_from_parameters(aa, tr, inv))
    # gradients reach the pose network's outputs
    a, t = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
    pose.transformation_from_parameters(a, t, True).sum().backward()
    assert a.grad is not None and t.grad is not None and torch.isfinite(a.grad).all()


def _rebalance_np(train_losses, cur_ptr, weights, state, lam, update_once=False):
    """manydepth/multilossmanager.py:58-84 restated in numpy fp32 (np.sum(a * w) as the reference
    writes it; on torch tensors that call raises with torch >= 2, see the pin script)."""
    mean = train_losses[:cur_ptr].mean(axis=0, dtype=np.float32)
    total = np.sum(mean * weights, dtype=np.float32)
    w = weights.copy()
    if not state["init"]:
        for k in range(len(w)):
            w[k] = (total * w[k]) / mean[k]
        state.update(init=True, prev_total=np.sum(mean * w, dtype=np.float32), prev=mean)
    elif not update_once:
        if state["prev_total"] > 0:
            for k in range(len(w)):
                adj = np.float32(1) + np.float32(lam) * ((total / state["prev_total"]) * (state["prev"][k] / mean[k]) - np.float32(1))
                adj = min(max(adj, np.float32(0.5)), np.float32(2.0))
                w[k] = w[k] * adj
        state.update(prev_total=np.sum(mean * w, dtype=np.float32), prev=mean)
    return w


def test_multilossmanager_follows_the_reference_arithmetic():
    B, n = 2, 2
    m = MultiLossManager(B, n, 8, "cpu")
    g = torch.Generator().manual_seed(4)
    weights = np.full(n, 0.5, np.float32)
    state = {"init": False}
    for epoch in range(4):
        rows = []
        for _ in range(3):
            losses = torch.rand(n, generator=g) + 0.1
            total, ptr = m.get_total_loss(losses, B)
            item = weights * losses.numpy()
            assert np.float32(total) == np.float32(item.sum(dtype=np.float32))
            rows += [item] * B
            assert ptr == len(rows)
        weights = _rebalance_np(np.stack(rows).astype(np.float32), len(rows), weights, state, 0.4)
        m.rebalancing(0.4, epoch)
        assert m.cur_ptr == 0
        np.testing.assert_array_equal(m.loss_weights.numpy(), weights)
    # weights_list overrides, update=False leaves the record alone
    total, ptr = m.get_total_loss(torch.tensor([1.0, 2.0]), B, update=False, weights_list=torch.tensor([0.25, 0.75]))
    assert float(total) == 1.75 and ptr == 0


def test_loss_balancing_zero_loss_term_follows_the_reference():
    """distil_loss == 0 in the first step: the reference divides by the zero mean (loss_utils.py:326), the
    weight becomes inf and previous_total_loss NaN (weights frozen for a step: NaN > 0 is False at :338), then
    inf / inf makes both weights NaN.
    The product mirrors that state for state (the oracle's LossBalancing is pinned against the reference)."""
    from mal_b200.loss_utils import LossBalancing
    ours, ora = LossBalancing(2, 64, 2), O.LossBalancing(2, 64, 2)
    seq = [(0.7, 0.0), (0.6, 0.01), (0.5, 0.02)]
    for it, (a, b) in enumerate(seq):
        ll = [torch.tensor(a), torch.tensor(b)]
        assert torch.equal(torch.as_tensor(ours.compute_loss(ll, it)), torch.as_tensor(ora.compute_loss(ll, it)))
        with np.errstate(divide="ignore", invalid="ignore"):
            want = ora.update_weight(it, 0.3)
        got = ours.update_weight(it, 0.3)
        assert np.array_equal(np.array(got), np.array(want), equal_nan=True), (it, got, want)
        assert np.array_equal(ours.previous_total_loss, ora.previous_total_loss, equal_nan=True)
        if it == 0:
            assert got[0] == 0.25 and np.isinf(got[1]) and np.isnan(ours.previous_total_loss)
    # step 1: NaN > 0 is False, the weights stay; step 2: inf / inf -> NaN reaches both weights
    assert np.isnan(got[0]) and np.isnan(got[1])


def test_pose_vec2mat_both_rotation_modes():
    """mal_b200.rigid_warp.pose_vec2mat / quat2mat against the pinned oracle (dynamicdepth/rigid_warp.py:243-284);
    a quaternion's matrix is orthonormal with determinant 1."""
    import torch
    from mal_b200 import rigid_warp
    from oracle import mal_oracle as O
    vec = torch.randn(7, 6, generator=torch.Generator().manual_seed(3)) * 0.4
    for mode in ("euler", "quat"):
        got, want = rigid_warp.pose_vec2mat(vec, mode), O.pose_vec2mat(vec, mode)
        assert torch.equal(got, want), mode
        R = got[:, :, :3]
        assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3).expand(7, 3, 3), atol=1e-5)
        assert torch.allclose(torch.linalg.det(R), torch.ones(7), atol=1e-5)
    import pytest
    with pytest.raises(ValueError):
        rigid_warp.pose_vec2mat(vec, "axis")
