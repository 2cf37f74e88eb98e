"""world_size-2 gloo run of the data-parallel training step (mal_b200/ddp.py) on CPU.

The kernels run through the host-emulated twin (test hook, see conftest.op_device); what is
checked is the N>1 host logic: per-rank batches, per-rank masked means, DDP averaging of the
network gradients == the mean of the two single-process gradients, identical parameters after
the optimizer step on both ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mal_b200 import ddp, ops
from mal_b200.step import default_opt


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _use_emulator():
    from tests.emu.emu_lib import emu
    handle = emu()
    ops._lib = lambda t: handle   # test hook: the package itself refuses CPU tensors


def _opt():
    return default_opt(1, 32, 64, num_depth_bins=8, matching_channels=16, loss_blc=False)


def _grads(model):
    return [p.grad.clone() for p in model.parameters()]


def _single(seed_rank):
    torch.manual_seed(0)
    opt = _opt()
    model = ddp.StandInNets(opt.matching_channels, opt.num_depth_bins)
    inputs, bins = ddp.synthetic_inputs(opt, 100 + seed_rank, torch.device("cpu"))
    noises = [torch.randn(1, 1, 32, 64, generator=torch.Generator().manual_seed(7 + seed_rank)) for _ in range(2)]
    losses = ddp.train_step(model, inputs, bins, opt, noises=noises)
    return float(losses["loss"]), _grads(model)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    _use_emulator()
    r, w, device = ddp.init_distributed("gloo")
    assert (r, w, device.type) == (rank, world, "cpu")
    torch.manual_seed(0)
    opt = _opt()
    model = ddp.wrap(ddp.StandInNets(opt.matching_channels, opt.num_depth_bins), device)
    optim = torch.optim.SGD(model.parameters(), 0.1)
    inputs, bins = ddp.synthetic_inputs(opt, 100 + rank, device)      # each rank has its own batch
    noises = [torch.randn(1, 1, 32, 64, generator=torch.Generator().manual_seed(7 + rank)) for _ in range(2)]
    losses = ddp.train_step(model, inputs, bins, opt, optim, noises=noises)
    grads = _grads(model.module)
    params = [p.detach().clone() for p in model.module.parameters()]
    torch.save({"loss": float(losses["loss"]), "grads": grads, "params": params}, os.path.join(out, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_training_step(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"r{i}.pt") for i in range(2))
    real = ops._lib
    _use_emulator()
    try:
        (l0, g0), (l1, g1) = _single(0), _single(1)
    finally:
        ops._lib = real   # drop the test hook
    # rank-local losses are the single-process losses of that rank's batch (no loss collective)
    assert abs(r0["loss"] - l0) <= 1e-6 * abs(l0) and abs(r1["loss"] - l1) <= 1e-6 * abs(l1)
    assert abs(l0 - l1) > 1e-6          # the two ranks really saw different data
    for a, b, x, y in zip(r0["grads"], r1["grads"], g0, g1):
        assert torch.allclose(a, b, rtol=0, atol=0)                        # all-reduced: identical on both ranks
        assert torch.allclose(a, (x + y) / 2, rtol=1e-5, atol=1e-8)        # DDP averages the per-rank gradients
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)


def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, device = ddp.init_distributed("nccl")
    torch.manual_seed(0)
    opt = default_opt(2, 64, 96, num_depth_bins=16, matching_channels=32, loss_blc=False)
    model = ddp.wrap(ddp.StandInNets(opt.matching_channels, opt.num_depth_bins, ballast_params=200_000).to(device), device)
    optim = torch.optim.SGD(model.parameters(), 0.05)
    inputs, bins = ddp.synthetic_inputs(opt, 100 + rank, device)      # each rank has its own batch
    losses = []
    for it in range(3):   # first the op-by-op autograd path, then the fused schedule
        step = ddp.train_step if it == 0 else ddp.train_step_fused
        losses.append(float(step(model, inputs, bins, opt, optim)["loss"]))
    params = [p.detach().cpu().clone() for p in model.module.parameters()]
    torch.save({"losses": losses, "params": params}, os.path.join(out, f"n{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_two_rank_nccl_training_steps_end_with_identical_parameters(tmp_path):
    """The data-parallel step on hardware: two ranks (one GPU each) under NCCL, three optimizer steps with the
    CUDA kernels on the path; rank-local losses differ (no loss collective), parameters end identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mp.spawn(_nccl_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"n{i}.pt") for i in range(2))
    assert all(abs(a - b) > 1e-7 for a, b in zip(r0["losses"], r1["losses"]))
    assert all(l == l for l in r0["losses"] + r1["losses"])
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)
