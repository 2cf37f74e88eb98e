"""Data-parallel training-step harness around the hot path.

The reference trains with HuggingFace accelerate -> torch DistributedDataParallel over NCCL
(manydepth/trainer.py:309,469): one process per GPU, batch sharded, conv-net gradients
all-reduced (averaged) during backward.  The hot-path kernels need no exchange - every loss term
is per-sample and the masked means are per-rank (loss_utils.py:112-113) - so this module only
wires them into that topology:

    StandInNets      small conv stand-ins for RepDepth's networks (manydepth/networks/repdepth.py:
                     247-338): teacher disparity, student disparity (fed by the cost volume), matching
                     features and the two relative poses.  cuDNN does the convolutions; the real
                     ResNet encoders/decoders are out of scope.
    train_step       networks -> cost-volume head -> MAL losses (trainer_ops.process_batch_losses)
                     -> backward (DDP all-reduce) -> optimizer step.
    init_distributed torch.distributed from the torchrun environment (nccl on GPUs, gloo on CPU).

Launch:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 -m mal_b200.ddp
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import loss_utils, ops, trainer_ops
from .pose import transformation_from_parameters


class StandInNets(nn.Module):
    """Tiny convolutional stand-ins producing the tensors the hot path consumes."""

    def __init__(self, matching_channels=64, num_depth_bins=96, ballast_params=0):
        """`ballast_params` > 0 pads the pose head with dense layers of about that many parameters, so that the
        gradient all-reduce has the reference's volume (3 x ResNet18 + decoders, ~40 M fp32 parameters = 160 MB,
        SURVEY.md section 8e) without rebuilding its conv nets; the layers sit on the path to the poses and get
        real gradients."""
        super().__init__()
        conv = lambda i, o, s=1: nn.Sequential(nn.Conv2d(i, o, 3, s, 1), nn.ELU())
        self.mono = nn.Sequential(conv(3, 16, 2), conv(16, 16), nn.Upsample(scale_factor=2, mode="nearest"),
                                  nn.Conv2d(16, 1, 3, 1, 1))
        self.feat = nn.Sequential(conv(3, 32, 2), conv(32, matching_channels, 2))           # 1/4 resolution
        self.multi = nn.Sequential(conv(matching_channels + num_depth_bins, 32),
                                   nn.Upsample(scale_factor=4, mode="nearest"), nn.Conv2d(32, 1, 3, 1, 1))
        head = [nn.Linear(16, 6)]
        if ballast_params > 0:
            d1 = 4096
            d2 = max(64, int(ballast_params // d1))
            head = [nn.Linear(16, d1), nn.ELU(), nn.Linear(d1, d2), nn.ELU(), nn.Linear(d2, 6)]
        self.pose = nn.Sequential(conv(6, 16, 4), conv(16, 16, 4), nn.AdaptiveAvgPool2d(1), nn.Flatten(), *head)

    def predict_pose(self, a, b, invert):
        out = 0.01 * self.pose(torch.cat([a, b], 1)).view(-1, 1, 2, 3)   # repdepth.py:141-170, pose_decoder scale
        return transformation_from_parameters(out[:, :, 0], out[:, :, 1], invert=invert)

    def forward(self, inputs, bins, opt):
        """-> (mono_outputs, outputs) with the reference's dict keys (repdepth.py:247-338)."""
        cur = inputs[("color", 0, 0)]
        mono_outputs, outputs = {}, {}
        T = {-1: self.predict_pose(inputs[("color", -1, 0)], cur, invert=True),
             1: self.predict_pose(cur, inputs[("color", 1, 0)], invert=False)}
        for f in (-1, 1):
            mono_outputs[("cam_T_cam", 0, f)] = outputs[("cam_T_cam", 0, f)] = T[f]
        mono_outputs[("disp", 0)] = torch.sigmoid(self.mono(cur))
        feats = self.feat(cur)
        with torch.no_grad():                                               # resnet_encoder.py:292-307
            look = self.feat(inputs[("color", -1, 0)]).unsqueeze(1)
            cv, _, conf, _, low = ops.cost_volume(feats, look, T[-1].detach().unsqueeze(1), inputs[("K", 2)],
                                                  inputs[("inv_K", 2)], bins, apply_confidence=True)
        outputs[("disp", 0)] = torch.sigmoid(self.multi(torch.cat([feats, cv], 1)))
        outputs["lowest_cost"], outputs["consistency_mask"] = low, conf      # matching resolution
        outputs["cost_volume"] = cv
        B = cur.shape[0]
        outputs["augmentation_mask"] = torch.zeros(B, 1, 1, 1, device=cur.device)
        return mono_outputs, outputs


def init_distributed(backend=None):
    """Process group from the torchrun environment; returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cuda = torch.cuda.is_available()
    device = torch.device("cuda", local) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(device)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if cuda else "gloo"))
    return rank, world, device


def wrap(model, device):
    if dist.is_initialized() and dist.get_world_size() > 1:
        ids = [device.index] if device.type == "cuda" else None
        return nn.parallel.DistributedDataParallel(model, device_ids=ids)
    return model


def train_step(model, inputs, bins, opt, optimizer=None, loss_blc=None, index_iter=0, noises=None):
    """One data-parallel step.  Returns the losses dict (rank-local values, like the reference logs)."""
    mono_outputs, outputs = model(inputs, bins, opt)
    for f in (-1, 1):   # synthetic temporal-hint images stand in for dyn_utils.image_synthesis
        mono_outputs[("syn", f, 0)] = outputs[("syn", f, 0)] = inputs[("syn", f, 0)]
    outputs, losses = trainer_ops.process_batch_losses(inputs, mono_outputs, outputs, opt, has_ins=True,
                                                       loss_blc=loss_blc, index_iter=index_iter, noises=noises)
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    losses["loss"].backward()            # DDP all-reduces (averages) the network gradients here
    if optimizer is not None:
        optimizer.step()
    return losses


def train_step_fused(model, inputs, bins, opt, optimizer=None, loss_blc=None, index_iter=0, noises=None,
                     lambda_for_adjust=0.0):
    """The same step with the hot path as the fused libmal_b200 schedule (mal_b200.step.fused_step: ~17 launches,
    no autograd graph through the losses): the networks' outputs are the leaves, the kernels return d total / d leaf
    and autograd continues from there into the networks (and DDP's all-reduce).  CUDA only."""
    from . import _capi, step as S
    net = model.module if hasattr(model, "module") else model
    mono_outputs, outputs = model(inputs, bins, opt)
    B, H, W = opt.batch_size, opt.height, opt.width
    leaves = [mono_outputs[("disp", 0)], outputs[("disp", 0)], outputs[("cam_T_cam", 0, -1)], outputs[("cam_T_cam", 0, 1)]]
    dev = leaves[0].device
    noises = noises or [torch.randn(B, 1, H, W).to(dev, non_blocking=True) for _ in range(2)]   # CPU draw, like the reference
    b = {"color_0": inputs[("color", 0, 0)], "color_-1": inputs[("color", -1, 0)], "color_1": inputs[("color", 1, 0)],
         "syn_-1": inputs[("syn", -1, 0)], "syn_1": inputs[("syn", 1, 0)], "K": inputs[("K", 0)], "inv_K": inputs[("inv_K", 0)],
         "mono_disp": leaves[0].detach(), "multi_disp": leaves[1].detach(), "T_-1": leaves[2].detach(),
         "T_1": leaves[3].detach(), "augmentation_mask": outputs["augmentation_mask"], "noise_mono": noises[0],
         "noise_main": noises[1]}
    head = {"cost_volume": outputs["cost_volume"], "confidence": outputs["consistency_mask"],
            "lowest_cost": outputs["lowest_cost"]}
    w = None
    if loss_blc is not None:
        w = torch.tensor(loss_blc.w_list, dtype=torch.float32).to(dev, non_blocking=True)
    with torch.no_grad():
        scalars, grads, outs = S.fused_step(_capi.lib(), b, opt, w, head=head)
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    torch.autograd.backward(leaves, [g.reshape(l.shape) for g, l in zip(grads, leaves)])   # DDP all-reduces here
    if optimizer is not None:
        optimizer.step()
    sc = scalars.cpu()
    if loss_blc is not None:
        loss_blc.record_scores(index_iter, [float(sc[1]), float(sc[2])])
        loss_blc.update_weight(index_iter, lambda_for_adjust)
    return {"loss": sc[0], "reproj_loss/0": sc[4], "distil_loss": sc[2], "_outputs": outs}


def synthetic_inputs(opt, seed, device):
    from .step import synthetic_batch
    b = synthetic_batch(opt, seed=seed)
    inputs = {("color", 0, 0): b["color_0"], ("color", -1, 0): b["color_-1"], ("color", 1, 0): b["color_1"],
              ("syn", -1, 0): b["syn_-1"], ("syn", 1, 0): b["syn_1"], ("K", 0): b["K"], ("inv_K", 0): b["inv_K"],
              ("K", 2): b["K2"], ("inv_K", 2): b["inv_K2"]}
    return {k: v.to(device) for k, v in inputs.items()}, b["bins"].to(device)


def main():
    from .step import default_opt
    rank, world, device = init_distributed()
    opt = default_opt(int(os.environ.get("MAL_BATCH", "12")))
    torch.manual_seed(0)
    model = wrap(StandInNets(opt.matching_channels, opt.num_depth_bins).to(device), device)
    optim = torch.optim.Adam(model.parameters(), 1e-4)
    blc = loss_utils.LossBalancing(2, 1 << 16, opt.batch_size)
    inputs, bins = synthetic_inputs(opt, 1234 + rank, device)
    for it in range(int(os.environ.get("MAL_STEPS", "5"))):
        losses = train_step(model, inputs, bins, opt, optim, blc, it)
        if rank == 0:
            print("step %d  loss %.5f  reproj %.5f  distil %.5f" % (
                it, float(losses["loss"]), float(losses["reproj_loss/0"]), float(losses["distil_loss"])), flush=True)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
