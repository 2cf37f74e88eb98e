"""Pin oracle/mal_oracle.py against the reference's own Python.  TEST INFRASTRUCTURE.

Runs ONLY in the build container, where /root/reference exists (the GPU box does not
have it).  It imports the reference's modules unmodified - with import stubs for
packages that are absent from this image (matplotlib, accelerate, detectron2, wandb,
skimage, torchmetrics, manydepth.pareto, manydepth.vis ...) and that the hot path never
executes - runs them on seeded synthetic inputs and asserts that the restatement in
mal_oracle.py is BITWISE identical.  tests/golden/make_golden.py reuses `load_reference`
to write the committed fixtures from the reference's outputs.

Usage:  python -m oracle.pin_against_reference
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from types import SimpleNamespace
from unittest import mock

import torch

REFERENCE_ROOT = os.environ.get("MAL_REFERENCE_ROOT", "/root/reference")

_STUB_PREFIXES = ("matplotlib", "accelerate", "detectron2", "wandb", "skimage", "torchmetrics",
                  "tensorboardX", "termcolor", "torch_sparse", "manydepth.pareto",
                  "manydepth.vis", "mask2former", "cv2", "IPython", "timm", "fvcore")


# stubbed even though a directory of that name exists: the vendored segmenter needs detectron2
_ALWAYS_STUB = ("manydepth.pareto", "manydepth.vis", "mask2former")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serve MagicMock-backed modules for imports the image lacks."""

    def find_spec(self, name, path=None, target=None):
        if any(name == p or name.startswith(p + ".") for p in _STUB_PREFIXES):
            try:
                for f in sys.meta_path:
                    if f is self:
                        continue
                    spec = f.find_spec(name, path, target) if hasattr(f, "find_spec") else None
                    if spec is not None and not name.startswith(_ALWAYS_STUB):
                        return None  # really installed
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__, m.__path__, m.__spec__ = spec.name, [], spec
        m.__loader__ = self
        # torchmetrics.Metric is subclassed by the trainer: give it a real base class.
        if spec.name == "torchmetrics":
            m.Metric = type("Metric", (torch.nn.Module,), {})
        return m

    def exec_module(self, module):
        pass


_loaded = None


def load_reference():
    """Import the reference packages (once) and return them in a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"{REFERENCE_ROOT} not present: pinning runs in the build container only")
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REFERENCE_ROOT)
    ns = SimpleNamespace()
    ns.layers = importlib.import_module("manydepth.layers")
    ns.loss_utils = importlib.import_module("manydepth.loss_utils")
    ns.resnet_encoder = importlib.import_module("manydepth.networks.resnet_encoder")
    ns.multilossmanager = importlib.import_module("manydepth.multilossmanager")
    ns.dualrefine_layers = importlib.import_module("dualrefine.layers")
    ns.dynamicdepth_layers = importlib.import_module("dynamicdepth.layers")
    try:
        ns.trainer = importlib.import_module("manydepth.trainer")
    except Exception as e:  # pragma: no cover - informational
        ns.trainer = None
        ns.trainer_error = repr(e)
    _loaded = ns
    return ns


# --------------------------------------------------------------------------
def reference_matcher(ref, height, width, bins):
    """A ResnetEncoderMatching shell with only the state match_features reads."""
    enc = ref.resnet_encoder.ResnetEncoderMatching.__new__(ref.resnet_encoder.ResnetEncoderMatching)
    torch.nn.Module.__init__(enc)
    h, w = height // 4, width // 4
    enc.num_depth_bins = len(bins)
    enc.matching_height, enc.matching_width = h, w
    enc.set_missing_to_max = True
    enc.depth_binning = "linear"
    enc.device = torch.device("cpu")
    enc.depth_bins = bins
    enc.warp_depths = torch.stack([torch.ones((1, h, w)) * d for d in bins], 0).float()
    enc.backprojector = ref.layers.BackprojectDepth(len(bins), h, w)
    enc.projector = ref.layers.Project3D(len(bins), h, w)
    return enc


def reference_trainer_shell(ref, batch, height, width, **opt):
    """A fake `self` good enough for Trainer's pure methods (manydepth/trainer.py)."""
    o = dict(sclm=0, v1_multiscale=False, height=height, width=width, min_depth=0.1,
             max_depth=100.0, frame_ids=[0, -1, 1], disable_automasking=False, temporal=False,
             main_temporal=False, no_ssim=False, disable_motion_masking=False,
             no_matching_augmentation=False, batch_size=batch, loss_pct=False, ensemble=False,
             disparity_smoothness=1e-3, debug=False)
    o.update(opt)
    shell = SimpleNamespace(opt=SimpleNamespace(**o), device=torch.device("cpu"),
                            ssim=ref.layers.SSIM(), has_ins=False, multi_has_ins=False, step=1,
                            is_main=False)
    shell.backproject_depth = {0: ref.layers.BackprojectDepth(batch, height, width)}
    shell.project_3d = {0: ref.layers.Project3D(batch, height, width)}
    T = ref.trainer.Trainer
    for name in ("generate_images_pred", "generate_images_pred_ensemble", "compute_reprojection_loss",
                 "compute_losses", "compute_matching_mask"):
        setattr(shell, name, types.MethodType(getattr(T, name), shell))
    shell.compute_loss_masks = T.compute_loss_masks
    return shell


def _eq(name, a, b):
    ok = torch.equal(a, b)
    print(f"  {'OK ' if ok else 'FAIL'} {name}: shape {tuple(a.shape)}"
          + ("" if ok else f"  max|d|={float((a - b).abs().max()):.3e}"))
    return ok


def reference_forward_warp():
    """dynamicdepth.rigid_warp.forward_warp from the reference, with the one third-party call it
    makes (torch_sparse.coalesce(op='max'), not installed, version unpinned) replaced by a
    restatement of its published semantics: sort the (row, col) indices, reduce duplicates by max."""
    load_reference()
    rw = importlib.import_module("dynamicdepth.rigid_warp")

    def coalesce(index, value, m, n, op="add"):
        assert op == "max"
        flat = index[0] * n + index[1]
        uniq, inv = torch.unique(flat, sorted=True, return_inverse=True)
        out = torch.full((uniq.numel(),), float("-inf"), dtype=value.dtype).scatter_reduce(0, inv, value, "amax")
        return torch.stack([uniq // n, uniq % n]), out

    rw.coalesce = coalesce
    return rw.forward_warp


def forward_warp_inputs(batch=2, height=24, width=64, seed=31):
    from mal_b200.utils.synthetic import CITYSCAPES_K, make_photometric_inputs
    from . import mal_oracle as O
    inputs, t = make_photometric_inputs(batch, height, width, seed=seed, normalised_K=CITYSCAPES_K,
                                        translation_scale=0.3)
    img = inputs[("color", 0, 0)].clone()
    img[:, :, : height // 3] = 0          # a "doj_mask"-like hole: forward_warp is fed masked images
    depth = O.disp_to_depth(t[("mono_disp", 0)], 0.1, 100.0)[1]
    pose = t[("cam_T_cam", 0, -1)][:, :3, :].clone()
    K = inputs[("K", 0)][:, :3, :3].clone()
    return img, depth, pose, K


def run_pin(batch=2, height=96, width=160, seed=7):
    """Bitwise comparison of the restatement with the reference on one seeded batch."""
    from mal_b200.utils.synthetic import make_photometric_inputs, make_cost_volume_inputs
    from oracle import mal_oracle as O
    ref = load_reference()
    ok = True
    inputs, t = make_photometric_inputs(batch, height, width, num_scales=2, seed=seed)

    # --- geometry + warp + SSIM ------------------------------------------------
    disp = O.upsample_disp(t[("mono_disp", 1)], height, width)
    _, depth = ref.layers.disp_to_depth(disp, 0.1, 100.0)
    ok &= _eq("disp_to_depth", O.disp_to_depth(disp, 0.1, 100.0)[1], depth)
    bp, pj = ref.layers.BackprojectDepth(batch, height, width), ref.layers.Project3D(batch, height, width)
    cam = bp(depth, inputs[("inv_K", 0)])
    ok &= _eq("backproject", O.backproject(depth, inputs[("inv_K", 0)]), cam)
    T = t[("cam_T_cam", 0, 1)]
    grid = pj(cam, inputs[("K", 0)], T)
    ok &= _eq("project3d", O.project3d(cam, inputs[("K", 0)], T, height, width), grid)
    pj2 = ref.dualrefine_layers.Project3D(batch, height, width)
    ok &= _eq("project3d(dualrefine)", O.project3d(cam, inputs[("K", 0)], T, height, width, O.DUALREFINE),
              pj2(cam, inputs[("K", 0)], T))
    ssim = ref.layers.SSIM()
    x, y = inputs[("color", 1, 0)], inputs[("color", 0, 0)]
    ok &= _eq("ssim", O.ssim(x, y), ssim(x, y))
    ok &= _eq("reprojection_loss", O.reprojection_loss(x, y), ref.loss_utils.compute_reprojection_loss(ssim, x, y))
    ok &= _eq("smooth_loss", O.smooth_loss(disp, y), ref.layers.get_smooth_loss(disp, y))
    for inv in (False, True):
        ok &= _eq(f"transformation_from_parameters(invert={inv})",
                  O.transformation_from_parameters(t[("axisangle", 1)], t[("translation", 1)], inv),
                  ref.layers.transformation_from_parameters(t[("axisangle", 1)], t[("translation", 1)], inv))

    # --- trainer glue + MAL losses --------------------------------------------
    if ref.trainer is None:
        print("  SKIP trainer glue:", ref.trainer_error)
        ok = False
    else:
        shell = reference_trainer_shell(ref, batch, height, width)
        mono_ref = {("disp", 0): t[("mono_disp", 0)].clone().requires_grad_(True)}
        mono_ora = {("disp", 0): t[("mono_disp", 0)].clone().requires_grad_(True)}
        for f in (-1, 1):
            mono_ref[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
            mono_ora[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
            mono_ref[("syn", f, 0)] = t[("syn", f, 0)]
            mono_ora[("syn", f, 0)] = t[("syn", f, 0)]
        shell.generate_images_pred(inputs, mono_ref)
        O.images_pred(inputs, mono_ora, height=height, width=width)
        for f in (-1, 1):
            ok &= _eq(f"generate_images_pred color {f}", mono_ora[("color", f, 0)], mono_ref[("color", f, 0)])
        for temporal in (False, True):
            torch.manual_seed(11)
            l_ref, mr_ref = ref.loss_utils.compute_mono_losses(ssim, inputs, mono_ref, temporal, True)
            torch.manual_seed(11)
            l_ora, mr_ora, _ = O.mono_losses(inputs, mono_ora, temporal, True)
            ok &= _eq(f"compute_mono_losses(temporal={temporal}) loss", l_ora["loss"], l_ref["loss"])
            ok &= _eq(f"compute_mono_losses(temporal={temporal}) mono_reproj", mr_ora, mr_ref)
        g_ref, = torch.autograd.grad(l_ref["loss"], mono_ref[("disp", 0)])
        g_ora, = torch.autograd.grad(l_ora["loss"], mono_ora[("disp", 0)])
        ok &= _eq("d mono loss / d disp", g_ora, g_ref)

        # student / multi path
        def multi_outputs():
            o = {("disp", 0): t[("multi_disp", 0)].clone().requires_grad_(True),
                 "consistency_mask": t["consistency_mask"], "augmentation_mask": t["augmentation_mask"],
                 ("mono_depth", 0, 0): mono_ref[("depth", 0, 0)].detach(),
                 ("mono_disp", 0): t[("mono_disp", 0)]}
            for f in (-1, 1):
                o[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
                o[("syn", f, 0)] = t[("syn", f, 0)]
            return o
        out_ref, out_ora = multi_outputs(), multi_outputs()
        disp_ens = (t[("mono_disp", 0)] + t[("multi_disp", 0)]) / 2.0
        ens_ref = shell.generate_images_pred_ensemble(inputs, t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)], disp_ens)
        ens_ora = O.images_pred_ensemble(inputs, t[("cam_T_cam", 0, -1)], t[("cam_T_cam", 0, 1)], disp_ens,
                                         height=height, width=width)
        ok &= _eq("generate_images_pred_ensemble", ens_ora, ens_ref)
        shell.generate_images_pred(inputs, out_ref, is_multi=True)
        O.images_pred(inputs, out_ora, height=height, width=width, is_multi=True)
        low = 1 / (mono_ref[("depth", 0, 0)].detach()[:, 0] * (0.5 + t["consistency_mask"] * 2.0))
        for o in (out_ref, out_ora):
            o["lowest_cost"] = low
        ok &= _eq("compute_matching_mask", O.matching_mask(out_ora), shell.compute_matching_mask(out_ref))
        for ens, has_ins, blc in ((ens_ref, False, True), (None, True, False)):
            opt = SimpleNamespace(batch_size=batch, dual_distil=False, learn_ens=False, pareto=False,
                                  loss_blc=blc, min_depth=0.1, max_depth=100.0)
            torch.manual_seed(5)
            l_ref, _, ll_ref = ref.loss_utils.compute_main_losses(ssim, inputs, out_ref, mr_ref.detach(), ens, opt,
                                                                  None, None, has_ins)
            torch.manual_seed(5)
            l_ora, _, ll_ora, _ = O.main_losses(inputs, out_ora, mr_ora.detach(), ens, batch_size=batch,
                                                multi_has_ins=has_ins, loss_blc=blc)
            tag = f"compute_main_losses(ens={'y' if ens is not None else 'n'},ins={has_ins},blc={blc})"
            for k in ("loss", "distil_loss", "reproj_loss/0", "consistency_loss/0"):
                ok &= _eq(f"{tag} {k}", l_ora[k], l_ref[k])
            tot_ref = l_ref["loss"] + (0.5 * ll_ref[1] if blc else 0)
            tot_ora = l_ora["loss"] + (0.5 * ll_ora[1] if blc else 0)
            g_ref, = torch.autograd.grad(tot_ref, out_ref[("disp", 0)], retain_graph=True)
            g_ora, = torch.autograd.grad(tot_ora, out_ora[("disp", 0)], retain_graph=True)
            ok &= _eq(f"{tag} d/d disp", g_ora, g_ref)

        # Trainer.compute_losses (non-distil path), 2 scales
        shell2 = reference_trainer_shell(ref, batch, height, width, sclm=1)
        def pyr(name):
            o = {("disp", s): t[(name + "_disp", s)] for s in range(2)}
            for f in (-1, 1):
                o[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
            return o
        p_ref, p_ora = pyr("mono"), pyr("mono")
        shell2.generate_images_pred(inputs, p_ref)
        O.images_pred(inputs, p_ora, num_scales=2, height=height, width=width)
        torch.manual_seed(3)
        l_ref, _ = shell2.compute_losses(inputs, p_ref, is_multi=False)
        torch.manual_seed(3)
        l_ora, _ = O.trainer_compute_losses(inputs, p_ora, num_scales=2, batch_size=batch)
        ok &= _eq("Trainer.compute_losses(2 scales) loss", l_ora["loss"], l_ref["loss"])

        # the two options that change what compute_losses compares: disable_automasking (identity still
        # compared, no tie-break noise: trainer.py:1292-1311) and v1_multiscale (:1089-1095, :1260-1263)
        import torch.nn.functional as F
        for s in (1,):
            for f in (-1, 1):
                inputs[("color", f, s)] = F.interpolate(inputs[("color", f, 0)], scale_factor=1 / 2 ** s, mode="area")
        for v1, no_auto in ((False, True), (True, False), (True, True)):
            shell3 = reference_trainer_shell(ref, batch, height, width, sclm=1, v1_multiscale=v1,
                                             disable_automasking=no_auto)
            for s in (1,):
                shell3.backproject_depth[s] = ref.layers.BackprojectDepth(batch, height // 2 ** s, width // 2 ** s)
                shell3.project_3d[s] = ref.layers.Project3D(batch, height // 2 ** s, width // 2 ** s)
            p_ref, p_ora = pyr("mono"), pyr("mono")
            shell3.generate_images_pred(inputs, p_ref)
            O.images_pred(inputs, p_ora, num_scales=2, height=height, width=width, v1_multiscale=v1)
            torch.manual_seed(3)
            l_ref, _ = shell3.compute_losses(inputs, p_ref, is_multi=False)
            torch.manual_seed(3)
            l_ora, _ = O.trainer_compute_losses(inputs, p_ora, num_scales=2, batch_size=batch, automask=not no_auto,
                                                v1_multiscale=v1)
            for k in ("loss", "reproj_loss/0", "reproj_loss/1"):
                ok &= _eq(f"Trainer.compute_losses(v1_multiscale={v1}, disable_automasking={no_auto}) {k}",
                          l_ora[k], l_ref[k])

    # --- cost volume -----------------------------------------------------------
    cv = make_cost_volume_inputs(batch, height, width, channels=16, num_bins=24, seed=seed,
                                 zero_pose_sample=batch - 1)
    enc = reference_matcher(ref, height, width, cv["bins"])
    cv_ref, miss_ref = enc.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"],
                                          cv["K"], cv["inv_K"])
    cv_ora, miss_ora = O.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"],
                                        cv["K"], cv["inv_K"], cv["bins"])
    ok &= _eq("match_features volume", cv_ora, cv_ref)
    ok &= _eq("match_features missing", miss_ora, miss_ref)
    ok &= _eq("compute_confidence_mask", O.confidence_mask(cv_ora * (1 - miss_ora)),
              enc.compute_confidence_mask(cv_ref * (1 - miss_ref)))
    viz = cv_ref.clone(); viz[viz == 0] = 100
    _, am = torch.min(viz, 1)
    ok &= _eq("lowest_cost", O.lowest_cost(cv_ora, cv["bins"])[0], enc.indices_to_disparity(am))

    # --- LossBalancing ---------------------------------------------------------
    lb_ref = ref.loss_utils.LossBalancing(2, 64, 4)
    lb_ora = O.LossBalancing(2, 64, 4)
    good = True
    gen = torch.Generator().manual_seed(1)
    for it in range(6):
        ll = [torch.rand((), generator=gen) + 0.1, torch.rand((), generator=gen) * 0.01 + 1e-3]
        with mock.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self):
            a = lb_ref.compute_loss(ll, it)
        b = lb_ora.compute_loss(ll, it)
        good &= bool(torch.equal(torch.as_tensor(a), torch.as_tensor(b)))
        good &= lb_ref.update_weight(it, 0.3) == lb_ora.update_weight(it, 0.3)
    print(f"  {'OK ' if good else 'FAIL'} LossBalancing (6 iterations, weights + loss)")
    ok &= good

    ok &= pin_host_side(ref, t)

    # ---- DynamicDepth forward_warp --------------------------------------------------------
    import warnings
    fw = reference_forward_warp()
    img, depth_fw, pose, K3 = forward_warp_inputs()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = fw(img, depth_fw, pose, K3, upscale=3)
    got = O.forward_warp(img, depth_fw, pose, K3, upscale=3)
    for n, a, b in zip(("img_w", "depth_w", "valid"), got, want):
        ok &= _eq("forward_warp " + n, a, b)
    # pose_vec2mat, both rotation modes (rigid_warp.py:243-284)
    rw = importlib.import_module("dynamicdepth.rigid_warp")
    vec = torch.randn(5, 6, generator=torch.Generator().manual_seed(12)) * 0.3
    for mode in ("euler", "quat"):
        ok &= _eq("pose_vec2mat " + mode, O.pose_vec2mat(vec, mode), rw.pose_vec2mat(vec, mode))
    return bool(ok)




def pin_host_side(ref, t):
    """The product's own host-side ports (no kernel behind them) against the reference modules:
    mal_b200/pose.py vs manydepth/layers.py:26-100, mal_b200/multilossmanager.py vs
    manydepth/multilossmanager.py:6-88."""
    from mal_b200 import pose
    from mal_b200.multilossmanager import MultiLossManager
    ok = True
    for inv in (False, True):
        ok &= _eq(f"mal_b200.pose.transformation_from_parameters(invert={inv})",
                  pose.transformation_from_parameters(t[("axisangle", 1)], t[("translation", 1)], inv),
                  ref.layers.transformation_from_parameters(t[("axisangle", 1)], t[("translation", 1)], inv))
    ok &= _eq("mal_b200.pose.rot_from_axisangle", pose.rot_from_axisangle(t[("axisangle", -1)]),
              ref.layers.rot_from_axisangle(t[("axisangle", -1)]))
    # MultiLossManager.rebalancing calls np.sum(tensor * tensor) (multilossmanager.py:62,71,83), which raises
    # TypeError with torch >= 2 / numpy 2 (numpy forwards axis=/out= to Tensor.sum).  The class is dead code in
    # the reference; pin everything else by running it with that ONE call shimmed to Tensor.sum.
    R = ref.multilossmanager
    try:
        R.MultiLossManager(2, 2, 4, "cpu").rebalancing(0.1, 0)
        raised = False
    except TypeError:
        raised = True
    print(f"  note reference MultiLossManager.rebalancing raises TypeError as shipped: {raised}")
    shim = SimpleNamespace(sum=lambda x: x.sum())
    m_ref, m_ours = R.MultiLossManager(2, 2, 8, "cpu"), MultiLossManager(2, 2, 8, "cpu")
    gen = torch.Generator().manual_seed(9)
    good = True
    with mock.patch.object(R, "np", shim):
        for epoch in range(4):
            for _ in range(3):
                losses = torch.rand(2, generator=gen) + 0.1
                a, pa = m_ref.get_total_loss(losses, 2)
                b, pb = m_ours.get_total_loss(losses, 2)
                good &= bool(torch.equal(a, b)) and pa == pb
            m_ref.rebalancing(0.4, epoch)
            m_ours.rebalancing(0.4, epoch)
            good &= bool(torch.equal(m_ref.loss_weights, m_ours.loss_weights)) and m_ref.cur_ptr == m_ours.cur_ptr
            good &= bool(torch.equal(torch.as_tensor(m_ref.previous_total_loss), torch.as_tensor(m_ours.previous_total_loss)))
    print(f"  {'OK ' if good else 'FAIL'} MultiLossManager (4 epochs: totals, pointers, weights, previous_total_loss)")
    return ok and good


def pin_dynamicdepth_match_features(seed=77):
    """dynamicdepth/networks/resnet_encoder.py:148-249 against O.match_features_dynamic.  The
    reference hard-codes 96 bins x 64 channels at 48x128 (:160, :193), so this runs at full size."""
    from mal_b200.utils.synthetic import CITYSCAPES_K, make_cost_volume_inputs
    from . import mal_oracle as O
    load_reference()
    dd = importlib.import_module("dynamicdepth.networks.resnet_encoder")
    dl = importlib.import_module("dynamicdepth.layers")
    H, W = 192, 512
    h, w = H // 4, W // 4
    cv = make_cost_volume_inputs(2, H, W, channels=64, num_lookup=2, num_bins=96, seed=seed, normalised_K=CITYSCAPES_K,
                                 min_bin=0.5, max_bin=6.0, translation_scale=0.5)
    gen = torch.Generator().manual_seed(seed + 1)
    look_img = torch.rand(2, 3, H, W, generator=gen)
    look_img[:, :, 60:130, 100:260] = 0.0       # DOMD holes are black
    look_img[:, :, 20:50, 300:420] = 0.01
    aug = torch.zeros(2, 1, 1, 1)
    aug[1] = 1
    bins = cv["bins"]
    enc = dd.ResnetEncoderMatching.__new__(dd.ResnetEncoderMatching)
    torch.nn.Module.__init__(enc)
    enc.num_depth_bins, enc.matching_height, enc.matching_width, enc.set_missing_to_max = len(bins), h, w, True
    enc.depth_bins = bins
    enc.warp_depths = torch.stack([torch.ones((1, h, w)) * d for d in bins], 0).float()
    enc.backprojector, enc.projector = dl.BackprojectDepth(len(bins), h, w), dl.Project3D(len(bins), h, w)
    ok = True
    for cv_min, set_1, pool in ((True, False, True), (False, True, False), (True, False, False)):
        want = enc.match_features(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"], cv["inv_K"],
                                  look_img, cv_min, aug, set_1, pool, 1, 0.7)
        got = O.match_features_dynamic(cv["current_feats"], cv["lookup_feats"], cv["relative_poses"], cv["K"],
                                       cv["inv_K"], bins, look_img, cv_min, aug, set_1, pool, 1, 0.7)
        tag = f"dynamicdepth match_features(cv_min={cv_min}, set_1={set_1}, pool={pool})"
        ok &= _eq(tag + " volume", got[0], want[0])
        ok &= _eq(tag + " missing", got[1], want[1])
    return bool(ok)


def dynamicdepth_loss_case(batch=1, height=32, width=64, seed=55, holes=True):
    """Inputs for DynamicDepth's compute_losses: 4-scale disparities, materialised warps with black
    DOMD-style holes, identity frames."""
    from mal_b200.utils.synthetic import make_photometric_inputs
    from . import mal_oracle as O
    inputs, t = make_photometric_inputs(batch, height, width, num_scales=4, seed=seed, translation_scale=0.3)
    outs = {}
    for name in ("mono", "multi"):
        o = {("disp", s): t[(name + "_disp", s)] for s in range(4)}
        for f in (-1, 1):
            o[("cam_T_cam", 0, f)] = t[("cam_T_cam", 0, f)]
        O.images_pred(inputs, o, num_scales=4, height=height, width=width, is_multi=(name == "multi"))
        if holes:
            for s in range(4):
                o[("color", -1, s)] = o[("color", -1, s)].clone()
                o[("color", 1, s)] = o[("color", 1, s)].clone()
                o[("color", -1, s)][:, :, 4:12, 6:30] = 0.0
                o[("color", 1, s)][:, :, 8:20, 20:44] = 0.0
        outs[name] = o
    outs["multi"]["consistency_mask"] = t["consistency_mask"]
    outs["multi"]["augmentation_mask"] = t["augmentation_mask"]
    for s in range(4):
        outs["multi"][("mono_depth", 0, s)] = outs["mono"][("depth", 0, s)]
    return inputs, t, outs


def pin_dynamicdepth_losses():
    """dynamicdepth/trainer.py Trainer.compute_losses (:1006-1128) called unbound on a shell `self`."""
    from . import mal_oracle as O
    load_reference()
    tr = importlib.import_module("dynamicdepth.trainer")
    dl = importlib.import_module("dynamicdepth.layers")
    ok = True
    for is_multi, selec, zero in ((False, True, True), (True, True, True), (False, False, False)):
        res = []
        for who in ("ref", "oracle"):
            inputs, t, outs = dynamicdepth_loss_case()
            o = outs["multi" if is_multi else "mono"]
            noises = t["noise"] + t["noise"]
            if who == "ref":
                opt = SimpleNamespace(scales=[0, 1, 2, 3], v1_multiscale=False, frame_ids=[0, -1, 1],
                                      disable_automasking=False, no_teacher_warp=False, train_teacher_only=False,
                                      avg_reprojection=False, selec_reproj=selec, zero_img=zero, no_ssim="false",
                                      disable_motion_masking=False, no_matching_augmentation="false",
                                      disparity_smoothness=1e-3, feat_loss="false")
                shell = SimpleNamespace(opt=opt, ssim=dl.SSIM(), device=torch.device("cpu"), num_scales=4)
                shell.compute_reprojection_loss = lambda p, tg: tr.Trainer.compute_reprojection_loss(shell, p, tg)
                shell.compute_loss_masks = tr.Trainer.compute_loss_masks
                it = iter(noises)
                with mock.patch.object(torch, "randn", lambda *a, **k: next(it)):
                    losses = tr.Trainer.compute_losses(shell, inputs, o, is_multi=is_multi)
            else:
                losses, _ = O.dynamicdepth_compute_losses(inputs, o, (0, 1, 2, 3), is_multi=is_multi,
                                                          selec_reproj=selec, zero_img=zero, noises=noises)
            res.append((losses, inputs[("color", 0, 0)].clone()))
        tag = f"dynamicdepth compute_losses(multi={is_multi}, selec={selec}, zero={zero})"
        for k in res[0][0]:
            ok &= _eq(f"{tag} {k}", res[1][0][k].detach(), res[0][0][k].detach())
        ok &= _eq(tag + " mutated target", res[1][1], res[0][1])
    return bool(ok)


def pin_image_synthesis():
    """manydepth/dyn_utils.py generate_dynamic_instance / fill_dynamic_obj (TorchScript) against the
    oracle restatement, on synthetic Mask2Former-shaped matched masks."""
    from mal_b200.utils.synthetic import make_instance_masks
    from . import mal_oracle as O
    load_reference()
    du = importlib.import_module("manydepth.dyn_utils")
    ok = True
    for N, H, W, seed, empty, replace in ((6, 48, 96, 1, None, False), (5, 40, 64, 2, 1, True),
                                          (17, 32, 48, 3, None, False), (3, 24, 40, 4, 0, False)):
        ml, mn = make_instance_masks(N, H, W, seed=seed, max_shift=9, empty=empty)
        g = torch.Generator().manual_seed(seed)
        il, inx = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
        gh, gw = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        want = du.generate_dynamic_instance(gh, gw, ml, mn, il, inx, replace)
        got = O.generate_dynamic_instance(ml, mn, il, inx, replace)
        ok &= _eq(f"generate_dynamic_instance N={N} last", got[0], want[0])
        ok &= _eq(f"generate_dynamic_instance N={N} next", got[1], want[1])
        dx, dy = torch.randint(-5, 6, (N,), generator=g), torch.randint(-7, 8, (N,), generator=g)
        ok &= _eq(f"fill_dynamic_obj N={N}", O.fill_dynamic_obj(ml, dx, dy, il, inx), du.fill_dynamic_obj(ml, dx, dy, il, inx))
    return bool(ok)


def corr_case(batch=2, channels=64, height=12, width=24, levels=3, samples=5, seed=91, spread=2.5):
    """Seeded inputs for the DualRefine correlation lookup: features >= 0 and epipolar candidates scattered
    around each pixel (some leave the map: zeros padding is exercised)."""
    g = torch.Generator().manual_seed(seed)
    fmap1 = torch.rand(batch, channels, height, width, generator=g)
    fmap2 = torch.rand(batch, channels, height, width, generator=g)
    ys, xs = torch.meshgrid(torch.arange(height).float(), torch.arange(width).float(), indexing="ij")
    base = torch.stack([xs, ys])[None, :, None, None]
    coords = base + spread * torch.randn(batch, 2, levels, samples, height, width, generator=g)
    return fmap1, fmap2, coords


def pin_corr_lookup():
    """dualrefine/networks/corr.py CoordSampler (imported by file path: its own imports are torch-only,
    SURVEY.md 8c) against the oracle restatement, with and without ATen's reduction tail (h*w*D % 32)."""
    import importlib.util
    from . import mal_oracle as O
    spec = importlib.util.spec_from_file_location("ref_dualrefine_corr", os.path.join(REFERENCE_ROOT, "dualrefine", "networks", "corr.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ok = True
    for (B, Cn, h, w, L, D, heads) in ((2, 64, 12, 24, 3, 5, 1), (1, 64, 8, 16, 2, 17, 1), (2, 32, 12, 20, 3, 4, 2)):
        fmap1, fmap2, coords = corr_case(B, Cn, h, w, L, D)
        cs = mod.CoordSampler(None)
        cs.register(fmap1, fmap2, num_levels=L)
        want = cs(coords, num_levels=L, num_head=heads)
        pyr = O.corr_pyramid(fmap2, L)
        ok &= _eq(f"corr pyramid {h}x{w}", torch.cat([p.reshape(-1) for p in pyr]),
                  torch.cat([p.reshape(-1) for p in cs.f2_pyramid]))
        ok &= _eq(f"corr_lookup {Cn}ch {h}x{w} L={L} D={D} heads={heads}", O.corr_lookup(fmap1, pyr, coords, heads), want)
        if heads == 1:
            ok &= _eq(f"__corr__ {h}x{w}", O.corr_lookup(fmap1, pyr, coords, 1), cs.__corr__(coords, num_levels=L))
    return bool(ok)


def pin_sample_tgt():
    """PoseUpdate.sample_tgt: dualrefine/networks/utils/utils.py does not import outside its package
    (SURVEY.md 8c), so the method's own source lines are cut out of the file and executed as they are."""
    import ast
    import textwrap
    import types
    import torch.nn.functional as F
    from . import mal_oracle as O
    path = os.path.join(REFERENCE_ROOT, "dualrefine", "networks", "utils", "utils.py")
    src = open(path).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "sample_tgt")
    code = textwrap.dedent("\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno]))
    ns = {"torch": torch, "F": F}
    exec(compile(code, path, "exec"), ns)
    g = torch.Generator().manual_seed(23)
    B, Cn, h, w = 2, 16, 10, 14
    feat, wmap = torch.rand(B, Cn, h, w, generator=g), torch.rand(B, 1, h, w, generator=g)
    ys, xs = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    c1 = torch.stack([xs, ys])[None, :, None, None] + 1.5 * torch.randn(B, 2, 1, 1, h, w, generator=g)
    delta = torch.tensor([[0., 1., -1., 0., 0.], [0., 0., 0., 1., -1.]]).reshape(1, 2, 1, 5, 1, 1)
    p2 = c1 + delta
    shell = types.SimpleNamespace(tgt_w=wmap)
    want_feat, want_grad = ns["sample_tgt"](shell, feat, p2)
    got = O.sample_tgt(feat, p2, wmap)
    ok = _eq("sample_tgt feat", got[0], want_feat)
    ok &= _eq("sample_tgt gradients", got[1], want_grad)
    ok &= _eq("sample_tgt weight", got[2], shell.warped_tgt_w)
    return bool(ok)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    good = run_pin()
    good &= pin_dynamicdepth_match_features()
    good &= pin_image_synthesis()
    good &= pin_dynamicdepth_losses()
    good &= pin_corr_lookup()
    good &= pin_sample_tgt()
    print("PINNED" if good else "PIN FAILED")
    sys.exit(0 if good else 1)
