"""The matching branch of ResnetEncoderMatching (manydepth/networks/resnet_encoder.py:71-353,
dualrefine/networks/resnet_encoder.py, dynamicdepth/networks/resnet_encoder.py) on the fused
plane-sweep kernel.

`CostVolumeMatcher` carries the matching state and methods with the reference's names:
    compute_depth_bins(min_depth_bin, max_depth_bin)             :121-148
    match_features(current_feats, lookup_feats, relative_poses, K, invK)   :151-233
    compute_confidence_mask(cost_volume, num_bins_threshold=None)         :255-262
    indices_to_disparity(indices)                                         :247-253
    matching_head(...)   = the no_grad block + argmin + masking of forward() :292-317 in ONE launch

`ResnetEncoderMatching` is the full encoder: torchvision ResNet blocks (cuDNN) around the matcher,
same constructor and forward() signature / return triple as the reference.

Differences that are deliberate: no (bins,1,h,w) `warp_depths` planes are built (the
kernel takes the 96 scalars), no host sync on `lookup_pose.sum() == 0` or `depth_bins[idx.cpu()]`.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops, raw


class CostVolumeMatcher:
    def __init__(self, num_depth_bins=96, min_depth_bin=0.1, max_depth_bin=20.0, adaptive_bins=False,
                 depth_binning="linear", convention=raw.CONV_MANYDEPTH):
        self.num_depth_bins = num_depth_bins
        self.adaptive_bins = adaptive_bins
        self.depth_binning = depth_binning
        self.set_missing_to_max = True
        self.convention = convention
        self.depth_bins = None
        self.compute_depth_bins(min_depth_bin, max_depth_bin)

    # -- bins ------------------------------------------------------------------------------
    def compute_depth_bins(self, min_depth_bin, max_depth_bin):
        """Depth hypotheses: linear in depth, in inverse depth, or log-spaced (host side, 96 scalars)."""
        as_float = lambda v: float(v.detach().reshape(()).item()) if torch.is_tensor(v) else float(v)
        lo, hi = as_float(min_depth_bin), as_float(max_depth_bin)
        n = self.num_depth_bins
        if self.depth_binning == "inverse":
            bins = 1 / np.linspace(1 / hi, 1 / lo, n)[::-1]
            bins = torch.from_numpy(np.ascontiguousarray(bins)).float()
        elif self.depth_binning == "linear":
            bins = torch.linspace(lo, hi, n)
        elif self.depth_binning == "log":
            base, it = torch.log(torch.tensor(lo)), torch.log(torch.tensor(hi / lo))
            bins = torch.exp(torch.Tensor([base + it * i / n for i in range(n)]))
        else:
            raise NotImplementedError(self.depth_binning)
        self.depth_bins = bins
        self._bins_dev = {}
        return bins

    def _bins_on(self, device):
        key = str(device)
        if key not in self._bins_dev:
            self._bins_dev[key] = self.depth_bins.to(device)
        return self._bins_dev[key]

    # -- volume ----------------------------------------------------------------------------
    def match_features(self, current_feats, lookup_feats, relative_poses, K, invK):
        """(B,C,h,w), (B,F,C,h,w), (B,F,4,4), (B,4,4) x2 -> cost volume and missing mask (B,bins,h,w)."""
        cv, missing, *_ = ops.cost_volume(current_feats, lookup_feats, relative_poses, K, invK,
                                          self._bins_on(current_feats.device), convention=self.convention,
                                          set_missing_to_max=self.set_missing_to_max)
        return cv, missing

    def compute_confidence_mask(self, cost_volume, num_bins_threshold=None):
        if num_bins_threshold is None:
            num_bins_threshold = self.num_depth_bins
        return ((cost_volume > 0).sum(1) == num_bins_threshold).float()

    def indices_to_disparity(self, indices):
        """arg-min bin indices (B,h,w) -> 1 / depth, gathered on the device."""
        bins = self._bins_on(indices.device)
        return 1 / bins[indices.reshape(-1).long()].reshape(indices.shape)

    def matching_head(self, current_feats, lookup_feats, relative_poses, K, invK):
        """forward() :292-317 in one launch -> (cost_volume * confidence, lowest_cost, confidence)."""
        cv, _, conf, _, low = ops.cost_volume(current_feats, lookup_feats, relative_poses, K, invK,
                                              self._bins_on(current_feats.device), convention=self.convention,
                                              set_missing_to_max=self.set_missing_to_max, apply_confidence=True)
        return cv, low, conf


class DynamicCostVolumeMatcher(CostVolumeMatcher):
    """DynamicDepth's matcher (dynamicdepth/networks/resnet_encoder.py:148-249): `match_features`
    takes the lookup image and the occlusion options; min over lookup frames (cv_min) and the
    set-to-1 / 3-D max-pool fill of occluded warped features happen inside the sweep kernel."""

    def occlusion_mask(self, lookup_images, h, w):
        """(B,3,H,W) DOMD-processed lookup image -> (B,h,w) {0,1}: black pixels (< 0.15 summed RGB),
        nearest-resized to the matching resolution (:160; the reference hard-codes 48x128)."""
        occ = torch.nn.functional.interpolate((lookup_images.sum(1).unsqueeze(1) < 0.15).float(), [h, w])
        return (occ[:, 0] > 0).float()

    def match_features(self, current_feats, lookup_feats, relative_poses, K, invK, lookup_images=None, cv_min=False,
                       aug_mask=None, set_1=False, pool=False, pool_r=1, pool_th=0.7):
        h, w = current_feats.shape[-2:]
        mode = raw.OCC_SET_1 if set_1 else (raw.OCC_POOL if pool else raw.OCC_NONE)
        occ = self.occlusion_mask(lookup_images, h, w) if mode != raw.OCC_NONE else None
        cv, missing, *_ = ops.cost_volume(current_feats, lookup_feats, relative_poses, K, invK,
                                          self._bins_on(current_feats.device), convention=self.convention,
                                          set_missing_to_max=self.set_missing_to_max, cv_min=bool(cv_min), occ=occ,
                                          occ_mode=mode, pool_radius=pool_r, pool_th=pool_th, aug_mask=aug_mask)
        return cv, missing


class ResnetEncoderMatching(nn.Module, CostVolumeMatcher):
    """ResNet encoder with the cost volume after the 2nd block (reference constructor/forward)."""

    def __init__(self, num_layers, pretrained, input_height, input_width, min_depth_bin=0.1,
                 max_depth_bin=20.0, num_depth_bins=96, adaptive_bins=False, depth_binning="linear",
                 convention=raw.CONV_MANYDEPTH):
        nn.Module.__init__(self)
        CostVolumeMatcher.__init__(self, num_depth_bins, min_depth_bin, max_depth_bin, adaptive_bins,
                                   depth_binning, convention)
        import torchvision.models as models
        if pretrained:
            raise ValueError("pretrained ImageNet weights need network access; load a state_dict instead")
        resnets = {18: models.resnet18, 34: models.resnet34, 50: models.resnet50, 101: models.resnet101,
                   152: models.resnet152}
        if num_layers not in resnets:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        self.matching_height, self.matching_width = input_height // 4, input_width // 4
        encoder = resnets[num_layers](weights=None)
        self.layer0 = nn.Sequential(encoder.conv1, encoder.bn1, encoder.relu)
        self.layer1 = nn.Sequential(encoder.maxpool, encoder.layer1)
        self.layer2, self.layer3, self.layer4 = encoder.layer2, encoder.layer3, encoder.layer4
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4
        self.reduce_conv = nn.Sequential(
            nn.Conv2d(int(self.num_ch_enc[1]) + self.num_depth_bins, out_channels=int(self.num_ch_enc[1]),
                      kernel_size=3, stride=1, padding=1), nn.ReLU(inplace=True))

    def feature_extraction(self, image, return_all_feats=False):
        image = (image - 0.45) / 0.225
        feats_0 = self.layer0(image)
        feats_1 = self.layer1(feats_0)
        return [feats_0, feats_1] if return_all_feats else feats_1

    def forward(self, current_image, lookup_images, poses, K, invK, min_depth_bin=None, max_depth_bin=None):
        self.features = self.feature_extraction(current_image, return_all_feats=True)
        current_feats = self.features[-1]
        with torch.no_grad():
            if self.adaptive_bins:
                self.compute_depth_bins(min_depth_bin, max_depth_bin)
            batch_size, num_frames, chns, height, width = lookup_images.shape
            lookup_feats = self.feature_extraction(lookup_images.reshape(batch_size * num_frames, chns, height, width))
            _, chns, height, width = lookup_feats.shape
            lookup_feats = lookup_feats.reshape(batch_size, num_frames, chns, height, width)
            cost_volume, lowest_cost, confidence_mask = self.matching_head(current_feats, lookup_feats, poses, K, invK)
        post_matching_feats = self.reduce_conv(torch.cat([self.features[-1], cost_volume], 1))
        self.features.append(self.layer2(post_matching_feats))
        self.features.append(self.layer3(self.features[-1]))
        self.features.append(self.layer4(self.features[-1]))
        return self.features, lowest_cost, confidence_mask
