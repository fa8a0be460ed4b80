"""Host-side mirror of the reference's DepthNet (models/TransMVSNet.py:33-109) on the fused kernels.

`DepthNet` keeps the reference's constructor, forward signature, return values and state_dict keys
(`pixel_wise_net.conv0.conv.weight`, ...), so it can replace `TransMVSNet.DepthNet` in the existing
cascade; only the body between the feature lists and the outputs runs on the sm_100a kernels:

  view loop  (:71-93)   -> ops.cost_volume  (fused warp + sampling + correlation + aggregation)
  read-out   (:99-103)  -> ops.softmax_wta  (softmax + WTA index/depth + confidence in one pass)

PixelwiseNet (:10-30), the 3-D CNN `cost_regularization` (:96) and everything outside DepthNet stay
PyTorch, as north_star prescribes.  In stage 1 (view_weights=None) the kernel emits the per-view
similarity, PixelwiseNet turns it into weights, and a small kernel aggregates.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .geometry import stage_rot_trans


class _ConvBnReLU3D(nn.Module):
    """1x1x1 Conv3d + BatchNorm3d + ReLU with the reference's parameter names (module.py:214-221)."""

    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, 1, stride=1, padding=0, bias=False)
        self.bn = nn.BatchNorm3d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)


class PixelwiseNet(nn.Module):
    """Learned per-pixel view weight, models/TransMVSNet.py:10-30 (stays in PyTorch)."""

    def __init__(self):
        super().__init__()
        self.conv0 = _ConvBnReLU3D(1, 16)
        self.conv1 = _ConvBnReLU3D(16, 8)
        self.conv2 = nn.Conv3d(8, 1, 1, stride=1, padding=0)
        self.output = nn.Sigmoid()

    def forward(self, x1):                       # [B,1,D,H,W] -> [B,1,H,W]
        x1 = self.conv2(self.conv1(self.conv0(x1))).squeeze(1)
        return torch.max(self.output(x1), dim=1, keepdim=True)[0]


class DepthNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.pixel_wise_net = PixelwiseNet()
        self._fold_key = None
        self._fold = None

    def _folded_mlp(self) -> torch.Tensor:
        """BatchNorm-folded PixelwiseNet parameters, recomputed only when a parameter or buffer changed."""
        tensors = list(self.pixel_wise_net.parameters()) + list(self.pixel_wise_net.buffers())
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self._fold_key:
            self._fold, self._fold_key = ops.fold_pixelwise_net(self.pixel_wise_net), key
        return self._fold

    def forward(self, features, proj_matrix, depth_values, num_depth, cost_regularization, view_weights=None):
        """Same contract as models/TransMVSNet.py:38-109.

        features: list of N [B,C,H,W]; proj_matrix [B,N,2,4,4]; depth_values [B,D,H,W];
        cost_regularization: nn.Module on [B,1,D,H,W]; view_weights None (stage 1) or [B,N-1,H,W].
        """
        assert len(features) == proj_matrix.shape[1], "Different number of images and projection matrices"
        assert depth_values.shape[1] == num_depth, \
            "depth_values.shape[1]:{}  num_depth:{}".format(depth_values.shape[1], num_depth)
        ref_feature, src_features = features[0], list(features[1:])
        rot_trans = stage_rot_trans(proj_matrix)                           # [Nsrc,B,12] on the host
        learned = view_weights is None
        if learned:
            _, sim_views = ops.cost_volume(ref_feature, src_features, rot_trans, depth_values, None, True)
            folded = (not self.training) and (not torch.is_grad_enabled())
            if folded:
                # inference: PixelwiseNet (BatchNorm folded) + aggregation in one kernel
                view_weights, similarity = ops.pixelwise_aggregate(sim_views, self._folded_mlp())
            else:
                weights = [self.pixel_wise_net(sim_views[i].unsqueeze(1)) for i in range(sim_views.shape[0])]
                view_weights = torch.cat(weights, dim=1)                   # [B,Nsrc,H,W]
                similarity = ops.aggregate(sim_views, view_weights)
        else:
            similarity, _ = ops.cost_volume(ref_feature, src_features, rot_trans, depth_values, view_weights, False)
        similarity = similarity.unsqueeze(1)                               # [B,1,D,H,W]

        cost_reg = cost_regularization(similarity)
        prob_volume_pre = cost_reg.squeeze(1)

        if torch.is_grad_enabled() and prob_volume_pre.requires_grad:
            # training: the loss needs d prob / d logits, keep that part in autograd; WTA via the kernel
            prob_volume = torch.exp(F.log_softmax(prob_volume_pre, dim=1))
            depth = ops.depth_wta(prob_volume, depth_values)
            with torch.no_grad():
                photo_confidence = torch.max(prob_volume, dim=1)[0]
        else:
            prob_volume, _, depth, photo_confidence = ops.softmax_wta(prob_volume_pre, depth_values)

        out = {"depth": depth, "photo_confidence": photo_confidence, "prob_volume": prob_volume,
               "depth_values": depth_values}
        if learned:
            return out, view_weights.detach()
        return out


def patch_reference(models_module=None, models_transmvsnet=None) -> None:
    """Bind the kernels into an imported reference tree (star import copies names, TransMVSNet.py:4)."""
    for mod in (models_module, models_transmvsnet):
        if mod is None:
            continue
        mod.homo_warping = ops.homo_warping
        mod.depth_wta = ops.depth_wta
        mod.depth_regression = ops.depth_regression
    if models_transmvsnet is not None:
        models_transmvsnet.DepthNet = DepthNet
