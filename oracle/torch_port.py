"""PyTorch-ops restatement of the reference's cost-volume path (CPU baseline / second oracle).

TEST + BENCH-BASELINE INFRASTRUCTURE ONLY (same rules as oracle/oracle.py).

The reference is Python that runs ATen kernels; /root/reference does not exist on the GPU
box, so the "reference PyTorch CPU path" that bench.py times beside every GPU number is this
port: the same ATen op sequence (inverse, matmul, elementwise grid build, F.grid_sample,
mul+mean, weighted accumulation, log_softmax+exp, argmax+gather, max), written from the
algorithm description in SURVEY.md section 3.3 -- not copied.  It is pinned against
tests/golden/*.npz (outputs of the real reference) by tests/test_oracle_golden.py.

Reference lines restated: models/module.py:284-322 (warp), models/TransMVSNet.py:71-93
(view loop), :99-103 (read-out), models/module.py:474-482 (WTA).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def sampling_grid(src_proj: torch.Tensor, ref_proj: torch.Tensor, depth_values: torch.Tensor,
                  height: int, width: int) -> torch.Tensor:
    """Normalised grid_sample grid [B, D*H, W, 2] of module.py:294-316."""
    b, d = depth_values.shape[0], depth_values.shape[1]
    dev = depth_values.device
    with torch.no_grad():
        rel = src_proj @ torch.inverse(ref_proj)
        rot, trans = rel[:, :3, :3], rel[:, :3, 3:4]
        ys = torch.arange(height, dtype=torch.float32, device=dev).view(height, 1).expand(height, width)
        xs = torch.arange(width, dtype=torch.float32, device=dev).view(1, width).expand(height, width)
        pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, device=dev)))
        rays = rot @ pix.unsqueeze(0).expand(b, 3, height * width)          # [B,3,HW]
        pts = rays.unsqueeze(2) * depth_values.reshape(b, 1, d, -1)          # [B,3,D,HW]
        pts = pts + trans.reshape(b, 3, 1, 1)
        z = pts[:, 2]
        behind = z < 1e-6
        u = pts[:, 0] / z
        v = pts[:, 1] / z
        u = u / ((width - 1) / 2) - 1
        v = v / ((height - 1) / 2) - 1
        u[behind] = -99.0
        v[behind] = -99.0
        return torch.stack((u, v), dim=3).reshape(b, d * height, width, 2)


def homo_warp(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor,
              depth_values: torch.Tensor) -> torch.Tensor:
    b, c, h, w = src_fea.shape
    d = depth_values.shape[1]
    grid = sampling_grid(src_proj, ref_proj, depth_values, h, w)
    vol = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    return vol.reshape(b, c, d, h, w)


def compose(proj_pair: torch.Tensor) -> torch.Tensor:
    out = proj_pair[:, 0].clone()
    out[:, :3, :4] = proj_pair[:, 1, :3, :3] @ proj_pair[:, 0, :3, :4]
    return out


def cost_volume(features: Sequence[torch.Tensor], proj_matrix: torch.Tensor, depth_values: torch.Tensor,
                view_weights: Optional[torch.Tensor]) -> Tuple[Optional[torch.Tensor], List[torch.Tensor]]:
    """View loop of DepthNet.forward with the view weights given (TransMVSNet.py:71-93).

    Returns (aggregated similarity [B,1,D,H,W] or None when view_weights is None,
             per-view similarities list of [B,1,D,H,W]).
    """
    views = torch.unbind(proj_matrix, 1)
    ref_fea, ref_p = features[0], compose(views[0])
    num, den = 0, 1e-5
    per_view = []
    for i, (fea, pm) in enumerate(zip(features[1:], views[1:])):
        vol = homo_warp(fea, compose(pm), ref_p, depth_values)
        sim = (vol * ref_fea.unsqueeze(2)).mean(1, keepdim=True)
        per_view.append(sim)
        if view_weights is not None:
            wgt = view_weights[:, i:i + 1].unsqueeze(1)
            num = num + sim * wgt
            den = den + wgt
        del vol
    agg = None if view_weights is None else num / den
    return agg, per_view


def read_out(logits: torch.Tensor, depth_values: torch.Tensor):
    """TransMVSNet.py:99-103: prob, WTA index (int64), WTA depth, confidence."""
    prob = torch.exp(F.log_softmax(logits, dim=1))
    idx = torch.argmax(prob, dim=1, keepdim=True).type(torch.long)
    depth = torch.gather(depth_values, 1, idx).squeeze(1)
    conf = torch.max(prob, dim=1)[0]
    return prob, idx.squeeze(1), depth, conf


def depth_wta(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    idx = torch.argmax(p, dim=1, keepdim=True).type(torch.long)
    return torch.gather(depth_values, 1, idx).squeeze(1)


def depth_regression(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """Upstream (CasMVSNet/MVSNet) definition; absent from this fork -> parity unpinned."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(p * depth_values, 1)


def hot_path(stage_inputs) -> dict:
    """One pass of the whole path over one StageInputs (what a bench 'step' times on the CPU)."""
    agg, _ = cost_volume(stage_inputs.features, stage_inputs.proj_matrix, stage_inputs.depth_values,
                         stage_inputs.view_weights)
    prob, idx, depth, conf = read_out(stage_inputs.logits, stage_inputs.depth_values)
    return {"similarity": agg, "prob_volume": prob, "index": idx, "depth": depth, "photo_confidence": conf}


def depth_hypotheses(cur_depth: torch.Tensor, ndepth: int, interval_pixel: float, image_hw, stage_scale: int) -> torch.Tensor:
    """Stage hypotheses as the reference builds them (models/TransMVSNet.py:174-190, 202-204 with
    models/module.py:606-634), restated with the same ATen ops: cur_depth is depth_values [B,192] (stage 1)
    or the previous stage's depth [B,hp,wp]."""
    h_img, w_img = image_hw
    b = cur_depth.shape[0]
    steps = torch.arange(0, ndepth, dtype=cur_depth.dtype, device=cur_depth.device)
    if cur_depth.dim() == 2:
        lo, hi = cur_depth[:, 0], cur_depth[:, -1]
        step = (hi - lo) / (ndepth - 1)
        vol = lo[:, None] + steps[None] * step[:, None]
        vol = vol[:, :, None, None].repeat(1, 1, h_img, w_img)
    else:
        up = F.interpolate(cur_depth.unsqueeze(1), [h_img, w_img], mode="bilinear", align_corners=False).squeeze(1)
        lo = up - ndepth / 2 * interval_pixel
        hi = up + ndepth / 2 * interval_pixel
        step = (hi - lo) / (ndepth - 1)
        vol = lo.unsqueeze(1) + steps.reshape(1, -1, 1, 1) * step.unsqueeze(1)
    return F.interpolate(vol.unsqueeze(1), [ndepth, h_img // stage_scale, w_img // stage_scale], mode="trilinear",
                         align_corners=False).squeeze(1)
