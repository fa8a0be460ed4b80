"""Seeded synthetic inputs for the cost-volume path.

Shapes and units mimic what the reference's datasets feed the path
(datasets/general_eval.py:66-98,185-209 for the camera/proj_matrix format and
depth_values; models/TransMVSNet.py:113-132 for the (C, D, scale) of each stage;
models/module.py:606-634 for the per-pixel depth hypotheses).  Everything is made
on the CPU with a torch.Generator so the same tensors can be rebuilt on any box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

# (channels, depth planes, downscale) per cascade stage: models/module.py:397,
# models/TransMVSNet.py:113-132
STAGES = ((32, 48, 4), (16, 32, 2), (8, 8, 1))
DEPTH_RATIOS = (4.0, 1.0, 0.5)  # models/TransMVSNet.py:114 depth_interals_ratio

# DTU camera (datasets/general_eval.py:73-76) at the 1600x1200 capture size
_DTU_FX, _DTU_FY, _DTU_CX, _DTU_CY = 2892.33, 2883.18, 823.205, 619.071


@dataclass
class StageInputs:
    """One cascade stage of one batch of reference views."""
    stage: int
    features: List[torch.Tensor]          # N x [B,C,h,w], features[0] is the reference view
    proj_matrix: torch.Tensor             # [B,N,2,4,4]  ([...,0]=extrinsic, [...,1,:3,:3]=intrinsic)
    depth_values: torch.Tensor            # [B,D,h,w] per-pixel hypotheses
    view_weights: torch.Tensor            # [B,Nsrc,h,w]
    logits: torch.Tensor                  # [B,D,h,w] stand-in for the 3-D CNN output
    num_depth: int = 0
    # what the cascade actually holds before it builds depth_values (models/TransMVSNet.py:174-190): the 192 global
    # planes (stage 1) or the previous stage's depth map; with interval_pixel and image_hw they regenerate hypotheses
    cur_depth: Optional[torch.Tensor] = None
    interval_pixel: float = 0.0
    image_hw: tuple = (0, 0)
    # (B, D, h, w) of the stage; lean scan jobs carry no depth_values (generated on the device from cur_depth) and, past
    # stage 1, no view_weights (the kernel reads stage 1's at its own resolution)
    bdhw: tuple = ()

    @property
    def voxel_views(self) -> int:
        b, d, h, w = self.bdhw if self.bdhw else self.depth_values.shape
        return b * d * h * w * (len(self.features) - 1)


def _look_at_extrinsic(center: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """World-to-camera 4x4 for a camera at `center` whose optical axis passes through `target`."""
    z = target - center
    z = z / z.norm()
    up = torch.tensor([0.0, 1.0, 0.0], dtype=torch.float64)
    x = torch.linalg.cross(up, z)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    rot = torch.stack([x, y, z])          # rows = camera axes in world coords
    ext = torch.eye(4, dtype=torch.float64)
    ext[:3, :3] = rot
    ext[:3, 3] = -rot @ center
    return ext


def make_cameras(batch: int, n_views: int, height: int, width: int, *, kind: str = "dtu",
                 seed: int = 0) -> Dict[str, torch.Tensor]:
    """proj_matrix per stage, in the reference's [B,N,2,4,4] format, plus depth_values[B,192].

    kind "dtu": DTU intrinsics rescaled to (height, width), depth 425 + 2.5*k mm, sources on
    50-150 mm baselines converging at z=680 mm (3-13 degrees).  kind "unit": fx=fy=0.6*W,
    depth range [0.5, 10] scene units (Tanks&Temples / BlendedMVS shaped).
    """
    g = torch.Generator().manual_seed(seed + 7919)
    if kind == "dtu":
        fx = _DTU_FX * width / 1600.0
        fy = _DTU_FY * height / 1200.0
        cx = _DTU_CX * width / 1600.0
        cy = _DTU_CY * height / 1200.0
        d_min, d_int, focus = 425.0, 2.5, 680.0
        base_lo, base_hi = 50.0, 150.0
    elif kind == "unit":
        fx = fy = 0.6 * width
        cx, cy = 0.5 * width, 0.5 * height
        d_min, d_int, focus = 0.5, 9.5 / 191.0, 3.0
        base_lo, base_hi = 0.2, 0.6
    else:
        raise ValueError(kind)
    k_full = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=torch.float64)
    depth_values = (d_min + d_int * torch.arange(192, dtype=torch.float64)).float()
    depth_values = depth_values[None].repeat(batch, 1)

    proj = {}
    ext_all = torch.zeros(batch, n_views, 4, 4, dtype=torch.float64)
    for b in range(batch):
        for v in range(n_views):
            if v == 0:
                ext_all[b, v] = torch.eye(4, dtype=torch.float64)
                continue
            ang = 2.0 * math.pi * (v - 1) / max(n_views - 1, 1) + 0.3 + 0.37 * b
            mag = base_lo + (base_hi - base_lo) * float(torch.rand((), generator=g))
            center = torch.tensor([mag * math.cos(ang), mag * math.sin(ang), 0.0], dtype=torch.float64)
            target = torch.tensor([0.0, 0.0, focus], dtype=torch.float64)
            ext_all[b, v] = _look_at_extrinsic(center, target)
    for s, (_, _, scale) in enumerate(STAGES, start=1):
        k = k_full.clone()
        k[:2, :] = k[:2, :] / scale
        pm = torch.zeros(batch, n_views, 2, 4, 4, dtype=torch.float32)
        pm[:, :, 0] = ext_all.float()
        pm[:, :, 1, :3, :3] = k.float()
        proj[f"stage{s}"] = pm
    proj["depth_values"] = depth_values
    return proj


def make_stage(stage: int, *, batch: int = 1, n_views: int = 5, height: int = 1152, width: int = 1600,
               kind: str = "dtu", seed: int = 0, channels: Optional[int] = None,
               num_depth: Optional[int] = None, cameras: Optional[Dict[str, torch.Tensor]] = None,
               stage1_weights: Optional[torch.Tensor] = None, features: Optional[List[torch.Tensor]] = None,
               logits: Optional[torch.Tensor] = None, lean: bool = False) -> StageInputs:
    """Synthetic inputs of cascade stage `stage` (1..3) for an image of (height, width).
    features / logits: use these tensors instead of drawing new ones (a scan shares feature maps between views).
    lean: leave out what the scan pipeline derives on the device (per-pixel hypotheses, upsampled view weights)."""
    c_def, d_def, scale = STAGES[stage - 1]
    c = channels or c_def
    d = num_depth or d_def
    h, w = height // scale, width // scale
    g = torch.Generator().manual_seed(seed * 1000 + stage)
    cams = cameras or make_cameras(batch, n_views, height, width, kind=kind, seed=seed)
    dv = cams["depth_values"]                                # [B,192]
    d_min, d_max = dv[:, 0], dv[:, -1]
    depth_interval = (d_max - d_min) / dv.shape[1]           # models/TransMVSNet.py:149
    cur_depth = dv
    hyp = None
    if stage == 1 and lean:
        pass
    elif stage == 1:
        # models/module.py:616-623 (2-D branch): the global range split into D planes
        new_int = (d_max - d_min) / (d - 1)
        hyp = d_min[:, None] + torch.arange(d, dtype=torch.float32)[None] * new_int[:, None]
        hyp = hyp[:, :, None, None].expand(batch, d, h, w).contiguous()
    else:
        # smooth surface in roughly the middle 70% of the range, then module.py:626-632
        yy = torch.linspace(0, 1, h)[:, None]
        xx = torch.linspace(0, 1, w)[None, :]
        span = (d_max - d_min)[:, None, None]
        mid = (d_min + 0.5 * (d_max - d_min))[:, None, None]
        surf = mid + 0.25 * span * torch.sin(3.0 * math.pi * xx) * torch.cos(2.0 * math.pi * yy)[None] \
            + 0.1 * span * (xx - 0.5)[None]
        if not lean:
            ipx = (DEPTH_RATIOS[stage - 1] * depth_interval)[:, None, None]
            cur_min = surf - d / 2 * ipx
            cur_max = surf + d / 2 * ipx
            new_int = (cur_max - cur_min) / (d - 1)
            hyp = cur_min[:, None] + torch.arange(d, dtype=torch.float32)[None, :, None, None] * new_int[:, None]
            hyp = hyp.contiguous()
        # the same surface at the previous stage's resolution: what the cascade would carry over
        hp, wp = height // STAGES[stage - 2][2], width // STAGES[stage - 2][2]
        yy = torch.linspace(0, 1, hp)[:, None]
        xx = torch.linspace(0, 1, wp)[None, :]
        cur_depth = (mid + 0.25 * span * torch.sin(3.0 * math.pi * xx) * torch.cos(2.0 * math.pi * yy)[None]
                     + 0.1 * span * (xx - 0.5)[None]).float().contiguous()
    feats = features if features is not None else [torch.randn(batch, c, h, w, generator=g) for _ in range(n_views)]
    if stage1_weights is None:
        h1, w1 = height // STAGES[0][2], width // STAGES[0][2]
        gw = torch.Generator().manual_seed(seed * 1000 + 17)
        stage1_weights = torch.sigmoid(torch.randn(batch, n_views - 1, h1, w1, generator=gw))
    vw = stage1_weights
    if lean:
        vw = stage1_weights if stage == 1 else None
    else:
        for _ in range(stage - 1):                           # models/TransMVSNet.py:193-194
            vw = torch.nn.functional.interpolate(vw, scale_factor=2, mode="nearest")
        if vw.shape[2] < h or vw.shape[3] < w:               # sizes not divisible by the stage scales
            vw = torch.sigmoid(torch.randn(batch, n_views - 1, h, w, generator=g))
        vw = vw[:, :, :h, :w].contiguous()
    if logits is None:
        logits = 3.0 * torch.randn(batch, d, h, w, generator=g)
    return StageInputs(stage=stage, features=feats, proj_matrix=cams[f"stage{stage}"],
                       depth_values=None if hyp is None else hyp.float(), view_weights=vw, logits=logits, num_depth=d,
                       cur_depth=cur_depth.float().contiguous(),
                       interval_pixel=float(DEPTH_RATIOS[stage - 1] * depth_interval[0]), image_hw=(height, width),
                       bdhw=(batch, d, h, w))


def make_cascade(*, batch: int = 1, n_views: int = 5, height: int = 1152, width: int = 1600,
                 kind: str = "dtu", seed: int = 0) -> List[StageInputs]:
    """All three stages of one batch of reference views (BASELINE.json configs 2-4)."""
    cams = make_cameras(batch, n_views, height, width, kind=kind, seed=seed)
    return [make_stage(s, batch=batch, n_views=n_views, height=height, width=width, kind=kind,
                       seed=seed, cameras=cams) for s in (1, 2, 3)]


@dataclass
class ScanInputs:
    """One scan as the reference's test loop walks it (datasets/general_eval.py:25-57, 133-138): V views, and for
    every view taken as the reference view its N-1 source views from the pairing.  A view's feature pyramid is ONE
    set of tensors, shared by every job it appears in (as reference view once, as source view N-1 times)."""
    pyramids: List[List[torch.Tensor]]                 # [view][stage] -> [1,C,h,w]
    pairs: List[tuple]                                 # (ref_view, [src_views])  -- the pair.txt of the scan
    jobs: List[List[StageInputs]]                      # [job][stage]; StageInputs.features alias the pyramids

    @property
    def voxel_views(self) -> int:
        return sum(st.voxel_views for job in self.jobs for st in job)


def scan_pairs(n_scan_views: int, n_views: int) -> List[tuple]:
    """Synthetic pair.txt: the N-1 nearest views on a ring (v+1, v-1, v+2, v-2, ...), so that -- as in DTU's pairing --
    every view is a source view of N-1 reference views."""
    pairs = []
    for v in range(n_scan_views):
        srcs, k = [], 1
        while len(srcs) < n_views - 1:
            for cand in ((v + k) % n_scan_views, (v - k) % n_scan_views):
                if len(srcs) < n_views - 1 and cand != v and cand not in srcs:
                    srcs.append(cand)
            k += 1
            if k > n_scan_views:                      # tiny scans: repeat the first source (general_eval.py:47-49)
                srcs += [srcs[0]] * (n_views - 1 - len(srcs))
        pairs.append((v, srcs))
    return pairs


def make_scan(n_scan_views: int = 49, *, n_views: int = 5, height: int = 1152, width: int = 1600, kind: str = "dtu",
              seed: int = 0, logits_pool: int = 4, lean: bool = True) -> ScanInputs:
    """A synthetic scan of `n_scan_views` views (DTU: 49) with the ring pairing above.  Feature pyramids are distinct per
    view (a rolled copy of one random pyramid: distinct bytes at the cost of a memcpy); every job has its own cameras,
    depth seeds and stage-1 view weights; the stand-in logits cycle through a pool of `logits_pool` distinct sets."""
    g = torch.Generator().manual_seed(seed * 7 + 3)
    base = [torch.randn(1, c, height // sc, width // sc, generator=g) for c, _, sc in STAGES]
    pyramids = [[torch.roll(m, shifts=(3 * v + 1, 7 * v + 1), dims=(2, 3)).contiguous() if v else m for m in base]
                for v in range(n_scan_views)]
    pool = [[3.0 * torch.randn(1, d, height // sc, width // sc, generator=g) for _, d, sc in STAGES]
            for _ in range(max(1, min(logits_pool, n_scan_views)))]
    pairs = scan_pairs(n_scan_views, n_views)
    jobs = []
    h1, w1 = height // STAGES[0][2], width // STAGES[0][2]
    for j, (ref, srcs) in enumerate(pairs):
        cams = make_cameras(1, n_views, height, width, kind=kind, seed=seed * 1000 + j)
        ids = [ref] + srcs
        w1s = torch.sigmoid(torch.randn(1, n_views - 1, h1, w1, generator=g))       # this job's stage-1 view weights
        jobs.append([make_stage(s, batch=1, n_views=n_views, height=height, width=width, kind=kind, seed=seed * 1000 + j,
                                cameras=cams, features=[pyramids[v][s - 1] for v in ids], logits=pool[j % len(pool)][s - 1],
                                stage1_weights=w1s, lean=lean)
                     for s in (1, 2, 3)])
    return ScanInputs(pyramids=pyramids, pairs=pairs, jobs=jobs)


def make_fusion_scene(n_views: int = 5, height: int = 64, width: int = 96, seed: int = 0, hole_fraction: float = 0.05,
                      outlier_fraction: float = 0.05):
    """Inputs of the depth-map fusion: `n_views` DTU-like cameras looking at a tilted plane, each
    with its exact depth map of the plane (float, not the 8-bit PNG quantisation), random colours, a few holes
    (depth 0 -> below fusibile's 425.001 floor) and a few outliers (depth off by 3-10 %, rejected by the consistency test).

    Returns (images [V,H,W,4] fp32: b, g, r, depth;  P [V,3,4] fp32 = K [R | t], the matrices test.py writes).
    """
    import numpy as np
    rng = np.random.default_rng(seed)
    cams = make_cameras(1, n_views, height, width, kind="dtu", seed=seed)
    pm = cams["stage3"][0].double().numpy()                      # [N,2,4,4]
    normal = np.array([0.12, -0.08, 1.0])
    normal /= np.linalg.norm(normal)
    d0 = 680.0 * normal[2]                                       # plane through (0, 0, 680)
    ys, xs = np.meshgrid(np.arange(height, dtype=np.float64), np.arange(width, dtype=np.float64), indexing="ij")
    u = np.stack([xs, ys, np.ones_like(xs)], -1)                 # [H,W,3]
    images = np.zeros((n_views, height, width, 4), np.float32)
    Ps = np.zeros((n_views, 3, 4), np.float32)
    for v in range(n_views):
        ext, K = pm[v, 0], pm[v, 1, :3, :3]
        P = (K @ ext[:3, :4]).astype(np.float32)
        Ps[v] = P
        Pd = P.astype(np.float64)
        Minv = np.linalg.inv(Pd[:, :3])
        # X(lambda) = Minv (lambda u - p4);  n . X = d0  ->  lambda
        num = d0 + normal @ (Minv @ Pd[:, 3])
        den = (u @ Minv.T) @ normal
        depth = num / den
        depth[(depth < 430.0) | (depth > 930.0)] = 0.0
        hole = rng.random((height, width)) < hole_fraction
        outl = rng.random((height, width)) < outlier_fraction
        depth = np.where(outl, depth * (1.0 + rng.uniform(0.03, 0.10, depth.shape)), depth)
        depth = np.where(hole, 0.0, depth)
        images[v, ..., :3] = rng.random((height, width, 3), dtype=np.float32)
        images[v, ..., 3] = depth.astype(np.float32)
    return torch.from_numpy(images), torch.from_numpy(Ps)
