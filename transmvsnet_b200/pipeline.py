"""The hot path as a user calls it: one reference view = three cascade stages of
(pack sources -> fused cost volume -> softmax/WTA read-out) on the sm_100a kernels.

`run_stage` / `run_cascade` take device tensors (what the PyTorch cascade holds);
`HostPipeline` is the end-to-end form used by bench.py's `e2e` figure: inputs start in pinned
host memory, results (depth + confidence maps) end in pinned host memory, copies included.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops
from .geometry import stage_rot_trans
from .synthetic import StageInputs

KERNELS_PER_STAGE = 3      # pack_sources, costvol_fwd, softmax_wta


def synthetic_scale(st: StageInputs) -> int:
    """Image-to-stage downscale of a StageInputs (4, 2 or 1)."""
    return st.image_hw[0] // st.depth_values.shape[2]


def stage_to_device(st: StageInputs, device, non_blocking: bool = True) -> Dict[str, object]:
    mv = lambda t: t.to(device, non_blocking=non_blocking)
    return {"features": [mv(f) for f in st.features], "depth_values": mv(st.depth_values),
            "view_weights": mv(st.view_weights), "logits": mv(st.logits),
            "rot_trans": stage_rot_trans(st.proj_matrix)}


def run_stage(dev: Dict[str, object], want_prob: bool = True, out_maps: Optional[torch.Tensor] = None
              ) -> Dict[str, torch.Tensor]:
    """One stage with the view weights given: 3 kernel launches, nothing else on the device.
    out_maps [B,2,H,W] (depth, confidence): where the read-out kernel writes its maps -- e.g. a peer-mapped slot of
    sharding.PeerMapSink on another GPU."""
    feats = dev["features"]
    packed = ops.pack_sources(feats[1:])
    sim, _ = ops.cost_volume_packed(feats[0], packed, dev["rot_trans"], dev["depth_values"], dev["view_weights"],
                                    False, True)
    out_d = out_c = None
    if out_maps is not None:
        if out_maps.shape[0] != 1:
            raise ValueError("out_maps is one view's [1,2,H,W] slot")
        out_d, out_c = out_maps[:, 0], out_maps[:, 1]          # contiguous for B = 1
    prob, idx, depth, conf = ops.softmax_wta(dev["logits"], dev["depth_values"], want_prob=want_prob,
                                             out_depth=out_d, out_conf=out_c)
    return {"similarity": sim, "prob_volume": prob, "index": idx, "depth": depth, "photo_confidence": conf}


def run_cascade(dev_stages: Sequence[Dict[str, object]], want_prob: bool = True,
                out_maps: Optional[torch.Tensor] = None) -> List[Dict[str, torch.Tensor]]:
    """out_maps: destination of the LAST stage's depth + confidence maps (see run_stage)."""
    last = len(dev_stages) - 1
    return [run_stage(d, want_prob, out_maps if n == last else None) for n, d in enumerate(dev_stages)]


def pin_stage(st: StageInputs) -> StageInputs:
    pin = lambda t: t.contiguous().pin_memory()
    return StageInputs(stage=st.stage, features=[pin(f) for f in st.features], proj_matrix=st.proj_matrix,
                       depth_values=pin(st.depth_values), view_weights=pin(st.view_weights), logits=pin(st.logits),
                       num_depth=st.num_depth, cur_depth=None if st.cur_depth is None else pin(st.cur_depth),
                       interval_pixel=st.interval_pixel, image_hw=st.image_hw)


def _uploads(st: StageInputs, first_stage: bool):
    """Tensors HostPipeline copies host->device for this stage: features, logits, the small depth seed the
    hypotheses are generated from (or the hypotheses themselves if there is none), stage-1 view weights."""
    ts = list(st.features) + [st.logits]
    ts.append(st.cur_depth if st.cur_depth is not None else st.depth_values)
    if first_stage:
        ts.append(st.view_weights)
    return ts


def stage_h2d_bytes(st: StageInputs, first_stage: bool = True) -> int:
    return sum(t.numel() * t.element_size() for t in _uploads(st, first_stage))


class HostPipeline:
    """End to end: pinned host inputs -> H2D -> kernels -> D2H of the depth and confidence maps."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._out: Dict[tuple, torch.Tensor] = {}
        self._copy = None
        self._dev: Dict[tuple, dict] = {}

    def _host_out(self, key, like: torch.Tensor) -> torch.Tensor:
        k = (key, tuple(like.shape))
        if k not in self._out:
            self._out[k] = torch.empty(like.shape, dtype=like.dtype).pin_memory()
        return self._out[k]

    def _device_inputs(self, st: StageInputs):
        """Persistent device buffers for a stage shape (allocated once; no allocator traffic in the steady state)."""
        key = (st.stage, tuple(st.features[0].shape), len(st.features), tuple(st.depth_values.shape))
        if key not in self._dev:
            mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.device)
            seed = st.cur_depth if st.cur_depth is not None else st.depth_values
            self._dev[key] = {"features": [mk(f) for f in st.features], "seed": mk(seed),
                              "view_weights": mk(st.view_weights), "logits": mk(st.logits), "done": None}
        return self._dev[key]

    def process_view(self, host_stages: Sequence[StageInputs]) -> List[Dict[str, torch.Tensor]]:
        """Returns per stage {"depth", "photo_confidence"} in pinned host memory (valid after a stream sync).

        What crosses PCIe is what the cascade receives from outside the path: the feature maps, the 3-D CNN logits,
        the stage-1 view weights, and the small depth seed of each stage (192 planes / previous depth map).  The
        per-pixel hypotheses [B,D,h,w] are generated on the device by the N1 kernel (as the cascade does,
        models/TransMVSNet.py:174-204) and the stage-2/3 weights are the nearest x2 upsample of stage 1 (:193-194).
        H2D copies run on a copy stream into persistent buffers; each stage's kernels wait only for their own inputs,
        so stage s+1's inputs cross PCIe while stage s computes.  The D2H of the result maps rides the compute stream.
        """
        compute = torch.cuda.current_stream(self.device)
        if self._copy is None:
            self._copy = torch.cuda.Stream(self.device)
        staged = []
        with torch.cuda.stream(self._copy):
            for n, st in enumerate(host_stages):
                dev = self._device_inputs(st)
                if dev["done"] is not None:
                    self._copy.wait_event(dev["done"])      # the previous view's kernels have consumed these buffers
                for d, h in zip(dev["features"], st.features):
                    d.copy_(h, non_blocking=True)
                dev["seed"].copy_(st.cur_depth if st.cur_depth is not None else st.depth_values, non_blocking=True)
                if n == 0:
                    dev["view_weights"].copy_(st.view_weights, non_blocking=True)
                dev["logits"].copy_(st.logits, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy)
                staged.append((st, dev, ev))
        results = []
        weights = None
        for n, (st, dev, ev) in enumerate(staged):
            compute.wait_event(ev)
            if st.cur_depth is not None:
                depth_values = ops.depth_hypotheses(dev["seed"], st.num_depth, st.interval_pixel, st.image_hw,
                                                    synthetic_scale(st))
            else:
                depth_values = dev["seed"]
            if n == 0:
                weights = dev["view_weights"]
            else:                                           # TransMVSNet.py:193-194
                weights = torch.nn.functional.interpolate(weights, scale_factor=2, mode="nearest")
                weights = weights[:, :, :depth_values.shape[2], :depth_values.shape[3]].contiguous()
            run_in = {"features": dev["features"], "depth_values": depth_values, "view_weights": weights,
                      "logits": dev["logits"], "rot_trans": stage_rot_trans(st.proj_matrix)}
            out = run_stage(run_in, want_prob=True)
            dev["done"] = torch.cuda.Event()
            dev["done"].record(compute)
            host = {}
            for key in ("depth", "photo_confidence"):
                buf = self._host_out((st.stage, key), out[key])
                buf.copy_(out[key], non_blocking=True)
                host[key] = buf
            results.append(host)
        return results

    @staticmethod
    def d2h_bytes(host_stages: Sequence[StageInputs]) -> int:
        return sum(2 * st.depth_values.shape[0] * st.depth_values.shape[2] * st.depth_values.shape[3] * 4
                   for st in host_stages)
