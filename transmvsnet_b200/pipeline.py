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


def stage_to_device(st: StageInputs, device, non_blocking: bool = True) -> Dict[str, object]:
    mv = lambda t: t.to(device, non_blocking=non_blocking)
    return {"features": [mv(f) for f in st.features], "depth_values": mv(st.depth_values),
            "view_weights": mv(st.view_weights), "logits": mv(st.logits),
            "rot_trans": stage_rot_trans(st.proj_matrix)}


def run_stage(dev: Dict[str, object], want_prob: bool = True) -> Dict[str, torch.Tensor]:
    """One stage with the view weights given: 3 kernel launches, nothing else on the device."""
    feats = dev["features"]
    packed = ops.pack_sources(feats[1:])
    sim, _ = ops.cost_volume_packed(feats[0], packed, dev["rot_trans"], dev["depth_values"], dev["view_weights"],
                                    False, True)
    prob, idx, depth, conf = ops.softmax_wta(dev["logits"], dev["depth_values"], want_prob=want_prob)
    return {"similarity": sim, "prob_volume": prob, "index": idx, "depth": depth, "photo_confidence": conf}


def run_cascade(dev_stages: Sequence[Dict[str, object]], want_prob: bool = True) -> List[Dict[str, torch.Tensor]]:
    return [run_stage(d, want_prob) for d in dev_stages]


def pin_stage(st: StageInputs) -> StageInputs:
    pin = lambda t: t.contiguous().pin_memory()
    return StageInputs(stage=st.stage, features=[pin(f) for f in st.features], proj_matrix=st.proj_matrix,
                       depth_values=pin(st.depth_values), view_weights=pin(st.view_weights), logits=pin(st.logits),
                       num_depth=st.num_depth)


def stage_h2d_bytes(st: StageInputs) -> int:
    ts = list(st.features) + [st.depth_values, st.view_weights, st.logits]
    return sum(t.numel() * t.element_size() for t in ts)


class HostPipeline:
    """End to end: pinned host inputs -> H2D -> kernels -> D2H of the depth and confidence maps."""

    def __init__(self, device):
        self.device = torch.device(device)
        self._out: Dict[tuple, torch.Tensor] = {}
        self._copy = None

    def _host_out(self, key, like: torch.Tensor) -> torch.Tensor:
        k = (key, tuple(like.shape))
        if k not in self._out:
            self._out[k] = torch.empty(like.shape, dtype=like.dtype).pin_memory()
        return self._out[k]

    def process_view(self, host_stages: Sequence[StageInputs]) -> List[Dict[str, torch.Tensor]]:
        """Returns per stage {"depth", "photo_confidence"} in pinned host memory (valid after a stream sync).

        All H2D copies are queued on a copy stream up front; each stage's kernels wait only for their own
        inputs, so stage s+1's inputs cross PCIe while stage s computes, and the D2H of the (small) result maps
        rides the compute stream behind the kernels that produce them.
        """
        compute = torch.cuda.current_stream(self.device)
        if self._copy is None:
            self._copy = torch.cuda.Stream(self.device)
        self._copy.wait_stream(compute)             # buffers of the previous call are free once its kernels ran
        staged = []
        with torch.cuda.stream(self._copy):
            for st in host_stages:
                dev = stage_to_device(st, self.device)
                ev = torch.cuda.Event()
                ev.record(self._copy)
                staged.append((st, dev, ev))
        results = []
        for st, dev, ev in staged:
            compute.wait_event(ev)
            for t in dev["features"] + [dev["depth_values"], dev["view_weights"], dev["logits"]]:
                t.record_stream(compute)
            out = run_stage(dev, want_prob=True)
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
