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
from .synthetic import ScanInputs, StageInputs

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


def pin_stage(st: StageInputs, memo: Optional[dict] = None) -> StageInputs:
    """memo: id(tensor) -> pinned copy, so tensors shared between stages / jobs of a scan are pinned once."""
    def pin(t):
        if t is None:
            return None
        if memo is None:
            return t.contiguous().pin_memory()
        if id(t) not in memo:
            memo[id(t)] = (t, t.contiguous().pin_memory())       # keep t alive: its id is the key
        return memo[id(t)][1]
    return StageInputs(stage=st.stage, features=[pin(f) for f in st.features], proj_matrix=st.proj_matrix,
                       depth_values=pin(st.depth_values), view_weights=pin(st.view_weights), logits=pin(st.logits),
                       num_depth=st.num_depth, cur_depth=pin(st.cur_depth),
                       interval_pixel=st.interval_pixel, image_hw=st.image_hw, bdhw=st.bdhw)


def pin_scan(scan: ScanInputs) -> ScanInputs:
    """The scan with every tensor in pinned host memory; a feature map shared by several jobs stays ONE buffer."""
    memo: dict = {}
    jobs = [[pin_stage(st, memo) for st in job] for job in scan.jobs]
    pyramids = [[memo[id(m)][1] if id(m) in memo else m.contiguous().pin_memory() for m in pyr] for pyr in scan.pyramids]
    return ScanInputs(pyramids=pyramids, pairs=scan.pairs, jobs=jobs)


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
        self._d2h = None
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
        vw1 = None
        for n, (st, dev, ev) in enumerate(staged):
            compute.wait_event(ev)
            if st.cur_depth is not None:
                depth_values = ops.depth_hypotheses(dev["seed"], st.num_depth, st.interval_pixel, st.image_hw,
                                                    synthetic_scale(st))
            else:
                depth_values = dev["seed"]
            if n == 0:
                vw1 = dev["view_weights"]
            # stages 2/3 read the stage-1 weights at their own resolution (nearest x2 of TransMVSNet.py:193-194 inside
            # the kernel: no upsampled copy, no PyTorch kernel on the path)
            packed = ops.pack_sources(dev["features"][1:])
            sim, _ = ops.cost_volume_packed(dev["features"][0], packed, stage_rot_trans(st.proj_matrix), depth_values,
                                            vw1, False, True, vw_shift=n)
            prob, idx, depth, conf = ops.softmax_wta(dev["logits"], depth_values, want_prob=True)
            out = {"similarity": sim, "prob_volume": prob, "index": idx, "depth": depth, "photo_confidence": conf}
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
        return sum(2 * (st.bdhw or st.depth_values.shape)[0] * (st.bdhw or st.depth_values.shape)[2] *
                   (st.bdhw or st.depth_values.shape)[3] * 4 for st in host_stages)

    # ------------------------------------------------------------------------------------------------ scan level
    def _scan_state(self, scan: ScanInputs):
        """Persistent device state for a scan shape: one slot per view for its feature pyramid (the H2D destination,
        read as the REFERENCE features) and for its packed form (read as a SOURCE view), two sets of per-job input
        buffers (the next job's inputs cross PCIe while this one computes) and one pinned output slot per job."""
        key = ("scan", len(scan.pyramids), tuple(tuple(m.shape) for m in scan.pyramids[0]),
               tuple(tuple(st.logits.shape) for st in scan.jobs[0]), len(scan.jobs))
        if key not in self._dev:
            mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.device)
            job0 = scan.jobs[0]
            state = {
                "nchw": [[mk(m) for m in pyr] for pyr in scan.pyramids],
                "packed": [[torch.empty((1, m.shape[2], (m.shape[3] + 7) // 8, (m.shape[1] + 3) // 4, 8, 4),
                                        dtype=torch.float32, device=self.device) for m in pyr] for pyr in scan.pyramids],
                "job": [{"logits": [mk(st.logits) for st in job0], "seed": [mk(st.cur_depth) for st in job0],
                         "vw1": mk(job0[0].view_weights), "done": None} for _ in range(2)],
                "out": [[{k: torch.empty(st.bdhw[0], st.bdhw[2], st.bdhw[3]).pin_memory()
                          for k in ("depth", "photo_confidence")} for st in job] for job in scan.jobs],
                # device-side result maps, two sets: job j's maps leave for the host while job j+1 computes
                "res": [[{k: torch.empty((st.bdhw[0], st.bdhw[2], st.bdhw[3]), dtype=torch.float32, device=self.device)
                          for k in ("depth", "photo_confidence")} for st in job0] for _ in range(2)],
                "res_free": [None, None],
                "rt": [[stage_rot_trans(st.proj_matrix) for st in job] for job in scan.jobs],   # host, by value
            }
            self._dev[key] = state
        return self._dev[key]

    def process_scan(self, scan: ScanInputs, cost_regularization=None, graphs: bool = False
                     ) -> List[List[Dict[str, torch.Tensor]]]:
        """Every view of a scan as the reference view once, with its source views from the scan's pairing -- the loop
        of the reference's test.py over datasets/general_eval.py:25-57 -- end to end from pinned host memory.

        A view is the reference view of one job and a source view of N-1 others, so its feature pyramid crosses PCIe
        ONCE per scan and is packed ONCE; both forms stay resident (49 DTU views: 5 + 5 GB of 180).  Per job only what
        is new moves: the pyramids of views not seen yet (one per job on average), the stand-in 3-D CNN logits, the
        stage-1 view weights and the depth seeds.  Per stage the device then runs the hypotheses kernel (N1), the fused
        cost volume fed from the cached packed maps with the stage-1 weights read at their own resolution
        (tmvs_costvol_fwd_cached: no upsampled copy), and the read-out.  Copies run on a copy stream one job ahead.
        Returns [job][stage] {"depth", "photo_confidence"} in pinned host memory (valid after a stream sync).

        cost_regularization: None -> the stand-in 3-D CNN logits of every job are INPUTS and cross PCIe with it (the
        conservative accounting bench.py's e2e figure uses); a list of three callables (one per stage, the reference's
        `cost_regularization` argument of DepthNet.forward, models/TransMVSNet.py:96-97) -> each maps the aggregated
        similarity [B,1,D,h,w] to logits on the device, as the cascade's own 3-D CNN does, and no logits are uploaded.
        graphs: capture the scan once as CUDA graphs (one per job) and replay them (see _process_scan_graphs).
        """
        state = self._scan_state(scan)
        if graphs:
            return self._process_scan_graphs(scan, state, cost_regularization)
        compute = torch.cuda.current_stream(self.device)
        if self._copy is None:
            self._copy = torch.cuda.Stream(self.device)
        copy = self._copy
        copy.wait_stream(compute)         # a previous scan's kernels are done with the per-view slots before we refill them
        resident = set()
        staged = []                        # per job: (buffers, ready event, views uploaded for it)
        self.h2d_bytes = 0
        n_jobs = len(scan.jobs)

        def upload(j):
            buf = state["job"][j & 1]
            with torch.cuda.stream(copy):
                if buf["done"] is not None:
                    copy.wait_event(buf["done"])            # job j-2 has consumed this buffer set
                new_views = self._upload_job(scan, state, j, resident, cost_regularization is None)
                ev = torch.cuda.Event()
                ev.record(copy)
            return buf, ev, new_views

        if self._d2h is None:
            self._d2h = torch.cuda.Stream(self.device)
        d2h = self._d2h
        staged.append(upload(0))
        results = []
        for j in range(n_jobs):
            if j + 1 < n_jobs:
                staged.append(upload(j + 1))                # one job ahead of the kernels
            buf, ev, new_views = staged[j]
            compute.wait_event(ev)
            if state["res_free"][j & 1] is not None:
                compute.wait_event(state["res_free"][j & 1])    # job j-2's maps have left this result set
            self._compute_job(scan, state, j, new_views, cost_regularization)
            buf["done"] = torch.cuda.Event()
            buf["done"].record(compute)
            d2h.wait_event(buf["done"])                     # the maps go home on their own stream, beside job j+1's kernels
            with torch.cuda.stream(d2h):
                results.append(self._download_job(scan, state, j))
                state["res_free"][j & 1] = torch.cuda.Event()
                state["res_free"][j & 1].record(d2h)
        compute.wait_stream(d2h)                            # "valid after a sync of the caller's stream"
        return results

    def _upload_job(self, scan: ScanInputs, state, j: int, resident: set, with_logits: bool) -> List[int]:
        """H2D copies of job j on the current stream: pyramids of views not yet resident, (logits,) seeds, stage-1 weights."""
        ref, srcs = scan.pairs[j]
        buf = state["job"][j & 1]
        new_views = []
        for v in [ref] + list(srcs):
            if v not in resident:
                resident.add(v)
                new_views.append(v)
                for d, h in zip(state["nchw"][v], scan.pyramids[v]):
                    d.copy_(h, non_blocking=True)
                    self.h2d_bytes += h.numel() * 4
        for s, st in enumerate(scan.jobs[j]):
            if with_logits:
                buf["logits"][s].copy_(st.logits, non_blocking=True)
                self.h2d_bytes += st.logits.numel() * 4
            buf["seed"][s].copy_(st.cur_depth, non_blocking=True)
            self.h2d_bytes += st.cur_depth.numel() * 4
        buf["vw1"].copy_(scan.jobs[j][0].view_weights, non_blocking=True)
        self.h2d_bytes += buf["vw1"].numel() * 4
        return new_views

    def _compute_job(self, scan: ScanInputs, state, j: int, new_views, cost_regularization):
        """The kernels of job j on the current stream; its maps land in the device result set j & 1."""
        buf = state["job"][j & 1]
        for v in new_views:                                 # layout pre-pass: once per view per scan
            for s in range(len(state["nchw"][v])):
                ops.pack_sources([state["nchw"][v][s]], out=state["packed"][v][s].unsqueeze(0))
        ref, srcs = scan.pairs[j]
        for s, st in enumerate(scan.jobs[j]):
            depth_values = ops.depth_hypotheses(buf["seed"][s], st.num_depth, st.interval_pixel, st.image_hw,
                                                st.image_hw[0] // st.bdhw[2])
            sim, _ = ops.cost_volume_packed(state["nchw"][ref][s], [state["packed"][v][s] for v in srcs],
                                            state["rt"][j][s], depth_values, buf["vw1"], False, True, vw_shift=s)
            logits = buf["logits"][s] if cost_regularization is None else \
                cost_regularization[s](sim.unsqueeze(1)).squeeze(1)
            res = state["res"][j & 1][s]
            ops.softmax_wta(logits, depth_values, want_prob=True, out_depth=res["depth"], out_conf=res["photo_confidence"])

    def _download_job(self, scan: ScanInputs, state, j: int):
        """D2H of job j's maps (current stream) from its device result set into its pinned output slot."""
        outs = []
        for s in range(len(scan.jobs[j])):
            host, res = state["out"][j][s], state["res"][j & 1][s]
            host["depth"].copy_(res["depth"], non_blocking=True)
            host["photo_confidence"].copy_(res["photo_confidence"], non_blocking=True)
            outs.append(host)
        return outs

    def _process_scan_graphs(self, scan: ScanInputs, state, cost_regularization):
        """process_scan as CUDA graphs: one graph per job (the uploads of job j+1 on a forked copy stream beside the
        kernels of job j), captured once per scan shape and replayed in order afterwards.  A job issues ~40 copies and
        launches through Python; with the 3-D CNN stand-in on the device the scan is launch-bound that way (2.46 ms per
        job against ~1.95 ms of kernels), and replaying graphs removes the host from the loop.  The buffers a graph
        touches are the persistent slots of _scan_state (per-view feature/packed slots, two ping-pong job-input sets, one
        pinned output slot per job), so the captured addresses stay valid; intermediates live in one memory pool shared
        by all graphs of the scan (they are replayed in capture order, never concurrently)."""
        key = ("graphs", id(scan), cost_regularization is None)
        compute = torch.cuda.current_stream(self.device)
        if self._copy is None:
            self._copy = torch.cuda.Stream(self.device)
        copy = self._copy
        if key not in state:
            torch.cuda.synchronize(self.device)
            self.h2d_bytes = 0
            resident: set = set()
            pool = torch.cuda.graph_pool_handle()
            graphs = []
            cap = torch.cuda.Stream(self.device)
            n_jobs = len(scan.jobs)
            with_logits = cost_regularization is None
            g0 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0, pool=pool, stream=cap):
                new_views = self._upload_job(scan, state, 0, resident, with_logits)
            graphs.append(g0)
            results = []
            if self._d2h is None:
                self._d2h = torch.cuda.Stream(self.device)
            d2h = self._d2h
            for j in range(n_jobs + 1):                     # graph j: kernels of job j, uploads of j+1, maps of j-1 going home
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=cap):
                    cur = torch.cuda.current_stream(self.device)
                    nxt = None
                    if j + 1 < n_jobs:                      # fork: job j+1's inputs cross PCIe beside job j's kernels
                        copy.wait_stream(cur)
                        with torch.cuda.stream(copy):
                            nxt = self._upload_job(scan, state, j + 1, resident, with_logits)
                    if j >= 1:                              # fork: job j-1's maps (the other result set) go to the host
                        d2h.wait_stream(cur)
                        with torch.cuda.stream(d2h):
                            results.append(self._download_job(scan, state, j - 1))
                    if j < n_jobs:
                        self._compute_job(scan, state, j, new_views, cost_regularization)
                    if j + 1 < n_jobs:
                        cur.wait_stream(copy)               # join
                        new_views = nxt
                    if j >= 1:
                        cur.wait_stream(d2h)                # join
                graphs.append(g)
            state[key] = {"graphs": graphs, "results": results, "h2d_bytes": self.h2d_bytes, "scan": scan}
        rec = state[key]
        self.h2d_bytes = rec["h2d_bytes"]
        for g in rec["graphs"]:
            g.replay()
        return rec["results"]
