#!/usr/bin/env python
"""Benchmark of the TransMVSNet cost-volume hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl tmvs|reference] [--workload dtu|tnt|bld]

A "step" is one pass of the hot path over ONE reference view of the named cascade: for each of
the three stages  pack sources -> fused cost volume (warp+sample+correlate+aggregate) -> softmax/WTA
read-out  = 9 kernel launches.  metric = cost-volume voxel-views/s (D*h*w*Nsrc summed over the stages,
BASELINE.json) and ms per reference view.  Under torchrun every rank processes its own reference
views (weak scaling, sharded by reference view) and the stage-3 depth/confidence maps are gathered
on rank 0: by default the read-out kernel stores them straight into rank 0's peer-mapped buffer over
NVLink (TMVS_GATHER=peer); TMVS_GATHER=copy uses the copy engine, TMVS_GATHER=nccl an NCCL all_gather.

--impl reference times the REFERENCE's own function for the path -- models.TransMVSNet.DepthNet.forward of the
unmodified model package staged under oracle/_ref by build() (oracle/build.py; /root/reference does not exist on the
GPU box) -- on the host cores at the SAME full-size configuration as the GPU arm: one step = the three DepthNet.forward
calls of one reference view, with the stand-in 3-D CNN logits and the stage-1 view weights given, as in the GPU arm.
The oracle port (oracle/torch_port.py) is only the fallback when oracle/_ref has not been staged.

The e2e figure walks a whole scan (49 views, every view the reference view once, ring pairing) through
HostPipeline.process_scan from pinned host memory; `workloads` carries the other BASELINE configs measured in the same
run (T&T-shaped N=7 forward, BlendedMVS-shaped B=8 forward + backward, three corners of the D x C x N sweep).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

WORKLOADS = {
    # BASELINE.json configs[1]: DTU test cascade 1152x1600, N=5, depths 48/32/8
    "dtu": dict(height=1152, width=1600, n_views=5, kind="dtu", batch=1,
                name="DTU 1152x1600 N=5 D=48/32/8 C=32/16/8 fp32 forward, 1 reference view per step"),
    # configs[2]: Tanks&Temples-shaped 1920x1056, N=7
    "tnt": dict(height=1056, width=1920, n_views=7, kind="unit", batch=1,
                name="T&T-shaped 1056x1920 N=7 D=48/32/8 fp32 forward, 1 reference view per step"),
    # configs[3] forward part: BlendedMVS-shaped 768x576, N=7, batch 8
    "bld": dict(height=576, width=768, n_views=7, kind="unit", batch=8,
                name="BlendedMVS-shaped 576x768 N=7 B=8 D=48/32/8 fp32 forward"),
    # seconds on a CPU: exercises both arms' plumbing in the test suite, never a reported number
    "tiny": dict(height=64, width=96, n_views=3, kind="dtu", batch=1,
                 name="tiny 64x96 N=3 D=48/32/8 (plumbing test only)"),
}
SCAN_VIEWS = 49         # views per scan in the e2e figure (DTU: 49 per scan, datasets/general_eval.py:25-57)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tmvs", choices=["tmvs", "reference"])
    ap.add_argument("--workload", default="dtu", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--scan-views", type=int, default=SCAN_VIEWS)
    return ap.parse_args()


def voxel_views(stages) -> int:
    return sum(s.voxel_views for s in stages)


def make_config(workload: dict, vv: int, world: int, gather: str) -> dict:
    """The `config` object of the JSON line -- the same for both arms (the reference arm times the same workload)."""
    how = {"peer": ", depth+conf maps written by the read-out kernel into rank 0's buffer over NVLink peer memory",
           "copy": ", depth+conf maps pushed into rank 0's peer-mapped buffer by the DMA engines over NVLink (side stream)",
           "nccl": ", NCCL all_gather of depth+conf", "": ""}[gather if world > 1 else ""]
    return {"workload": workload["name"], "voxel_views_per_step": vv,
            "l2": "inputs larger than L2 (>= 1 GB touched per step)", "view_weights": "given as inputs",
            "sharding": "by reference view, one process per GPU" + how}


def algorithmic_bytes(st) -> dict:
    """SURVEY.md 8(d) per-stage algorithmic bytes (fp32) for the three kernels of one stage."""
    b, d, h, w = st.depth_values.shape
    n_src = len(st.features) - 1
    c = st.features[0].shape[1]
    hw = h * w
    return {
        # layout pre-pass (read Nsrc*C*h*w NCHW, write the same packed): the kernel's own traffic.  It is NOT part of
        # SURVEY 8(d)'s algorithmic bytes of the path (those count the features once, in costvol_fwd below).
        "pack_sources": 4 * b * (2 * n_src * c * hw),
        # (1+Nsrc)*C*h*w features + D*h*w hypotheses + Nsrc*h*w weights in, D*h*w similarity out
        "costvol_fwd": 4 * b * ((1 + n_src) * c * hw + d * hw + n_src * hw + d * hw),
        # logits + hypotheses in, prob out, index(int64)+depth+conf out
        "softmax_wta": 4 * b * (3 * d * hw + 4 * hw),
    }


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while `active` is set."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.active, self.stop_flag = index, threading.Event(), threading.Event()
        self.sm, self.mask, self.sm_max, self.err = [], 0, None, None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag.is_set():
                if self.active.is_set():
                    self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    self.mask |= int(get_reasons(h))
                time.sleep(0.002)
        except Exception as e:          # noqa: BLE001 -- the clocks record is best effort, never fatal
            self.err = repr(e)

    def summary(self) -> dict:
        reasons = [n for bit, n in self.REASONS.items() if self.mask & bit]
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(self.sm), **({"error": self.err} if self.err else {})}


# ----------------------------------------------------------------------------------------- CPU arms
class _GivenLogits:
    """Stand-in for the 3-D CNN (`cost_regularization`, not part of the path): hands DepthNet.forward the same
    precomputed logits the GPU arm's read-out kernel consumes."""

    def __init__(self, logits):
        self.logits = logits

    def __call__(self, similarity):
        return self.logits.unsqueeze(1)


def cpu_hot_path_time(workload: dict, steps: int, warmup: int, min_seconds: float = 0.0):
    """The reference's own DepthNet.forward (oracle/_ref, staged by build()) on the host cores at the FULL size of the
    workload: one step = one reference view = the three stage calls.  Falls back to the oracle's port of the same ATen
    op sequence (oracle/torch_port.py) when the reference has not been staged.  Runs `steps` timed passes, then keeps
    going until `min_seconds` of timed CPU work have accumulated."""
    import torch
    from oracle import build as oracle_build
    from transmvsnet_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    stages = synthetic.make_cascade(batch=1, n_views=workload["n_views"], height=workload["height"], width=workload["width"],
                                    kind=workload["kind"], seed=0)
    vv = voxel_views(stages)
    ref = oracle_build.import_reference()
    if ref is not None:
        kind = "reference"
        net = ref[1].DepthNet().eval()                # PixelwiseNet parameters unused: the view weights are given

        def one_pass():
            for st in stages:
                net(st.features, st.proj_matrix, st.depth_values, st.num_depth, _GivenLogits(st.logits),
                    view_weights=st.view_weights)
        what = "models.TransMVSNet.DepthNet.forward of the unmodified reference (oracle/_ref)"
    else:
        from oracle import torch_port
        kind = "port"

        def one_pass():
            for st in stages:
                torch_port.hot_path(st)
        what = "oracle/torch_port.py (the reference's ATen op sequence; oracle/_ref not staged)"
    with torch.no_grad():
        for _ in range(warmup):
            one_pass()
        times = []
        while len(times) < steps or sum(times) < min_seconds:
            t0 = time.perf_counter()
            one_pass()
            times.append(time.perf_counter() - t0)
    sample = (f"{what}, full size {workload['height']}x{workload['width']} N={workload['n_views']}, "
              f"{vv / 1e6:.2f} M voxel-views per step (one reference view), torch {torch.__version__} CPU ops, "
              f"view weights and logits given")
    return vv, times, cores, sample, kind


def run_reference_arm(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)   # exactly K timed steps of the full-size workload
    vv, times, cores, sample, kind = cpu_hot_path_time(workload, steps, warmup)
    total = sum(times)
    value = vv * len(times) / total
    line = {
        "impl": "reference", "metric": "cost_volume_voxel_views_per_s", "value": value, "unit": "voxel-views/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(workload, vv, args.gpus, os.environ.get("TMVS_GATHER", "peer")),
        "cpu_baseline": {"value": value, "unit": "voxel-views/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "voxel-views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------- other BASELINE configs
def _timed(fn, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure_workloads(dev, rank: int, world: int):
    """BASELINE.json configs 3-5 in the same run, a few steps each, on every rank's own data (weak scaling by
    reference view / batch shard): max over ranks of the per-rank time, aggregate = world x per-rank work.
      tnt   T&T-shaped 1056x1920 N=7 forward cascade (pack + cost volume + read-out, 3 stages)
      bld   BlendedMVS-shaped 576x768 N=7 batch 8: forward cascade, and forward + backward of the cost volume through
            torch.autograd (atomic-free grad_ref / grad_src kernels) per stage
      sweep three corners of D x C x N at a 288x400 map
      stage1_learned  DTU stage 1 as the cascade runs it at inference: per-view similarity -> folded PixelwiseNet ->
            aggregation (SURVEY 8(d) asks for the with-PixelwiseNet figure separately)."""
    import torch
    import torch.distributed as dist
    import transmvsnet_b200 as tm
    from transmvsnet_b200 import ops, pipeline, synthetic

    def agg_ms(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    out = {}
    # ---- config 3
    w = WORKLOADS["tnt"]
    st = synthetic.make_cascade(batch=1, n_views=w["n_views"], height=w["height"], width=w["width"], kind=w["kind"], seed=rank)
    ds = [pipeline.stage_to_device(x, dev) for x in st]
    ms = agg_ms(_timed(lambda: pipeline.run_cascade(ds), 10))
    vv = voxel_views(st)
    out["tnt_forward"] = {"workload": w["name"], "ms_per_ref_view": round(ms, 4), "voxel_views_per_step": vv,
                          "value": world * vv / (ms * 1e-3), "unit": "voxel-views/s", "n_gpus": world}
    del ds, st
    torch.cuda.empty_cache()
    # ---- config 4
    w = WORKLOADS["bld"]
    st = synthetic.make_cascade(batch=w["batch"], n_views=w["n_views"], height=w["height"], width=w["width"], kind=w["kind"],
                                seed=rank)
    ds = [pipeline.stage_to_device(x, dev) for x in st]
    ms_f = agg_ms(_timed(lambda: pipeline.run_cascade(ds), 5))

    def fwd_bwd():
        for d in ds:
            fs = [f.detach().requires_grad_(True) for f in d["features"]]
            agg, _ = ops.cost_volume(fs[0], fs[1:], d["rot_trans"], d["depth_values"], d["view_weights"])
            agg.backward(torch.ones_like(agg))
    ms_fb = agg_ms(_timed(fwd_bwd, 3, warm=1))
    vv = voxel_views(st)
    out["bld_forward"] = {"workload": w["name"], "ms_per_batch": round(ms_f, 4), "voxel_views_per_step": vv,
                          "value": world * vv / (ms_f * 1e-3), "unit": "voxel-views/s", "n_gpus": world}
    out["bld_forward_backward"] = {
        "workload": "BlendedMVS-shaped 576x768 N=7 B=8: cost volume forward + backward wrt all feature maps through "
                    "torch.autograd, 3 stages (grad of ones)", "ms_per_batch": round(ms_fb, 4),
        "voxel_views_per_step": vv, "value": world * vv / (ms_fb * 1e-3), "unit": "voxel-views/s (forward count)",
        "n_gpus": world}
    del ds, st
    torch.cuda.empty_cache()
    # ---- config 5 (corners)
    rows = []
    for d_, c_, n_ in ((48, 8, 3), (96, 16, 5), (192, 32, 11)):
        s1 = synthetic.make_stage(1, batch=1, n_views=n_, height=1152, width=1600, channels=c_, num_depth=d_, seed=rank)
        d1 = pipeline.stage_to_device(s1, dev)
        ms = agg_ms(_timed(lambda: pipeline.run_stage(d1), 10))
        rows.append({"D": d_, "C": c_, "N": n_, "map": "288x400", "ms": round(ms, 4), "voxel_views": s1.voxel_views,
                     "value": world * s1.voxel_views / (ms * 1e-3), "unit": "voxel-views/s", "n_gpus": world})
        del d1, s1
    out["sweep_corners"] = rows
    torch.cuda.empty_cache()
    # ---- DTU stage 1 with the view weights LEARNED (eval-mode PixelwiseNet folded, N2)
    s1 = synthetic.make_stage(1, batch=1, n_views=5, height=1152, width=1600, seed=rank)
    d1 = pipeline.stage_to_device(s1, dev)
    net = tm.DepthNet().to(dev).eval()
    ident = torch.nn.Identity()
    pm = s1.proj_matrix

    def learned():
        with torch.no_grad():
            net(d1["features"], pm, d1["depth_values"], s1.num_depth, ident, view_weights=None)
    ms = agg_ms(_timed(learned, 10))
    out["dtu_stage1_learned_weights"] = {
        "what": "DepthNet.forward(view_weights=None), eval: pack + per-view similarity + folded PixelwiseNet + "
                "aggregation + read-out (Identity in place of the 3-D CNN)", "ms": round(ms, 4),
        "voxel_views": s1.voxel_views, "value": world * s1.voxel_views / (ms * 1e-3), "unit": "voxel-views/s"}
    del d1, s1, net
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------- GPU arm
def run_tmvs_arm(args, workload):
    import torch
    import torch.distributed as dist
    from transmvsnet_b200 import _lib, ops, pipeline, sharding, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl tmvs needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    host = [pipeline.pin_stage(s) for s in synthetic.make_cascade(
        batch=workload["batch"], n_views=workload["n_views"], height=workload["height"], width=workload["width"],
        kind=workload["kind"], seed=rank)]
    dev_stages = [pipeline.stage_to_device(s, dev) for s in host]
    vv = voxel_views(host)
    torch.cuda.synchronize()
    # Multi-GPU: the stage-3 depth + confidence maps of every view end up on rank 0.  Default transport: the read-out
    # kernel writes them straight into rank 0's buffer through NVLink peer memory (sharding.PeerMapSink) -- no
    # collective, no extra kernel.  TMVS_GATHER=nccl keeps the NCCL all_gather on a side stream instead.
    # TMVS_GATHER: "peer" (default) = the read-out kernel stores the maps into rank 0's peer-mapped buffer itself;
    # "copy" = they are pushed there by the DMA engines on a side stream (the same time at 8 GPUs: 1.8064 vs 1.8063 ms
    # per view, profiles/r2_n8_gather_transports.json); "nccl" = all_gather on a side stream.
    gather = os.environ.get("TMVS_GATHER", "peer")
    use_peer = world > 1 and workload["batch"] == 1 and gather in ("peer", "copy")
    comm = torch.cuda.Stream() if world > 1 else None
    sink = None
    if use_peer:
        h3, w3 = host[-1].depth_values.shape[2:]
        ok = torch.ones(1, device=dev)
        try:
            sink = sharding.PeerMapSink(2 * world, (h3, w3), dev, dst=0, sync=False)
        except Exception as e:      # no peer access between these GPUs: every rank falls back to the NCCL transport
            print(f"[bench] rank {rank}: peer-mapped gather unavailable ({e}); using NCCL all_gather", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            sink, use_peer, gather = None, False, "nccl"
    step_no = [0]
    copy_done = [None, None]
    local_maps = None
    if sink is not None and gather == "copy":
        h3, w3 = host[-1].depth_values.shape[2:]
        local_maps = [torch.zeros(1, 2, h3, w3, device=dev) for _ in range(2)]

    def step():
        if sink is not None:    # ring of two step-slots per rank: this step's maps land in slot (parity, rank)
            slot = (step_no[0] & 1) * world + rank
            step_no[0] += 1
            if gather == "peer":
                return pipeline.run_cascade(dev_stages, out_maps=sink.slot(slot))
            # "copy": the read-out kernel writes a persistent LOCAL map pair (two sets, alternating), the DMA engines move
            # it into rank 0's buffer on the side stream; set p is not rewritten before its previous copy has left
            p = step_no[0] & 1
            if copy_done[p] is not None:
                torch.cuda.current_stream().wait_event(copy_done[p])
            outs = pipeline.run_cascade(dev_stages, out_maps=local_maps[p])
            ready = torch.cuda.Event()
            ready.record()
            comm.wait_event(ready)
            sink.push(slot, local_maps[p][:, 0], local_maps[p][:, 1], comm)
            copy_done[p] = torch.cuda.Event()
            copy_done[p].record(comm)
            return outs
        outs = pipeline.run_cascade(dev_stages)
        if world > 1:       # gather this view's stage-3 depth + confidence on rank 0, off the compute stream
            maps = torch.stack([outs[-1]["depth"], outs[-1]["photo_confidence"]], 1)    # [B,2,H,W]
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                maps.record_stream(comm)
                sharding.gather_maps(maps, maps.shape[0] * world)
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident throughput: exactly K steps, CUDA events on the launching stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.LAUNCHES
    sampler.active.set()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    if comm is not None:
        torch.cuda.current_stream().wait_stream(comm)
    e1.record()
    barrier()
    sampler.active.clear()
    launches = _lib.LAUNCHES - launches0
    elapsed_ms = e0.elapsed_time(e1)
    per_rank = None
    if world > 1:
        # every rank's own device time and median SM clock: the max is the reported time; the spread shows how much of
        # the distance to N x the 1-GPU figure is GPU-to-GPU variation rather than the gather
        mine = torch.tensor([elapsed_ms / args.steps, float(statistics.median(sampler.sm)) if sampler.sm else 0.0], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(t[0]), 4) for t in allr], "sm_mhz": [float(t[1]) for t in allr]}
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * vv * args.steps / (elapsed_ms * 1e-3)

    # ---- per-kernel durations (CUDA events around each launch, same stream), for the roofline block
    names = ["pack_sources", "costvol_fwd", "softmax_wta"]
    reps = max(3, min(args.steps, 20))
    dur = {(s, n): 0.0 for s in range(3) for n in names}
    evs = []
    sampler.active.set()
    for _ in range(reps):
        for si, d in enumerate(dev_stages):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            packed = ops.pack_sources(d["features"][1:])
            ev[1].record()
            ops.cost_volume_packed(d["features"][0], packed, d["rot_trans"], d["depth_values"], d["view_weights"], False, True)
            ev[2].record()
            ops.softmax_wta(d["logits"], d["depth_values"])
            ev[3].record()
            evs.append((si, ev))
    torch.cuda.synchronize()
    sampler.active.clear()
    for si, ev in evs:
        for k, n in enumerate(names):
            dur[(si, n)] += ev[k].elapsed_time(ev[k + 1]) / reps
    kernels = []
    for si, st in enumerate(host):
        ab = algorithmic_bytes(st)
        for n in names:
            ms = dur[(si, n)]
            kernels.append({"kernel": f"{n}/stage{si + 1}", "ms": round(ms, 4), "algorithmic_mb": round(ab[n] / 1e6, 2),
                            "gbps": round(ab[n] / 1e9 / (ms * 1e-3), 1) if ms > 0 else None})
    top = max(kernels, key=lambda k: k["ms"])
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    traffic = None
    tr_path = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get(top["kernel"])
    # The fused cost-volume kernel is bound by the SM's load data path, not HBM: every voxel-view
    # moves 4 taps x C channels x 4 bytes from L1 into registers.  Peak = LDG.128 from L1-resident data on all SMs,
    # measured on this GPU pool by scripts/l1_peak.cu (profiles/r1_l1_peak.json); reported beside the HBM fraction.
    sm_load = None
    l1_path = os.path.join(REPO, "profiles", "r1_l1_peak.json")
    if top["kernel"].startswith("costvol_fwd") and os.path.exists(l1_path):
        l1_peak = [t for t in json.load(open(l1_path))["tests"] if t["name"] == "ldg128_aligned"][0]["GBps"]
        st = host[int(top["kernel"][-1]) - 1]
        tap_bytes = 16.0 * st.features[0].shape[1] * st.voxel_views
        l1_gbps = tap_bytes / 1e9 / (top["ms"] * 1e-3)
        sm_load = {"what": "bilinear-tap bytes delivered from L1 to registers (4 taps x C x 4 B per voxel-view)",
                   "bytes": int(tap_bytes), "achieved": round(l1_gbps, 1), "peak": l1_peak, "unit": "GB/s",
                   "frac": round(l1_gbps / l1_peak, 4),
                   "peak_source": "profiles/r1_l1_peak.json ldg128_aligned (scripts/l1_peak.cu, B200, 148 SMs)"}
    roofline = {"bound": "hbm", "kernel": top["kernel"], "achieved": top["gbps"], "peak": peak, "unit": "GB/s",
                "frac": round(top["gbps"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": top["ms"], "algorithmic_bytes": int(top["algorithmic_mb"] * 1e6),
                "all_kernels": kernels, "sm_load_path": sm_load,
                # SURVEY 8(d) bytes only: cost volume + read-out (the layout pre-pass is the kernels' own traffic)
                "step_algorithmic_mb": round(sum(k["algorithmic_mb"] for k in kernels
                                                 if not k["kernel"].startswith("pack_sources")), 2),
                "step_algorithmic_frac": round(sum(k["algorithmic_mb"] for k in kernels
                                                   if not k["kernel"].startswith("pack_sources")) * 1e6 / 1e9 /
                                               (elapsed_ms * 1e-3 / args.steps) / peak, 4)}

    # ---- end to end through the public API: a whole scan from pinned host memory (H2D + kernels + D2H for every view)
    e2e = None
    if not args.no_e2e and workload["batch"] == 1:
        del dev_stages
        torch.cuda.empty_cache()
        scan = pipeline.pin_scan(synthetic.make_scan(args.scan_views, n_views=workload["n_views"], height=workload["height"],
                                                     width=workload["width"], kind=workload["kind"], seed=rank))
        pipe = pipeline.HostPipeline(dev)
        pipe.process_scan(scan)                              # warm-up scan: allocates the resident slots
        barrier()
        n_scans = max(1, round(args.steps / len(scan.jobs)))
        sampler.active.set()
        barrier()
        e0.record()
        for _ in range(n_scans):
            pipe.process_scan(scan)
        e1.record()
        barrier()
        sampler.active.clear()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        k_e2e = n_scans * len(scan.jobs)
        h2d_host_logits = pipe.h2d_bytes // len(scan.jobs)
        # the same scan with the 3-D CNN stand-in ON THE DEVICE (a fixed gain, as tests/golden/make_golden.py uses): what
        # the pipeline moves when the logits are produced where the reference produces them.  Reported beside the
        # headline, never instead of it.
        gain = [lambda x: x * 25.0] * 3
        use_graphs = os.environ.get("TMVS_SCAN_GRAPHS", "1") == "1"
        try:
            pipe.process_scan(scan, cost_regularization=gain, graphs=use_graphs)      # captures the per-job CUDA graphs
        except Exception as e:      # noqa: BLE001 -- a capture problem must not cost the bench line: fall back to eager
            print(f"[bench] rank {rank}: CUDA-graph capture of the scan failed ({e!r}); eager launches", file=sys.stderr)
            use_graphs = False
            torch.cuda.synchronize()
            pipe.process_scan(scan, cost_regularization=gain)
        barrier()
        e0.record()
        for _ in range(n_scans):
            pipe.process_scan(scan, cost_regularization=gain, graphs=use_graphs)
        e1.record()
        barrier()
        ms_dev = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_dev], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_dev = float(t.item())
        device_reg = {"value": world * scan.voxel_views * n_scans / (ms_dev * 1e-3), "unit": "voxel-views/s",
                      "ms_per_step": ms_dev / k_e2e, "h2d_bytes_per_step": pipe.h2d_bytes // len(scan.jobs),
                      "cuda_graphs": use_graphs,
                      "what": "same scan, cost_regularization = x * 25 on the device (the reference's DepthNet.forward takes the "
                              "3-D CNN as a module argument: its logits never exist on the host), no logits uploaded"}
        e2e = {"value": world * scan.voxel_views * n_scans / (ms * 1e-3), "unit": "voxel-views/s",
               "h2d_bytes_per_step": h2d_host_logits,
               "h2d_what": "per reference view, averaged over the scan: ONE new feature pyramid (each view crosses PCIe and "
                           "is packed once per scan, then stays resident) + the stand-in 3-D CNN logits + stage-1 view "
                           "weights + depth seeds; hypotheses generated on the device",
               "d2h_bytes_per_step": pipeline.HostPipeline.d2h_bytes(scan.jobs[0]), "steps": k_e2e,
               "ms_per_step": ms / k_e2e,
               # what the host link delivered, summed over the ranks (one rank alone: ~54 GB/s of PCIe; the 8-GPU box
               # saturates near 165-180 GB/s in total, which is what bounds the end-to-end figure at 8 GPUs)
               "h2d_gbps_all_ranks": round(world * h2d_host_logits / (ms / k_e2e * 1e-3) / 1e9, 1),
               "with_device_side_regulariser": device_reg,
               "scan": f"{len(scan.pyramids)} views, every view the reference view once, {workload['n_views'] - 1} source "
                       f"views each (ring pairing), {n_scans} scan(s) timed after one warm-up scan; one scan per rank"}
        del scan, pipe
        torch.cuda.empty_cache()
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    extra = None
    if not args.no_workloads:
        extra = measure_workloads(dev, rank, world)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cvv, times, cores, sample, kind = cpu_hot_path_time(workload, steps=2, warmup=1, min_seconds=10.0)
        cpu = {"value": cvv * len(times) / sum(times), "unit": "voxel-views/s", "cores": cores, "kind": kind,
               "sample": sample + f"; {len(times)} passes, {sum(times):.1f} s of CPU work",
               "best_pass_value": cvv / min(times)}

    if rank == 0:
        line = {
            "metric": "cost_volume_voxel_views_per_s", "value": value, "unit": "voxel-views/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "ms_per_ref_view": elapsed_ms / args.steps / workload["batch"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(workload, vv, world, gather if use_peer else "nccl"),
            "per_rank": per_rank, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "workloads": extra,
            "clocks": sampler.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    # exactly ONE line may reach stdout (the JSON): libraries such as NCCL print banners there, so fd 1 is
    # pointed at stderr for the run and the JSON line is written to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    workload = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, workload)
    else:
        run_tmvs_arm(args, workload)


if __name__ == "__main__":
    main()
