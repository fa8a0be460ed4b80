"""Scan-level pipeline and the boundary options added in round 2, on the GPU through the C ABI.

* rot/trans read from DEVICE memory (TMVS_F_RT_DEVICE) == passed by value from the host, bit for bit, forward and
  backward; DepthNet.forward with the projection matrices on the GPU (as the reference's test.py holds them) runs
  without a single host synchronisation (torch.cuda.set_sync_debug_mode("error")).
* view weights read at the stage-1 resolution inside the kernel (vw_shift) == the materialised nearest x2 upsampling of
  models/TransMVSNet.py:193-194; per-view packed maps (tmvs_costvol_fwd_cached) == the contiguous pack.
* HostPipeline.process_scan (each view uploaded and packed once per scan) == the per-view cascade on the same inputs.
"""
import numpy as np
import pytest
import torch

import transmvsnet_b200 as tm
from transmvsnet_b200 import geometry, ops, pipeline, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(t):
    return t.to(DEV)


def test_device_rot_trans_equals_host_rot_trans_forward_and_backward():
    for stage, hw in ((1, (96, 160)), (2, (72, 104)), (3, (40, 72))):
        st = synthetic.make_stage(stage, batch=3, n_views=4, height=hw[0], width=hw[1], seed=31)
        rt_host = geometry.stage_rot_trans(st.proj_matrix)
        assert not rt_host.is_cuda
        rt_dev = rt_host.to(DEV)
        feats = [cu(f) for f in st.features]
        outs = []
        for rt in (rt_host, rt_dev):
            fs = [f.detach().requires_grad_(True) for f in feats]
            agg, views = tm.cost_volume(fs[0], fs[1:], rt, cu(st.depth_values), cu(st.view_weights), want_views=True)
            (agg.sum() + 0.5 * views.sum()).backward()
            outs.append([agg.detach(), views.detach()] + [f.grad for f in fs])
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        # the 4x4 algebra itself on the device: the matrices the reference would have computed there
        rt_gpu_algebra = geometry.stage_rot_trans(cu(st.proj_matrix))
        assert rt_gpu_algebra.is_cuda and tuple(rt_gpu_algebra.shape) == tuple(rt_host.shape)
        assert float((rt_gpu_algebra.cpu() - rt_host).abs().max()) <= 1e-3 * float(rt_host.abs().max())
        warp_h = ops.homo_warp_packed(ops.pack_sources(feats[1:2])[0], rt_host[0], cu(st.depth_values), feats[1].shape[1],
                                      feats[1].shape[3])
        warp_d = ops.homo_warp_packed(ops.pack_sources(feats[1:2])[0], rt_dev[0], cu(st.depth_values), feats[1].shape[1],
                                      feats[1].shape[3])
        assert torch.equal(warp_h, warp_d)


def test_depthnet_forward_with_gpu_projections_never_synchronises():
    """models/TransMVSNet.py:198-215 hands DepthNet the projection matrices on the GPU (test.py: tocuda(sample)).  The
    round-1 path ended the 4x4 algebra in .cpu() -- three blocking round trips per reference view; now nothing in
    DepthNet.forward may synchronise (inference with given weights: stages 2/3 of the cascade)."""
    st = synthetic.make_stage(2, batch=1, n_views=4, height=96, width=128, seed=33)
    net = tm.DepthNet().to(DEV).eval()
    feats = [cu(f) for f in st.features]
    pm, dv, vw = cu(st.proj_matrix), cu(st.depth_values), cu(st.view_weights)
    ident = torch.nn.Identity()
    with torch.no_grad():
        ref = net(feats, pm, dv, st.num_depth, ident, view_weights=vw)          # warm-up: allocations, lazy init
        torch.cuda.synchronize()
        torch.cuda.set_sync_debug_mode("error")
        try:
            out = net(feats, pm, dv, st.num_depth, ident, view_weights=vw)
        finally:
            torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert torch.equal(out["depth"], ref["depth"]) and torch.equal(out["prob_volume"], ref["prob_volume"])


def test_view_weights_read_at_stage1_resolution_and_per_view_packed_maps():
    cascade = synthetic.make_cascade(batch=2, n_views=4, height=96, width=160, seed=35)
    vw1 = cu(cascade[0].view_weights)
    for s, st in enumerate(cascade):
        rt = geometry.stage_rot_trans(st.proj_matrix)
        feats = [cu(f) for f in st.features]
        packed = ops.pack_sources(feats[1:])
        full, _ = ops.cost_volume_packed(feats[0], packed, rt, cu(st.depth_values), cu(st.view_weights), False, True)
        shifted, _ = ops.cost_volume_packed(feats[0], packed, rt, cu(st.depth_values), vw1, False, True, vw_shift=s)
        assert torch.equal(full, shifted), f"stage {s + 1}: in-kernel nearest upsampling"
        # the same views as separate allocations, in another order of memory
        singles = [ops.pack_sources([f])[0].clone() for f in reversed(feats[1:])][::-1]
        cached, views_c = ops.cost_volume_packed(feats[0], singles, rt, cu(st.depth_values), vw1, True, True, vw_shift=s)
        _, views_p = ops.cost_volume_packed(feats[0], packed, rt, cu(st.depth_values), None, True, False)
        assert torch.equal(full, cached) and torch.equal(views_c, views_p)


def test_packed_maps_of_another_shape_are_refused():
    st = synthetic.make_stage(2, batch=1, n_views=3, height=64, width=96, seed=36)
    feats = [cu(f) for f in st.features]
    rt = geometry.stage_rot_trans(st.proj_matrix)
    wrong = ops.pack_sources([f[:, :8].contiguous() for f in feats[1:]])             # packed for C = 8, features C = 16
    with pytest.raises(RuntimeError, match="packed sources"):
        ops.cost_volume_packed(feats[0], wrong, rt, cu(st.depth_values), cu(st.view_weights), False, True)
    with pytest.raises(RuntimeError, match="rot_trans"):
        ops.cost_volume_packed(feats[0], ops.pack_sources(feats[1:]), rt[:1], cu(st.depth_values), cu(st.view_weights),
                               False, True)
    with pytest.raises(RuntimeError, match="do not match"):
        tm.cost_volume(feats[0], [feats[1][:, :, :-1].contiguous()], rt[:1], cu(st.depth_values), cu(st.view_weights)[:, :1])
    with pytest.raises(RuntimeError, match="not divisible"):
        ops.depth_hypotheses(cu(st.cur_depth), st.num_depth, st.interval_pixel, (65, 96), 2)


def test_process_scan_equals_the_per_view_cascade():
    """Same inputs through (a) HostPipeline.process_scan -- one upload + one pack per view, cached packed maps, weights
    read at stage-1 resolution -- and (b) the plain per-view path (pack the four sources, materialised weights):
    identical maps for every job and stage, twice in a row (the second scan reuses the resident slots)."""
    scan = synthetic.make_scan(7, n_views=5, height=96, width=128, seed=41, logits_pool=3, lean=False)
    pinned = pipeline.pin_scan(scan)
    # shared feature maps are pinned once
    assert pinned.jobs[0][0].features[1] is pinned.pyramids[scan.pairs[0][1][0]][0]
    pipe = pipeline.HostPipeline(DEV)
    for rep in range(2):
        res = pipe.process_scan(pinned)
        torch.cuda.synchronize()
        got = [[{k: v.clone() for k, v in stage.items()} for stage in job] for job in res]
        for j, job in enumerate(scan.jobs):
            for s, st in enumerate(job):
                dev = pipeline.stage_to_device(st, DEV)
                dev["depth_values"] = ops.depth_hypotheses(cu(st.cur_depth), st.num_depth, st.interval_pixel, st.image_hw,
                                                           st.image_hw[0] // st.bdhw[2])
                want = pipeline.run_stage(dev)
                assert torch.equal(got[j][s]["depth"], want["depth"].cpu()), (rep, j, s)
                assert torch.equal(got[j][s]["photo_confidence"], want["photo_confidence"].cpu()), (rep, j, s)
    # every view crossed PCIe once: 7 pyramids + 7 x (logits + seeds + stage-1 weights)
    pyr = sum(m.numel() * 4 for m in scan.pyramids[0])
    per_job = sum((st.logits.numel() + st.cur_depth.numel()) * 4 for st in scan.jobs[0]) + scan.jobs[0][0].view_weights.numel() * 4
    assert pipe.h2d_bytes == 7 * (pyr + per_job)
    # the same scan captured as CUDA graphs (one per job, uploads of the next job forked beside the kernels) and
    # replayed twice: identical maps, identical byte count
    eager = [[{k: v.clone() for k, v in stage.items()} for stage in job] for job in res]
    for rep in range(2):
        for job in res:
            for stage in job:
                for v in stage.values():
                    v.zero_()
        res_g = pipe.process_scan(pinned, graphs=True)
        torch.cuda.synchronize()
        for j in range(len(scan.jobs)):
            for s in range(3):
                for k in ("depth", "photo_confidence"):
                    assert torch.equal(res_g[j][s][k], eager[j][s][k]), (rep, j, s, k)
        assert pipe.h2d_bytes == 7 * (pyr + per_job)


def test_channel_pass_kernel_agrees_with_the_one_pass_kernel():
    """C = 32 can run in two channel passes of 16 (TMVS_F_FWD_SPLIT: 64 registers, 4 CTAs per SM; measured slower, so
    opt-in) instead of one pass over all 32.  Same taps, same weights; only the channel sum is re-associated: <= 2e-6 of the range, for
    the aggregated and the per-view outputs, per-pixel and per-plane hypotheses, ragged sizes and the image rim."""
    from transmvsnet_b200 import _lib
    for hw, batch in (((96, 160), 2), ((148, 204), 1)):
        st = synthetic.make_stage(1, batch=batch, n_views=4, height=hw[0], width=hw[1], seed=51)
        rt = geometry.stage_rot_trans(st.proj_matrix)
        for dv in (st.depth_values, st.depth_values[:, :, 0, 0].contiguous()):
            args = (cu(st.features[0]), [cu(f) for f in st.features[1:]], rt, cu(dv), cu(st.view_weights))
            agg_1, views_1 = tm.cost_volume(*args, want_views=True)
            with ops.extra_flags(_lib.F_FWD_SPLIT):
                agg_s, views_s = tm.cost_volume(*args, want_views=True)
            for a, b in ((agg_s, agg_1), (views_s, views_1)):
                assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())


def test_epipolar_sweep_kernel_is_bit_identical_to_the_four_tap_kernel():
    """TMVS_F_FWD_SWEEP re-indexes the plane loop by the source columns the epipolar walk crosses and reuses the channel
    dot products of a source pixel between the planes whose footprints share it.  Same positions, same dots, same blend:
    the aggregated volume must be BIT-identical to costvol_fwd_kernel's -- on cascade shapes (x- and y-major walks, both
    directions: the four source cameras sit around the reference), ragged sizes and the image rim, per-plane [B,D]
    hypotheses, a long-step stage-1 shape (warps fall back to the generic path), exaggerated geometry (walks that jump,
    turn steep or leave the image), and hypotheses that run backwards."""
    from transmvsnet_b200 import _lib
    cases = []
    for stage, hw, batch, nv in ((2, (144, 200), 2, 5), (3, (96, 136), 1, 5), (2, (74, 106), 1, 4), (3, (40, 72), 2, 7),
                                 (3, (8, 8), 1, 3)):
        st = synthetic.make_stage(stage, batch=batch, n_views=nv, height=hw[0], width=hw[1], seed=61)
        cases.append((st, geometry.stage_rot_trans(st.proj_matrix), st.depth_values))
    st = synthetic.make_stage(2, batch=1, n_views=5, height=128, width=192, seed=62)
    cases.append((st, geometry.stage_rot_trans(st.proj_matrix), st.depth_values[:, :, 0, 0].contiguous()))      # [B,D]
    cases.append((st, geometry.stage_rot_trans(st.proj_matrix), st.depth_values.flip(1).contiguous()))          # far -> near
    rt = geometry.stage_rot_trans(st.proj_matrix).clone()
    rt[:, :, 0:6] *= 3.0                                                                                          # 3x zoom: jumps
    rt[:, :, 9:11] *= 3.0
    cases.append((st, rt, st.depth_values))
    rt = geometry.stage_rot_trans(st.proj_matrix).clone()
    rt[:, :, 9] += 40.0                                                                                           # pushed off the image
    rt[:, :, 11] -= 300.0                                                                                         # some z < 1e-6
    cases.append((st, rt, st.depth_values))
    st16 = synthetic.make_stage(1, batch=1, n_views=4, height=128, width=160, channels=16, num_depth=24, seed=63)
    cases.append((st16, geometry.stage_rot_trans(st16.proj_matrix), st16.depth_values))                          # ~2 px per plane
    for st, rt, dv in cases:
        feats = [cu(f) for f in st.features]
        packed = ops.pack_sources(feats[1:])
        for arith in ("cuda", "cpu"):
            want, _ = ops.cost_volume_packed(feats[0], packed, rt, cu(dv), cu(st.view_weights), False, True, arith=arith)
            with ops.extra_flags(_lib.F_FWD_SWEEP):
                got, _ = ops.cost_volume_packed(feats[0], packed, rt, cu(dv), cu(st.view_weights), False, True, arith=arith)
            assert torch.equal(got, want), (st.stage, tuple(dv.shape), arith, float((got - want).abs().max()))


def test_process_scan_with_the_regulariser_on_the_device():
    """cost_regularization given as callables (the reference's DepthNet.forward argument): logits are produced on the
    device from the aggregated similarity, nothing but features, seeds and stage-1 weights crosses PCIe."""
    scan = synthetic.make_scan(5, n_views=4, height=64, width=96, seed=43, lean=False)
    pinned = pipeline.pin_scan(scan)
    pipe = pipeline.HostPipeline(DEV)
    gain = [lambda x: x * 25.0] * 3
    res = pipe.process_scan(pinned, cost_regularization=gain)
    torch.cuda.synchronize()
    for j, job in enumerate(scan.jobs):
        for s, st in enumerate(job):
            dev = pipeline.stage_to_device(st, DEV)
            dv = ops.depth_hypotheses(cu(st.cur_depth), st.num_depth, st.interval_pixel, st.image_hw, st.image_hw[0] // st.bdhw[2])
            packed = ops.pack_sources(dev["features"][1:])
            sim, _ = ops.cost_volume_packed(dev["features"][0], packed, dev["rot_trans"], dv, dev["view_weights"], False, True)
            _, _, depth, conf = ops.softmax_wta(sim * 25.0, dv)
            assert torch.equal(res[j][s]["depth"], depth.cpu()) and torch.equal(res[j][s]["photo_confidence"], conf.cpu())
    pyr = sum(m.numel() * 4 for m in scan.pyramids[0])
    per_job = sum(st.cur_depth.numel() * 4 for st in scan.jobs[0]) + scan.jobs[0][0].view_weights.numel() * 4
    assert pipe.h2d_bytes == 5 * (pyr + per_job)


def test_backward_full_size_against_the_references_cuda_autograd():
    """Config-2 stage 3 at full size (1152 x 1600, N = 5): grad_ref / grad_src of the fused path against the reference's
    own op sequence differentiated by autograd on the same GPU (ATen grid_sampler_2d_backward: float atomics, so the
    comparison is a tolerance, 1e-4, not bit equality) -- and ours twice, bit for bit."""
    from conftest import assert_costvol_close
    from oracle import torch_port
    st = synthetic.make_stage(3, batch=1, n_views=5, height=1152, width=1600, seed=0)
    feats = [cu(f) for f in st.features]
    pm, dv, vw = cu(st.proj_matrix), cu(st.depth_values), cu(st.view_weights)
    g = torch.randn(dv.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(11))
    fs = [f.clone().requires_grad_(True) for f in feats]
    agg, _ = torch_port.cost_volume(fs, pm, dv, vw)
    want = torch.autograd.grad(agg.squeeze(1), fs, g)
    del agg
    rt = geometry.stage_rot_trans(pm)                      # on the device: read in place by the kernels
    runs = []
    for _ in range(2):
        fs2 = [f.clone().requires_grad_(True) for f in feats]
        agg2, _ = tm.cost_volume(fs2[0], fs2[1:], rt, dv, vw)
        runs.append(torch.autograd.grad(agg2, fs2, g))
    for v, (a, b, w) in enumerate(zip(runs[0], runs[1], want)):
        assert torch.equal(a, b), f"view {v}: backward is not bit-reproducible"
        assert_costvol_close(a.cpu().numpy(), w.cpu().numpy(), f"full-size stage 3 grad of view {v}")


def test_bench_line_on_a_tiny_workload():
    """bench.py end to end on the GPU with a tiny workload: one JSON line carrying the contract's keys, an e2e figure
    measured over a scan with declared copies, and a launch count that matches 9 launches per step."""
    import json
    import os
    import subprocess
    import sys
    from conftest import REPO
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--workload", "tiny", "--steps", "4", "--warmup", "3",
                          "--scan-views", "5", "--no-workloads", "--no-cpu-baseline"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-800:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "gpu_launches", "roofline", "e2e", "clocks"):
        assert key in line, key
    assert line["gpu_launches"] == 9 * 4 and line["steps"] == 4 and line["n_gpus"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] > 0
    assert line["e2e"]["with_device_side_regulariser"]["h2d_bytes_per_step"] < line["e2e"]["h2d_bytes_per_step"]
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3


def test_homo_warping_backward_under_strong_minification_uses_the_fallback_and_matches_autograd():
    """A 5x zoom-out makes ~25 reference pixels share every source pixel: far more than a cell's 4 + 2 slots, so the
    footprints flag their tiles and the general backward goes through the tile-scan fallback (one channel at a time).
    Compared with the reference's op sequence differentiated by autograd on the same GPU (float atomics there: 1e-4),
    for per-pixel and per-plane hypotheses, and bit-reproducible."""
    from conftest import assert_costvol_close
    from oracle import torch_port
    st = synthetic.make_stage(2, batch=2, n_views=3, height=96, width=128, channels=12, num_depth=6, seed=71)
    projs = [geometry.compose_projection(v) for v in torch.unbind(st.proj_matrix, 1)]
    src_proj = projs[1].clone()
    src_proj[:, :2, :] *= 0.2                       # the source image sees the scene five times smaller
    for dv in (st.depth_values, st.depth_values[:, :, 3, 5].contiguous()):
        g = torch.randn(2, 12, dv.shape[1], 48, 64, generator=torch.Generator().manual_seed(9))
        src_ref = cu(st.features[1]).requires_grad_(True)
        out_ref = torch_port.homo_warp(src_ref, cu(src_proj), cu(projs[0]), cu(dv))
        out_ref.backward(cu(g))
        grads = []
        for _ in range(2):
            src = cu(st.features[1]).requires_grad_(True)
            out = tm.homo_warping(src, cu(src_proj), cu(projs[0]), cu(dv), arith="cuda")
            out.backward(cu(g))
            grads.append(src.grad)
        assert_costvol_close(out.detach().cpu().numpy(), out_ref.detach().cpu().numpy(), "minified warp forward")
        assert torch.equal(grads[0], grads[1])
        assert_costvol_close(grads[0].cpu().numpy(), src_ref.grad.cpu().numpy(), "minified warp backward")
        # most of the source image receives nothing: the warped reference frustum covers a fifth of it per axis
        assert float((grads[0] == 0).float().mean()) > 0.5
