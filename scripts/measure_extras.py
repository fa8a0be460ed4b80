"""CUDA-event timings of the kernels outside the headline step: drop-in homo_warping, per-view (stage-1) cost
volume, the atomic-free backward, depth_regression -- at BASELINE config-2 / config-4 sizes.  Prints JSON lines."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import geometry, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cu(t):
    return t.to(dev)


out = []
# ---- config 2, per stage: drop-in warp (one view), per-view similarity (stage-1 form), read-out variants
for stage in (1, 2, 3):
    st = synthetic.make_stage(stage, batch=1, n_views=5, height=1152, width=1600, seed=0)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    feats = [cu(f) for f in st.features]
    dv, vw, lg = cu(st.depth_values), cu(st.view_weights), cu(st.logits)
    b, d, h, w = st.depth_values.shape
    c = feats[0].shape[1]
    packed = ops.pack_sources(feats[1:])
    ms = timed(lambda: ops.homo_warp_packed(packed[0], rt[0], dv, c, w))
    out.append({"kernel": f"homo_warp_fwd/stage{stage}", "ms": ms, "algorithmic_mb": 4 * (c * h * w + d * h * w + c * d * h * w) / 1e6})
    ms = timed(lambda: ops.cost_volume_packed(feats[0], packed, rt, dv, None, True, False))
    out.append({"kernel": f"costvol_fwd(per-view out)/stage{stage}", "ms": ms,
                "algorithmic_mb": 4 * (5 * c * h * w + d * h * w + 4 * d * h * w) / 1e6})
    views = ops.cost_volume_packed(feats[0], packed, rt, dv, None, True, False)[1]
    ms = timed(lambda: ops._aggregate_fwd(views, vw))
    out.append({"kernel": f"aggregate_fwd/stage{stage}", "ms": ms, "algorithmic_mb": 4 * (4 * d * h * w + 4 * h * w + d * h * w) / 1e6})
    prob = ops.softmax_wta(lg, dv)[0]
    ms = timed(lambda: ops.depth_wta_index(prob, dv))
    out.append({"kernel": f"depth_wta/stage{stage}", "ms": ms, "algorithmic_mb": 4 * (d * h * w + 4 * h * w) / 1e6})
    ms = timed(lambda: ops.depth_regression(prob, dv))
    out.append({"kernel": f"depth_regression_fwd/stage{stage}", "ms": ms, "algorithmic_mb": 4 * (2 * d * h * w + h * w) / 1e6})
    del views, prob, packed
    torch.cuda.empty_cache()

# ---- backward at config 2 (B=1, N=5) and config 4 (BlendedMVS-shaped, B=8, N=7)
for name, kw in (("dtu B=1 N=5 1152x1600", dict(batch=1, n_views=5, height=1152, width=1600, kind="dtu")),
                 ("bld B=8 N=7 576x768", dict(batch=8, n_views=7, height=576, width=768, kind="unit"))):
    for stage in (1, 2, 3):
        st = synthetic.make_stage(stage, seed=0, **kw)
        rt = geometry.stage_rot_trans(st.proj_matrix)
        feats = [cu(f) for f in st.features]
        dv, vw = cu(st.depth_values), cu(st.view_weights)
        packed = ops.pack_sources(feats[1:])
        n = len(feats) - 1
        gv = torch.randn(n, *dv.shape, device=dev)
        f_ms = timed(lambda: ops.cost_volume_packed(feats[0], packed, rt, dv, vw, False, True), reps=3, warm=1)
        r_ms = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, True, False), reps=3, warm=1)
        s_ms = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, False, True), reps=3, warm=1)
        out.append({"kernel": f"backward {name} stage{stage}", "fwd_ms": f_ms, "bwd_grad_ref_ms": r_ms,
                    "bwd_grad_src_ms": s_ms, "voxel_views": st.voxel_views})
        del packed, gv, feats
        torch.cuda.empty_cache()
for o in out:
    if "algorithmic_mb" in o:
        o["gbps"] = round(o["algorithmic_mb"] / 1e3 / (o["ms"] * 1e-3), 1)
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in o.items()}))
