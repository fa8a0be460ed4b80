"""Time grad_src (and grad_ref) at the DTU stage sizes through both scatter paths: the cell tables (default) and the
tile-scan kernels (TMVS_BWD_SRC_PATH=scan), and report how far apart their results are (fp32 re-association only)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import geometry, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
height, width, views, batch = 1152, 1600, 5, 1
if len(sys.argv) > 1:
    height, width, views, batch = (int(v) for v in sys.argv[1:5])
kind = "dtu" if len(sys.argv) <= 5 else sys.argv[5]


def timed(fn, reps=5):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


rows = []
for stage in (1, 2, 3):
    st = synthetic.make_stage(stage, batch=batch, n_views=views, height=height, width=width, kind=kind, seed=0)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    feats = [f.to(dev) for f in st.features]
    dv = st.depth_values.to(dev)
    packed = ops.pack_sources(feats[1:])
    gv = torch.randn(views - 1, *dv.shape, device=dev)
    os.environ.pop("TMVS_BWD_SRC_PATH", None)
    t_cells, (_, g_cells) = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, False, True))
    t_ref, _ = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, True, False))
    os.environ["TMVS_BWD_SRC_PATH"] = "scan"
    t_scan, (_, g_scan) = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, False, True))
    os.environ.pop("TMVS_BWD_SRC_PATH", None)
    diff = float((g_cells - g_scan).abs().max() / g_scan.abs().max())
    rows.append({"stage": stage, "grad_src_cells_ms": round(t_cells, 3), "grad_src_scan_ms": round(t_scan, 3),
                 "grad_ref_ms": round(t_ref, 3), "max_rel_diff_cells_vs_scan": diff})
    print(rows[-1], flush=True)
print(json.dumps({"workload": f"{kind} {height}x{width} N={views} B={batch}", "rows": rows}))
