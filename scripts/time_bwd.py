"""Time grad_ref and grad_src at a workload's stage sizes (CUDA events): the cell tables (default) and, with --scan, the
tile-scan kernels (TMVS_F_BWD_SCAN), and how far apart their results are (fp32 re-association only).
   python scripts/time_bwd.py [dtu|bld] [--scan]      -> one JSON line (also used by scripts/tune_bwd.py)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import _lib, geometry, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
which = args[0] if args else "dtu"
cfg = {"dtu": dict(height=1152, width=1600, n_views=5, batch=1, kind="dtu"),
       "bld": dict(height=576, width=768, n_views=7, batch=8, kind="unit")}[which]


def timed(fn, reps=10):
    # the result of the call before stays alive while the next one runs: warm up with the same pattern, or the second timed
    # call pays a cudaMalloc (1-90 ms of host stall) for the second output block
    for _ in range(3):
        out = fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


rows = []
for stage in (1, 2, 3):
    st = synthetic.make_stage(stage, seed=0, **cfg)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    feats = [f.to(dev) for f in st.features]
    dv = st.depth_values.to(dev)
    packed = ops.pack_sources(feats[1:])
    gv = torch.randn(cfg["n_views"] - 1, *dv.shape, device=dev)
    t_cells, (_, g_cells) = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, False, True))
    t_ref, _ = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, True, False))
    row = {"stage": stage, "grad_src_cells_ms": round(t_cells, 3), "grad_ref_ms": round(t_ref, 3)}
    if "--scan" in sys.argv:
        with ops.extra_flags(_lib.F_BWD_SCAN):
            t_scan, (_, g_scan) = timed(lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, False, True))
        row["grad_src_scan_ms"] = round(t_scan, 3)
        row["max_rel_diff_cells_vs_scan"] = float((g_cells - g_scan).abs().max() / g_scan.abs().max())
    rows.append(row)
    del feats, dv, packed, gv, g_cells
    torch.cuda.empty_cache()
print(json.dumps({"workload": which, "lib": os.path.basename(_lib.LIB_PATH), "rows": rows}))
