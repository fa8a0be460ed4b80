"""One backward (grad_ref + grad_src) per stage at a workload's sizes, for an ncu launch list:
   ncu --metrics gpu__time_duration.sum -k regex:. python scripts/profile_bwd.py dtu|bld [stages]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import geometry, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "dtu"
cfg = {"dtu": dict(height=1152, width=1600, n_views=5, batch=1, kind="dtu"),
       "bld": dict(height=576, width=768, n_views=7, batch=8, kind="unit")}[which]
stages = [int(a) for a in sys.argv[2:]] or [1, 2, 3]
for stage in stages:
    st = synthetic.make_stage(stage, seed=0, **cfg)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    feats = [f.to(dev) for f in st.features]
    dv = st.depth_values.to(dev)
    packed = ops.pack_sources(feats[1:])
    gv = torch.randn(cfg["n_views"] - 1, *dv.shape, device=dev)
    for _ in range(2):
        ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, True, True)
    torch.cuda.synchronize()
    print("stage", stage, "done", flush=True)
