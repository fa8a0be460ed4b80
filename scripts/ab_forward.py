"""A/B timing of forward cost-volume variants at the config-2 stage sizes (CUDA events, 20 reps after warm-up).
   python scripts/ab_forward.py            -> gpurun_out/ab_forward.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import _lib, ops, pipeline, synthetic  # noqa: E402

dev = torch.device("cuda:0")
rows = []
variants = {"default": 0, "split": _lib.F_FWD_SPLIT, "sweep": _lib.F_FWD_SWEEP}
for stage in (1, 2, 3):
    st = synthetic.make_stage(stage, batch=1, n_views=5, height=1152, width=1600, seed=0)
    d = pipeline.stage_to_device(st, dev)
    packed = ops.pack_sources(d["features"][1:])
    for name, bits in variants.items():
        if (stage != 1 and name == "split") or (stage == 1 and name == "sweep"):
            continue
        for want_views in ((False, True) if stage == 1 else (False,)):
            def run():
                with ops.extra_flags(bits):
                    ops.cost_volume_packed(d["features"][0], packed, d["rot_trans"], d["depth_values"], d["view_weights"],
                                           want_views, not want_views)
            for _ in range(5):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run()
            e1.record()
            torch.cuda.synchronize()
            rows.append({"stage": stage, "variant": name, "output": "per-view" if want_views else "aggregated",
                         "ms": round(e0.elapsed_time(e1) / 20, 4)})
            print(rows[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/ab_forward.json", "w"), indent=1)
