"""BASELINE.json configs[4]: throughput sweep D in {48,96,192} x C in {8,16,32} x N in {3,5,11} at a 288x400 map
(one GPU; the path shards by reference view, so N-GPU throughput is N x this -- see profiles/r1_bench_final_n4.json).
Prints one JSON line per shape: voxel-views/s of pack + fused cost volume + read-out."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import pipeline, synthetic  # noqa: E402

dev = torch.device("cuda:0")
rows = []
for d in (48, 96, 192):
    for c in (8, 16, 32):
        for n in (3, 5, 11):
            st = synthetic.make_stage(1, batch=1, n_views=n, height=1152, width=1600, channels=c, num_depth=d, seed=0)
            dv = pipeline.stage_to_device(st, dev)
            for _ in range(3):
                pipeline.run_stage(dv)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                pipeline.run_stage(dv)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            rows.append({"D": d, "C": c, "N": n, "map": "288x400", "ms": round(ms, 4), "voxel_views": st.voxel_views,
                         "gvv_per_s": round(st.voxel_views / ms / 1e6, 2)})
            print(json.dumps(rows[-1]), flush=True)
            del dv
            torch.cuda.empty_cache()
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "sweep_config5.json")
json.dump({"note": "one B200, pack + fused cost volume + read-out per stage-shaped launch set, fp32", "rows": rows},
          open(out, "w"), indent=1)
