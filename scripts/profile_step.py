"""One or two passes of the hot path (config-2 cascade by default) for ncu / compute-sanitizer runs."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from transmvsnet_b200 import pipeline, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--height", type=int, default=1152)
ap.add_argument("--width", type=int, default=1600)
ap.add_argument("--views", type=int, default=5)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--passes", type=int, default=2)
ap.add_argument("--backward", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda:0")
stages = synthetic.make_cascade(batch=args.batch, n_views=args.views, height=args.height, width=args.width, seed=0)
dev_stages = [pipeline.stage_to_device(s, dev) for s in stages]
torch.cuda.synchronize()
for _ in range(args.passes):
    outs = pipeline.run_cascade(dev_stages)
    if args.backward:
        from transmvsnet_b200 import ops
        for d in dev_stages:
            feats = [f.clone().requires_grad_(True) for f in d["features"]]
            agg, _ = ops.cost_volume(feats[0], feats[1:], d["rot_trans"], d["depth_values"], d["view_weights"])
            agg.backward(torch.ones_like(agg))
torch.cuda.synchronize()
print("profile_step ok", [tuple(o["similarity"].shape) for o in outs])
