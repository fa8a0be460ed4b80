"""Time tmvs_fusibile_fwd at the size of a DTU scan (49 views of 1152x1600 by default) and the CPU restatement on a
small sample.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import oracle  # noqa: E402
from transmvsnet_b200 import fusion, synthetic  # noqa: E402

views, height, width = 49, 1152, 1600
if len(sys.argv) > 1:
    views, height, width = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda:0")
images, Ps = synthetic.make_fusion_scene(n_views=views, height=height, width=width, seed=0)
cams = fusion.camera_records(Ps.numpy())
d_img = images.to(dev)
pts = fusion.fuse_depth_maps(d_img, cams)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    pts = fusion.fuse_depth_maps(d_img, cams)
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / reps * 1e3
# CPU restatement on a sample: a 1/8 x 1/8 crop-free downscale of the same scene
s_img, s_P = synthetic.make_fusion_scene(n_views=min(views, 12), height=height // 8, width=width // 8, seed=0)
s_cams = fusion.camera_records(s_P.numpy())
t0 = time.perf_counter()
s_pts = oracle.fusibile(s_img, s_cams)
cpu_s = time.perf_counter() - t0
pix = views * height * width
s_pix = s_img.shape[0] * s_img.shape[1] * s_img.shape[2]
print(json.dumps({"workload": f"{views} views of {height}x{width}, plane scene, carry-over on",
                  "points": int(len(pts)), "gpu_ms": round(gpu_ms, 2), "gpu_Mpixel_views_per_s": round(pix / gpu_ms / 1e3, 1),
                  "workspace_GB": round(fusion._lib.load().tmvs_fusibile_workspace_bytes(views, height, width) / 1e9, 2),
                  "cpu_oracle_sample": f"{s_img.shape[0]} views of {s_img.shape[1]}x{s_img.shape[2]}",
                  "cpu_oracle_s": round(cpu_s, 2), "cpu_Mpixel_views_per_s": round(s_pix / cpu_s / 1e6, 3),
                  "cpu_cores": os.cpu_count()}))
