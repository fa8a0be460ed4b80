"""NVLink traffic of the fused gather, one process, two GPUs: the config-2 cascade runs on cuda:1 and its stage-3 read-out
kernel stores the depth + confidence maps straight into a buffer that lives on cuda:0 (peer access over NVLink) -- what
sharding.PeerMapSink arranges between processes through CUDA IPC.  Run under ncu:
   ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:softmax_wta python scripts/nvlink_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from transmvsnet_b200 import pipeline, synthetic  # noqa: E402

assert torch.cuda.device_count() >= 2
d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
torch.zeros(8, device=d0)
torch.cuda.set_device(d1)
torch.zeros(8, device=d1)
# kernels on cuda:1 may dereference cuda:0 allocations only after peer access is enabled explicitly
from cuda import cudart  # noqa: E402  (cuda-python)
assert torch.cuda.can_device_access_peer(1, 0), "no peer access between the two GPUs"
(err,) = cudart.cudaDeviceEnablePeerAccess(0, 0)
assert int(err) in (0, 704), f"cudaDeviceEnablePeerAccess failed: {err}"     # 704 = already enabled
stages = synthetic.make_cascade(batch=1, n_views=5, height=1152, width=1600, seed=1)
dev_stages = [pipeline.stage_to_device(s, d1) for s in stages]
h, w = stages[-1].depth_values.shape[2:]
remote = torch.zeros(1, 2, h, w, device=d0)            # rank 0's slot
local = torch.zeros(1, 2, h, w, device=d1)
for _ in range(2):
    out_r = pipeline.run_cascade(dev_stages, out_maps=remote)[-1]
    out_l = pipeline.run_cascade(dev_stages, out_maps=local)[-1]
torch.cuda.synchronize(d1)
torch.cuda.synchronize(d0)
print("remote == local:", bool(torch.equal(remote.to(d1), local)), flush=True)
