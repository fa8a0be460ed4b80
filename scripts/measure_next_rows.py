"""CUDA-event timings of the SURVEY 8(f) kernels (N1 hypotheses, N2 folded PixelwiseNet + aggregation, N3 finalize)
at config-2 sizes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transmvsnet_b200 as tm
from transmvsnet_b200 import synthetic
dev = torch.device("cuda:0")
def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
st = synthetic.make_cascade(batch=1, n_views=5, height=1152, width=1600, seed=0)
for s in st:
    cur = s.cur_depth.to(dev)
    scale = s.image_hw[0] // s.depth_values.shape[2]
    ms = timed(lambda: tm.depth_hypotheses(cur, s.num_depth, s.interval_pixel, s.image_hw, scale))
    print(json.dumps({"kernel": f"depth_hypotheses/stage{s.stage}", "ms": round(ms, 4), "out_mb": s.depth_values.numel() * 4 / 1e6}))
s1 = st[0]
views = torch.randn(4, *s1.depth_values.shape, device=dev) * 0.2
mlp = tm.fold_pixelwise_net(tm.PixelwiseNet().eval())
ms = timed(lambda: tm.pixelwise_aggregate(views, mlp))
print(json.dumps({"kernel": "pixelwise_aggregate/stage1 (4 views, D=48, 288x400)", "ms": round(ms, 4), "in_mb": views.numel() * 4 / 1e6}))
net = tm.PixelwiseNet().eval().to(dev)
def torch_pwn():
    with torch.no_grad():
        w = torch.cat([net(views[i].unsqueeze(1)) for i in range(4)], 1)
        return tm.aggregate(views, w)
ms = timed(torch_pwn, reps=3, warm=1)
print(json.dumps({"kernel": "PyTorch PixelwiseNet (cuDNN, TF32 allowed) + aggregate kernel, same input", "ms": round(ms, 4)}))
d3, c3 = torch.rand(1, 1152, 1600, device=dev) * 500 + 425, torch.rand(1, 1152, 1600, device=dev)
c1, c2 = torch.rand(1, 288, 400, device=dev), torch.rand(1, 576, 800, device=dev)
ms = timed(lambda: tm.finalize_maps(d3, c3, c1, c2))
print(json.dumps({"kernel": "finalize_maps 1152x1600", "ms": round(ms, 4)}))
