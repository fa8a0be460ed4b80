import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import oracle
from transmvsnet_b200 import fusion, synthetic
import test_gpu_fusion as T
DEV = torch.device("cuda:0")
rng = np.random.default_rng(7)
h, w = 24, 32
img = torch.from_numpy(rng.random((h, w, 4), dtype=np.float32) * np.array([1, 1, 1, 500], np.float32) + np.array([0, 0, 0, 430], np.float32)).float().to(DEV).contiguous()
uv = (rng.random((20000, 2)) * [w - 2, h - 2] + 1.0).astype(np.float32)
a = T._probe(img, uv, 0); p = T._probe(img, uv, 2); e = oracle.tex_linear(img.cpu().numpy(), uv)
sc = np.array([1, 1, 1, 930], np.float32)
print("array vs pitch: max", (np.abs(a - p) / sc).max(), "equal frac", (a == p).mean())
print("array vs model: max", (np.abs(a - e) / sc).max(), "equal frac", (a == e).mean())
print("pitch vs model: max", (np.abs(p - e) / sc).max(), "equal frac", (p == e).mean())
for shape in [(6, 96, 128, 4), (5, 64, 96, 5)]:
    v, hh, ww, seed = shape
    images, Ps = synthetic.make_fusion_scene(n_views=v, height=hh, width=ww, seed=seed)
    cams = fusion.camera_records(Ps.numpy())
    ref = T._reference_fusibile(images, cams)
    for name, pl in (("array", False), ("pitch", True)):
        got = fusion.fuse_depth_maps(images.to(DEV), cams, carry_over=True, pitch_linear=pl).cpu().numpy()
        if len(got) == len(ref):
            d = np.abs(got - ref)
            print(shape, name, "n", len(got), "max diff", d.max(), "rows identical", (got == ref).all(1).mean(), "cols max", d.max(0))
        else:
            print(shape, name, "count", len(got), "vs", len(ref))
    orc = oracle.fusibile(images, cams, carry_over=True)
    print(shape, "oracle count", len(orc), (np.abs(orc - ref).max() if len(orc) == len(ref) else None), ((orc == ref).all(1).mean() if len(orc) == len(ref) else None))
