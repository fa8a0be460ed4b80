"""Which 3-term evaluation order does torch.matmul(rot[B,3,3], xyz[B,3,HW]) use on this device?  (bitwise probe)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from transmvsnet_b200 import geometry, synthetic
SIZES = [(1152, 1600), (576, 800), (288, 400), (264, 480), (144, 192)] if "--sizes" in sys.argv else [(1152, 1600)]
for devname, (h, w) in [(d, hw) for d in ("cpu", "cuda") for hw in SIZES]:
    dev = torch.device(devname)
    print(f"--- {devname} map {h}x{w}")
    st = synthetic.make_stage(3, batch=1, n_views=3, height=1152, width=1600, seed=0)
    rt = geometry.stage_rot_trans(st.proj_matrix.to(dev)).to(dev)
    rot = rt[0, :, :9].reshape(1, 3, 3)
    y, x = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev), torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(h * w, device=dev)))[None]
    ref = torch.matmul(rot, xyz)[0].double().cpu().numpy().astype(np.float32)          # [3,HW] fp32 bits
    r = rot[0].cpu().numpy().astype(np.float64)
    X, Y = x.reshape(-1).cpu().numpy().astype(np.float64), y.reshape(-1).cpu().numpy().astype(np.float64)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    fma = lambda a, b, c: f32(a * b + c)            # exact in float64 for fp32 inputs (products of 24-bit mantissas)
    cands = {}
    for row in range(3):
        r0, r1, r2 = r[row]
        cands.setdefault("fma(r0,x,fma(r1,y,r2))", []).append(fma(r0, X, fma(r1, Y, np.full_like(X, r2))))
        cands.setdefault("fma(r2,1,fma(r1,y,r0*x))", []).append(fma(r2, 1.0, fma(r1, Y, f32(r0 * X))))
        cands.setdefault("fma(r1,y,r0*x)+r2 unfused last", []).append(f32(fma(r1, Y, f32(r0 * X)) + r2))
        cands.setdefault("((r0*x)+(r1*y))+r2 unfused", []).append(f32(f32(f32(r0 * X) + f32(r1 * Y)) + r2))
        cands.setdefault("fma(r0,x,r2) then fma(r1,y,.)", []).append(fma(r1, Y, fma(r0, X, np.full_like(X, r2))))
        cands.setdefault("exact (fp64 sum, one rounding)", []).append(f32(r0 * X + r1 * Y + r2))
    for name, rows in cands.items():
        got = np.stack(rows).astype(np.float32)
        mism = (got != ref).mean()
        print(f"{devname:5s} {name:38s} mismatching elements: {mism:.6f}")
