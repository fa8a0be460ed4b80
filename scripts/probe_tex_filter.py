"""Probe of the texture unit's bilinear filter (B200) through tmvs_fusibile_tex_probe: ramp textures return the
quantised sample position, so the weight rounding rule can be read off.  Output: gpurun_out/probe_tex.json."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transmvsnet_b200 import _lib  # noqa: E402

DEV = torch.device("cuda:0")
MODE = 2 if "--array" in sys.argv else 0       # 0: pitch-linear texture over the buffer, 2: cudaArray texture (the reference's)
h, w = 32, 64
img = torch.zeros(h, w, 4)
img[..., 0] = torch.arange(w)[None, :].float()
img[..., 1] = torch.arange(h)[:, None].float()
rng = np.random.default_rng(0)
img[..., 2] = torch.from_numpy(rng.random((h, w), dtype=np.float32))
img[..., 3] = torch.from_numpy(rng.random((h, w), dtype=np.float32)) * 500 + 430
img = img.to(DEV).contiguous()
n = 20000
uv = (rng.random((n, 2)) * [w - 4, h - 4] + 2).astype(np.float32)
lib = _lib.load()
uv_d = torch.from_numpy(uv).to(DEV)
out = torch.empty((n, 4), device=DEV)
rc = lib.tmvs_fusibile_tex_probe(ctypes.c_void_p(img.data_ptr()), h, w, ctypes.c_void_p(uv_d.data_ptr()),
                                 ctypes.c_void_p(out.data_ptr()), n, MODE, None)
assert rc == 0
r = out.cpu().numpy().astype(np.float64)
res = {}
for axis, name in ((0, "x"), (1, "y")):
    b = uv[:, axis].astype(np.float64) - 0.5
    for model, q in (("round", np.floor(b * 256 + 0.5) / 256), ("trunc", np.floor(b * 256) / 256)):
        res[f"{name}_{model}_mismatches"] = int((np.abs(r[:, axis] - q) > 1e-9).sum())
# value channels: blend with the rounded weights in double vs what the unit returned
im = img.cpu().numpy().astype(np.float64)
xb, yb = uv[:, 0].astype(np.float64) - 0.5, uv[:, 1].astype(np.float64) - 0.5
i, j = np.floor(xb).astype(int), np.floor(yb).astype(int)
a, bq = np.floor((xb - i) * 256 + 0.5) / 256, np.floor((yb - j) * 256 + 0.5) / 256
for ch in (2, 3):
    t = ((1 - a) * (1 - bq) * im[j, i, ch] + a * (1 - bq) * im[j, i + 1, ch] + (1 - a) * bq * im[j + 1, i, ch]
         + a * bq * im[j + 1, i + 1, ch])
    res[f"channel{ch}_max_abs_err_vs_double_blend"] = float(np.abs(r[:, ch] - t).max())
    res[f"channel{ch}_max_rel_err"] = float((np.abs(r[:, ch] - t) / np.abs(t)).max())
    res[f"channel{ch}_bit_exact_vs_float_of_double_blend"] = int((r[:, ch].astype(np.float32) == t.astype(np.float32)).sum())
res["n"] = n
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
res["texture"] = "cudaArray" if MODE else "pitch-linear"
json.dump(res, open(f"gpurun_out/probe_tex_{'pitch' if MODE else 'array'}.json", "w"), indent=1)
np.savez(f"gpurun_out/probe_tex_{'pitch' if MODE else 'array'}.npz", img=img.cpu().numpy(), uv=uv, out=out.cpu().numpy())
