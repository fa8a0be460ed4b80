"""Time tmvs_fusibile_fwd at the size of a DTU scan (49 views of 1152x1600 by default), with textures over the caller's buffer
(default) and with the reference's texture set-up (one cudaArray per view, allocated and filled inside the call), and --
when oracle/_ref/libfusibile_ref.so is present -- the reference's own kernel + per-camera host loop on the same inputs.
Prints one JSON line (measurement script: not part of the product path)."""
import ctypes
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from transmvsnet_b200 import fusion, synthetic  # noqa: E402

views, height, width = 49, 1152, 1600
if len(sys.argv) > 1:
    views, height, width = (int(v) for v in sys.argv[1:4])
dev = torch.device("cuda:0")
images, Ps = synthetic.make_fusion_scene(n_views=views, height=height, width=width, seed=0)
cams = fusion.camera_records(Ps.numpy())
d_img = images.to(dev)
out = {"workload": f"{views} views of {height}x{width}, plane scene, carry-over on"}
for name, kw in (("in_place_textures", {}), ("array_textures", {"array_textures": True})):
    pts = fusion.fuse_depth_maps(d_img, cams, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        pts = fusion.fuse_depth_maps(d_img, cams, **kw)
    torch.cuda.synchronize()
    out[name + "_ms"] = round((time.perf_counter() - t0) / reps * 1e3, 2)
    out["points"] = int(len(pts))
out["workspace_GB"] = round(fusion._lib.load().tmvs_fusibile_workspace_bytes(views, height, width) / 1e9, 2)
ref_lib = os.path.join(REPO, "oracle", "_ref", "libfusibile_ref.so")
if os.path.exists(ref_lib):
    lib = ctypes.CDLL(ref_lib)
    lib.fusibile_ref_run.restype = ctypes.c_int
    lib.fusibile_ref_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                     ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
    img = np.ascontiguousarray(images.numpy(), np.float32)
    cam = np.ascontiguousarray(cams, np.float32)
    cap = views * height * width
    buf = np.zeros((cap, 8), np.float32)
    n = ctypes.c_longlong(0)
    t0 = time.perf_counter()
    rc = lib.fusibile_ref_run(img.ctypes.data, cam.ctypes.data, views, height, width, 0.25, 3, buf.ctypes.data, cap, ctypes.byref(n))
    out["reference_binary_s"] = round(time.perf_counter() - t0, 2)
    out["reference_binary_points"] = int(n.value)
    out["reference_binary_what"] = ("gipuma/fusibile/fusibile.cu compiled for sm_100a with its own flags: upload + one launch, "
                                    "synchronise and host scan per camera over managed memory + copy-out (wall clock)")
print(json.dumps(out))
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(REPO, "gpurun_out", "fusion_timing.json"), "w"), indent=1)
