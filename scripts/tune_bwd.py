"""Rebuild tmvs_costvol_bwd.cu / tmvs_costvol_bwd_cells.cu with different tunables and time grad_src at the DTU stage sizes (GPU box)."""
import ctypes, json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from transmvsnet_b200 import _lib, build, geometry, ops, synthetic
CSRC = os.path.join(REPO, "transmvsnet_b200", "csrc")
SCRATCH = os.path.join(REPO, "gpurun_out", "tune")
os.makedirs(SCRATCH, exist_ok=True)
dev = torch.device("cuda:0")
flags0 = [f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
objs = {}
for src in build.SOURCES:
    if src not in ("tmvs_costvol_bwd.cu", "tmvs_costvol_bwd_cells.cu"):
        o = os.path.join(SCRATCH, src + ".o")
        subprocess.run(["nvcc", *flags0, "-c", os.path.join(CSRC, src), "-o", o], check=True, capture_output=True)
        objs[src] = o
cases = []
for stage in (1, 2, 3):
    st = synthetic.make_stage(stage, batch=1, n_views=5, height=1152, width=1600, seed=0)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    cases.append((st, rt))
for defs in json.loads(sys.argv[1]):
    o = os.path.join(SCRATCH, "bwd.o")
    o2 = os.path.join(SCRATCH, "bwd_cells.o")
    r = subprocess.run(["nvcc", *flags0, *[f"-D{k}={v}" for k, v in defs.items()], "-c", os.path.join(CSRC, "tmvs_costvol_bwd.cu"), "-o", o], capture_output=True, text=True)
    r2 = subprocess.run(["nvcc", *flags0, *[f"-D{k}={v}" for k, v in defs.items()], "-c", os.path.join(CSRC, "tmvs_costvol_bwd_cells.cu"), "-o", o2], capture_output=True, text=True)
    if r.returncode or r2.returncode:
        print(defs, "build failed", r.stderr[-200:], r2.stderr[-200:]); continue
    lib = os.path.join(SCRATCH, "lib.so")
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs.values(), o, o2, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    import shutil; lib2 = os.path.join(SCRATCH, f"lib_{abs(hash(str(defs)))}.so"); shutil.copy(lib, lib2)
    L = ctypes.CDLL(lib2)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(L, name); fn.restype, fn.argtypes = res, args
    _lib._LIB = L
    out = []
    for st, rt in cases:
        feats = [f.to(dev) for f in st.features]; dv = st.depth_values.to(dev)
        packed = ops.pack_sources(feats[1:]); gv = torch.randn(4, *dv.shape, device=dev)
        f = lambda: ops.costvol_backward_packed(feats[0], packed, rt, dv, gv, False, True)
        f(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); f(); e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1) / 2, 3))
    print(defs, out, flush=True)
