"""Rebuild tmvs_pack.cu with different tunables and time pack_sources at the config-2 stage sizes (GPU box)."""
import ctypes, json, os, shutil, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from transmvsnet_b200 import _lib, build, ops
CSRC = os.path.join(REPO, "transmvsnet_b200", "csrc")
SCRATCH = os.path.join(REPO, "gpurun_out", "tune")
os.makedirs(SCRATCH, exist_ok=True)
dev = torch.device("cuda:0")
flags0 = [f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
objs = {}
for src in build.SOURCES:
    if src != "tmvs_pack.cu":
        o = os.path.join(SCRATCH, src + ".o")
        subprocess.run(["nvcc", *flags0, "-c", os.path.join(CSRC, src), "-o", o], check=True, capture_output=True)
        objs[src] = o
shapes = [(32, 288, 400), (16, 576, 800), (8, 1152, 1600)]
feats = [[torch.randn(1, c, h, w, device=dev) for _ in range(4)] for c, h, w in shapes]
for n, defs in enumerate(json.loads(sys.argv[1])):
    o = os.path.join(SCRATCH, "pack.o")
    subprocess.run(["nvcc", *flags0, *[f"-D{k}={v}" for k, v in defs.items()], "-c", os.path.join(CSRC, "tmvs_pack.cu"), "-o", o], check=True, capture_output=True)
    lib = os.path.join(SCRATCH, f"libp{n}.so")
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs.values(), o, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    L = ctypes.CDLL(lib)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(L, name); fn.restype, fn.argtypes = res, args
    _lib._LIB = L
    out = []
    for f in feats:
        for _ in range(3): ops.pack_sources(f)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.pack_sources(f)
        e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1) / 20, 4))
    print(defs, out, "sum", round(sum(out), 4), flush=True)
