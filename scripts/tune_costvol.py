"""Rebuild the forward cost-volume kernels with different tunables and time them on the config-2 cascade.

Run on the GPU box:  python scripts/tune_costvol.py            (uses nvcc there; ~20 s per configuration)
Only tmvs_costvol.cu is rebuilt (TMVS_FAST_BUILD: the three exact kernels) into a scratch library.
"""
import ctypes
import itertools
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from transmvsnet_b200 import _lib, build, ops, pipeline, synthetic  # noqa: E402

CSRC = os.path.join(REPO, "transmvsnet_b200", "csrc")
SCRATCH = os.path.join(REPO, "gpurun_out", "tune")
os.makedirs(SCRATCH, exist_ok=True)


def build_variant(tag, defs):
    objs = []
    for src in build.SOURCES:
        obj = os.path.join(SCRATCH, f"{tag}_{src}.o") if src == "tmvs_costvol.cu" else os.path.join(SCRATCH, f"base_{src}.o")
        if src != "tmvs_costvol.cu" and os.path.exists(obj):
            objs.append(obj)
            continue
        flags = [f"-D{k}={v}" for k, v in defs.items()] + ["-DTMVS_FAST_BUILD=1"] if src == "tmvs_costvol.cu" else []
        subprocess.run(["nvcc", *[f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c",
                        os.path.join(CSRC, src), "-o", obj], check=True, capture_output=True)
        objs.append(obj)
    lib = os.path.join(SCRATCH, f"lib_{tag}.so")
    subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    return lib


def time_variant(lib_path, dev_stages, reps=10):
    lib = ctypes.CDLL(lib_path)
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib._LIB = lib
    out, outs = [], []
    for d in dev_stages:
        packed = ops.pack_sources(d["features"][1:])
        run = lambda: ops.cost_volume_packed(d["features"][0], packed, d["rot_trans"], d["depth_values"],
                                             d["view_weights"], False, True)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / reps)
        outs.append(run()[0])
    return out, outs


def main():
    dev = torch.device("cuda:0")
    stages = synthetic.make_cascade(batch=1, n_views=5, height=1152, width=1600, seed=0)
    dev_stages = [pipeline.stage_to_device(s, dev) for s in stages]
    torch.cuda.synchronize()
    grid = []
    for ty, dc, unroll in itertools.product((8, 4), (8, 16), (1, 2)):
        for mb in ((2, 3, 4), (2, 2, 3), (2, 2, 2), (2, 4, 5)):
            grid.append(dict(TMVS_TILE_Y=ty, TMVS_DC=dc, TMVS_UNROLL=unroll, TMVS_MINB8=mb[0], TMVS_MINB4=mb[1], TMVS_MINB2=mb[2]))
    if len(sys.argv) > 1:       # explicit list: '[{"TMVS_FWD_V":3}, ...]' merged over the defaults
        import json
        base = dict(TMVS_TILE_Y=8, TMVS_DC=8, TMVS_UNROLL=2, TMVS_MINB8=2, TMVS_MINB4=4, TMVS_MINB2=5, TMVS_FFMA2=1)
        grid = [{**base, **g} for g in json.loads(sys.argv[1])]
    print("variant                                          s1_ms   s2_ms   s3_ms   sum   | max rel diff vs first variant", flush=True)
    first = None
    for i, defs in enumerate(grid):
        try:
            lib = build_variant(f"v{i}", defs)
            t, outs = time_variant(lib, dev_stages)
            if first is None:
                first = outs
            diff = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, first)]
            tag = " ".join(f"{k[5:]}={v}" for k, v in defs.items())
            print(f"{tag:48s} {t[0]:7.4f} {t[1]:7.4f} {t[2]:7.4f} {sum(t):7.4f} | " + " ".join(f"{d:.1e}" for d in diff), flush=True)
        except subprocess.CalledProcessError as e:
            print("build failed", defs, e.stderr[-300:] if e.stderr else "", flush=True)


if __name__ == "__main__":
    main()
