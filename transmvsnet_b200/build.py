"""Build libtmvs_sm100a.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m transmvsnet_b200.build [--force] [--sass]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.path.join(LIB_DIR, "libtmvs_sm100a.so")
SOURCES = ["tmvs_pack.cu", "tmvs_costvol.cu", "tmvs_costvol_bwd.cu", "tmvs_readout.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG, "..", "include", "tmvs.h")]
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".ptxas.log", "w") as f:
            f.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{res.stderr}")
        if verbose:
            print(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
    return LIB


def dump_sass(out_dir: str) -> None:
    """Write one SASS listing per translation unit (committed under profiles/sass/)."""
    os.makedirs(out_dir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        res = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True)
        with open(os.path.join(out_dir, src.replace(".cu", ".sass")), "w") as f:
            f.write(res.stdout)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--sass" in sys.argv:
        dump_sass(os.path.join(PKG, "..", "profiles", "sass"))
