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
SOURCES = ["tmvs_pack.cu", "tmvs_pack_tma.cu", "tmvs_costvol.cu", "tmvs_costvol_sweep.cu", "tmvs_costvol_tma.cu", "tmvs_costvol_bwd.cu", "tmvs_costvol_bwd_cells.cu", "tmvs_readout.cu", "tmvs_fusion.cu"]
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


# kernels whose full SASS is committed (the hot instantiations); every other kernel gets an opcode histogram
HOT_KERNELS = ("costvol_fwd_kernelILi8ELb1ELb1ELb0ELb1ELb1ELi1E", "costvol_fwd_kernelILi4ELb1ELb1ELb0ELb1ELb1ELi1E",
               "costvol_fwd_kernelILi2ELb1ELb1ELb0ELb1ELb1ELi1E", "costvol_fwd_kernelILi8ELb1ELb1ELb1ELb0ELb1ELi1E",
               "costvol_fwd_kernelILi4ELb1ELb1ELb0ELb1ELb1ELi2E", "costvol_fwd_sweep_kernelILi4ELi16ELb1ELb1E",
               "costvol_fwd_sweep_kernelILi2ELi8ELb1ELb1E", "cells_gather_warp_kernelILi8E",
               "pack_sources_nchw4_kernel", "homo_warp_fwd_kernelILb1E", "softmax_wta_kernelILi48E",
               "softmax_wta_kernelILi32E", "softmax_wta_kernelILi8E", "depth_wta_kernel",
               "bwd_src_kernelILi8ELb1ELb1E", "bwd_ref_kernelILi8ELb1ELb1E", "bwd_bbox_kernelILb1E",
               "costvol_tma_kernelILi4ELb1ELb0ELb1E", "cells_register_kernelILb1E", "cells_fixup_kernelILi1E",
               "cells_gather_kernelILi4ELb1E", "fuse_points_kernel", "pixelwise_weight_kernel", "pack_sources_tma_kernelILi8E",
               "softmax_wta_split_lean_kernelILi4ELi8ELb1E", "softmax_wta_lean_kernelILi8ELb1E")


def dump_sass(out_dir: str) -> None:
    """Compact SASS evidence under profiles/sass/: full listings (encodings stripped) of the hot kernels and a
    per-kernel opcode histogram of everything in the library."""
    import collections
    import re
    os.makedirs(out_dir, exist_ok=True)
    hist_lines = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        res = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True)
        keep, name, ops, hot = [], None, None, False
        for line in res.stdout.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if name:
                    hist_lines.append(f"{name}: " + " ".join(f"{o}={c}" for o, c in ops.most_common(14)))
                name, ops = m.group(1), collections.Counter()
                hot = any(h in name for h in HOT_KERNELS)
                if hot:
                    keep.append("\n" + line.strip())
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
            if m and name:
                ins = m.group(2).strip()
                toks = ins.split()
                op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
                ops[op.split(".")[0]] += 1
                if hot:
                    keep.append(f"  /*{m.group(1)}*/ {ins} ;")
        if name:
            hist_lines.append(f"{name}: " + " ".join(f"{o}={c}" for o, c in ops.most_common(14)))
        with open(os.path.join(out_dir, src.replace(".cu", ".sass")), "w") as f:
            f.write("\n".join(keep) + "\n")
    with open(os.path.join(out_dir, "opcode_histogram.txt"), "w") as f:
        f.write("\n".join(hist_lines) + "\n")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--sass" in sys.argv:
        dump_sass(os.path.join(PKG, "..", "profiles", "sass"))
