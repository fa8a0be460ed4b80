"""Build the CPU oracle (test infrastructure): oracle/_build/libtmvs_oracle.so and oracle/_ref/.

* libtmvs_oracle.so -- the repo's own C restatement (tmvs_oracle.c), compiled with gcc.
* oracle/_ref/      -- the REAL reference, staged for the GPU box.  The reference is pure Python (PyTorch): there is
  nothing to compile, but /root/reference does not exist on the GPU box, so build() copies the reference's model
  package (models/*.py, unmodified) from where it lies into the git-ignored oracle/_ref/models/ -- the Python analogue
  of compiling a C reference into oracle/_ref.  It travels with the gpurun snapshot (git-ignored, not gpurun-ignored)
  and never enters the history.  Users: tests/test_gpu_dropin_reference.py (the drop-in bound INTO the reference's own
  TransMVSNet.forward, patched against unpatched) and bench.py --impl reference (the reference's own functions on the
  host cores).  When /root/reference is absent (the GPU box) the staged copy is used as it is.
* oracle/_ref/libfusibile_ref.so -- the reference's own depth-map fusion kernel (gipuma/fusibile/fusibile.cu), compiled
  for sm_100a from where it lies (build_fusibile_ref); pins tmvs_fusibile_fwd in tests/test_gpu_fusion.py.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/models"
REF_DIR = os.path.join(HERE, "_ref")
REF_FILES = ("__init__.py", "module.py", "TransMVSNet.py", "FMT.py", "dcn.py", "position_encoding.py")


def build_ref() -> str:
    """Stage the reference's model package under oracle/_ref/models (copy if /root/reference is present)."""
    dst = os.path.join(REF_DIR, "models")
    if os.path.isdir(REF_SRC):
        os.makedirs(dst, exist_ok=True)
        for name in REF_FILES:
            src = os.path.join(REF_SRC, name)
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(dst, name))
    return dst if os.path.exists(os.path.join(dst, "module.py")) else ""


FUSE_SRC = "/root/reference/gipuma/fusibile"
FUSE_REF = os.path.join(REF_DIR, "libfusibile_ref.so")             # the reference's own flags: -O3 --use_fast_math
FUSE_REF_IEEE = os.path.join(REF_DIR, "libfusibile_ref_ieee.so")   # the same source without --use_fast_math


def build_fusibile_ref(force: bool = False) -> str:
    """Compile the reference's OWN fusion kernel file (gipuma/fusibile/fusibile.cu, from where it lies) for sm_100a into
    oracle/_ref/libfusibile_ref.so, behind oracle/fusibile_ref/harness.cu.  The reference's CMake build needs OpenCV;
    the kernel file does not -- camera.h only pulls <opencv2/core/core.hpp> in for a host-side struct, for which
    oracle/fusibile_ref/opencv2/core/core.hpp is a 15-line stand-in.  nvcc cross-compiles without a GPU; the library
    travels to the GPU box with the snapshot.  Returns "" when neither the sources nor a built library are present."""
    harness = os.path.join(HERE, "fusibile_ref", "harness.cu")
    kernel = os.path.join(FUSE_SRC, "fusibile.cu")
    if not os.path.exists(kernel):
        return FUSE_REF if os.path.exists(FUSE_REF) else ""
    newest = max(os.path.getmtime(harness), os.path.getmtime(kernel))
    if not force and all(os.path.exists(f) and os.path.getmtime(f) >= newest for f in (FUSE_REF, FUSE_REF_IEEE)):
        return FUSE_REF
    os.makedirs(REF_DIR, exist_ok=True)
    nvcc = "/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc"
    # the reference's CUDA_NVCC_FLAGS (gipuma/fusibile/CMakeLists.txt:10) with the architecture replaced by sm_100a
    for out, flags in ((FUSE_REF, ["-O3", "--use_fast_math"]), (FUSE_REF_IEEE, ["-O3"])):
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", *flags, "-std=c++11", "-shared", "-Xcompiler", "-fPIC",
               "-I", os.path.join(HERE, "fusibile_ref"), "-I", FUSE_SRC, harness, kernel, "-o", out]
        subprocess.run(cmd, check=True)
    return FUSE_REF


def import_reference():
    """(models.module, models.TransMVSNet) of the staged reference, or None when it has not been staged.
    Test / bench infrastructure only -- the product package never imports this."""
    if not os.path.exists(os.path.join(REF_DIR, "models", "module.py")):
        return None
    import importlib
    import warnings
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module("models.module")
        net = importlib.import_module("models.TransMVSNet")
    return mod, net

SRC = os.path.join(HERE, "tmvs_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libtmvs_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
           "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_ref() or "reference not staged (/root/reference absent)")
    print(build_fusibile_ref(force="--force" in sys.argv) or "fusibile reference not built (/root/reference absent)")
