"""Build the CPU oracle (test infrastructure): oracle/_build/libtmvs_oracle.so and oracle/_ref/.

* libtmvs_oracle.so -- the repo's own C restatement (tmvs_oracle.c), compiled with gcc.
* oracle/_ref/      -- the REAL reference, staged for the GPU box.  The reference is pure Python (PyTorch): there is
  nothing to compile, but /root/reference does not exist on the GPU box, so build() copies the reference's model
  package (models/*.py, unmodified) from where it lies into the git-ignored oracle/_ref/models/ -- the Python analogue
  of compiling a C reference into oracle/_ref.  It travels with the gpurun snapshot (git-ignored, not gpurun-ignored)
  and never enters the history.  Users: tests/test_gpu_dropin_reference.py (the drop-in bound INTO the reference's own
  TransMVSNet.forward, patched against unpatched) and bench.py --impl reference (the reference's own functions on the
  host cores).  When /root/reference is absent (the GPU box) the staged copy is used as it is.
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
