"""Memory-safety pass without compute-sanitizer (closed on the development GPU pool): build the library with
-DTMVS_CHECK_BOUNDS -- every computed offset into a packed image, cell table or position map is tested against its
extent and the kernel traps on a violation -- and run the whole GPU test suite against that build.

    python scripts/check_bounds.py            (GPU box; ~1 min to build, then the suite)
"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from transmvsnet_b200 import build  # noqa: E402

out_dir = os.path.join(REPO, "gpurun_out", "checked")
os.makedirs(out_dir, exist_ok=True)
flags = [f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + ["-DTMVS_CHECK_BOUNDS=1"]
procs = []
for src in build.SOURCES:
    obj = os.path.join(out_dir, src.replace(".cu", ".o"))
    procs.append((obj, subprocess.Popen(["nvcc", *flags, "-c", os.path.join(build.CSRC, src), "-o", obj])))
objs = []
for obj, p in procs:
    if p.wait() != 0:
        sys.exit(f"nvcc failed for {obj}")
    objs.append(obj)
lib = os.path.join(out_dir, "libtmvs_sm100a_checked.so")
subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], check=True)
env = dict(os.environ, TMVS_LIB_PATH=lib)
rc = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REPO, "tests"), "-m", "gpu", "-q", "-x",
                     "--deselect", "tests/test_capi_symbols.py"], env=env, cwd=REPO).returncode
print("bounds-checked build:", "all GPU tests passed, no trap" if rc == 0 else f"FAILED (pytest exit {rc})")
import shutil
shutil.rmtree(out_dir, ignore_errors=True)
sys.exit(rc)
