"""Build variants of the backward kernels (occupancy tunables) HERE (nvcc cross-compiles without a GPU) into
build/variants/, then time each on the GPU box:
   python scripts/tune_bwd.py build        # in the build container
   python scripts/tune_bwd.py time [dtu|bld]     # on the GPU box (the variant libraries travel with the snapshot)"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from transmvsnet_b200 import build  # noqa: E402

OUT = os.path.join(REPO, "build", "variants")
VARIANTS = {
    "base": {},                                   # TMVS_GATHER_PF defaults to 2 (cells + registrants one plane ahead)
    "pf0": {"TMVS_GATHER_PF": 0},
    "pf1": {"TMVS_GATHER_PF": 1},
    "pf1d2": {"TMVS_GATHER_PF": 1, "TMVS_GATHER_PFD": 2},
}


def build_all():
    os.makedirs(OUT, exist_ok=True)
    flags = [f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    for tag, defs in VARIANTS.items():
        objs = []
        for src in build.SOURCES:
            tuned = src in ("tmvs_costvol_bwd.cu", "tmvs_costvol_bwd_cells.cu")
            obj = os.path.join(OUT, f"{tag if tuned else 'common'}_{src}.o")
            if tuned or not os.path.exists(obj):
                d = [f"-D{k}={v}" for k, v in defs.items()] if tuned else []
                subprocess.run(["nvcc", *flags, *d, "-c", os.path.join(build.CSRC, src), "-o", obj], check=True)
            objs.append(obj)
        subprocess.run(["nvcc", "-shared", "-o", os.path.join(OUT, f"libtmvs_{tag}.so"), *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a"], check=True)
        print("built", tag, flush=True)


def time_all(which):
    rows = []
    order = os.environ.get("TUNE_ORDER")
    for tag in (order.split(",") if order else VARIANTS):
        lib = os.path.join(OUT, f"libtmvs_{tag.split('@')[0]}.so")
        if not os.path.exists(lib):
            continue
        res = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "time_bwd.py"), which],
                             env=dict(os.environ, TMVS_LIB_PATH=lib), capture_output=True, text=True)
        line = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-300:]
        print(tag, line, flush=True)
        try:
            rows.append({"variant": tag, "defs": VARIANTS.get(tag.split("@")[0], {}), **json.loads(line)})
        except ValueError:
            pass
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(REPO, "gpurun_out", f"tune_bwd_{which}.json"), "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build_all()
    else:
        time_all(sys.argv[2] if len(sys.argv) > 2 else "dtu")
