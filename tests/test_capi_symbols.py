"""The C-ABI library loads and exports every symbol include/tmvs.h declares (no GPU needed)."""
import ctypes
import os
import re
import subprocess

from conftest import REPO
from transmvsnet_b200 import _lib


def _declared():
    text = open(os.path.join(REPO, "include", "tmvs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tmvs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    names = _declared()
    assert len(names) >= 12
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/tmvs.h but not exported"
    assert sorted(_lib.SIGNATURES) == names       # the ctypes table covers the header exactly


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for n in _declared():
        assert n in exported                       # unmangled => extern "C"


def test_version_and_error_strings():
    lib = _lib.load()
    assert lib.tmvs_version() == 200
    assert lib.tmvs_error_string(0) == b"ok"
    assert b"NULL" in lib.tmvs_error_string(-1)
    assert lib.tmvs_packed_bytes(4, 1, 32, 288, 400) == 4 * 8 * 288 * 400 * 16
    assert lib.tmvs_packed_bytes(4, 1, 30, 2, 2) == 4 * 2 * 8 * 8 * 16      # C -> 8 groups, W -> one 8-pixel block


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call (so this is safe on a CPU-only box)."""
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    assert lib.tmvs_softmax_wta_fwd(null, null, null, null, null, null, 1, 8, 4, 4, null) == -1
    assert lib.tmvs_softmax_wta_fwd(one, one, null, one, one, one, 1, 0, 4, 4, null) == -2
    assert lib.tmvs_softmax_wta_fwd(one, one, null, one, one, one, 1, 1000, 4, 4, null) == -2
    assert lib.tmvs_costvol_fwd(null, 0, 0, 0, 0, null, null, null, 1, null, null, null, 1, 8, 8, 4, 4, 2, 0, null) == -1
    assert lib.tmvs_costvol_fwd(one, 0, 0, 0, 0, one, one, one, 1, null, one, null, 1, 8, 8, 4, 4, 99, 0, null) == -2
    assert lib.tmvs_costvol_fwd(one, 0, 0, 0, 0, ctypes.c_void_p(20), one, one, 1, null, one, null,
                                1, 8, 8, 4, 4, 2, 0, null) == -3
    # per-view packed maps (scan cache): a NULL or misaligned entry of the host pointer array, view weights too small
    ptrs = (ctypes.c_void_p * 2)(16, 0)
    arr = ctypes.cast(ptrs, ctypes.c_void_p)
    assert lib.tmvs_costvol_fwd_cached(one, 0, 0, 0, 0, arr, one, one, 1, null, 0, 4, 4, one, null, 1, 8, 8, 4, 4, 2, 0, null) == -1
    ptrs[1] = 20
    assert lib.tmvs_costvol_fwd_cached(one, 0, 0, 0, 0, arr, one, one, 1, null, 0, 4, 4, one, null, 1, 8, 8, 4, 4, 2, 0, null) == -3
    ptrs[1] = 32
    assert lib.tmvs_costvol_fwd_cached(one, 0, 0, 0, 0, arr, one, one, 1, one, 1, 1, 2, null, one, 1, 8, 8, 4, 4, 2, 0, null) == -2
    # the drop-in warp's backward
    assert lib.tmvs_homo_warp_bwd(null, one, 1, one, one, one, 1 << 20, 1, 8, 8, 4, 4, 0, null) == -1
    assert lib.tmvs_homo_warp_bwd(one, one, 1, one, one, one, 16, 1, 8, 8, 4, 4, 0, null) == -2       # workspace too small
    assert lib.tmvs_homo_warp_bwd_workspace_bytes(1, 8, 8, 16, 24) > 0 and lib.tmvs_homo_warp_bwd_workspace_bytes(1, 0, 8, 16, 24) == 0
    assert lib.tmvs_depth_wta(one, one, null, null, 1, 8, 4, 4, null) == -1
    # fusion / peer-buffer entry points: NULL and shape checks come before any CUDA call
    assert lib.tmvs_fusibile_fwd(null, null, 4, 8, 8, 0.25, 3, 1, null, 16, null, null, 0, null) == -1
    assert lib.tmvs_fusibile_fwd(one, one, 1, 8, 8, 0.25, 3, 1, one, 16, one, one, 0, null) == -2        # one view: nothing to fuse
    assert lib.tmvs_fusibile_workspace_bytes(0, 8, 8) == 0 and lib.tmvs_fusibile_workspace_bytes(4, 8, 8) > 4 * 64 * 32
    assert lib.tmvs_fusibile_tex_probe(null, 8, 8, one, one, 4, 0, null) == -1
    assert lib.tmvs_peer_buffer_create(0, null, null) == -1 and lib.tmvs_peer_buffer_open(null, null) == -1
    assert lib.tmvs_peer_buffer_release(null, 1) == -1
    assert lib.tmvs_costvol_bwd_workspace_bytes(1, 8, 8, 16, 24, 2, 0) > 0 and lib.tmvs_costvol_bwd_workspace_bytes(0, 8, 8, 16, 24, 2, 0) == 0
    # the table cap travels with the call (no environment variable): a 1 MiB cap shrinks the workspace
    assert lib.tmvs_costvol_bwd_workspace_bytes(4, 8, 8, 64, 96, 4, _lib.f_table_mb(1)) < lib.tmvs_costvol_bwd_workspace_bytes(4, 8, 8, 64, 96, 4, 0)
