"""ctypes binding of libtmvs_sm100a.so (the C ABI declared in include/tmvs.h).

There is deliberately no fallback: if the library is missing the import of any op raises, and
every op raises on non-CUDA tensors (north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int64, c_size_t, c_uint, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# TMVS_LIB_PATH: load another build of the same library (scripts/check_bounds.py uses it for the bounds-checking build)
LIB_PATH = os.environ.get("TMVS_LIB_PATH") or os.path.join(_PKG, "lib", "libtmvs_sm100a.so")

# every symbol include/tmvs.h declares: name -> (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    "tmvs_version": (c_int, []),
    "tmvs_peer_buffer_create": (c_int, [c_size_t, _P, _P]),
    "tmvs_peer_buffer_open": (c_int, [_P, _P]),
    "tmvs_peer_buffer_release": (c_int, [_P, c_int]),
    "tmvs_peer_copy_async": (c_int, [_P, _P, c_size_t, _P]),
    "tmvs_error_string": (ctypes.c_char_p, [c_int]),
    "tmvs_packed_bytes": (c_size_t, [c_int] * 5),
    "tmvs_pack_sources": (c_int, [_P, c_int, c_int64, c_int64, c_int64, c_int64, _P, c_int, c_int, c_int, c_int, c_uint, _P]),
    "tmvs_homo_warp_fwd": (c_int, [_P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_uint, _P]),
    "tmvs_homo_warp_bwd": (c_int, [_P, _P, c_int, _P, _P, _P, c_size_t, c_int, c_int, c_int, c_int, c_int, c_uint, _P]),
    "tmvs_homo_warp_bwd_workspace_bytes": (c_size_t, [c_int] * 5),
    "tmvs_costvol_fwd": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_int, _P, _P, _P,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_uint, _P]),
    "tmvs_costvol_fwd_cached": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_int, _P, c_int, c_int, c_int,
                                        _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_uint, _P]),
    "tmvs_aggregate_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "tmvs_finalize_maps_fwd": (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, c_int, ctypes.c_float, ctypes.c_float,
                                       ctypes.c_float, _P, _P, _P, c_int, c_int, c_int, _P]),
    "tmvs_depth_hypotheses_fwd": (c_int, [_P, c_int, c_int, c_int, ctypes.c_float, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "tmvs_pixelwise_aggregate_fwd": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "tmvs_fusibile_fwd": (c_int, [_P, _P, c_int, c_int, c_int, ctypes.c_float, c_int, c_int, _P, ctypes.c_longlong, _P, _P,
                                  c_size_t, _P]),
    "tmvs_fusibile_workspace_bytes": (c_size_t, [c_int] * 3),
    "tmvs_fusibile_tex_probe": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, _P]),
    "tmvs_costvol_bwd": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, c_int, _P, _P, _P, _P,
                                 c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_uint, _P]),
    "tmvs_costvol_bwd_workspace_bytes": (c_size_t, [c_int] * 6 + [c_uint]),
    "tmvs_softmax_wta_fwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "tmvs_depth_wta": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "tmvs_depth_regression_fwd": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
    "tmvs_depth_regression_bwd": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P]),
}

# per-call option bits (include/tmvs.h): the library keeps no process-wide state and reads no environment variable
F_ARITH_ATEN_CUDA = 0x1
F_RT_DEVICE = 0x2
F_FWD_TMA = 0x4
F_BWD_SCAN = 0x8
F_PACK_LDG = 0x10
F_FWD_SPLIT = 0x20
F_FWD_SWEEP = 0x40
F_RAY_UNFUSED = 0x80


def f_table_mb(mb: int) -> int:
    return int(mb) << 16


_LIB = None
LAUNCHES = 0          # C-ABI kernel launches issued by this process (bench.py reports it)


class TmvsError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise TmvsError(
                f"{LIB_PATH} is missing: build it with `python -m transmvsnet_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)            # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


def check(code: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if code != 0:
        msg = load().tmvs_error_string(code).decode()
        raise TmvsError(f"{what} failed: {msg} (code {code})")
