"""numpy/ctypes front-end of the CPU oracle (oracle/tmvs_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by transmvsnet_b200/.
Parity status: pinned against tests/golden/*.npz (outputs of the real reference);
depth_regression is "parity unpinned" (absent from the reference fork).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

from . import build as _build

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = _build.OUT
        if not os.path.exists(path) or (os.path.exists(_build.SRC)
                                        and os.path.getmtime(path) < os.path.getmtime(_build.SRC)):
            path = _build.build()
        _LIB = ctypes.CDLL(path)
        _LIB.tmvs_oracle_version.restype = ctypes.c_int
    return _LIB


def _f(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(_f32p)


def _c(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def homo_warp(src, rot_trans, depth) -> np.ndarray:
    """src [B,C,H,W], rot_trans [B,12], depth [B,D] or [B,D,H,W] -> [B,C,D,H,W]."""
    src, rot_trans, depth = _c(src), _c(rot_trans), _c(depth)
    b, c, h, w = src.shape
    d = depth.shape[1]
    out = np.empty((b, c, d, h, w), np.float32)
    lib().tmvs_oracle_homo_warp(_f(src), _f(rot_trans), _f(depth), int(depth.ndim == 4), _f(out),
                                b, c, d, h, w)
    return out


def costvol_fwd(ref, srcs, rot_trans, depth, weights=None, want_views=True):
    """ref [B,C,H,W]; srcs [Nsrc,B,C,H,W]; rot_trans [Nsrc,B,12]; weights [B,Nsrc,H,W] or None.

    Returns (sim_views [Nsrc,B,D,H,W] or None, agg [B,D,H,W] or None).
    """
    ref, srcs, rot_trans, depth = _c(ref), _c(srcs), _c(rot_trans), _c(depth)
    n, b, c, h, w = srcs.shape
    d = depth.shape[1]
    weights = None if weights is None else _c(weights)
    views = np.empty((n, b, d, h, w), np.float32) if want_views else None
    agg = np.empty((b, d, h, w), np.float32) if weights is not None else None
    lib().tmvs_oracle_costvol_fwd(_f(ref), _f(srcs), _f(rot_trans), _f(depth), int(depth.ndim == 4),
                                  _f(weights), _f(views), _f(agg), b, c, d, h, w, n)
    return views, agg


def aggregate_fwd(sim_views, weights) -> np.ndarray:
    sim_views, weights = _c(sim_views), _c(weights)
    n, b, d, h, w = sim_views.shape
    agg = np.empty((b, d, h, w), np.float32)
    lib().tmvs_oracle_aggregate_fwd(_f(sim_views), _f(weights), _f(agg), b, d, h, w, n)
    return agg


def costvol_bwd(ref, srcs, rot_trans, depth, grad_views):
    """grad_views = dL/d sim_i [Nsrc,B,D,H,W] -> (grad_ref [B,C,H,W], grad_src [Nsrc,B,C,H,W])."""
    ref, srcs, rot_trans, depth, grad_views = map(_c, (ref, srcs, rot_trans, depth, grad_views))
    n, b, c, h, w = srcs.shape
    d = depth.shape[1]
    gref = np.empty((b, c, h, w), np.float32)
    gsrc = np.empty((n, b, c, h, w), np.float32)
    lib().tmvs_oracle_costvol_bwd(_f(ref), _f(srcs), _f(rot_trans), _f(depth), int(depth.ndim == 4),
                                  _f(grad_views), _f(gref), _f(gsrc), b, c, d, h, w, n)
    return gref, gsrc


def softmax_wta(logits, depth_values, want_prob=True):
    """-> (prob [B,D,H,W] or None, idx int64 [B,H,W], depth [B,H,W], conf [B,H,W])."""
    logits, depth_values = _c(logits), _c(depth_values)
    b, d, h, w = logits.shape
    prob = np.empty((b, d, h, w), np.float32) if want_prob else None
    idx = np.empty((b, h, w), np.int64)
    dep = np.empty((b, h, w), np.float32)
    conf = np.empty((b, h, w), np.float32)
    lib().tmvs_oracle_softmax_wta(_f(logits), _f(depth_values), _f(prob), idx.ctypes.data_as(_i64p),
                                  _f(dep), _f(conf), b, d, h, w)
    return prob, idx, dep, conf


def depth_wta(p, depth_values):
    """-> (idx int64 [B,H,W], depth [B,H,W])."""
    p, depth_values = _c(p), _c(depth_values)
    b, d, h, w = p.shape
    idx = np.empty((b, h, w), np.int64)
    dep = np.empty((b, h, w), np.float32)
    lib().tmvs_oracle_depth_wta(_f(p), _f(depth_values), idx.ctypes.data_as(_i64p), _f(dep), b, d, h, w)
    return idx, dep


def depth_regression(p, depth_values) -> np.ndarray:
    p, depth_values = _c(p), _c(depth_values)
    b, d, h, w = p.shape
    dep = np.empty((b, h, w), np.float32)
    lib().tmvs_oracle_depth_regression(_f(p), _f(depth_values), int(depth_values.ndim == 4), _f(dep),
                                       b, d, h, w)
    return dep


def pixelwise_weights(sim_views, state) -> np.ndarray:
    """Eval-mode PixelwiseNet of every view.  sim_views [N,B,D,H,W]; state = the reference module's state_dict
    as numpy arrays (keys conv0.conv.weight, conv0.bn.weight, ...) -> view_weights [B,N,H,W]."""
    sim_views = _c(sim_views)
    n, b, d, h, w = sim_views.shape
    g = lambda k: _c(np.asarray(state[k]).reshape(-1))
    bn = lambda pre: _c(np.stack([np.asarray(state[pre + s]).reshape(-1) for s in
                                  (".weight", ".bias", ".running_mean", ".running_var")]))
    w0, w1, w2 = g("conv0.conv.weight"), g("conv1.conv.weight"), g("conv2.weight")
    bn0, bn1 = bn("conv0.bn"), bn("conv1.bn")
    b2 = float(np.asarray(state["conv2.bias"]).reshape(-1)[0])
    out = np.empty((b, n, h, w), np.float32)
    for i in range(n):
        wi = np.empty((b, h, w), np.float32)
        lib().tmvs_oracle_pixelwise_weight(_f(np.ascontiguousarray(sim_views[i])), _f(w0), _f(bn0), _f(w1), _f(bn1),
                                           _f(w2), ctypes.c_float(b2), ctypes.c_float(1e-5), _f(wi), b, d, h, w)
        out[:, i] = wi
    return out


def fusibile(images, cams, depth_threshold=0.25, consistent_threshold=3, carry_over=True):
    """images [V,H,W,4] (b, g, r, depth), cams [V,28] -> fused points [n,8]; gipuma/fusibile restated (parity unpinned)."""
    images, cams = _c(images), _c(cams)
    v, h, w, _ = images.shape
    cap = v * h * w
    pts = np.empty((cap, 8), np.float32)
    scratch = np.empty((h * w, 8), np.float32)
    fn = lib().tmvs_oracle_fusibile
    fn.restype = ctypes.c_longlong
    n = fn(_f(images), _f(cams), v, h, w, ctypes.c_float(depth_threshold), int(consistent_threshold), int(bool(carry_over)),
           _f(pts), ctypes.c_longlong(cap), _f(scratch))
    return pts[:n].copy()


def tex_linear(img, uv):
    """img [H,W,4], uv [n,2] unnormalised texture coordinates -> [n,4]: the oracle's model of the bilinear texture fetch."""
    img, uv = _c(img), _c(uv)
    out = np.empty((uv.shape[0], 4), np.float32)
    lib().tmvs_oracle_tex_linear(_f(img), img.shape[0], img.shape[1], _f(uv), _f(out), uv.shape[0])
    return out
