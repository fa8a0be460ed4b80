"""PyTorch-facing operators over the C ABI (include/tmvs.h); CUDA tensors only.

Reference signatures are kept:
  homo_warping(src_fea, src_proj, ref_proj, depth_values)      models/module.py:284
  depth_wta(p, depth_values)                                   models/module.py:474
  depth_regression(p, depth_values)                            north_star (absent in this fork)
plus the fused forms that replace the body of DepthNet.forward (models/TransMVSNet.py:71-103):
  cost_volume(...), aggregate(...), softmax_wta(...).
PyTorch owns every buffer; the library only launches kernels on the current stream.
"""
from __future__ import annotations

import contextlib
import contextvars
import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .geometry import relative_rot_trans

# Which of the reference's two fp32 arithmetics the kernels follow is a PER-CALL flag of the C ABI (include/tmvs.h).
# The ops below run on CUDA tensors, so their default is the arithmetic of the device the reference would have run on:
# ATen's CUDA kernels ("cuda": `x / ((W-1)/2)` as a reciprocal multiply).  "cpu" follows ATen's CPU kernels (true
# division) -- what the CPU-generated golden vectors (and the CPU checker of the test suite) pin.  Every op takes `arith=`; a caller that
# cannot pass it (code under test behind the reference's signatures) scopes it with `reference_arithmetic(...)`,
# which is context-local (contextvars), not process-wide.  TMVS_ARITH in the environment only seeds that default for
# test runs; the library itself reads no environment variable.
_ARITH = contextvars.ContextVar("tmvs_arith", default=os.environ.get("TMVS_ARITH", "cuda"))
_ARITH_FLAG = {"cpu": 0, "ieee": 0, "cuda": _lib.F_ARITH_ATEN_CUDA}
# experiment switches of the same kind (scripts/, tests): extra flag bits OR-ed into every call made in the context
_EXTRA_FLAGS = contextvars.ContextVar("tmvs_extra_flags", default=0)


@contextlib.contextmanager
def reference_arithmetic(which: str):
    """with reference_arithmetic("cpu"): ...   -- calls made inside follow ATen's CPU arithmetic by default."""
    if which not in _ARITH_FLAG:
        raise ValueError(f"arith must be one of {sorted(_ARITH_FLAG)}, got {which!r}")
    tok = _ARITH.set(which)
    try:
        yield
    finally:
        _ARITH.reset(tok)


@contextlib.contextmanager
def extra_flags(bits: int):
    """with extra_flags(_lib.F_FWD_TMA): ...   -- OR `bits` into the flags of every call made inside."""
    tok = _EXTRA_FLAGS.set(_EXTRA_FLAGS.get() | int(bits))
    try:
        yield
    finally:
        _EXTRA_FLAGS.reset(tok)


# torch.matmul(rot[B,3,3], xyz[B,3,HW]) (models/module.py:305) is a library GEMM whose kernel -- and with it the order in
# which the K = 3 dot product is evaluated -- depends on the problem size: on B200 / cuBLAS 12 the small stage-1 maps are
# evaluated UNFUSED, ((r0*x) + (r1*y)) + r2, the larger ones as fma(r2, 1, fma(r1, y, r0*x)) (scripts/probe_matmul.py).
# To land on the reference's CUDA rays bit for bit at every size, the CUDA arithmetic asks the library itself: one
# torch.matmul of the call's shape per (device, B, H, W), compared on the host with both candidate orders, cached.
_RAY_ORDER: dict = {}


def _ray_bits(dev: torch.device, b: int, h: int, w: int, arith: Optional[str]) -> int:
    which = _ARITH.get() if arith is None else arith
    if which != "cuda":
        return 0                      # ATen's CPU path (MKL) is fused at every size
    key = (dev.index, b, h, w)
    if key not in _RAY_ORDER:
        import numpy as np
        with torch.no_grad():
            rot = torch.tensor([[1.0103, 0.0131, -7.7021], [-0.0113, 0.9907, 5.3009], [1.1e-5, -2.3e-5, 1.0007]],
                               dtype=torch.float32, device=dev)[None].repeat(b, 1, 1)
            ys = torch.arange(0, h, dtype=torch.float32, device=dev).view(h, 1).expand(h, w).reshape(-1)
            xs = torch.arange(0, w, dtype=torch.float32, device=dev).view(1, w).expand(h, w).reshape(-1)
            xyz = torch.stack((xs, ys, torch.ones_like(xs)))[None].repeat(b, 1, 1)
            got = torch.matmul(rot, xyz)[0]                                       # [3, HW], the library's answer
            idx = torch.linspace(0, h * w - 1, min(h * w, 8192), device=dev).long()
            got = got[:, idx].cpu().numpy()
            x, y = xs[idx].cpu().numpy().astype(np.float64), ys[idx].cpu().numpy().astype(np.float64)
            r = rot[0].cpu().numpy().astype(np.float64)
        f32 = lambda a: a.astype(np.float32).astype(np.float64)      # products of fp32 values are exact in fp64
        fused = np.stack([f32(f32(r[k, 1] * y + f32(r[k, 0] * x)) + r[k, 2]) for k in range(3)]).astype(np.float32)
        unfused = np.stack([f32(f32(f32(r[k, 0] * x) + f32(r[k, 1] * y)) + r[k, 2]) for k in range(3)]).astype(np.float32)
        _RAY_ORDER[key] = bool((unfused == got).all() and not (fused == got).all())
    return _lib.F_RAY_UNFUSED if _RAY_ORDER[key] else 0


def _flags(arith: Optional[str] = None, extra: int = 0) -> int:
    which = _ARITH.get() if arith is None else arith
    if which not in _ARITH_FLAG:
        raise ValueError(f"arith must be one of {sorted(_ARITH_FLAG)}, got {which!r}")
    return _ARITH_FLAG[which] | _EXTRA_FLAGS.get() | extra


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.TmvsError("tmvs ops run on CUDA tensors only (no CPU fallback)")
        if t.dtype != torch.float32:
            raise _lib.TmvsError(f"tmvs ops are fp32; got {t.dtype}")
        if dev is not None and t.device != dev:
            raise _lib.TmvsError("all tensors must be on the same device")
        dev = t.device
    return dev


def _rt_arg(rot_trans, shape: Tuple[int, ...], dev: torch.device) -> Tuple[torch.Tensor, int]:
    """rot/trans for the C ABI: (contiguous fp32 tensor, flag bits).  A host tensor is passed to the kernels by value;
    a tensor already on the device is read by them in place (TMVS_F_RT_DEVICE) -- no copy, no synchronisation."""
    rt = torch.as_tensor(rot_trans).detach()
    if tuple(rt.shape) != tuple(shape):
        raise _lib.TmvsError(f"rot_trans must be {list(shape)}, got {list(rt.shape)}")
    rt = rt.to(torch.float32).contiguous()
    if rt.is_cuda:
        if rt.device != dev:
            raise _lib.TmvsError("rot_trans lives on another device than the features")
        return rt, _lib.F_RT_DEVICE
    return rt, 0


def _check_packed(packed: torch.Tensor, n: Optional[int], b: int, c: int, h: int, w: int) -> None:
    """A packed map built for other (B, C, H, W) would be indexed with the wrong strides: refuse it."""
    want = (b, h, (w + 7) // 8, (c + 3) // 4, 8, 4)
    if n is not None:
        want = (n,) + want
    if tuple(packed.shape) != want or not packed.is_contiguous():
        raise _lib.TmvsError(f"packed sources must be a contiguous {list(want)} tensor (pack_sources of maps shaped like "
                             f"the reference features [{b},{c},{h},{w}]), got {list(packed.shape)}")


def _feature_strides(feats: Sequence[torch.Tensor]) -> Tuple[int, int, int, int]:
    st = feats[0].stride()
    for f in feats:
        if f.stride() != st or f.shape != feats[0].shape:
            raise _lib.TmvsError("source feature maps must share shape and strides")
    return st


def pack_sources(src_feas: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """N x [B,C,H,W] (any common strides) -> packed [N,B,H,Wb,C4,8,4] (kernel-native blocked channel-last).
    out: an existing buffer of that shape to pack into (e.g. a slot of a scan-level cache)."""
    lib = _lib.load()
    dev = _need_cuda(*src_feas)
    b, c, h, w = src_feas[0].shape
    n = len(src_feas)
    sb, sc, sh, sw = _feature_strides(src_feas)
    shape = (n, b, h, (w + 7) // 8, (c + 3) // 4, 8, 4)
    if out is not None:
        _need_cuda(out)
        if tuple(out.shape) != shape or not out.is_contiguous():
            raise _lib.TmvsError(f"pack_sources: out must be a contiguous {list(shape)} tensor, got {list(out.shape)}")
    packed = out if out is not None else torch.empty(shape, dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in src_feas])
    with torch.cuda.device(dev):
        rc = lib.tmvs_pack_sources(ctypes.cast(ptrs, ctypes.c_void_p), n, sb, sc, sh, sw, _ptr(packed),
                                   b, c, h, w, _flags("cpu") & _lib.F_PACK_LDG, _stream())
    _lib.check(rc, "tmvs_pack_sources")
    return packed


def _depth_mode(depth_values: torch.Tensor, b: int, h: int, w: int) -> int:
    if depth_values.dim() == 2:
        return 0
    if depth_values.dim() == 4 and tuple(depth_values.shape[2:]) == (h, w):
        return 1
    raise _lib.TmvsError(f"depth_values must be [B,D] or [B,D,{h},{w}], got {tuple(depth_values.shape)}")


def homo_warp_packed(packed_view: torch.Tensor, rot_trans, depth_values: torch.Tensor, channels: int,
                     width: int, arith: Optional[str] = None) -> torch.Tensor:
    """packed_view: one view's slice [B,H,Wb,C4,8,4] of pack_sources(); width = the unpadded W."""
    lib = _lib.load()
    dev = _need_cuda(packed_view, depth_values)
    b, h, w = packed_view.shape[0], packed_view.shape[1], width
    _check_packed(packed_view, None, b, channels, h, w)
    d = depth_values.shape[1]
    if depth_values.shape[0] != b:
        raise _lib.TmvsError(f"depth_values has batch {depth_values.shape[0]}, the features {b}")
    mode = _depth_mode(depth_values, b, h, w)
    depth_values = depth_values.contiguous()
    rt, rt_flag = _rt_arg(rot_trans, (b, 12), dev)
    out = torch.empty((b, channels, d, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_homo_warp_fwd(_ptr(packed_view), ctypes.c_void_p(rt.data_ptr()), _ptr(depth_values), mode,
                                    _ptr(out), b, channels, d, h, w,
                                    _flags(arith, rt_flag | _ray_bits(dev, b, h, w, arith)), _stream())
    _lib.check(rc, "tmvs_homo_warp_fwd")
    return out


class _HomoWarp(torch.autograd.Function):
    """Drop-in warp with the gradient of F.grid_sample wrt the source features (module.py:318-320); the grid is built
    under no_grad in the reference (:294-316), so cameras and depth hypotheses get no gradient there either."""

    @staticmethod
    def forward(ctx, src_fea, rt, depth_values, arith):
        packed = pack_sources([src_fea.detach()])
        ctx.save_for_backward(rt, depth_values)
        ctx.arith, ctx.shape = arith, tuple(src_fea.shape)
        return homo_warp_packed(packed[0], rt, depth_values, src_fea.shape[1], src_fea.shape[3], arith)

    @staticmethod
    def backward(ctx, grad_out):
        rt, depth_values = ctx.saved_tensors
        return homo_warp_backward(rt, depth_values, grad_out, ctx.shape, ctx.arith), None, None, None


def homo_warp_backward(rot_trans, depth_values: torch.Tensor, grad_out: torch.Tensor, src_shape,
                       arith: Optional[str] = None) -> torch.Tensor:
    """grad_out [B,C,D,H,W] -> grad_src [B,C,H,W]: the grid_sample scatter as a deterministic, atomic-free gather."""
    lib = _lib.load()
    dev = _need_cuda(depth_values, grad_out)
    b, c, h, w = src_shape
    d = depth_values.shape[1]
    if tuple(grad_out.shape) != (b, c, d, h, w):
        raise _lib.TmvsError(f"grad_out must be [{b},{c},{d},{h},{w}], got {list(grad_out.shape)}")
    mode = _depth_mode(depth_values, b, h, w)
    depth_values, grad_out = depth_values.contiguous(), grad_out.contiguous()
    rt, rt_flag = _rt_arg(rot_trans, (b, 12), dev)
    gsrc = torch.empty((b, c, h, w), dtype=torch.float32, device=dev)
    ws_bytes = lib.tmvs_homo_warp_bwd_workspace_bytes(b, c, d, h, w)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_homo_warp_bwd(ctypes.c_void_p(rt.data_ptr()), _ptr(depth_values), mode, _ptr(grad_out), _ptr(gsrc),
                                    _ptr(ws), ws_bytes, b, c, d, h, w,
                                    _flags(arith, rt_flag | _ray_bits(dev, b, h, w, arith)), _stream())
    _lib.check(rc, "tmvs_homo_warp_bwd")
    return gsrc


def homo_warping(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor,
                 depth_values: torch.Tensor, arith: Optional[str] = None) -> torch.Tensor:
    """Drop-in for models/module.py:284-322, differentiable wrt src_fea like the reference.

    src_fea [B,C,H,W]; src_proj, ref_proj [B,4,4]; depth_values [B,D] or [B,D,H,W] -> [B,C,D,H,W].
    The 4x4 algebra runs with the reference's torch ops on the device the projections live on and its result is
    handed to the kernel where it is (no host round trip).  Training through this function alone materialises the
    warped volume and its gradient, as the reference does; the fused cost_volume / DepthNet path never forms either.
    """
    _need_cuda(src_fea, depth_values)
    if depth_values.shape[0] != src_fea.shape[0]:
        raise _lib.TmvsError(f"depth_values has batch {depth_values.shape[0]}, src_fea {src_fea.shape[0]}")
    with torch.no_grad():
        rt = relative_rot_trans(src_proj.float(), ref_proj.float())
    if torch.is_grad_enabled() and src_fea.requires_grad:
        return _HomoWarp.apply(src_fea, rt, depth_values.detach(), arith)
    with torch.no_grad():
        packed = pack_sources([src_fea.detach()])
        return homo_warp_packed(packed[0], rt, depth_values.detach(), src_fea.shape[1], src_fea.shape[3], arith)


def cost_volume_packed(ref_fea: torch.Tensor, packed, rot_trans, depth_values: torch.Tensor,
                       view_weights: Optional[torch.Tensor], want_views: bool, want_agg: bool,
                       arith: Optional[str] = None, vw_shift: int = 0
                       ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """packed: pack_sources() of the N source maps [N,B,H,Wb,C4,8,4], or a list of N per-view packed maps
    [B,H,Wb,C4,8,4] living anywhere on the device (a scan-level cache).  vw_shift: view_weights are given at the
    resolution of a coarser stage [B,N,ceil(H/2^s),ceil(W/2^s)] and read as their nearest-x2 upsampling
    (models/TransMVSNet.py:193-194) without materialising it."""
    lib = _lib.load()
    per_view = isinstance(packed, (list, tuple))
    dev = _need_cuda(ref_fea, depth_values, view_weights, *(packed if per_view else [packed]))
    b, c, h, w = ref_fea.shape
    n = len(packed) if per_view else packed.shape[0]
    if per_view:
        for pv in packed:
            _check_packed(pv, None, b, c, h, w)
    else:
        _check_packed(packed, n, b, c, h, w)
    if depth_values.shape[0] != b:
        raise _lib.TmvsError(f"depth_values has batch {depth_values.shape[0]}, the features {b}")
    d = depth_values.shape[1]
    mode = _depth_mode(depth_values, b, h, w)
    depth_values = depth_values.contiguous()
    rt, rt_flag = _rt_arg(rot_trans, (n, b, 12), dev)
    vw_h, vw_w = h, w
    if want_agg:
        if view_weights is None:
            raise _lib.TmvsError("aggregation needs view_weights")
        sh = 1 << vw_shift
        vw_h, vw_w = (h + sh - 1) // sh, (w + sh - 1) // sh
        if view_weights.dim() != 4 or tuple(view_weights.shape[:2]) != (b, n) or view_weights.shape[2] < vw_h \
                or view_weights.shape[3] < vw_w or (vw_shift == 0 and tuple(view_weights.shape[2:]) != (h, w)):
            raise _lib.TmvsError(f"view_weights must be [{b},{n},{vw_h},{vw_w}] (vw_shift={vw_shift}), "
                                 f"got {tuple(view_weights.shape)}")
        view_weights = view_weights.contiguous()
        vw_h, vw_w = view_weights.shape[2], view_weights.shape[3]
    views = torch.empty((n, b, d, h, w), dtype=torch.float32, device=dev) if want_views else None
    agg = torch.empty((b, d, h, w), dtype=torch.float32, device=dev) if want_agg else None
    rb, rc_, rh, rw = ref_fea.stride()
    flags = _flags(arith, rt_flag | _ray_bits(dev, b, h, w, arith))
    with torch.cuda.device(dev):
        if per_view or vw_shift:
            ptrs = (ctypes.c_void_p * n)(*[(packed[i] if per_view else packed[i]).data_ptr() for i in range(n)])
            rc = lib.tmvs_costvol_fwd_cached(_ptr(ref_fea), rb, rc_, rh, rw, ctypes.cast(ptrs, ctypes.c_void_p),
                                             ctypes.c_void_p(rt.data_ptr()), _ptr(depth_values), mode,
                                             _ptr(view_weights if want_agg else None), vw_shift, vw_h, vw_w,
                                             _ptr(views), _ptr(agg), b, c, d, h, w, n, flags, _stream())
        else:
            rc = lib.tmvs_costvol_fwd(_ptr(ref_fea), rb, rc_, rh, rw, _ptr(packed), ctypes.c_void_p(rt.data_ptr()),
                                      _ptr(depth_values), mode, _ptr(view_weights if want_agg else None), _ptr(views),
                                      _ptr(agg), b, c, d, h, w, n, flags, _stream())
    _lib.check(rc, "tmvs_costvol_fwd")
    return agg, views


def _aggregate_fwd(sim_views: torch.Tensor, view_weights: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    dev = _need_cuda(sim_views, view_weights)
    n, b, d, h, w = sim_views.shape
    if tuple(view_weights.shape) != (b, n, h, w):
        raise _lib.TmvsError(f"view_weights must be [{b},{n},{h},{w}], got {tuple(view_weights.shape)}")
    sim_views, view_weights = sim_views.contiguous(), view_weights.contiguous()
    agg = torch.empty((b, d, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_aggregate_fwd(_ptr(sim_views), _ptr(view_weights), _ptr(agg), b, d, h, w, n, _stream())
    _lib.check(rc, "tmvs_aggregate_fwd")
    return agg


class _Aggregate(torch.autograd.Function):
    """agg = sum_i s_i w_i / (1e-5 + sum_i w_i); gradients to both s_i and w_i (stage-1 training)."""

    @staticmethod
    def forward(ctx, sim_views, view_weights):
        agg = _aggregate_fwd(sim_views.detach(), view_weights.detach())
        ctx.save_for_backward(sim_views.detach(), view_weights.detach(), agg)
        return agg

    @staticmethod
    def backward(ctx, g):
        sim_views, vw, agg = ctx.saved_tensors
        wsum = vw.sum(1, keepdim=True) + 1e-5                               # [B,1,H,W]
        gs = gw = None
        if ctx.needs_input_grad[0]:
            gs = g.unsqueeze(0) * (vw / wsum).permute(1, 0, 2, 3).unsqueeze(2)   # [N,B,D,H,W]
        if ctx.needs_input_grad[1]:
            gw = ((sim_views - agg.unsqueeze(0)) * g.unsqueeze(0)).sum(2).permute(1, 0, 2, 3) / wsum
        return gs, gw


def aggregate(sim_views: torch.Tensor, view_weights: torch.Tensor) -> torch.Tensor:
    """sim_views [N,B,D,H,W], view_weights [B,N,H,W] -> [B,D,H,W] (TransMVSNet.py:71-72,88-93)."""
    return _Aggregate.apply(sim_views, view_weights)


def depth_hypotheses(cur_depth: torch.Tensor, ndepth: int, depth_interval_pixel: float, image_hw: Tuple[int, int],
                     stage_scale: int) -> torch.Tensor:
    """Depth hypotheses of a cascade stage at the stage resolution.

    One kernel for models/TransMVSNet.py:174-190 + 202-204: cur_depth is depth_values [B,192] (stage 1, the
    2-D branch of get_depth_samples) or the previous stage's depth map [B,hp,wp] (bilinear upsample to the
    image size, +- ndepth/2 * interval, trilinear resample to [ndepth, H/scale, W/scale]).
    """
    lib = _lib.load()
    dev = _need_cuda(cur_depth)
    cur = cur_depth.detach().contiguous()
    b = cur.shape[0]
    if image_hw[0] % stage_scale or image_hw[1] % stage_scale:
        # the kernel blends the two centre pixels with weight 1/2, which is what F.interpolate(align_corners=False)
        # does only for an integer ratio (models/TransMVSNet.py:202-204)
        raise _lib.TmvsError(f"image size {tuple(image_hw)} is not divisible by the stage scale {stage_scale}")
    h, w = image_hw[0] // stage_scale, image_hw[1] // stage_scale
    out = torch.empty((b, ndepth, h, w), dtype=torch.float32, device=dev)
    planes, hp, wp = (cur.shape[1], 0, 0) if cur.dim() == 2 else (0, cur.shape[1], cur.shape[2])
    with torch.cuda.device(dev):
        rc = lib.tmvs_depth_hypotheses_fwd(_ptr(cur), planes, hp, wp, float(depth_interval_pixel), _ptr(out), b,
                                           ndepth, h, w, int(stage_scale), _stream())
    _lib.check(rc, "tmvs_depth_hypotheses_fwd")
    return out


def finalize_maps(depth: torch.Tensor, conf3: torch.Tensor, conf1: torch.Tensor, conf2: torch.Tensor,
                  conf_threshold: float = 0.01, depth_min: float = 425.0, depth_max: float = 935.0):
    """Read-out -> wire format on the device (test.py:119-158, utils.py:11-21).

    depth, conf3 [B,H,W]; conf1 [B,H/4,W/4]; conf2 [B,H/2,W/2] ->
    (masked depth [B,H,W], final confidence [B,H,W], 8-bit alpha [B,H,W] uint8).
    """
    lib = _lib.load()
    dev = _need_cuda(depth, conf3, conf1, conf2)
    b, h, w = depth.shape
    depth, conf3, conf1, conf2 = (t.detach().contiguous() for t in (depth, conf3, conf1, conf2))
    d_out = torch.empty_like(depth)
    c_out = torch.empty_like(depth)
    a_out = torch.empty((b, h, w), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_finalize_maps_fwd(_ptr(depth), _ptr(conf3), _ptr(conf1), conf1.shape[1], conf1.shape[2],
                                        _ptr(conf2), conf2.shape[1], conf2.shape[2], float(conf_threshold),
                                        float(depth_min), float(depth_max), _ptr(d_out), _ptr(c_out), _ptr(a_out),
                                        b, h, w, _stream())
    _lib.check(rc, "tmvs_finalize_maps_fwd")
    return d_out, c_out, a_out


def fold_pixelwise_net(pwn: torch.nn.Module) -> torch.Tensor:
    """PixelwiseNet (models/TransMVSNet.py:10-30) -> 177 floats with the eval-mode BatchNorm folded into the
    1x1x1 convolutions: w0[16], b0[16], w1[8,16], b1[8], w2[8], b2  (host tensor, passed by value)."""
    def fold(conv_bn):
        w = conv_bn.conv.weight.detach().double().flatten(1)                 # [out, in]
        bn = conv_bn.bn
        scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        return w * scale[:, None], bn.bias.detach().double() - bn.running_mean.detach().double() * scale
    w0, b0 = fold(pwn.conv0)
    w1, b1 = fold(pwn.conv1)
    w2 = pwn.conv2.weight.detach().double().flatten()
    b2 = pwn.conv2.bias.detach().double().flatten()
    flat = torch.cat([w0.flatten(), b0, w1.flatten(), b1, w2, b2]).float().cpu().contiguous()
    assert flat.numel() == 177
    return flat


def pixelwise_aggregate(sim_views: torch.Tensor, folded_mlp: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode PixelwiseNet + aggregation in one kernel.

    sim_views [N,B,D,H,W], folded_mlp = fold_pixelwise_net(net) -> (view_weights [B,N,H,W], agg [B,D,H,W]).
    """
    lib = _lib.load()
    dev = _need_cuda(sim_views)
    n, b, d, h, w = sim_views.shape
    sim_views = sim_views.detach().contiguous()
    mlp = folded_mlp.detach().to("cpu", torch.float32).contiguous()
    vw = torch.empty((b, n, h, w), dtype=torch.float32, device=dev)
    agg = torch.empty((b, d, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_pixelwise_aggregate_fwd(_ptr(sim_views), ctypes.c_void_p(mlp.data_ptr()), _ptr(vw), _ptr(agg),
                                              b, d, h, w, n, _stream())
    _lib.check(rc, "tmvs_pixelwise_aggregate_fwd")
    return vw, agg


def costvol_backward_packed(ref_fea: torch.Tensor, packed: torch.Tensor, rot_trans, depth_values: torch.Tensor,
                            grad_views: torch.Tensor, need_ref: bool = True, need_src: bool = True,
                            arith: Optional[str] = None, extra: int = 0
                            ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """dL/d sim_i [N,B,D,H,W] -> (grad_ref [B,C,H,W], grad_src [N,B,C,H,W]); deterministic, no fp atomics."""
    lib = _lib.load()
    dev = _need_cuda(ref_fea, packed, depth_values, grad_views)
    b, c, h, w = ref_fea.shape
    n = packed.shape[0]
    _check_packed(packed, n, b, c, h, w)
    if depth_values.shape[0] != b:
        raise _lib.TmvsError(f"depth_values has batch {depth_values.shape[0]}, the features {b}")
    d = depth_values.shape[1]
    if tuple(grad_views.shape) != (n, b, d, h, w):
        raise _lib.TmvsError(f"grad_views must be [{n},{b},{d},{h},{w}], got {list(grad_views.shape)}")
    mode = _depth_mode(depth_values, b, h, w)
    depth_values, grad_views = depth_values.contiguous(), grad_views.contiguous()
    rt, rt_flag = _rt_arg(rot_trans, (n, b, 12), dev)
    flags = _flags(arith, rt_flag | extra | _ray_bits(dev, b, h, w, arith))
    gref = torch.empty((b, c, h, w), dtype=torch.float32, device=dev) if need_ref else None
    gsrc = torch.empty((n, b, c, h, w), dtype=torch.float32, device=dev) if need_src else None
    ws_bytes = lib.tmvs_costvol_bwd_workspace_bytes(b, c, d, h, w, n, flags)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    rb, rc_, rh, rw = ref_fea.stride()
    with torch.cuda.device(dev):
        rc = lib.tmvs_costvol_bwd(_ptr(ref_fea), rb, rc_, rh, rw, _ptr(packed), ctypes.c_void_p(rt.data_ptr()),
                                  _ptr(depth_values), mode, _ptr(grad_views), _ptr(gref), _ptr(gsrc), _ptr(ws),
                                  ws_bytes, b, c, d, h, w, n, flags, _stream())
    _lib.check(rc, "tmvs_costvol_bwd")
    return gref, gsrc


class _CostVolume(torch.autograd.Function):
    """Fused cost volume with autograd to the features: saves the inputs, never the volume."""

    @staticmethod
    def forward(ctx, rot_trans, depth_values, view_weights, want_views, arith, ref_fea, *src_feas):
        for s in src_feas:
            if s.shape != ref_fea.shape:
                raise _lib.TmvsError(f"source features {list(s.shape)} do not match the reference {list(ref_fea.shape)}")
        packed = pack_sources([s.detach() for s in src_feas])
        want_agg = view_weights is not None
        arith = _ARITH.get() if arith is None else arith       # the backward may run in another context
        agg, views = cost_volume_packed(ref_fea.detach(), packed, rot_trans, depth_values.detach(),
                                        None if view_weights is None else view_weights.detach(),
                                        want_views, want_agg, arith)
        ctx.rot_trans = rot_trans
        ctx.arith = arith
        ctx.has_agg, ctx.has_views = want_agg, want_views
        ctx.save_for_backward(ref_fea.detach(), packed, depth_values.detach(),
                              None if view_weights is None else view_weights.detach())
        outs = []
        if want_agg:
            outs.append(agg)
        if want_views:
            outs.append(views)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        ref_fea, packed, depth_values, view_weights = ctx.saved_tensors
        n = packed.shape[0]
        gi = 0
        grad_views = None
        if ctx.has_agg:
            g = grads[gi]
            gi += 1
            if g is not None:
                # d agg / d sim_i = w_i / (1e-5 + sum w)   (TransMVSNet.py:88-93; weights are inputs here)
                wsum = view_weights.sum(1, keepdim=True) + 1e-5           # [B,1,H,W]
                coef = (view_weights / wsum).permute(1, 0, 2, 3)          # [N,B,H,W]
                grad_views = g.unsqueeze(0) * coef.unsqueeze(2)           # [N,B,D,H,W]
        if ctx.has_views:
            g = grads[gi]
            if g is not None:
                grad_views = g if grad_views is None else grad_views + g
        if grad_views is None:
            return (None,) * (6 + n)
        need_ref = ctx.needs_input_grad[5]
        need_src = any(ctx.needs_input_grad[6:])
        gref, gsrc = costvol_backward_packed(ref_fea, packed, ctx.rot_trans, depth_values, grad_views.contiguous(),
                                             need_ref, need_src, ctx.arith)
        src_grads = [gsrc[i] if (need_src and ctx.needs_input_grad[6 + i]) else None for i in range(n)]
        return (None, None, None, None, None, gref, *src_grads)


def cost_volume(ref_fea: torch.Tensor, src_feas: Sequence[torch.Tensor], rot_trans, depth_values: torch.Tensor,
                view_weights: Optional[torch.Tensor] = None, want_views: bool = False, arith: Optional[str] = None
                ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused view loop of DepthNet.forward (models/TransMVSNet.py:71-93).

    Returns (aggregated similarity [B,D,H,W] or None, per-view similarity [N,B,D,H,W] or None).
    Differentiable wrt ref_fea / src_feas (not wrt cameras, depths or weights -- as the reference,
    whose grid is built under no_grad and whose stage-2/3 weights are detached).
    """
    want_agg = view_weights is not None
    if not want_agg and not want_views:
        raise _lib.TmvsError("cost_volume: nothing to compute (no view_weights and want_views=False)")
    if want_agg and view_weights.requires_grad and torch.is_grad_enabled():
        # learned weights in the graph (stage-1 training): keep them differentiable
        _, views = cost_volume(ref_fea, src_feas, rot_trans, depth_values, None, True, arith)
        return aggregate(views, view_weights), (views if want_views else None)
    outs = _CostVolume.apply(rot_trans, depth_values, view_weights, want_views, arith, ref_fea, *src_feas)
    agg = outs[0] if want_agg else None
    views = outs[-1] if want_views else None
    return agg, views


def softmax_wta(logits: torch.Tensor, depth_values: torch.Tensor, want_prob: bool = True,
                out_depth: Optional[torch.Tensor] = None, out_conf: Optional[torch.Tensor] = None):
    """logits, depth_values [B,D,H,W] -> (prob or None, index int64 [B,H,W], depth [B,H,W], conf [B,H,W]).

    One pass for models/TransMVSNet.py:99-103 + models/module.py:474-482 (forward only).
    out_depth / out_conf: contiguous fp32 [B,H,W] tensors the kernel writes the maps into.  They may live on ANOTHER
    GPU of the box (a peer-mapped slot of sharding.PeerMapSink): the kernel's stores then cross NVLink themselves and
    the multi-GPU gather of the maps needs no collective.
    """
    lib = _lib.load()
    dev = _need_cuda(logits, depth_values)
    b, d, h, w = logits.shape
    if tuple(depth_values.shape) != (b, d, h, w):
        raise _lib.TmvsError("softmax_wta: depth_values must match logits [B,D,H,W]")
    logits, depth_values = logits.detach().contiguous(), depth_values.detach().contiguous()
    prob = torch.empty_like(logits) if want_prob else None
    index = torch.empty((b, h, w), dtype=torch.int64, device=dev)
    for name, t in (("out_depth", out_depth), ("out_conf", out_conf)):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32 or tuple(t.shape) != (b, h, w)
                              or not t.is_contiguous()):
            raise _lib.TmvsError(f"softmax_wta: {name} must be a contiguous fp32 CUDA tensor [{b},{h},{w}]")
    depth = out_depth if out_depth is not None else torch.empty((b, h, w), dtype=torch.float32, device=dev)
    conf = out_conf if out_conf is not None else torch.empty((b, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_softmax_wta_fwd(_ptr(logits), _ptr(depth_values), _ptr(prob), _ptr(index), _ptr(depth),
                                      _ptr(conf), b, d, h, w, _stream())
    _lib.check(rc, "tmvs_softmax_wta_fwd")
    return prob, index, depth, conf


def depth_wta_index(p: torch.Tensor, depth_values: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    dev = _need_cuda(p, depth_values)
    b, d, h, w = p.shape
    if tuple(depth_values.shape) != (b, d, h, w):
        raise _lib.TmvsError("depth_wta: depth_values must match p [B,D,H,W]")
    p, depth_values = p.detach().contiguous(), depth_values.detach().contiguous()
    index = torch.empty((b, h, w), dtype=torch.int64, device=dev)
    depth = torch.empty((b, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_depth_wta(_ptr(p), _ptr(depth_values), _ptr(index), _ptr(depth), b, d, h, w, _stream())
    _lib.check(rc, "tmvs_depth_wta")
    return index, depth


def depth_wta(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """Drop-in for models/module.py:474-482: winner-take-all depth [B,H,W]."""
    return depth_wta_index(p, depth_values)[1]


class _DepthRegression(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, depth_values):
        lib = _lib.load()
        dev = _need_cuda(p, depth_values)
        b, d, h, w = p.shape
        mode = _depth_mode(depth_values, b, h, w)
        pc, dv = p.detach().contiguous(), depth_values.detach().contiguous()
        depth = torch.empty((b, h, w), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.tmvs_depth_regression_fwd(_ptr(pc), _ptr(dv), mode, _ptr(depth), b, d, h, w, _stream())
        _lib.check(rc, "tmvs_depth_regression_fwd")
        ctx.save_for_backward(dv)
        ctx.mode = mode
        ctx.shape = (b, d, h, w)
        return depth

    @staticmethod
    def backward(ctx, grad_depth):
        lib = _lib.load()
        (dv,) = ctx.saved_tensors
        b, d, h, w = ctx.shape
        gd = grad_depth.contiguous()
        gp = torch.empty((b, d, h, w), dtype=torch.float32, device=gd.device)
        with torch.cuda.device(gd.device):
            rc = lib.tmvs_depth_regression_bwd(_ptr(gd), _ptr(dv), ctx.mode, _ptr(gp), b, d, h, w, _stream())
        _lib.check(rc, "tmvs_depth_regression_bwd")
        return gp, None


def depth_regression(p: torch.Tensor, depth_values: torch.Tensor) -> torch.Tensor:
    """north_star signature (upstream MVSNet): sum_d p * depth_values -> [B,H,W]; differentiable wrt p."""
    return _DepthRegression.apply(p, depth_values)
