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

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .geometry import relative_rot_trans


def set_reference_arithmetic(which: str) -> None:
    """"cpu" (default): follow ATen's CPU arithmetic (IEEE division by (W-1)/2), the one the golden vectors pin;
    "cuda": follow ATen's CUDA arithmetic (multiply by the reciprocal) -- see include/tmvs.h."""
    mode = {"cpu": 0, "ieee": 0, "cuda": 1}[which]
    _lib.check(_lib.load().tmvs_set_reference_arithmetic(mode), "tmvs_set_reference_arithmetic")


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


def _host_rt(rot_trans) -> torch.Tensor:
    """rot/trans as a contiguous fp32 CPU tensor (it is passed to the kernels by value)."""
    rt = torch.as_tensor(rot_trans).detach().to("cpu", torch.float32).contiguous()
    return rt


def _feature_strides(feats: Sequence[torch.Tensor]) -> Tuple[int, int, int, int]:
    st = feats[0].stride()
    for f in feats:
        if f.stride() != st or f.shape != feats[0].shape:
            raise _lib.TmvsError("source feature maps must share shape and strides")
    return st


def pack_sources(src_feas: Sequence[torch.Tensor]) -> torch.Tensor:
    """N x [B,C,H,W] (any common strides) -> packed [N,B,H,Wb,C4,8,4] (kernel-native blocked channel-last)."""
    lib = _lib.load()
    dev = _need_cuda(*src_feas)
    b, c, h, w = src_feas[0].shape
    n = len(src_feas)
    sb, sc, sh, sw = _feature_strides(src_feas)
    packed = torch.empty((n, b, h, (w + 7) // 8, (c + 3) // 4, 8, 4), dtype=torch.float32, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[f.data_ptr() for f in src_feas])
    with torch.cuda.device(dev):
        rc = lib.tmvs_pack_sources(ctypes.cast(ptrs, ctypes.c_void_p), n, sb, sc, sh, sw, _ptr(packed),
                                   b, c, h, w, _stream())
    _lib.check(rc, "tmvs_pack_sources")
    return packed


def _depth_mode(depth_values: torch.Tensor, b: int, h: int, w: int) -> int:
    if depth_values.dim() == 2:
        return 0
    if depth_values.dim() == 4 and tuple(depth_values.shape[2:]) == (h, w):
        return 1
    raise _lib.TmvsError(f"depth_values must be [B,D] or [B,D,{h},{w}], got {tuple(depth_values.shape)}")


def homo_warp_packed(packed_view: torch.Tensor, rot_trans, depth_values: torch.Tensor, channels: int,
                     width: int) -> torch.Tensor:
    """packed_view: one view's slice [B,H,Wb,C4,8,4] of pack_sources(); width = the unpadded W."""
    lib = _lib.load()
    dev = _need_cuda(packed_view, depth_values)
    b, h, w = packed_view.shape[0], packed_view.shape[1], width
    d = depth_values.shape[1]
    mode = _depth_mode(depth_values, b, h, w)
    depth_values = depth_values.contiguous()
    rt = _host_rt(rot_trans)
    out = torch.empty((b, channels, d, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_homo_warp_fwd(_ptr(packed_view), ctypes.c_void_p(rt.data_ptr()), _ptr(depth_values), mode,
                                    _ptr(out), b, channels, d, h, w, _stream())
    _lib.check(rc, "tmvs_homo_warp_fwd")
    return out


def homo_warping(src_fea: torch.Tensor, src_proj: torch.Tensor, ref_proj: torch.Tensor,
                 depth_values: torch.Tensor) -> torch.Tensor:
    """Drop-in for models/module.py:284-322 (forward only; the fused path carries the autograd).

    src_fea [B,C,H,W]; src_proj, ref_proj [B,4,4]; depth_values [B,D] or [B,D,H,W] -> [B,C,D,H,W].
    Raises when a gradient would have to flow through the warped volume (training with only this function patched):
    the volume's general gradient does not have the rank-1 form the fused backward exploits, and returning a detached
    tensor would silently zero the feature gradients.  Training goes through cost_volume / DepthNet (patch_reference).
    """
    _need_cuda(src_fea, depth_values)
    if torch.is_grad_enabled() and (src_fea.requires_grad or depth_values.requires_grad):
        raise _lib.TmvsError("homo_warping is forward-only: for training use transmvsnet_b200.cost_volume / DepthNet "
                             "(patch_reference), whose autograd runs the atomic-free backward kernels")
    with torch.no_grad():
        rt = relative_rot_trans(src_proj.float(), ref_proj.float())
        packed = pack_sources([src_fea.detach()])
        return homo_warp_packed(packed[0], rt, depth_values.detach(), src_fea.shape[1], src_fea.shape[3])


def cost_volume_packed(ref_fea: torch.Tensor, packed: torch.Tensor, rot_trans, depth_values: torch.Tensor,
                       view_weights: Optional[torch.Tensor], want_views: bool, want_agg: bool
                       ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    lib = _lib.load()
    dev = _need_cuda(ref_fea, packed, depth_values, view_weights)
    b, c, h, w = ref_fea.shape
    n = packed.shape[0]
    d = depth_values.shape[1]
    mode = _depth_mode(depth_values, b, h, w)
    depth_values = depth_values.contiguous()
    rt = _host_rt(rot_trans)
    if tuple(rt.shape) != (n, b, 12):
        raise _lib.TmvsError(f"rot_trans must be [{n},{b},12], got {tuple(rt.shape)}")
    if want_agg:
        if view_weights is None:
            raise _lib.TmvsError("aggregation needs view_weights")
        if tuple(view_weights.shape) != (b, n, h, w):
            raise _lib.TmvsError(f"view_weights must be [{b},{n},{h},{w}], got {tuple(view_weights.shape)}")
        view_weights = view_weights.contiguous()
    views = torch.empty((n, b, d, h, w), dtype=torch.float32, device=dev) if want_views else None
    agg = torch.empty((b, d, h, w), dtype=torch.float32, device=dev) if want_agg else None
    rb, rc_, rh, rw = ref_fea.stride()
    with torch.cuda.device(dev):
        rc = lib.tmvs_costvol_fwd(_ptr(ref_fea), rb, rc_, rh, rw, _ptr(packed), ctypes.c_void_p(rt.data_ptr()),
                                  _ptr(depth_values), mode, _ptr(view_weights if want_agg else None), _ptr(views),
                                  _ptr(agg), b, c, d, h, w, n, _stream())
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
                            grad_views: torch.Tensor, need_ref: bool = True, need_src: bool = True
                            ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """dL/d sim_i [N,B,D,H,W] -> (grad_ref [B,C,H,W], grad_src [N,B,C,H,W]); deterministic, no fp atomics."""
    lib = _lib.load()
    dev = _need_cuda(ref_fea, packed, depth_values, grad_views)
    b, c, h, w = ref_fea.shape
    n = packed.shape[0]
    d = depth_values.shape[1]
    mode = _depth_mode(depth_values, b, h, w)
    depth_values, grad_views = depth_values.contiguous(), grad_views.contiguous()
    rt = _host_rt(rot_trans)
    gref = torch.empty((b, c, h, w), dtype=torch.float32, device=dev) if need_ref else None
    gsrc = torch.empty((n, b, c, h, w), dtype=torch.float32, device=dev) if need_src else None
    ws_bytes = lib.tmvs_costvol_bwd_workspace_bytes(b, c, d, h, w, n)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    rb, rc_, rh, rw = ref_fea.stride()
    with torch.cuda.device(dev):
        rc = lib.tmvs_costvol_bwd(_ptr(ref_fea), rb, rc_, rh, rw, _ptr(packed), ctypes.c_void_p(rt.data_ptr()),
                                  _ptr(depth_values), mode, _ptr(grad_views), _ptr(gref), _ptr(gsrc), _ptr(ws),
                                  ws_bytes, b, c, d, h, w, n, _stream())
    _lib.check(rc, "tmvs_costvol_bwd")
    return gref, gsrc


class _CostVolume(torch.autograd.Function):
    """Fused cost volume with autograd to the features: saves the inputs, never the volume."""

    @staticmethod
    def forward(ctx, rot_trans, depth_values, view_weights, want_views, ref_fea, *src_feas):
        packed = pack_sources([s.detach() for s in src_feas])
        want_agg = view_weights is not None
        agg, views = cost_volume_packed(ref_fea.detach(), packed, rot_trans, depth_values.detach(),
                                        None if view_weights is None else view_weights.detach(),
                                        want_views, want_agg)
        ctx.rot_trans = rot_trans
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
            return (None,) * (5 + n)
        need_ref = ctx.needs_input_grad[4]
        need_src = any(ctx.needs_input_grad[5:])
        gref, gsrc = costvol_backward_packed(ref_fea, packed, ctx.rot_trans, depth_values, grad_views.contiguous(),
                                             need_ref, need_src)
        src_grads = [gsrc[i] if (need_src and ctx.needs_input_grad[5 + i]) else None for i in range(n)]
        return (None, None, None, None, gref, *src_grads)


def cost_volume(ref_fea: torch.Tensor, src_feas: Sequence[torch.Tensor], rot_trans, depth_values: torch.Tensor,
                view_weights: Optional[torch.Tensor] = None, want_views: bool = False
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
        _, views = cost_volume(ref_fea, src_feas, rot_trans, depth_values, None, True)
        return aggregate(views, view_weights), (views if want_views else None)
    outs = _CostVolume.apply(rot_trans, depth_values, view_weights, want_views, ref_fea, *src_feas)
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
