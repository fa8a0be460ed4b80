"""fusibile depth-map fusion on the device (gipuma/fusibile).

Host-side mirror of what gipuma/fusibile/main.cpp and cameraGeometryUtils.h do before the kernel runs -- camera
records from the 3x4 `K [R | t]` matrices that test.py writes (test.py:40-66 `write_cam`), float4 images from the
BGRA PNGs whose alpha channel carries the quantised depth (utils.py:11-21, main.cpp:128-141) -- and the call into
`tmvs_fusibile_fwd`.  PyTorch owns every buffer; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import numpy as np
import torch

from . import _lib

CAM_FLOATS = 28


def _det3(m: np.ndarray) -> float:
    """cv::determinant of a 3x3 CV_32F matrix: float elements, double arithmetic."""
    m = m.astype(np.float64)
    return (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
            + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))


def camera_record(P) -> np.ndarray:
    """One camera's TMVS_FUSE_CAM_FLOATS floats from its 3x4 projection matrix (cameraGeometryUtils.h:104-156).

    P (12) | RK_inv = inverse(P[:, :3]) (9) | centre C (3) | P[:, 3] (3) | K[0][0] of the RQ-decomposed P[:, :3] (1)
    """
    P = np.asarray(P, dtype=np.float32).reshape(3, 4)
    M = P[:, :3]
    # cv::Mat::inv() of a 3x3 float matrix: adjugate / determinant in double, stored as float
    d = _det3(M)
    Md = M.astype(np.float64)
    adj = np.array([[Md[1, 1] * Md[2, 2] - Md[1, 2] * Md[2, 1], Md[0, 2] * Md[2, 1] - Md[0, 1] * Md[2, 2], Md[0, 1] * Md[1, 2] - Md[0, 2] * Md[1, 1]],
                    [Md[1, 2] * Md[2, 0] - Md[1, 0] * Md[2, 2], Md[0, 0] * Md[2, 2] - Md[0, 2] * Md[2, 0], Md[0, 2] * Md[1, 0] - Md[0, 0] * Md[1, 2]],
                    [Md[1, 0] * Md[2, 1] - Md[1, 1] * Md[2, 0], Md[0, 1] * Md[2, 0] - Md[0, 0] * Md[2, 1], Md[0, 0] * Md[1, 1] - Md[0, 1] * Md[1, 0]]])
    rk_inv = (adj * (1.0 / d)).astype(np.float32)
    # getCameraCenter (cameraGeometryUtils.h:21-49): signed 3x3 minors of P, then / C[3] in float
    c = np.array([_det3(P[:, [1, 2, 3]]), -_det3(P[:, [0, 2, 3]]), _det3(P[:, [0, 1, 3]]), -_det3(P[:, [0, 1, 2]])],
                 dtype=np.float32)
    centre = (c[:3] / c[3]).astype(np.float32)
    # cv::decomposeProjectionMatrix -> RQDecomp3x3 (double internally): upper-triangular K with K[0][0], K[1][1] > 0
    from scipy.linalg import rq
    K, _ = rq(Md)
    k00 = np.float32(abs(K[0, 0]))
    return np.concatenate([P.reshape(-1), rk_inv.reshape(-1), centre, P[:, 3], [k00]]).astype(np.float32)


def camera_records(Ps: Sequence) -> np.ndarray:
    return np.stack([camera_record(P) for P in Ps]).astype(np.float32)


def images_from_bgra(bgra_u8: torch.Tensor) -> torch.Tensor:
    """[V,H,W,4] uint8 BGRA (what cv::imread(IMREAD_UNCHANGED) returns for the PNGs of test.py) -> the float4 images of
    main.cpp:128-141: colour / 255, depth = 425 + 512 * (alpha / 255)."""
    img = bgra_u8.to(torch.float64) * (1.0 / 255.0)
    img = img.to(torch.float32)
    img[..., 3] = (425.0 + 512.0 * img[..., 3].to(torch.float64)).to(torch.float32)
    return img.contiguous()


def fuse_depth_maps(images: torch.Tensor, cams, depth_threshold: float = 0.25, consistent_threshold: int = 3,
                    carry_over: bool = True, capacity: int | None = None, array_textures: bool = False, ieee: bool = False) -> torch.Tensor:
    """images [V,H,W,4] fp32 CUDA (b, g, r, depth), cams [V,28] (camera_records) -> fused points [n,8]
    (x, y, z, 0, b, g, r, 0) in the reference's order.  carry_over=True reproduces the reference's output, including
    the points every later camera re-emits because the per-pixel buffer is never cleared (fusibile.cu:165-166,188).
    The images are sampled in place through pitch-linear textures when the buffer allows it (512-byte aligned, W even,
    H*W a multiple of 32), else -- or with array_textures=True -- through one cudaArray per view like the reference;
    both give the same samples.  ieee=True: the reference's source compiled WITHOUT its
    --use_fast_math flag (IEEE division / square root); the default mirrors the reference's own build."""
    lib = _lib.load()
    if not images.is_cuda:
        raise _lib.TmvsError("tmvs ops run on CUDA tensors only (no CPU fallback)")
    if images.dtype != torch.float32 or images.dim() != 4 or images.shape[3] != 4:
        raise _lib.TmvsError("images must be fp32 [V,H,W,4]")
    images = images.contiguous()
    v, h, w, _ = images.shape
    cams_h = torch.as_tensor(np.asarray(cams, dtype=np.float32)).contiguous()
    if tuple(cams_h.shape) != (v, CAM_FLOATS):
        raise _lib.TmvsError(f"cams must be [{v},{CAM_FLOATS}]")
    dev = images.device
    if capacity is None:
        capacity = v * h * w
    points = torch.empty((capacity, 8), dtype=torch.float32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = lib.tmvs_fusibile_workspace_bytes(v, h, w)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = lib.tmvs_fusibile_fwd(ctypes.c_void_p(images.data_ptr()), ctypes.c_void_p(cams_h.data_ptr()), v, h, w,
                                   float(depth_threshold), int(consistent_threshold),
                                   int(bool(carry_over)) | (2 if array_textures else 0) | (4 if ieee else 0),
                                   ctypes.c_void_p(points.data_ptr()), int(capacity), ctypes.c_void_p(count.data_ptr()),
                                   ctypes.c_void_p(ws.data_ptr()), ws_bytes,
                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "tmvs_fusibile_fwd")
    n = int(count.item())
    if n > capacity:
        raise _lib.TmvsError(f"fused point cloud has {n} points, capacity was {capacity}")
    return points[:n]
