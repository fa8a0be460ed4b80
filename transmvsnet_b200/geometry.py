"""Host-side camera algebra of the path (tiny 4x4 work; stays in PyTorch by design).

The kernels take 12 floats per (view, batch item): the 3x3 `rot` (row-major) and the
3-vector `trans` of   proj = src_proj @ inverse(ref_proj)   exactly as the reference
forms them (models/module.py:295-297), from projections composed the way
DepthNet.forward does (models/TransMVSNet.py:75-78).  They are computed with the same
torch ops as the reference so the kernels see bit-identical matrices.
"""
from __future__ import annotations

from typing import Sequence

import torch


def compose_projection(proj_pair: torch.Tensor) -> torch.Tensor:
    """[B,2,4,4] (extrinsic, intrinsic) -> [B,4,4] with the top 3x4 = K[:3,:3] @ E[:3,:4].

    Mirrors models/TransMVSNet.py:75-76.
    """
    out = proj_pair[:, 0].clone()
    out[:, :3, :4] = torch.matmul(proj_pair[:, 1, :3, :3], proj_pair[:, 0, :3, :4])
    return out


def _inverse(m: torch.Tensor) -> torch.Tensor:
    """torch.inverse(m) without its host synchronisation: torch.inverse is linalg.inv_ex followed by a check of the
    `info` tensor on the host (to raise on a singular matrix), which on a CUDA tensor blocks until the stream drains.
    inv_ex alone is the same computation, bit for bit; a singular projection yields inf/nan exactly as the check-free
    kernels downstream would propagate them."""
    return torch.linalg.inv_ex(m)[0]


def relative_rot_trans(src_proj: torch.Tensor, ref_proj: torch.Tensor, ref_inv: torch.Tensor = None) -> torch.Tensor:
    """[B,4,4] x2 -> [B,12] = (rot row-major, trans) of src_proj @ inverse(ref_proj).

    Mirrors models/module.py:295-297.  ref_inv: torch.inverse(ref_proj) if the caller already has it (the reference
    recomputes the same inverse for every source view).
    """
    proj = torch.matmul(src_proj, _inverse(ref_proj) if ref_inv is None else ref_inv)
    rot = proj[:, :3, :3].reshape(-1, 9)
    trans = proj[:, :3, 3]
    return torch.cat([rot, trans], dim=1).contiguous()


def stage_rot_trans(proj_matrix: torch.Tensor) -> torch.Tensor:
    """proj_matrix [B,N,2,4,4] -> [Nsrc,B,12] for every source view against view 0, ON THE DEVICE OF proj_matrix.

    The 4x4 algebra runs with the reference's torch ops where the matrices live, so the kernels see the matrices the
    reference would have computed on that device.  A CPU result is passed to the kernels by value; a CUDA result
    (the reference's test.py / train.py move the sample to the GPU first) is read by them in place -- there is no
    device-to-host copy and no synchronisation on the way (the round-1 version ended in .cpu()).
    """
    with torch.no_grad():
        pm = proj_matrix.detach().float()
        views = torch.unbind(pm, 1)
        ref = compose_projection(views[0])
        ref_inv = _inverse(ref)
        rts = torch.stack([relative_rot_trans(compose_projection(v), ref, ref_inv) for v in views[1:]], 0)
    return rts.contiguous()
