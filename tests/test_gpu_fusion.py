"""SURVEY.md 8(f) N4 on the GPU: tmvs_fusibile_fwd (texture-unit sampling, one launch for all cameras, device-side
compaction) against the CPU restatement of gipuma/fusibile.  Parity of the restatement itself is UNPINNED (the
reference program needs CUDA + OpenCV and cannot run where the oracle was written)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import oracle
from transmvsnet_b200 import _lib, fusion, synthetic

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _probe(img: torch.Tensor, uv: np.ndarray, mode: int = 0) -> np.ndarray:     # mode 2: cudaArray texture
    lib = _lib.load()
    uv_d = torch.from_numpy(np.ascontiguousarray(uv, np.float32)).to(DEV)
    out = torch.empty((uv.shape[0], 4), dtype=torch.float32, device=DEV)
    rc = lib.tmvs_fusibile_tex_probe(ctypes.c_void_p(img.data_ptr()), img.shape[0], img.shape[1],
                                     ctypes.c_void_p(uv_d.data_ptr()), ctypes.c_void_p(out.data_ptr()), uv.shape[0], mode,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "tmvs_fusibile_tex_probe")
    return out.cpu().numpy()


def test_texture_unit_matches_the_oracles_filter_model():
    """The oracle models the hardware's bilinear fetch as: xB = x - 0.5, fractions rounded to 8 bits, ONE rounded
    product weight (w11 = round(a*b*256)/256) with the other three derived by subtraction, clamp to edge
    (scripts/probe_tex_filter.py found the rule).  Here the model meets the texture unit on random positions, texel
    centres, texel borders and outside coordinates: equal to 1 ulp, bit-identical on > 99 % of the samples."""
    rng = np.random.default_rng(7)
    h, w = 24, 32
    img = torch.from_numpy(rng.random((h, w, 4), dtype=np.float32) * np.array([1, 1, 1, 500], np.float32)
                           + np.array([0, 0, 0, 430], np.float32)).float().to(DEV).contiguous()
    uv = np.concatenate([
        rng.random((20000, 2)) * [w + 2, h + 2] - 1.0,
        np.array([[x + 0.5, y + 0.5] for y in range(h) for x in range(w)]),
        np.array([[x + 0.0, y + 0.25] for y in range(h) for x in range(w)]),
        (rng.integers(0, 256 * w, (20000, 1)) / 256.0 + 0.5) * [1, 0] + rng.random((20000, 2)) * [0, h],
    ]).astype(np.float32)
    hw = _probe(img, uv)
    emu = oracle.tex_linear(img.cpu().numpy(), uv)
    scale = np.array([1, 1, 1, 930], np.float32)
    err = np.abs(hw - emu) / scale
    assert err.max() <= 2e-7, (float(err.max()), uv[np.unravel_index(err.argmax(), err.shape)[0]])
    assert (hw == emu).mean() >= 0.99


@pytest.mark.parametrize("carry", [True, False])
def test_fusion_matches_oracle(carry):
    images, Ps = synthetic.make_fusion_scene(n_views=6, height=96, width=128, seed=4)
    cams = fusion.camera_records(Ps.numpy())
    ref = oracle.fusibile(images, cams, carry_over=carry)
    got = fusion.fuse_depth_maps(images.to(DEV), cams, carry_over=carry, ieee=True).cpu().numpy()   # the CPU restatement is IEEE
    assert len(ref) > 1000
    # borderline consistency decisions may flip where the texture unit and its model differ in the last bit
    assert abs(len(got) - len(ref)) <= max(2, len(ref) // 2000), (len(got), len(ref))
    if len(got) == len(ref):
        diff = np.abs(got - ref)
        close = diff.max(axis=1) <= 1e-3                      # mm / colour units; same pixel => same consistent set
        assert close.mean() >= 0.999
        assert np.array_equal(got[:, 3], ref[:, 3]) and np.array_equal(got[:, 7], ref[:, 7])


def test_fusion_is_deterministic_and_ordered():
    images, Ps = synthetic.make_fusion_scene(n_views=5, height=64, width=96, seed=5)
    cams = fusion.camera_records(Ps.numpy())
    a = fusion.fuse_depth_maps(images.to(DEV), cams)
    b = fusion.fuse_depth_maps(images.to(DEV), cams)
    assert torch.equal(a, b)
    own = fusion.fuse_depth_maps(images.to(DEV), cams, carry_over=False)
    assert 0 < len(own) < len(a)
    # without carry-over camera 0's block is a prefix of the carried output as well
    n0 = len(fusion.fuse_depth_maps(images[:1].repeat(5, 1, 1, 1).to(DEV), cams, consistent_threshold=0, carry_over=False)) // 5
    assert n0 > 0


def test_fusion_edge_cases():
    images, Ps = synthetic.make_fusion_scene(n_views=4, height=32, width=64, seed=6, hole_fraction=0.0, outlier_fraction=0.0)
    cams = fusion.camera_records(Ps.numpy())
    dev_img = images.to(DEV)
    assert len(fusion.fuse_depth_maps(dev_img, cams, consistent_threshold=4)) == 0            # only 3 other views exist
    dead = dev_img.clone()
    dead[..., 3] = 425.0
    assert len(fusion.fuse_depth_maps(dead, cams)) == 0                                        # depth floor, fusibile.cu:110
    with pytest.raises(_lib.TmvsError):
        fusion.fuse_depth_maps(dev_img, cams, capacity=16)                                     # more points than capacity
    with pytest.raises(_lib.TmvsError):
        fusion.fuse_depth_maps(images, cams)                                                   # CPU tensor: no fallback
    odd = fusion.fuse_depth_maps(dev_img[:, :, :63].contiguous(), cams)       # odd width: falls back to array textures
    assert len(odd) > 0


# ------------------------------------------------------------------------------- the pin: the reference's own kernel
def _reference_fusibile(images: torch.Tensor, cams: np.ndarray, depth_threshold=0.25, consistent_threshold=3, ieee=False):
    """gipuma/fusibile/fusibile.cu itself (kernel :89-173, copy_pc_to_host :175-210, per-camera loop :216-285), compiled
    for sm_100a by oracle/build.py build_fusibile_ref() into oracle/_ref/libfusibile_ref.so."""
    import os
    from oracle import build as oracle_build
    path = oracle_build.FUSE_REF_IEEE if ieee else oracle_build.FUSE_REF
    if not os.path.exists(path):
        if os.path.isdir(oracle_build.FUSE_SRC):
            pytest.fail("oracle/_ref/libfusibile_ref*.so is missing although the reference is present: run __graft_entry__.build()")
        pytest.skip("oracle/_ref/libfusibile_ref*.so is not in this tree and the reference's sources do not exist on this "
                    "box: it is compiled by __graft_entry__.build() in the build container and travels with the snapshot")
    lib = ctypes.CDLL(path)
    lib.fusibile_ref_run.restype = ctypes.c_int
    lib.fusibile_ref_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_float, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p]
    img = np.ascontiguousarray(images.cpu().numpy(), np.float32)
    v, h, w, _ = img.shape
    cams = np.ascontiguousarray(cams, np.float32)
    cap = v * h * w
    pts = np.zeros((cap, 8), np.float32)
    n = ctypes.c_longlong(0)
    rc = lib.fusibile_ref_run(img.ctypes.data, cams.ctypes.data, v, h, w, depth_threshold, consistent_threshold,
                              pts.ctypes.data, cap, ctypes.byref(n))
    assert rc == 0, rc
    return pts[:n.value]


@pytest.mark.parametrize("ieee", [False, True])
@pytest.mark.parametrize("shape", [(6, 96, 128, 4), (9, 288, 400, 11), (5, 64, 96, 5), (4, 33, 45, 6)])
def test_fusion_matches_the_reference_kernel(shape, ieee):
    """tmvs_fusibile_fwd against the reference's OWN compiled kernel on the same GPU, same float4 images, same camera
    records -- built with the reference's flags (-O3 --use_fast_math, the default arithmetic of the entry point) and
    without --use_fast_math (TMVS_FUSE_IEEE): the same points, in the same order (camera, y, x), BIT FOR BIT."""
    v, h, w, seed = shape
    images, Ps = synthetic.make_fusion_scene(n_views=v, height=h, width=w, seed=seed)
    cams = fusion.camera_records(Ps.numpy())
    ref = _reference_fusibile(images, cams, ieee=ieee)
    got = fusion.fuse_depth_maps(images.to(DEV), cams, carry_over=True, ieee=ieee).cpu().numpy()
    assert len(ref) > 500
    print(f"{shape} ieee={ieee}: reference {len(ref)} points, ours {len(got)}")
    assert len(got) == len(ref), (len(got), len(ref))
    exact = float((got == ref).all(axis=1).mean())
    print(f"   max diff {np.abs(got - ref).max():.3e}; rows bit-identical {exact:.4%}")
    assert np.array_equal(got, ref)
    # the reference's own texture set-up (one cudaArray per view) samples identically to the in-place textures
    again = fusion.fuse_depth_maps(images.to(DEV), cams, carry_over=True, ieee=ieee, array_textures=True).cpu().numpy()
    assert np.array_equal(again, got)
