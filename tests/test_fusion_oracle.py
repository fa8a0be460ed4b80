"""SURVEY.md 8(f) N4 on the CPU: the fusibile restatement (oracle) and the host-side camera records."""
import numpy as np

from oracle import oracle
from transmvsnet_b200 import fusion, synthetic


def test_camera_record_identities():
    _, Ps = synthetic.make_fusion_scene(n_views=4, height=48, width=64, seed=3)
    for P in Ps.numpy():
        rec = fusion.camera_record(P)
        assert rec.shape == (fusion.CAM_FLOATS,) and rec.dtype == np.float32
        M, rk_inv, centre, p34, k00 = P[:, :3], rec[12:21].reshape(3, 3), rec[21:24], rec[24:27], rec[27]
        assert np.allclose(rk_inv.astype(np.float64) @ M.astype(np.float64), np.eye(3), atol=1e-4)
        assert np.allclose(P.astype(np.float64) @ np.append(centre, 1.0), 0.0, atol=0.5)      # P C = 0 (pixels * mm scale)
        assert np.array_equal(p34, P[:, 3])
        fx = synthetic._DTU_FX * 64 / 1600.0
        assert abs(k00 - fx) < 1e-3 * fx


def test_oracle_fuses_points_onto_the_plane():
    images, Ps = synthetic.make_fusion_scene(n_views=5, height=48, width=64, seed=1)
    cams = fusion.camera_records(Ps.numpy())
    own = oracle.fusibile(images, cams, carry_over=False)
    carried = oracle.fusibile(images, cams, carry_over=True)
    assert len(own) > 0.3 * 5 * 48 * 64                      # most pixels see the plane consistently in >= 3 other views
    assert len(carried) > len(own)                           # the reference re-emits earlier cameras' points
    normal = np.array([0.12, -0.08, 1.0]); normal /= np.linalg.norm(normal)
    dist = own[:, :3].astype(np.float64) @ normal - 680.0 * normal[2]
    # accepted neighbours are within 0.25 px of disparity: at this 64-pixel-wide toy size (f = 116 px, 100 mm baselines)
    # that is ~10 mm of depth, and the neighbour's point is taken at its integer pixel (fusibile.cu:154)
    assert np.abs(dist).max() < 12.0 and np.median(np.abs(dist)) < 1.0
    assert np.all(own[:, 3] == 0) and np.all(own[:, 7] == 0) # the reference's float4 operators zero w
    assert own[:, 4:7].min() >= 0 and own[:, 4:7].max() <= 1
    # first camera's block is identical with and without carry-over; outliers and holes never produce a point
    n0 = int((images[0, ..., 3] > 425.001).sum())
    assert len(own) < 5 * n0
    first = oracle.fusibile(images[:], cams, consistent_threshold=3, carry_over=False)
    assert np.array_equal(first, own)


def test_oracle_consistency_threshold_and_depth_floor():
    images, Ps = synthetic.make_fusion_scene(n_views=4, height=32, width=48, seed=2, hole_fraction=0.0, outlier_fraction=0.0)
    cams = fusion.camera_records(Ps.numpy())
    assert len(oracle.fusibile(images, cams, consistent_threshold=4, carry_over=False)) == 0      # only 3 other views
    n3 = len(oracle.fusibile(images, cams, consistent_threshold=3, carry_over=False))
    n1 = len(oracle.fusibile(images, cams, consistent_threshold=1, carry_over=False))
    assert 0 < n3 <= n1
    dead = images.clone()
    dead[..., 3] = 425.0                                     # <= 425.001: every pixel is skipped (fusibile.cu:110)
    assert len(oracle.fusibile(dead, cams, carry_over=True)) == 0


def test_oracle_texture_model_at_texel_centres_and_edges():
    rng = np.random.default_rng(0)
    img = rng.random((6, 8, 4), dtype=np.float32)
    centres = np.array([[x + 0.5, y + 0.5] for y in range(6) for x in range(8)], np.float32)
    assert np.array_equal(oracle.tex_linear(img, centres), img.reshape(-1, 4))                 # exact texels
    mid = oracle.tex_linear(img, np.array([[1.0, 0.5]], np.float32))[0]                         # halfway between texels 0 and 1
    assert np.allclose(mid, 0.5 * (img[0, 0] + img[0, 1]), atol=1e-7)
    edge = oracle.tex_linear(img, np.array([[0.0, 0.0], [8.0, 6.0]], np.float32))              # clamp to edge
    assert np.allclose(edge[0], img[0, 0]) and np.allclose(edge[1], img[5, 7])


def test_images_from_bgra_follows_main_cpp():
    """main.cpp:128-141: colour / 255, depth = 425 + 512 * (alpha / 255); the 8-bit alpha written by finalize_maps
    (utils.py:11-21) therefore decodes to 2 mm steps of the 425-935 mm range."""
    import torch
    bgra = torch.zeros(1, 2, 3, 4, dtype=torch.uint8)
    bgra[0, 0, 0] = torch.tensor([255, 128, 0, 0], dtype=torch.uint8)
    bgra[0, 1, 2] = torch.tensor([1, 2, 3, 255], dtype=torch.uint8)
    img = fusion.images_from_bgra(bgra)
    assert img.dtype == torch.float32 and tuple(img.shape) == (1, 2, 3, 4)
    assert float(img[0, 0, 0, 0]) == 1.0 and abs(float(img[0, 0, 0, 1]) - 128 / 255) < 1e-7
    assert float(img[0, 0, 0, 3]) == 425.0 and float(img[0, 1, 2, 3]) == 937.0
    assert float(img[0, 0, 1, 3]) == 425.0                    # alpha 0 -> below fusibile's 425.001 floor: skipped
