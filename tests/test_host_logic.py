"""Host-side logic: camera algebra, synthetic inputs, CPU tensors are refused (no fallback)."""
import numpy as np
import pytest
import torch

from conftest import golden
import transmvsnet_b200 as tm
from transmvsnet_b200 import _lib, geometry, synthetic


def test_rot_trans_matches_reference_matrices():
    g = golden("depthnet_s2_given")
    rts = geometry.stage_rot_trans(torch.tensor(g["proj_matrix"]))
    assert rts.shape == g["rot_trans"].shape
    assert np.array_equal(rts.numpy(), g["rot_trans"])       # same torch ops as the reference -> same bits


def test_identity_relative_pose():
    p = torch.eye(4)[None].repeat(2, 1, 1)
    p[:, 0, 3] = 3.0
    rt = geometry.relative_rot_trans(p, p)
    assert torch.allclose(rt[:, :9].reshape(2, 3, 3), torch.eye(3)[None].repeat(2, 1, 1), atol=1e-6)
    assert torch.allclose(rt[:, 9:], torch.zeros(2, 3), atol=1e-5)


def test_synthetic_cascade_shapes_and_counts():
    stages = synthetic.make_cascade(batch=1, n_views=5, height=64, width=96, seed=0)
    assert [s.features[0].shape[1] for s in stages] == [32, 16, 8]
    assert [tuple(s.depth_values.shape) for s in stages] == [(1, 48, 16, 24), (1, 32, 32, 48), (1, 8, 64, 96)]
    assert stages[0].voxel_views == 48 * 16 * 24 * 4
    # stage-2/3 weights are the nearest x2 upsample of stage 1 (TransMVSNet.py:193-194)
    assert torch.equal(stages[1].view_weights[:, :, ::2, ::2], stages[0].view_weights)
    # hypotheses are increasing in d
    for s in stages:
        assert bool((s.depth_values[:, 1:] > s.depth_values[:, :-1]).all())
    again = synthetic.make_cascade(batch=1, n_views=5, height=64, width=96, seed=0)
    assert torch.equal(again[2].features[3], stages[2].features[3])          # seeded


def test_config2_voxel_views():
    # 140.08 M voxel-views per reference view at DTU 1152x1600, N=5
    total = sum(d * (1152 // s) * (1600 // s) * 4 for (_, d, s) in synthetic.STAGES)
    assert total == 140_083_200


def test_cpu_tensors_are_refused():
    x = torch.zeros(1, 8, 4, 4)
    with pytest.raises(_lib.TmvsError):
        tm.homo_warping(x, torch.eye(4)[None], torch.eye(4)[None], torch.ones(1, 2))
    with pytest.raises(_lib.TmvsError):
        tm.depth_wta(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(_lib.TmvsError):
        tm.depth_regression(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2))
    with pytest.raises(_lib.TmvsError):
        tm.softmax_wta(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 4))


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtmvs_sm100a.so")
    with pytest.raises(_lib.TmvsError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_depthnet_fold_cache_tracks_parameter_updates():
    net = tm.DepthNet().eval()
    a = net._folded_mlp()
    assert net._folded_mlp() is a                                   # cached
    with torch.no_grad():
        net.pixel_wise_net.conv2.bias.add_(1.0)                      # in-place update bumps the version counter
    b = net._folded_mlp()
    assert b is not a and float(b[-1] - a[-1]) == pytest.approx(1.0)
    net.load_state_dict(tm.DepthNet().state_dict())
    assert net._folded_mlp() is not b


def test_depthnet_state_dict_keys_match_reference():
    g = golden("depthnet_s1_learned")
    ref_keys = sorted(k[4:] for k in g if k.startswith("pwn."))
    ours = sorted(k[len("pixel_wise_net."):] for k in tm.DepthNet().state_dict())
    assert ours == ref_keys
