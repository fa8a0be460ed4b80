"""The oracle against the reference's own outputs (tests/golden, made by make_golden.py).

CPU only.  This is what pins the oracle: every function of oracle/tmvs_oracle.c and
oracle/torch_port.py is compared with vectors produced by the real reference code.
"""
import numpy as np
import pytest
import torch

from conftest import (COSTVOL_REL, DEPTH_FRAC, DEPTHNET_GIVEN, PROB_ABS, WARP_CASES, assert_costvol_close, golden,
                      rel_err)
from oracle import oracle, torch_port


@pytest.mark.parametrize("name", WARP_CASES)
def test_c_oracle_homo_warp(name):
    g = golden(name)
    out = oracle.homo_warp(g["src"], g["rot_trans"], g["depth"])
    assert out.shape == g["out"].shape
    assert_costvol_close(out, g["out"], name)
    # zero padding is exact: wherever the reference sampled nothing, so does the oracle
    assert np.array_equal(out == 0, g["out"] == 0)


@pytest.mark.parametrize("name", WARP_CASES)
def test_torch_port_homo_warp_bit_exact(name):
    g = golden(name)
    out = torch_port.homo_warp(torch.tensor(g["src"]), torch.tensor(g["src_proj"]), torch.tensor(g["ref_proj"]),
                               torch.tensor(g["depth"])).numpy()
    assert np.array_equal(out, g["out"])          # same ATen ops as the reference -> same bits


def test_identity_warp_reproduces_source():
    g = golden("warp_identity")
    out = oracle.homo_warp(g["src"], g["rot_trans"], g["depth"])
    for d in range(out.shape[2]):
        assert_costvol_close(out[:, :, d], g["src"], f"plane {d}")


def test_integer_shift_is_a_copy():
    g = golden("warp_integer")
    out = oracle.homo_warp(g["src"], g["rot_trans"], g["depth"])
    src = g["src"]
    # depth 8 -> shift (+1, -2): out[y, x] = src[y-2, x+1]
    d = list(g["depth"][0]).index(8.0)
    h, w = src.shape[2:]
    # (the normalise -> un-normalise round trip of module.py:311 + ATen is not exact, so neither is the copy)
    assert_costvol_close(out[0, :, d, 2:h, 0:w - 1], src[0, :, 0:h - 2, 1:w], "integer shift")
    assert_costvol_close(g["out"][0, :, d, 2:h, 0:w - 1], src[0, :, 0:h - 2, 1:w], "integer shift (reference)")


@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_c_oracle_cost_volume_and_readout(name):
    g = golden(name)
    feats = g["features"]
    views, agg = oracle.costvol_fwd(feats[0], feats[1:], g["rot_trans"], g["depth_values"], g["view_weights"])
    assert_costvol_close(agg, g["similarity"][:, 0], name)
    assert_costvol_close(oracle.aggregate_fwd(views, g["view_weights"]), g["similarity"][:, 0], name + " two-step")
    prob, idx, dep, conf = oracle.softmax_wta(g["similarity"][:, 0] * g["gain"], g["depth_values"])
    assert np.abs(prob - g["prob_volume"]).max() <= PROB_ABS
    assert np.array_equal(idx, g["index"])                       # integer output: bit-exact
    assert np.array_equal(dep, g["depth"])                       # gather of an input: bit-exact
    assert np.abs(conf - g["photo_confidence"]).max() <= PROB_ABS
    # depth_wta on the reference's own probabilities: bit-exact
    idx2, dep2 = oracle.depth_wta(g["prob_volume"], g["depth_values"])
    assert np.array_equal(idx2, g["index"]) and np.array_equal(dep2, g["depth"])


@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_torch_port_cost_volume(name):
    g = golden(name)
    feats = [torch.tensor(f) for f in g["features"]]
    agg, _ = torch_port.cost_volume(feats, torch.tensor(g["proj_matrix"]), torch.tensor(g["depth_values"]),
                                    torch.tensor(g["view_weights"]))
    assert np.array_equal(agg.numpy(), g["similarity"])
    prob, idx, dep, conf = torch_port.read_out(agg.squeeze(1) * float(g["gain"]), torch.tensor(g["depth_values"]))
    assert np.array_equal(prob.numpy(), g["prob_volume"]) and np.array_equal(idx.numpy(), g["index"])
    assert np.array_equal(dep.numpy(), g["depth"]) and np.array_equal(conf.numpy(), g["photo_confidence"])


@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_c_oracle_backward(name):
    g = golden(name)
    feats, vw = g["features"], g["view_weights"]
    coef = vw / (1e-5 + vw.sum(1, keepdims=True))                                   # d agg / d sim_i
    gviews = (g["grad_similarity"][:, 0][None] * coef.transpose(1, 0, 2, 3)[:, :, None]).astype(np.float32)
    gref, gsrc = oracle.costvol_bwd(feats[0], feats[1:], g["rot_trans"], g["depth_values"], gviews)
    assert_costvol_close(gref, g["grad_features"][0], name + " grad_ref")
    assert_costvol_close(gsrc, g["grad_features"][1:], name + " grad_src")


def test_c_oracle_stage1_learned_weights():
    """Stage 1: per-view similarity -> PixelwiseNet (PyTorch, weights from the fixture) -> aggregate."""
    from transmvsnet_b200.depthnet import PixelwiseNet
    g = golden("depthnet_s1_learned")
    feats = g["features"]
    views, _ = oracle.costvol_fwd(feats[0], feats[1:], g["rot_trans"], g["depth_values"], None)
    pwn = PixelwiseNet().eval()
    pwn.load_state_dict({k[4:]: torch.tensor(v) for k, v in g.items() if k.startswith("pwn.")})
    with torch.no_grad():
        vw = torch.cat([pwn(torch.tensor(views[i])[:, None]) for i in range(views.shape[0])], 1).numpy()
    assert np.abs(vw - g["view_weights"]).max() <= 1e-5
    assert_costvol_close(oracle.aggregate_fwd(views, vw), g["similarity"][:, 0], "stage-1 aggregate")


def test_c_oracle_pixelwise_net_matches_reference():
    """The plain-C PixelwiseNet (conv, then batch-norm, unfused) against the reference module's own output."""
    g = golden("depthnet_s1_learned")
    feats = g["features"]
    views, _ = oracle.costvol_fwd(feats[0], feats[1:], g["rot_trans"], g["depth_values"], None)
    state = {k[4:]: v for k, v in g.items() if k.startswith("pwn.")}
    vw = oracle.pixelwise_weights(views, state)
    assert np.abs(vw - g["view_weights"]).max() <= 1e-5


def test_folded_pixelwise_params_shape():
    from transmvsnet_b200 import fold_pixelwise_net, PixelwiseNet
    assert fold_pixelwise_net(PixelwiseNet().eval()).shape == (177,)


def test_torch_port_depth_hypotheses_bit_exact():
    g = golden("hypotheses")
    hw = tuple(int(v) for v in g["image_hw"])
    iv = float(g["depth_interval"])
    for stage, (nd, ratio, scale) in enumerate(((48, 4.0, 4), (32, 1.0, 2), (8, 0.5, 1)), start=1):
        cur = torch.tensor(g["depth_values"] if stage == 1 else g[f"prev{stage}"])
        out = torch_port.depth_hypotheses(cur, nd, ratio * iv, hw, scale).numpy()
        assert np.array_equal(out, g[f"hyp{stage}"]), stage


def test_wta_ties_first_maximal():
    g = golden("wta_ties")
    idx, dep = oracle.depth_wta(g["p"], g["depth_values"])
    assert np.array_equal(idx, g["index"]) and np.array_equal(dep, g["depth"])
    assert idx[0, 0, 0] == 0 and idx[1, 2, 3] == 2


def test_depth_regression_unpinned_definition():
    """depth_regression is absent from the reference fork: pinned only to the upstream definition."""
    g = golden("regression_unpinned")
    rng = float(g["depth_values_4d"].max() - g["depth_values_4d"].min())
    assert np.abs(oracle.depth_regression(g["p"], g["depth_values_4d"]) - g["depth_4d"]).max() <= DEPTH_FRAC * rng
    assert np.abs(oracle.depth_regression(g["p"], g["depth_values_2d"]) - g["depth_2d"]).max() <= DEPTH_FRAC * rng
    t = torch_port.depth_regression(torch.tensor(g["p"]), torch.tensor(g["depth_values_2d"])).numpy()
    assert np.array_equal(t, g["depth_2d"])
