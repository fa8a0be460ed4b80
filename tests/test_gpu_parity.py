"""Parity of the sm_100a kernels (through the C ABI) with the reference's golden outputs and the oracle.

Run on the B200 box:  python -m pytest tests -m gpu
Tolerances are the ones in conftest.py (north_star): 1e-4 relative for cost volumes, 1e-3 of the depth
range for depths, bit-exact integer indices.
"""
import numpy as np
import pytest
import torch

from conftest import (COSTVOL_REL, DEPTH_FRAC, DEPTHNET_GIVEN, PROB_ABS, WARP_CASES, assert_costvol_close, golden,
                      rel_err)
import transmvsnet_b200 as tm
from transmvsnet_b200 import _lib, geometry, ops, pipeline, synthetic
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _cpu_reference_arithmetic():
    """The golden vectors and the C oracle were produced by the reference's CPU execution: every test of this file
    compares against them, so the kernels are asked (per call, include/tmvs.h TMVS_F_ARITH_ATEN_CUDA off) to follow
    ATen's CPU arithmetic.  The CUDA arithmetic -- the drop-in's default -- is pinned against the reference's stock
    CUDA execution in test_gpu_vs_stock_cuda.py / test_gpu_dropin_reference.py."""
    with ops.reference_arithmetic("cpu"):
        yield


def cu(a):
    return torch.as_tensor(a).to(DEV)


# ----------------------------------------------------------------------------- drop-in homo_warping
@pytest.mark.parametrize("name", WARP_CASES)
def test_homo_warping_golden(name):
    g = golden(name)
    out = tm.homo_warping(cu(g["src"]), cu(g["src_proj"]), cu(g["ref_proj"]), cu(g["depth"]))
    assert tuple(out.shape) == g["out"].shape
    assert_costvol_close(out.cpu().numpy(), g["out"], name)
    assert np.array_equal(out.cpu().numpy() == 0, g["out"] == 0)       # zero padding / z<1e-6 exactly


def test_pack_tma_and_register_transpose_paths_agree():
    """The layout pre-pass has a TMA-engine kernel (contiguous NCHW, C in {8,16,32,64}, W % 4 == 0) and the
    register-transpose kernels it falls back to: identical packed features wherever a pixel exists (the padding pixels
    of a row's last 8-pixel block are never read)."""
    torch.manual_seed(4)
    for b, c, h, w in ((1, 32, 24, 64), (2, 16, 17, 100), (1, 8, 9, 132), (2, 64, 5, 8), (1, 32, 288, 400)):
        feats = [torch.randn(b, c, h, w, device=DEV) for _ in range(3)]
        with ops.extra_flags(_lib.F_PACK_LDG):
            ref = ops.pack_sources(feats)
        got = ops.pack_sources(feats)                           # [N,B,H,Wb,C4,8,4]
        valid = (torch.arange(ref.shape[3], device=DEV)[:, None] * 8 + torch.arange(8, device=DEV)[None, :]) < w
        m = valid[None, None, None, :, None, :, None].expand_as(ref)
        assert torch.equal(got[m], ref[m]), (b, c, h, w)
        # and back to NCHW: channel 4g+k of pixel 8*blk+j
        nchw = got.permute(0, 1, 4, 6, 2, 3, 5).reshape(3, b, c, h, ref.shape[3] * 8)[..., :w]
        assert torch.equal(nchw, torch.stack(feats, 0))


@pytest.mark.parametrize("name", WARP_CASES)
def test_homo_warping_is_differentiable_like_the_reference(name):
    """models/module.py:318-320 is differentiable wrt src_fea through F.grid_sample; so is the drop-in, for an
    ARBITRARY upstream gradient [B,C,D,H,W] (golden: the reference's autograd on the CPU), deterministically."""
    g = golden(name)
    src = cu(g["src"]).requires_grad_(True)
    out = tm.homo_warping(src, cu(g["src_proj"]), cu(g["ref_proj"]), cu(g["depth"]))
    assert out.requires_grad
    out.backward(cu(g["grad_out"]))
    assert_costvol_close(src.grad.cpu().numpy(), g["grad_src"], name + " grad_src")
    src2 = cu(g["src"]).requires_grad_(True)
    tm.homo_warping(src2, cu(g["src_proj"]), cu(g["ref_proj"]), cu(g["depth"])).backward(cu(g["grad_out"]))
    assert torch.equal(src.grad, src2.grad)                          # no float atomics: bit-reproducible
    with torch.no_grad():
        assert tm.homo_warping(src, cu(g["src_proj"]), cu(g["ref_proj"]), cu(g["depth"])).requires_grad is False


def test_homo_warping_channels_last_and_oracle():
    st = synthetic.make_stage(2, batch=2, n_views=3, height=64, width=96, seed=9)
    src = st.features[1]
    rt = geometry.stage_rot_trans(st.proj_matrix)[0]
    want = oracle.homo_warp(src, rt, st.depth_values)
    packed = ops.pack_sources([cu(src).contiguous(memory_format=torch.channels_last)])
    got = ops.homo_warp_packed(packed[0], rt, cu(st.depth_values), src.shape[1], src.shape[3])
    assert_costvol_close(got.cpu().numpy(), want, "channels_last")


# ----------------------------------------------------------------------------- fused cost volume
@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_cost_volume_golden(name):
    g = golden(name)
    feats = [cu(f) for f in g["features"]]
    agg, views = tm.cost_volume(feats[0], feats[1:], g["rot_trans"], cu(g["depth_values"]), cu(g["view_weights"]),
                                want_views=True)
    assert_costvol_close(agg.cpu().numpy(), g["similarity"][:, 0], name)
    o_views, o_agg = oracle.costvol_fwd(g["features"][0], g["features"][1:], g["rot_trans"], g["depth_values"],
                                        g["view_weights"])
    assert_costvol_close(views.cpu().numpy(), o_views, name + " per-view vs oracle")
    assert_costvol_close(agg.cpu().numpy(), o_agg, name + " agg vs oracle")
    # the two-kernel stage-1 form gives the same aggregate
    agg2 = tm.aggregate(views, cu(g["view_weights"]))
    assert_costvol_close(agg2.cpu().numpy(), g["similarity"][:, 0], name + " two-step")
    # agg-only launch (the stage-2/3 hot configuration) is the same kernel family
    agg3, none = tm.cost_volume(feats[0], feats[1:], g["rot_trans"], cu(g["depth_values"]), cu(g["view_weights"]))
    assert none is None and torch.equal(agg3, agg)


def test_cost_volume_plane_hypotheses_and_batch_split():
    """[B,D] hypotheses (module.py:288) and a batch large enough to split across launches."""
    b, n = 9, 9            # 8 source views * 9 items > 64 parameter slots -> two launches
    st = synthetic.make_stage(1, batch=b, n_views=n, height=32, width=48, channels=8, num_depth=11, seed=4)
    planes = st.depth_values[:, :, 0, 0].contiguous()
    rt = geometry.stage_rot_trans(st.proj_matrix)
    srcs = torch.stack(st.features[1:], 0)
    o_views, o_agg = oracle.costvol_fwd(st.features[0], srcs, rt, planes, st.view_weights)
    agg, views = tm.cost_volume(cu(st.features[0]), [cu(f) for f in st.features[1:]], rt, cu(planes),
                                cu(st.view_weights), want_views=True)
    assert_costvol_close(views.cpu().numpy(), o_views, "planes per-view")
    assert_costvol_close(agg.cpu().numpy(), o_agg, "planes agg")


def test_cost_volume_stage_shapes_vs_oracle():
    """The three (C, D) kernel shapes of the cascade at a size the oracle finishes in seconds."""
    for stage in (1, 2, 3):
        st = synthetic.make_stage(stage, batch=1, n_views=5, height=160, width=224, seed=2)
        rt = geometry.stage_rot_trans(st.proj_matrix)
        _, o_agg = oracle.costvol_fwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values,
                                      st.view_weights, want_views=False)
        agg, _ = tm.cost_volume(cu(st.features[0]), [cu(f) for f in st.features[1:]], rt, cu(st.depth_values),
                                cu(st.view_weights))
        assert_costvol_close(agg.cpu().numpy(), o_agg, f"stage {stage}")


def test_cost_volume_full_size_stage3_vs_oracle():
    """BASELINE config-2 stage-3 size (1152x1600, the hardest case for fp32 coordinate rounding: ulp(1600) ~ 1e-4 px)
    against the CPU oracle, at the north_star tolerance."""
    st = synthetic.make_stage(3, batch=1, n_views=3, height=1152, width=1600, seed=7)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    _, o_agg = oracle.costvol_fwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values,
                                  st.view_weights, want_views=False)
    agg, _ = tm.cost_volume(cu(st.features[0]), [cu(f) for f in st.features[1:]], rt, cu(st.depth_values),
                            cu(st.view_weights))
    assert_costvol_close(agg.cpu().numpy(), o_agg, "full-size stage 3")


def test_tma_and_l1_paths_agree():
    """The TMA-staged shared-memory kernel and the L1 global-gather kernel land on the same sample positions and
    weights (same coordinate arithmetic); only the order of the channel sum differs (the L1 kernel accumulates even
    and odd channels separately for FFMA2), so they agree to fp32 re-association: <= 2e-6 of the volume's range.
    Cases: cascade shapes, a tiny image (box larger than the image) and a case whose window does not fit any box
    (6x zoom-in -> global path inside the TMA kernel)."""
    cases = []
    for stage, hw in ((1, (160, 224)), (2, (96, 136)), (3, (48, 72)), (3, (8, 8))):
        st = synthetic.make_stage(stage, batch=2, n_views=4, height=hw[0], width=hw[1], seed=13)
        cases.append((st, geometry.stage_rot_trans(st.proj_matrix)))
    st = synthetic.make_stage(2, batch=1, n_views=3, height=256, width=384, seed=14)
    rt = geometry.stage_rot_trans(st.proj_matrix).clone()
    rt[:, :, 0:6] *= 6.0
    rt[:, :, 9:11] *= 6.0
    cases.append((st, rt))
    for st, rt in cases:
        args = (cu(st.features[0]), [cu(f) for f in st.features[1:]], rt, cu(st.depth_values), cu(st.view_weights))
        with ops.extra_flags(_lib.F_FWD_TMA):
            agg_t, views_t = tm.cost_volume(*args, want_views=True)
        agg_l, views_l = tm.cost_volume(*args, want_views=True)
        for a, b in ((agg_t, agg_l), (views_t, views_l)):
            assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
        _, o_agg = oracle.costvol_fwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values,
                                      st.view_weights, want_views=False)
        assert_costvol_close(agg_t.cpu().numpy(), o_agg, "tma path vs oracle")


def test_cost_volume_linearity_full_size():
    """Size-independent property at the BASELINE config-2 stage-3 size: the volume is linear in the source
    features and in the reference features (checked without a CPU oracle)."""
    st = synthetic.make_stage(3, batch=1, n_views=3, height=1152, width=1600, seed=1)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    ref, srcs = cu(st.features[0]), [cu(f) for f in st.features[1:]]
    dv, vw = cu(st.depth_values), cu(st.view_weights)
    a, _ = tm.cost_volume(ref, srcs, rt, dv, vw)
    b2, _ = tm.cost_volume(ref, [2.0 * s for s in srcs], rt, dv, vw)
    assert torch.equal(b2, 2.0 * a)                                         # scaling by 2 is exact in fp32
    other = [torch.randn_like(s) for s in srcs]
    c, _ = tm.cost_volume(ref, other, rt, dv, vw)
    s, _ = tm.cost_volume(ref, [x + y for x, y in zip(srcs, other)], rt, dv, vw)
    e_max, e_l2 = rel_err((a + c).cpu().numpy(), s.cpu().numpy())
    assert e_max <= COSTVOL_REL and e_l2 <= COSTVOL_REL
    z, _ = tm.cost_volume(torch.zeros_like(ref), srcs, rt, dv, vw)
    assert not bool(z.any())


def test_identity_cameras_full_size_stage1():
    """src_proj == ref_proj at the config-2 stage-1 size: similarity_i = mean_c(ref*src) at every depth."""
    st = synthetic.make_stage(1, batch=1, n_views=2, height=1152, width=1600, seed=3)
    pm = st.proj_matrix.clone()
    pm[:, 1] = pm[:, 0]
    rt = geometry.stage_rot_trans(pm)
    ref, src = cu(st.features[0]), cu(st.features[1])
    _, views = tm.cost_volume(ref, [src], rt, cu(st.depth_values), None, want_views=True)
    want = (ref * src).mean(1)                                             # [B,H,W]
    got = views[0]
    for d in (0, 17, 47):
        e_max, e_l2 = rel_err(got[:, d].cpu().numpy(), want.cpu().numpy())
        assert e_max <= COSTVOL_REL and e_l2 <= COSTVOL_REL, (d, e_max, e_l2)


# ----------------------------------------------------------------------------- BASELINE.json configs 3-5 as parity cases
@pytest.mark.parametrize("n_views", [3, 5, 11])
@pytest.mark.parametrize("channels", [8, 16, 32])
@pytest.mark.parametrize("depths", [48, 96, 192])
def test_config5_sweep_vs_oracle(depths, channels, n_views):
    """configs[4]: D x C x N sweep (small map so the oracle takes a fraction of a second per case)."""
    st = synthetic.make_stage(1, batch=1, n_views=n_views, height=96, width=160, channels=channels,
                              num_depth=depths, seed=depths + channels + n_views)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    _, o_agg = oracle.costvol_fwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values,
                                  st.view_weights, want_views=False)
    agg, _ = tm.cost_volume(cu(st.features[0]), [cu(f) for f in st.features[1:]], rt, cu(st.depth_values),
                            cu(st.view_weights))
    assert_costvol_close(agg.cpu().numpy(), o_agg, f"D={depths} C={channels} N={n_views}")
    _, idx, dep, conf = tm.softmax_wta(cu(st.logits), cu(st.depth_values), want_prob=False)
    _, o_idx, o_dep, o_conf = oracle.softmax_wta(st.logits, st.depth_values, want_prob=False)
    assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(dep.cpu().numpy(), o_dep)
    assert np.abs(conf.cpu().numpy() - o_conf).max() <= PROB_ABS


def test_config3_tnt_shaped_vs_oracle():
    """configs[2]: Tanks&Temples-shaped cameras (fx=fy=0.6W, depth 0.5-10), N=7, at 1/4 x 1/4 of 1056x1920."""
    for st in synthetic.make_cascade(batch=1, n_views=7, height=264, width=480, kind="unit", seed=21):
        rt = geometry.stage_rot_trans(st.proj_matrix)
        _, o_agg = oracle.costvol_fwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values,
                                      st.view_weights, want_views=False)
        out = pipeline.run_stage(pipeline.stage_to_device(st, DEV))
        assert_costvol_close(out["similarity"].cpu().numpy(), o_agg, f"T&T-shaped stage {st.stage}")
        _, o_idx, o_dep, _ = oracle.softmax_wta(st.logits, st.depth_values, want_prob=False)
        assert np.array_equal(out["index"].cpu().numpy(), o_idx) and np.array_equal(out["depth"].cpu().numpy(), o_dep)


def test_config4_blendedmvs_batch8_forward_backward():
    """configs[3]: BlendedMVS-shaped N=7, batch 8, forward + atomic-free backward; 1/4-size maps against the
    oracle, then one full-size (576x768) batch-8 stage executed twice for bit-reproducibility."""
    st = synthetic.make_stage(2, batch=8, n_views=7, height=144, width=192, kind="unit", seed=31)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    feats = [cu(f).requires_grad_(True) for f in st.features]
    agg, _ = tm.cost_volume(feats[0], feats[1:], rt, cu(st.depth_values), cu(st.view_weights))
    g = torch.randn(agg.shape, generator=torch.Generator().manual_seed(1))
    agg.backward(cu(g))
    vw = st.view_weights
    coef = (vw / (1e-5 + vw.sum(1, keepdim=True))).permute(1, 0, 2, 3)
    gviews = (g[None] * coef[:, :, None]).contiguous()
    _, o_agg = oracle.costvol_fwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values, vw, want_views=False)
    o_ref, o_src = oracle.costvol_bwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values, gviews)
    assert_costvol_close(agg.detach().cpu().numpy(), o_agg, "B=8 forward")
    assert_costvol_close(feats[0].grad.cpu().numpy(), o_ref, "B=8 grad_ref")
    assert_costvol_close(torch.stack([f.grad for f in feats[1:]], 0).cpu().numpy(), o_src, "B=8 grad_src")
    big = synthetic.make_stage(3, batch=8, n_views=7, height=576, width=768, kind="unit", seed=32)
    rt = geometry.stage_rot_trans(big.proj_matrix)
    packed = ops.pack_sources([cu(f) for f in big.features[1:]])
    gv = torch.randn(6, *big.depth_values.shape, device=DEV)
    a = ops.costvol_backward_packed(cu(big.features[0]), packed, rt, cu(big.depth_values), gv)
    b2 = ops.costvol_backward_packed(cu(big.features[0]), packed, rt, cu(big.depth_values), gv)
    assert torch.equal(a[0], b2[0]) and torch.equal(a[1], b2[1])
    assert bool(torch.isfinite(a[0]).all()) and bool(torch.isfinite(a[1]).all())


# ----------------------------------------------------------------------------- backward (atomic-free)
@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_cost_volume_backward_golden(name):
    """Autograd of the fused volume against the reference's autograd (grid_sample backward + mean + mul)."""
    g = golden(name)
    feats = [cu(f).requires_grad_(True) for f in g["features"]]
    agg, _ = tm.cost_volume(feats[0], feats[1:], g["rot_trans"], cu(g["depth_values"]), cu(g["view_weights"]))
    agg.backward(cu(g["grad_similarity"][:, 0]))
    assert_costvol_close(feats[0].grad.cpu().numpy(), g["grad_features"][0], name + " grad_ref")
    for i, f in enumerate(feats[1:]):
        assert_costvol_close(f.grad.cpu().numpy(), g["grad_features"][1 + i], f"{name} grad_src[{i}]")


def _backward_once(st, rt, gviews):
    packed = ops.pack_sources([cu(f) for f in st.features[1:]])
    return ops.costvol_backward_packed(cu(st.features[0]), packed, rt, cu(st.depth_values), gviews)


def test_cost_volume_backward_vs_oracle_and_deterministic():
    for stage, hw in ((1, (96, 160)), (2, (72, 104)), (3, (40, 72))):
        st = synthetic.make_stage(stage, batch=2, n_views=3, height=hw[0], width=hw[1], seed=8)
        rt = geometry.stage_rot_trans(st.proj_matrix)
        n, (b, d, h, w) = len(st.features) - 1, st.depth_values.shape
        gv = torch.randn(n, b, d, h, w, generator=torch.Generator().manual_seed(3))
        o_ref, o_src = oracle.costvol_bwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values, gv)
        gref, gsrc = _backward_once(st, rt, cu(gv))
        assert_costvol_close(gref.cpu().numpy(), o_ref, f"stage {stage} grad_ref")
        assert_costvol_close(gsrc.cpu().numpy(), o_src, f"stage {stage} grad_src")
        gref2, gsrc2 = _backward_once(st, rt, cu(gv))
        assert torch.equal(gref, gref2) and torch.equal(gsrc, gsrc2)        # bit-reproducible: no float atomics


def test_cost_volume_backward_minification_fallback():
    """Source camera zoomed out 6x: many reference footprints share one source pixel (> 4 per cell), which
    takes the exhaustive ordered path of the scatter kernel."""
    st = synthetic.make_stage(2, batch=1, n_views=3, height=64, width=96, seed=12)
    rt = geometry.stage_rot_trans(st.proj_matrix).clone()
    rt[:, :, 0:6] /= 6.0            # rows 0-1 of rot
    rt[:, :, 9:11] /= 6.0           # rows 0-1 of trans
    n, (b, d, h, w) = 2, st.depth_values.shape
    gv = torch.randn(n, b, d, h, w, generator=torch.Generator().manual_seed(4))
    o_ref, o_src = oracle.costvol_bwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values, gv)
    gref, gsrc = _backward_once(st, rt, cu(gv))
    assert_costvol_close(gref.cpu().numpy(), o_ref, "minified grad_ref")
    assert_costvol_close(gsrc.cpu().numpy(), o_src, "minified grad_src")
    _, gsrc2 = _backward_once(st, rt, cu(gv))
    assert torch.equal(gsrc, gsrc2)


def test_grad_src_cell_table_and_tile_scan_paths_agree():
    """grad_src has two atomic-free implementations: the global cell table (default) and the tile-scan kernels it
    falls back to when a cell overflows (the TMVS_F_BWD_SCAN flag forces them).  Same contributions, different fixed
    summation orders: they agree to fp32 re-association, and each is bit-reproducible.  Cases: cascade shapes with
    per-pixel hypotheses (parity collisions + overflow slots in use), [B,D] hypotheses, a ragged 37x53 map."""
    cases = []
    for stage, hw in ((1, (128, 160)), (2, (144, 200)), (3, (96, 136))):
        st = synthetic.make_stage(stage, batch=2, n_views=4, height=hw[0], width=hw[1], seed=21)
        cases.append((st, st.depth_values))
    st = synthetic.make_stage(2, batch=1, n_views=3, height=74, width=106, seed=22)
    cases.append((st, st.depth_values))
    st = synthetic.make_stage(1, batch=2, n_views=3, height=96, width=128, seed=23)
    cases.append((st, st.depth_values[:, :, 0, 0].contiguous()))            # [B,D] hypotheses
    for st, dv in cases:
        rt = geometry.stage_rot_trans(st.proj_matrix)
        packed = ops.pack_sources([cu(f) for f in st.features[1:]])
        b, d, h, w = st.depth_values.shape
        gv = cu(torch.randn(len(st.features) - 1, b, d, h, w, generator=torch.Generator().manual_seed(5)))
        run = lambda: ops.costvol_backward_packed(cu(st.features[0]), packed, rt, cu(dv), gv, need_ref=False)[1]
        g_cells, g_cells2 = run(), run()
        with ops.extra_flags(_lib.F_BWD_SCAN):
            g_scan = run()
        assert torch.equal(g_cells, g_cells2)
        assert float((g_cells - g_scan).abs().max()) <= 3e-6 * float(g_scan.abs().max())
        o_ref, o_src = oracle.costvol_bwd(st.features[0], torch.stack(st.features[1:], 0), rt, dv, gv.cpu())
        assert_costvol_close(g_cells.cpu().numpy(), o_src, f"stage {st.stage} grad_src (cell table)")


def test_grad_src_cell_table_multi_group_and_multi_pass():
    """Plumbing of the cell-table path: a batch that needs two launch groups (rot/trans travel as kernel parameters,
    64 (view, batch) slots per launch) and a table workspace capped so that the pairs of a group go through the tables
    in several passes (TMVS_F_TABLE_MB) -- same result as the uncapped run, bit for bit, and equal to the oracle."""
    st = synthetic.make_stage(2, batch=23, n_views=4, height=32, width=48, seed=17)       # 3 x 23 = 69 pairs > 64
    rt = geometry.stage_rot_trans(st.proj_matrix)
    n, (b, d, h, w) = 3, st.depth_values.shape
    gv = torch.randn(n, b, d, h, w, generator=torch.Generator().manual_seed(6))
    o_ref, o_src = oracle.costvol_bwd(st.features[0], torch.stack(st.features[1:], 0), rt, st.depth_values, gv)
    gref, gsrc = _backward_once(st, rt, cu(gv))
    assert_costvol_close(gref.cpu().numpy(), o_ref, "two launch groups grad_ref")
    assert_costvol_close(gsrc.cpu().numpy(), o_src, "two launch groups grad_src")
    with ops.extra_flags(_lib.f_table_mb(1)):                    # ~1.3 MB per pair at this size: one pair per pass
        _, gsrc_capped = _backward_once(st, rt, cu(gv))
    assert torch.equal(gsrc, gsrc_capped)


def test_grad_src_local_overflow_falls_back_per_tile():
    """A steep ramp in the hypotheses of one image region makes dozens of reference pixels land on the same source
    pixel there (more than a cell's 4 + 2 slots): the footprints that find no slot flag only the 32x8 source tiles they
    touch, the tile-scan kernels redo exactly those tiles and the cell tables serve the rest.  Result: equal to the
    oracle, to the forced tile scan (fp32 re-association) and bit-reproducible."""
    st = synthetic.make_stage(2, batch=1, n_views=4, height=192, width=256, seed=19)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    dv = st.depth_values.clone()
    d, h, w = dv.shape[1:]
    # hypotheses for a 40x40 patch chosen so that, for source view 0, every pixel of a patch row projects onto the SAME
    # source column: x_src = (R0.u z + t0) / (R2.u z + t2) = X  =>  z = (X t2 - t0) / (R0.u - X R2.u)
    R = rt[0, 0, :9].double().reshape(3, 3).numpy()
    t = rt[0, 0, 9:].double().numpy()
    ys, xs = np.meshgrid(np.arange(30, 70), np.arange(40, 80), indexing="ij")
    u = np.stack([xs, ys, np.ones_like(xs)], -1).astype(np.float64)             # [40,40,3]
    centre = np.array([60.0, 50.0, 1.0])
    X = ((R[0] @ centre) * 680.0 + t[0]) / ((R[2] @ centre) * 680.0 + t[2])
    z = (X * t[2] - t[0]) / (u @ R[0] - X * (u @ R[2]))
    z = np.clip(z, 440.0, 920.0)
    dv[:, :, 30:70, 40:80] = torch.from_numpy(z).float()[None, None] + torch.linspace(-2, 2, d)[None, :, None, None]
    gv = torch.randn(3, 1, d, h, w, generator=torch.Generator().manual_seed(8))
    o_ref, o_src = oracle.costvol_bwd(st.features[0], torch.stack(st.features[1:], 0), rt, dv, gv)
    packed = ops.pack_sources([cu(f) for f in st.features[1:]])
    run = lambda: ops.costvol_backward_packed(cu(st.features[0]), packed, rt, cu(dv), cu(gv), need_ref=False)[1]
    a, b2 = run(), run()
    with ops.extra_flags(_lib.F_BWD_SCAN):
        scan = run()
    assert torch.equal(a, b2)
    assert_costvol_close(a.cpu().numpy(), o_src, "locally minified grad_src")
    assert float((a - scan).abs().max()) <= 3e-6 * float(scan.abs().max())
    # the two paths really were mixed: some tiles bit-equal to the scan result (redone by it), others not
    tiles_equal = (a == scan)[0, 0].reshape(-1, h // 8, 8, w // 32, 32).all(0).all(1).all(2)     # view 0, [tiles_y, tiles_x]
    assert bool(tiles_equal.any()) and not bool(tiles_equal.all())


def test_backward_adjoint_identity_full_size():
    """Size-independent property at the BlendedMVS training size (config 4, one stage-2 item):
    <G, J(src)> == <J^T(G), src> for the linear map src -> per-view similarity."""
    st = synthetic.make_stage(2, batch=1, n_views=3, height=576, width=768, kind="unit", seed=5)
    rt = geometry.stage_rot_trans(st.proj_matrix)
    ref, srcs, dv = cu(st.features[0]), [cu(f) for f in st.features[1:]], cu(st.depth_values)
    packed = ops.pack_sources(srcs)
    _, views = ops.cost_volume_packed(ref, packed, rt, dv, None, True, False)
    gv = torch.randn_like(views)
    _, gsrc = ops.costvol_backward_packed(ref, packed, rt, dv, gv, need_ref=False)
    lhs = float((gv.double() * views.double()).sum())
    rhs = float((gsrc.double() * torch.stack(srcs, 0).double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)


def test_stage1_training_graph_learned_weights():
    """Stage 1 with PixelwiseNet in the graph: gradients reach ref, src and the PixelwiseNet parameters."""
    g = golden("depthnet_s1_learned")
    net = tm.DepthNet().to(DEV).train()
    feats = [cu(f).requires_grad_(True) for f in g["features"]]
    out, vw = net(feats, cu(g["proj_matrix"]), cu(g["depth_values"]), 48, _Gain(5.0), view_weights=None)
    out["prob_volume"].square().sum().backward()
    assert all(f.grad is not None and bool(torch.isfinite(f.grad).all()) and float(f.grad.abs().sum()) > 0 for f in feats)
    assert float(net.pixel_wise_net.conv0.conv.weight.grad.abs().sum()) > 0
    assert not vw.requires_grad


# ----------------------------------------------------------------------------- read-out
@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_softmax_wta_golden(name):
    g = golden(name)
    logits = cu(g["similarity"][:, 0] * g["gain"])
    prob, idx, dep, conf = tm.softmax_wta(logits, cu(g["depth_values"]))
    assert np.abs(prob.cpu().numpy() - g["prob_volume"]).max() <= PROB_ABS
    assert np.abs(conf.cpu().numpy() - g["photo_confidence"]).max() <= PROB_ABS
    rng = float(g["depth_values"].max() - g["depth_values"].min())
    idx_c, gi = idx.cpu().numpy(), g["index"]
    mism = idx_c != gi
    if mism.any():
        # allowed only where the reference's two candidates are within a few ulp
        pv = g["prob_volume"]
        b, y, x = np.nonzero(mism)
        assert np.all(np.abs(pv[b, idx_c[b, y, x], y, x] - pv[b, gi[b, y, x], y, x]) <= PROB_ABS)
    assert mism.mean() <= 1e-4
    assert np.abs(dep.cpu().numpy() - g["depth"])[~mism].max() <= DEPTH_FRAC * rng
    # not materialising prob gives the same maps
    none, idx2, dep2, conf2 = tm.softmax_wta(logits, cu(g["depth_values"]), want_prob=False)
    assert none is None and torch.equal(idx2, idx) and torch.equal(dep2, dep) and torch.equal(conf2, conf)


@pytest.mark.parametrize("name", DEPTHNET_GIVEN + ["wta_ties"])
def test_depth_wta_bit_exact(name):
    g = golden(name)
    p = g["prob_volume"] if "prob_volume" in g else g["p"]
    idx, dep = ops.depth_wta_index(cu(p), cu(g["depth_values"]))
    assert np.array_equal(idx.cpu().numpy(), g["index"])
    assert np.array_equal(dep.cpu().numpy(), g["depth"])
    assert np.array_equal(tm.depth_wta(cu(p), cu(g["depth_values"])).cpu().numpy(), g["depth"])


def test_softmax_wta_generic_depths_vs_oracle():
    gen = torch.Generator().manual_seed(5)
    for d in (1, 5, 16, 33, 48, 64, 96, 192):
        logits = 3 * torch.randn(2, d, 9, 13, generator=gen)
        dv = 425 + 2.5 * torch.arange(d, dtype=torch.float32)[None, :, None, None] + torch.rand(2, d, 9, 13, generator=gen)
        prob, idx, dep, conf = tm.softmax_wta(cu(logits), cu(dv))
        o_prob, o_idx, o_dep, o_conf = oracle.softmax_wta(logits, dv)
        assert np.abs(prob.cpu().numpy() - o_prob).max() <= PROB_ABS
        assert np.array_equal(idx.cpu().numpy(), o_idx), d
        assert np.array_equal(dep.cpu().numpy(), o_dep)
        assert torch.allclose(prob.sum(1), torch.ones(2, 9, 13, device=DEV), atol=1e-5)


def test_readout_full_size_properties():
    """Config-2 stage-2 size: probabilities sum to 1, conf is their max, depth is a hypothesis, index agrees
    with torch.argmax on the kernel's own probabilities (integer output: bit-exact)."""
    st = synthetic.make_stage(2, batch=1, n_views=2, height=1152, width=1600, seed=6)
    logits, dv = cu(st.logits), cu(st.depth_values)
    prob, idx, dep, conf = tm.softmax_wta(logits, dv)
    assert torch.allclose(prob.sum(1), torch.ones_like(conf), atol=1e-5)
    assert torch.equal(conf, prob.max(1)[0])
    assert torch.equal(idx, torch.argmax(prob, 1))
    assert torch.equal(dep, torch.gather(dv, 1, idx[:, None]).squeeze(1))


def test_depth_regression_fwd_bwd():
    g = golden("regression_unpinned")
    rng = float(g["depth_values_4d"].max() - g["depth_values_4d"].min())
    for key_dv, key_out in (("depth_values_4d", "depth_4d"), ("depth_values_2d", "depth_2d")):
        p = cu(g["p"]).requires_grad_(True)
        dv = cu(g[key_dv])
        out = tm.depth_regression(p, dv)
        assert np.abs(out.detach().cpu().numpy() - g[key_out]).max() <= DEPTH_FRAC * rng
        go = torch.randn_like(out)
        out.backward(go)
        dv4 = dv if dv.dim() == 4 else dv[:, :, None, None]
        assert torch.allclose(p.grad, go[:, None] * dv4.expand_as(p), rtol=1e-6, atol=0)


# ----------------------------------------------------------------------------- DepthNet drop-in
class _Gain(torch.nn.Module):
    def __init__(self, gain):
        super().__init__()
        self.gain = gain

    def forward(self, x):
        return x * self.gain


@pytest.mark.parametrize("name", DEPTHNET_GIVEN)
def test_depthnet_forward_given_weights(name):
    g = golden(name)
    net = tm.DepthNet().to(DEV).eval()
    feats = [cu(f) for f in g["features"]]
    with torch.no_grad():
        out = net(feats, cu(g["proj_matrix"]), cu(g["depth_values"]), g["depth_values"].shape[1],
                  _Gain(float(g["gain"])), view_weights=cu(g["view_weights"]))
    assert np.abs(out["prob_volume"].cpu().numpy() - g["prob_volume"]).max() <= 5e-5   # gain 25 amplifies 1e-6 of sim
    rng = float(g["depth_values"].max() - g["depth_values"].min())
    idx = torch.argmax(out["prob_volume"], 1).cpu().numpy()
    same = idx == g["index"]
    assert same.mean() >= 0.999
    assert np.abs(out["depth"].cpu().numpy() - g["depth"])[same].max() <= DEPTH_FRAC * rng
    assert np.abs(out["photo_confidence"].cpu().numpy() - g["photo_confidence"]).max() <= 5e-5


def test_host_pipeline_end_to_end():
    """The end-to-end form bench.py times: pinned host inputs -> H2D -> N1 hypotheses -> pack / cost volume /
    read-out -> D2H, against the oracle fed with the torch-port hypotheses."""
    from oracle import torch_port
    host = [pipeline.pin_stage(s) for s in synthetic.make_cascade(batch=1, n_views=3, height=96, width=160, seed=17)]
    pipe = pipeline.HostPipeline(DEV)
    for _ in range(2):                                   # second call reuses the persistent buffers / events
        res = pipe.process_view(host)
        torch.cuda.synchronize()
    for st, r in zip(host, res):
        scale = pipeline.synthetic_scale(st)
        hyp = torch_port.depth_hypotheses(st.cur_depth, st.num_depth, st.interval_pixel, st.image_hw, scale)
        _, o_idx, o_dep, o_conf = oracle.softmax_wta(st.logits, hyp, want_prob=False)
        rng = float(hyp.max() - hyp.min())
        assert np.abs(r["depth"].numpy() - o_dep).max() <= DEPTH_FRAC * rng
        assert np.abs(r["photo_confidence"].numpy() - o_conf).max() <= PROB_ABS


def test_finalize_maps_wire_format():
    """confidence product + cv2-style resize + masking + 8-bit depth, against the reference's own
    numpy/cv2 statements (golden)."""
    g = golden("finalize")
    d, c, a = tm.finalize_maps(cu(g["depth"]), cu(g["conf3"]), cu(g["conf1"]), cu(g["conf2"]))
    c_ref, d_ref, a_ref = g["conf_out"], g["depth_out"], g["alpha_out"]
    assert np.abs(c.cpu().numpy() - c_ref).max() <= 1e-6
    near_thr = np.abs(c_ref - 0.01) <= 1e-6                                 # a last-bit difference may flip the mask
    assert np.array_equal(d.cpu().numpy()[~near_thr], d_ref[~near_thr])
    assert a.dtype == torch.uint8 and np.array_equal(a.cpu().numpy()[~near_thr], a_ref[~near_thr])   # bytes: bit-exact
    assert near_thr.mean() < 1e-3


def test_depth_hypotheses_kernel():
    """stage hypotheses in one kernel against the reference's interpolate -> get_depth_samples ->
    interpolate chain (golden), and against the torch port at the DTU image size (stage 2)."""
    from oracle import torch_port
    g = golden("hypotheses")
    hw = tuple(int(v) for v in g["image_hw"])
    iv = float(g["depth_interval"])
    rng = float(g["depth_values"].max() - g["depth_values"].min())
    for stage, (nd, ratio, scale) in enumerate(((48, 4.0, 4), (32, 1.0, 2), (8, 0.5, 1)), start=1):
        cur = cu(g["depth_values"] if stage == 1 else g[f"prev{stage}"])
        out = tm.depth_hypotheses(cur, nd, ratio * iv, hw, scale)
        assert tuple(out.shape) == g[f"hyp{stage}"].shape
        assert np.abs(out.cpu().numpy() - g[f"hyp{stage}"]).max() <= 1e-6 * rng       # a few ulp of ~900 mm
    prev = 500 + 300 * torch.rand(1, 288, 400, generator=torch.Generator().manual_seed(2))
    want = torch_port.depth_hypotheses(prev, 32, 2.4869792, (1152, 1600), 2)
    got = tm.depth_hypotheses(cu(prev), 32, 2.4869792, (1152, 1600), 2)
    assert np.abs(got.cpu().numpy() - want.numpy()).max() <= 1e-6 * rng


@pytest.mark.parametrize("d,hw", [(8, (64, 96)), (32, (40, 64)), (48, (24, 32)), (24, (16, 32))])
def test_readout_nan_inf_and_ties_follow_torch(d, hw):
    """The read-out on rows that break the happy path -- a NaN logit, +inf, -inf, all-equal logits (ties), a huge
    spread -- against the reference's own statements run by torch on the same GPU (models/TransMVSNet.py:99-103,
    models/module.py:474-482): probabilities NaN where torch's are, first maximal index, gathered depth, confidence.
    D = 8 / 32 / 48 take the lean kernels, D = 24 the general one."""
    torch.manual_seed(d)
    h, w = hw
    logits = 3 * torch.randn(2, d, h, w, device=DEV)
    logits[0, 3, 0, :8] = float("nan")
    logits[0, d - 1, 1, :8] = float("inf")
    logits[0, 0, 2, :8] = float("-inf")
    logits[1, :, 3, :8] = 1.25                                   # exact ties: the first plane wins
    logits[1, :, 4, :8] *= 40.0                                  # exp underflow for most planes
    logits[1, 1, 5, :8] = logits[1, 6, 5, :8]                    # a two-way tie somewhere in the middle
    dv = 425 + 500 * torch.rand(2, d, h, w, device=DEV)
    prob, idx, dep, conf = tm.softmax_wta(logits, dv)
    want_p = torch.exp(torch.log_softmax(logits, 1))
    nan_rows = torch.isnan(want_p).any(1)
    assert torch.equal(torch.isnan(prob), torch.isnan(want_p))
    assert float((prob - want_p)[~torch.isnan(want_p)].abs().max()) <= PROB_ABS
    assert torch.equal(idx, torch.argmax(prob, 1))               # integer output: bit-exact on the kernel's own prob
    assert torch.equal(idx[nan_rows], torch.zeros_like(idx[nan_rows]))      # torch: first element of an all-NaN row
    assert torch.equal(dep, torch.gather(dv, 1, idx[:, None])[:, 0])
    assert torch.equal(torch.isnan(conf), nan_rows) and torch.equal(conf[~nan_rows], prob.max(1)[0][~nan_rows])
    assert torch.equal(idx[1, 3, :8], torch.zeros_like(idx[1, 3, :8]))
    none, idx2, dep2, conf2 = tm.softmax_wta(logits, dv, want_prob=False)
    assert none is None and torch.equal(idx2, idx) and torch.equal(dep2, dep)
    assert torch.equal(torch.isnan(conf2), torch.isnan(conf)) and torch.equal(conf2[~nan_rows], conf[~nan_rows])


def test_readout_writes_maps_into_caller_buffers_and_peer_sink_slots():
    """softmax_wta(out_depth=, out_conf=) and run_cascade(out_maps=): the read-out kernel stores its maps where the
    caller says (on a multi-GPU box: a peer-mapped slot of sharding.PeerMapSink, checked by scripts/check_peer_sink.py;
    here the single-process form of the sink, whose slots are local memory)."""
    from transmvsnet_b200 import sharding
    stages = synthetic.make_cascade(batch=1, n_views=3, height=128, width=192, seed=9)
    dev_stages = [pipeline.stage_to_device(s, DEV) for s in stages]
    want = pipeline.run_cascade(dev_stages)[-1]
    h, w = stages[-1].depth_values.shape[2:]
    sink = sharding.PeerMapSink(3, (h, w), DEV)
    got = pipeline.run_cascade(dev_stages, out_maps=sink.slot(1))[-1]
    assert torch.equal(sink.result()[1, 0], want["depth"][0]) and torch.equal(sink.result()[1, 1], want["photo_confidence"][0])
    assert got["depth"].data_ptr() == sink.slot(1)[:, 0].data_ptr()             # no staging copy
    assert float(sink.result()[0].abs().sum()) == 0.0 and float(sink.result()[2].abs().sum()) == 0.0
    with pytest.raises(RuntimeError):
        tm.softmax_wta(dev_stages[-1]["logits"], dev_stages[-1]["depth_values"], out_depth=torch.empty(1, h, w + 1, device=DEV))
    with pytest.raises(RuntimeError):
        tm.softmax_wta(dev_stages[-1]["logits"], dev_stages[-1]["depth_values"], out_conf=torch.empty(1, h, w))


def test_pixelwise_aggregate_folded_kernel():
    """eval-mode PixelwiseNet folded into the aggregation kernel, against the reference's
    view weights / aggregated similarity (golden) and the unfused C oracle."""
    g = golden("depthnet_s1_learned")
    pwn = tm.PixelwiseNet().eval()
    pwn.load_state_dict({k[4:]: torch.tensor(v) for k, v in g.items() if k.startswith("pwn.")})
    feats = [cu(f) for f in g["features"]]
    _, views = tm.cost_volume(feats[0], feats[1:], g["rot_trans"], cu(g["depth_values"]), None, want_views=True)
    vw, agg = tm.pixelwise_aggregate(views, tm.fold_pixelwise_net(pwn))
    assert np.abs(vw.cpu().numpy() - g["view_weights"]).max() <= 1e-5
    assert_costvol_close(agg.cpu().numpy(), g["similarity"][:, 0], "folded PixelwiseNet aggregate")
    o_vw = oracle.pixelwise_weights(views.cpu().numpy(), {k[4:]: v for k, v in g.items() if k.startswith("pwn.")})
    assert np.abs(vw.cpu().numpy() - o_vw).max() <= 1e-5
    # non-trivial batch-norm statistics (a trained net has them): fold vs the PyTorch module itself
    torch.manual_seed(3)
    net = tm.PixelwiseNet()
    for bn in (net.conv0.bn, net.conv1.bn):
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.data.normal_(1, 0.2)
        bn.bias.data.normal_(0, 0.2)
    net = net.eval()                  # the module on the CPU: cuDNN convolutions default to TF32 on the GPU
    with torch.no_grad():
        vc = views.cpu()
        want = torch.cat([net(vc[i].unsqueeze(1)) for i in range(vc.shape[0])], 1)
    got, _ = tm.pixelwise_aggregate(views, tm.fold_pixelwise_net(net))
    assert float((got.cpu() - want).abs().max()) <= 1e-5


def test_pixelwise_weight_table_matches_the_mlp_in_double():
    """The kernel evaluates eval-mode PixelwiseNet (a ReLU network of ONE scalar) as a piecewise-linear table built on
    the host.  Random folded parameters -- including dead first-layer units (w = 0), duplicated hinges and large
    magnitudes -- and similarities spanning far beyond every breakpoint, against the MLP evaluated in float64."""
    rng = np.random.default_rng(11)
    for trial in range(6):
        mlp = rng.normal(0, 1.0 + trial, 177).astype(np.float32)
        if trial >= 2:
            mlp[rng.integers(0, 16, 3)] = 0.0                       # dead first-layer units
            mlp[5], mlp[16 + 5] = mlp[4], mlp[16 + 4]               # two identical hinges
        if trial == 5:
            mlp[:16] *= 1e-3                                        # hinges far away from the data
        n, b, d, h, w = 2, 1, 24, 16, 64
        sims = (rng.normal(0, 1, (n, b, d, h, w)) * rng.choice([0.01, 1.0, 50.0], (n, b, 1, h, w))).astype(np.float32)
        vw, agg = tm.pixelwise_aggregate(cu(torch.from_numpy(sims)), torch.from_numpy(mlp))
        x = sims.astype(np.float64)[..., None]
        w0, b0 = mlp[:16].astype(np.float64), mlp[16:32].astype(np.float64)
        w1, b1 = mlp[32:160].astype(np.float64).reshape(8, 16), mlp[160:168].astype(np.float64)
        w2, b2 = mlp[168:176].astype(np.float64), float(mlp[176])
        h0 = np.maximum(x * w0 + b0, 0.0)
        h1 = np.maximum(h0 @ w1.T + b1, 0.0)
        logit = h1 @ w2 + b2                                        # [n,b,d,h,w]
        want = 1.0 / (1.0 + np.exp(-logit.max(axis=2)))             # [n,b,h,w]
        got = vw.cpu().numpy().transpose(1, 0, 2, 3)                # [b,n,h,w] -> [n,b,h,w]
        scale = max(1.0, float(np.abs(logit).max()))
        # sigmoid is 1/4-Lipschitz: a logit error of eps moves the weight by <= eps/4; fp32 logits carry ~1e-7 * scale
        assert np.abs(got - want).max() <= 2e-6 * scale, (trial, float(np.abs(got - want).max()))
        wsum = 1e-5 + got.sum(0)
        ref_agg = (sims.astype(np.float64) * got[:, :, None]).sum(0) / wsum[:, None]
        assert np.abs(agg.cpu().numpy() - ref_agg).max() <= 1e-5 * max(1.0, float(np.abs(ref_agg).max()))


def test_depthnet_forward_learned_weights():
    g = golden("depthnet_s1_learned")
    net = tm.DepthNet().eval()
    net.pixel_wise_net.load_state_dict({k[4:]: torch.tensor(v) for k, v in g.items() if k.startswith("pwn.")})
    net = net.to(DEV)
    feats = [cu(f) for f in g["features"]]
    with torch.no_grad():
        out, vw = net(feats, cu(g["proj_matrix"]), cu(g["depth_values"]), 48, _Gain(float(g["gain"])), view_weights=None)
    assert np.abs(vw.cpu().numpy() - g["view_weights"]).max() <= 1e-5
    assert np.abs(out["prob_volume"].cpu().numpy() - g["prob_volume"]).max() <= 5e-5
    same = torch.argmax(out["prob_volume"], 1).cpu().numpy() == g["index"]
    assert same.mean() >= 0.999
