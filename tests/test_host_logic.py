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


def test_scan_pairs_every_view_is_a_source_n_minus_1_times():
    """The synthetic pair.txt has the property of DTU's that the feature cache relies on (datasets/general_eval.py:
    25-57): V jobs, each view the reference view once and a source view of N-1 others."""
    pairs = synthetic.scan_pairs(49, 5)
    assert [r for r, _ in pairs] == list(range(49))
    use = np.zeros(49, int)
    for r, srcs in pairs:
        assert len(srcs) == 4 and r not in srcs and len(set(srcs)) == 4
        for v in srcs:
            use[v] += 1
    assert (use == 4).all()
    # a scan smaller than N views repeats its first source (general_eval.py:47-49 "fill to nviews")
    r, srcs = synthetic.scan_pairs(3, 5)[0]
    assert len(srcs) == 4 and set(srcs) <= {1, 2}


def test_scan_jobs_alias_the_view_pyramids():
    scan = synthetic.make_scan(6, n_views=5, height=32, width=48, seed=1)
    assert len(scan.jobs) == 6 and scan.voxel_views == 6 * sum(s.voxel_views for s in scan.jobs[0])
    ref, srcs = scan.pairs[2]
    for s in range(3):
        job = scan.jobs[2][s]
        assert job.features[0] is scan.pyramids[ref][s]
        assert all(job.features[1 + k] is scan.pyramids[v][s] for k, v in enumerate(srcs))
        assert job.depth_values is None and job.bdhw[1] == synthetic.STAGES[s][1]      # lean: generated on the device
    assert scan.jobs[2][0].view_weights is not None and scan.jobs[2][1].view_weights is None
    assert not torch.equal(scan.pyramids[0][0], scan.pyramids[1][0])


def test_reference_arithmetic_is_context_local_not_process_wide():
    """The arithmetic choice is a per-call flag of the C ABI; the Python default is scoped with contextvars, so a thread
    that asks for the CPU arithmetic does not change what another thread's calls get."""
    import threading
    from transmvsnet_b200 import ops
    seen = {}

    def worker():
        seen["other thread"] = ops._flags()
    with ops.reference_arithmetic("cpu"):
        t = threading.Thread(target=worker)
        t.start()
        t.join()
        seen["inside"] = ops._flags()
    seen["after"] = ops._flags()
    assert seen["inside"] & _lib.F_ARITH_ATEN_CUDA == 0
    assert seen["other thread"] & _lib.F_ARITH_ATEN_CUDA and seen["after"] & _lib.F_ARITH_ATEN_CUDA     # default: CUDA
    assert ops._flags("cpu") == 0 and ops._flags("cuda") == _lib.F_ARITH_ATEN_CUDA
    with pytest.raises(ValueError):
        ops._flags("fp16")


def test_library_reads_no_environment_and_keeps_no_arithmetic_state():
    """no getenv dispatch, no process-global mode in the library."""
    import glob
    import os
    import subprocess
    csrc = os.path.join(os.path.dirname(_lib.__file__), "csrc")
    for path in glob.glob(os.path.join(csrc, "*.cu*")):          # (the statically linked CUDA runtime has its own getenv)
        text = open(path).read()
        assert "getenv" not in text and "g_arith" not in text, path
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "tmvs_set_reference_arithmetic" not in syms and "tmvs_costvol_fwd_cached" in syms


def test_bench_reference_arm_runs_the_staged_reference_and_prints_the_contract_line():
    """bench.py --impl reference on the tiny workload (seconds on a CPU): ONE JSON line on stdout with the keys the driver
    reads, timed through the reference's own DepthNet.forward when oracle/_ref is staged (kind "reference")."""
    import json
    import os
    import subprocess
    import sys
    from conftest import REPO
    from oracle import build as oracle_build
    oracle_build.build_ref()
    res = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-500:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    staged = oracle_build.import_reference() is not None
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert line["config"]["workload"].startswith("tiny") and "timed_sample" not in line["config"]
