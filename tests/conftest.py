import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")

# Tolerances of north_star / SURVEY.md section 8d -- the only numbers the parity tests use.
COSTVOL_REL = 1e-4        # max|a-b| / max|b|  and  ||a-b|| / ||b||  for cost volumes / warped volumes
DEPTH_FRAC = 1e-3         # regressed depth: |a-b| <= DEPTH_FRAC * (dmax - dmin)
PROB_ABS = 2e-6           # softmax probabilities (values in [0,1]); fp32 exp/log differ by a few ulp


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(a, b):
    """(max-norm relative error, L2 relative error) of a against b, in float64."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mx = np.abs(b).max()
    l2 = np.linalg.norm(b)
    return (np.abs(a - b).max() / (mx if mx > 0 else 1.0), np.linalg.norm(a - b) / (l2 if l2 > 0 else 1.0))


def assert_costvol_close(a, b, what=""):
    e_max, e_l2 = rel_err(a, b)
    assert e_max <= COSTVOL_REL and e_l2 <= COSTVOL_REL, f"{what}: max-rel {e_max:.3e}, l2-rel {e_l2:.3e}"


WARP_CASES = ["warp_perpixel", "warp_planes", "warp_identity", "warp_integer", "warp_behind"]
DEPTHNET_GIVEN = ["depthnet_s1_given", "depthnet_s2_given", "depthnet_s3_given", "depthnet_odd_given"]
