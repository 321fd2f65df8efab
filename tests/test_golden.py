"""Independent pin of the oracle: tests/golden/numpy_costs.npz holds cost values computed by a plain NumPy restatement
of the residual functors written from the reference's source (tools/make_golden_numpy.py: reprojection + Cauchy loss,
ImuError with a fresh preintegration, PoseError, SpeedAndBiasError -- no line of it goes through liboracle.so or the
CUDA library).  The C++ oracle must reproduce them at every fixture state; tests/test_gpu_golden.py asks the same of
the device kernels."""
import dataclasses
import os

import numpy as np
import pytest

from okvis_b200 import abi, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "numpy_costs.npz")
WINDOWS = {"cfg1_w0": (1, 0), "cfg1_w3": (1, 3)}


def fixture_states():
    g = np.load(GOLDEN)
    for tag, (cfg, idx) in WINDOWS.items():
        w = synthetic.make_window(cfg, idx)
        for i in range(4):
            st = dataclasses.replace(w, poses=np.ascontiguousarray(g["%s_s%d_poses" % (tag, i)]),
                                     speed_bias=np.ascontiguousarray(g["%s_s%d_speed_bias" % (tag, i)]),
                                     landmarks=np.ascontiguousarray(g["%s_s%d_landmarks" % (tag, i)]))
            yield "%s_s%d" % (tag, i), st, g["%s_s%d_cost" % (tag, i)]


def reprojection_only(w):
    return dataclasses.replace(w, imu_terms=np.zeros(0, abi.imu_term_dtype), pose_priors=np.zeros(0, abi.pose_prior_dtype),
                               sb_priors=np.zeros(0, abi.sb_prior_dtype))


def test_fixture_is_present_and_described():
    g = np.load(GOLDEN)
    assert "tools/make_golden_numpy.py" in str(g["layout"])
    assert len([k for k in g.files if k.endswith("_cost")]) == 8


@pytest.mark.parametrize("name,state,cost", list(fixture_states()), ids=lambda x: x if isinstance(x, str) else "")
def test_oracle_cost_equals_the_numpy_cost(oracle, name, state, cost):
    total, reproj = float(cost[0]), float(cost[1])
    assert abs(oracle.OracleProblem(state).cost() - total) <= 1e-12 * total
    assert abs(oracle.OracleProblem(reprojection_only(state)).cost() - reproj) <= 1e-12 * reproj


def test_generator_reproduces_the_fixture():
    """The committed vectors are what the committed script produces (first state of the first window; fast)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_golden_numpy as G
    nw = G.NumpyWindow(synthetic.make_window(1, 0))
    c = G.family_costs(nw, nw.x0())
    g = np.load(GOLDEN)["cfg1_w0_s0_cost"]
    assert abs(c["total"] - g[0]) <= 1e-13 * g[0] and abs(c["reprojection"] - g[1]) <= 1e-13 * g[1]
