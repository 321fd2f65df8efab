"""The device kernels against the independent NumPy cost values (tests/golden/numpy_costs.npz, see test_golden.py):
okb_optimize with max_iterations = 0 linearises once and reports the cost at the uploaded state -- k_linearize
(reprojection + Cauchy), k_imu (preintegration + ImuError), the priors in k_solve."""
import numpy as np
import pytest

from test_golden import fixture_states, reprojection_only

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 1)
    yield c
    c.close()


@pytest.mark.parametrize("name,state,cost", list(fixture_states()), ids=lambda x: x if isinstance(x, str) else "")
def test_device_cost_equals_the_numpy_cost(ctx, name, state, cost):
    total, reproj = float(cost[0]), float(cost[1])
    ctx.upload(0, state)
    s = ctx.optimize(0, 1, max_iterations=0)[0]
    assert abs(s["initial_cost"] - total) <= 1e-11 * total
    ctx.upload(0, reprojection_only(state))
    s = ctx.optimize(0, 1, max_iterations=0)[0]
    assert abs(s["initial_cost"] - reproj) <= 1e-11 * reproj
