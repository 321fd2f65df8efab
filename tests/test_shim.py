"""The C++ host shim (include/okvis_b200_estimator.hpp, the okvis::Estimator mirror): compiles against
the C-ABI, fails loudly without a GPU, and -- on the GPU -- drives the addStates / addLandmark /
addObservation / optimize cycle of ThreadedKFVio to convergence."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "shim", "shim_test.bin")


def build_shim():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(BIN)


def test_shim_compiles_and_fails_loudly_without_gpu():
    build_shim()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    out = subprocess.run([BIN, "--no-gpu"], capture_output=True, text=True, timeout=60).stdout
    r = json.loads(out.strip().splitlines()[-1])
    assert r["created"] is False and "no CUDA device" in r["error"]


@pytest.mark.gpu
def test_shim_estimator_cycle_converges():
    build_shim()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["frames"] == 5 and r["landmarks"] > 50
    # exact measurements: what remains is the first-frame speed prior (v = 0, sigma = 1) against the true
    # 0.5 m/s: 0.5 * 0.5^2 = 0.125
    assert r["final_cost"] < 0.2, r
    assert r["pos_err"] < 2e-3            # exact measurements: the estimate returns to the true trajectory
    assert abs(r["vy"] - 0.5) < 2e-2      # velocity recovered although the first-frame prior says 0
    assert 0.0 < r["lm_quality"] <= 1.0
    assert r["dup_obs"] == 0              # duplicate observation returns NULL like the reference


@pytest.mark.gpu
def test_shim_survives_the_threaded_kfvio_cycle_with_marginalisation():
    """30 frames of addStates -> addObservation -> optimize -> applyMarginalizationStrategy(5, 3): the window holds
    numKeyframes + numImuFrames frames, the estimate stays on the true trajectory (exact measurements) although
    frames / landmarks are marginalised every step, and only a frame's worth of data is uploaded per optimize."""
    build_shim()
    out = subprocess.run([BIN, "--sliding", "30"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr + out.stdout
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert "error" not in r, r
    assert r["marg_calls"] == 30 and r["max_frames"] <= 8 and r["frames"] == 8
    assert r["imu_window_ok"] == 1
    assert r["removed_landmarks"] > 50            # landmarks leave the field of view and are marginalised / dropped
    # the first frames still carry the zero-velocity prior against 1.5 m/s; afterwards the estimate sits on the trajectory
    assert r["max_pos_err"] < 5e-2 and r["pos_err"] < 2e-3
    assert abs(r["vy"] - 1.5) < 2e-2
    assert r["steady_upload_bytes"] * 4 < r["first_upload_bytes"] or r["steady_upload_bytes"] < 100000


@pytest.mark.gpu
def test_frontend_and_dense_matcher_shims():
    """okvis_b200::Frontend (detectAndDescribe, propagation, parameter accessors) and okvis_b200::DenseMatcher
    (match<ALGORITHM> with the reference's epilogue, candidate lists) on a stereo pair with 12 px disparity."""
    build_shim()
    out = subprocess.run([BIN, "--frontend"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert min(r["keypoints"]) > 100 and r["initialized"] == 1
    assert r["matches"] > 60 and r["consistent"] >= 0.7 * r["matches"]     # the blocky texture repeats: some corners are ambiguous
    assert r["candidates"] >= r["matches"]
    # the 2D-2D gate (matchGated): what survives is geometrically consistent, and the consistent matches survive
    assert r["gated_matches"] > 60 and r["gated_consistent"] >= 0.95 * r["gated_matches"]
    assert r["gated_consistent"] >= r["consistent"] - 3
    assert r["propagation_ok"] == 1 and abs(r["propagated_x"] - 0.2 * 0.2) < 1e-6     # 0.2 m/s for 0.2 s
