"""Edge cases through the C-ABI on the GPU: empty and one-element descriptor lists, everything skipped, a blank image,
keypoint caps, a window whose speed/bias blocks are NOT numbered along the IMU chain (the reduced solve then takes its
plain dense path instead of the chain elimination, okb_chol.cuh), and a window without any IMU term."""
import dataclasses

import numpy as np
import pytest

from okvis_b200 import abi, images, synthetic
from test_gpu_solver import compare
from test_oracle_frontend import CAM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 2)
    yield c
    c.close()


def test_empty_and_single_descriptor_lists(ctx, okb, oracle):
    rng = np.random.default_rng(3)
    A = rng.integers(0, 256, (5, 48), dtype=np.uint8)
    E = np.zeros((0, 48), np.uint8)
    r = ctx.hamming_match(E, A)                      # nothing to match from
    assert len(r["topk"]) == 0 and np.all(r["pairs"]["index_a"] == -1) and len(r["pairs"]) == 5
    r = ctx.hamming_match(A, E)                      # nothing to match against
    assert len(r["pairs"]) == 0 and np.all(r["topk"]["index_a"] == -1)
    rp, col, dist = ctx.hamming_candidates(A, E)
    assert list(rp) == [0] * 6 and len(col) == 0
    rp, col, dist = ctx.hamming_candidates(E, A)
    assert list(rp) == [0] and len(col) == 0
    # one against one; identical descriptors -> distance 0 < threshold
    r = ctx.hamming_match(A[:1], A[:1])
    assert r["pairs"]["index_a"][0] == 0 and r["pairs"]["distance"][0] == 0.0
    o = oracle.match_hamming(A[:1], A[:1], None, None, threshold=60.0, num_best=4, use_ratio=False)
    assert np.array_equal(r["pairs"]["index_a"], o["pairs"]["index_a"])
    # everything skipped on one side: no pair, like the reference's skipA / skipB
    r = ctx.hamming_match(A, A, skipA=np.ones(5, np.uint8))
    assert np.all(r["pairs"]["index_a"] == -1)
    r = ctx.hamming_match(A, A, skipB=np.ones(5, np.uint8))
    assert np.all(r["pairs"]["index_a"] == -1)
    # threshold 0: a distance must be strictly below it (DenseMatcher.hpp:162)
    r = ctx.hamming_match(A, A, threshold=0.0)
    assert np.all(r["pairs"]["index_a"] == -1)


def test_blank_image_and_keypoint_cap(ctx, oracle):
    blank = np.full((480, 752), 117, np.uint8)
    kp, desc = ctx.detect_describe(blank, CAM, np.eye(3))
    assert len(kp) == 0 and len(desc) == 0
    # the cap keeps the strongest keypoints, in the same order as the oracle
    img = images.textured_image(0x0B200 + 3000, n_shapes=2600)
    for maxk in (1, 7, 64):
        kp, desc = ctx.detect_describe(img, CAM, np.eye(3), uniformity_radius=15.0, max_keypoints=maxk)
        okp, odesc = oracle.detect_describe(img, CAM, np.eye(3), uniformity_radius=15.0, max_keypoints=maxk)
        assert len(kp) == maxk == len(okp)
        assert np.array_equal(kp["x"], okp["x"]) and np.array_equal(kp["y"], okp["y"]) and np.array_equal(desc, odesc)


def permuted_speed_bias(w, perm):
    """The same window with speed/bias block k stored at index perm[k] (terms and priors re-pointed)."""
    perm = np.asarray(perm)
    sb = np.empty_like(w.speed_bias)
    sb[perm] = w.speed_bias
    terms = w.imu_terms.copy()
    terms["sb0"] = perm[w.imu_terms["sb0"]]
    terms["sb1"] = perm[w.imu_terms["sb1"]]
    sp = w.sb_priors.copy()
    sp["sb_idx"] = perm[w.sb_priors["sb_idx"]]
    return dataclasses.replace(w, speed_bias=np.ascontiguousarray(sb), imu_terms=terms, sb_priors=sp)


def test_speed_bias_blocks_off_the_chain_order(ctx, oracle):
    """IMU terms that link non-neighbouring speed/bias indices: the chain elimination does not apply (its structure test
    fails) and the reduced system goes through the dense blocked Cholesky; parity with the oracle on the same window, and
    the same estimate as the chain-ordered window up to the rounding of a different elimination order."""
    w = synthetic.make_window(1, 0)
    perm = [2, 0, 4, 1, 3]
    wp = permuted_speed_bias(w, perm)
    compare(ctx, oracle, wp)
    gp = ctx.download(0)
    ctx.upload(1, w)
    ctx.optimize(1, 1, max_iterations=10)
    g = ctx.download(1)
    assert np.abs(gp["poses"] - g["poses"]).max() < 1e-6
    assert np.abs(gp["speed_bias"][perm] - g["speed_bias"]).max() < 1e-6


def test_window_without_imu_terms(ctx, oracle):
    """Vision-only window (no ImuError): the speed/bias blocks only see the first-frame prior and the trust-region
    damping; same decisions and estimates as the oracle."""
    w = synthetic.make_window(1, 1)
    sp = np.zeros(len(w.speed_bias), abi.sb_prior_dtype)      # a prior per block keeps every block observable
    for k in range(len(sp)):
        sp[k] = w.sb_priors[0]
        sp[k]["sb_idx"] = k
        sp[k]["meas"] = w.speed_bias[k]
    pp = np.zeros(2, abi.pose_prior_dtype)                    # two pose priors fix the gauge the IMU would have fixed
    pp[0] = w.pose_priors[0]
    pp[1] = w.pose_priors[0]
    pp[1]["pose_idx"] = len(w.poses) - 1
    pp[1]["meas"] = w.poses[-1]
    pp[1]["sqrt_info"] = np.diag([1e2] * 6).reshape(-1)
    wv = dataclasses.replace(w, imu_terms=w.imu_terms[:0].copy(), imu_samples=w.imu_samples[:1].copy(), sb_priors=sp, pose_priors=pp)
    compare(ctx, oracle, wv)

