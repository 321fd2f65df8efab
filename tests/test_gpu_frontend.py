"""GPU parity of the frontend path through the C-ABI: bit-exact keypoints / descriptors / match indices
against the CPU oracle (north_star: "bit-exact for match indices"), and the reference's matcher
known-answer tests (okvis_matcher/test/testMatcher.cpp:69-155) through the Hamming kernel."""
import numpy as np
import pytest

from okvis_b200 import abi, images
from test_oracle_frontend import CAM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 1)
    yield c
    c.close()


def unary(values, nbytes=48):
    """Descriptor whose Hamming distance to another unary code is |va - vb|."""
    out = np.zeros((len(values), nbytes * 8), np.uint8)
    for i, v in enumerate(values):
        out[i, :int(v)] = 1
    return np.packbits(out, axis=1, bitorder="little")


def test_matcher_known_answer_through_hamming(ctx, okb, oracle):
    # testMatcher.cpp:69-110 with all values x10 (threshold 4 -> 40): expect {1->2, 2->1, 3->3}
    A = unary([10, 30, 20, 9])
    B = unary([180, 21, 40, 10])
    res = ctx.hamming_match(A, B, skipA=[1, 0, 0, 0], threshold=40.0, num_best=4)
    got = {(a, b) for a, b, _ in okb.matches_from_pairs(res["pairs"], res["topk"], 40.0)}
    assert got == {(1, 2), (2, 1), (3, 3)}
    # testMatcher.cpp:112-155: ratio test, expect {1->3, 3->1}
    A = unary([80, 10, 30, 20, 9])
    B = unary([180, 21, 40, 10, 70])
    res = ctx.hamming_match(A, B, skipA=[1, 0, 0, 0, 0], threshold=40.0, num_best=4, use_ratio=True, ratio_threshold=3.0)
    got = {(a, b) for a, b, _ in okb.matches_from_pairs(res["pairs"], res["topk"], 40.0, True, 3.0)}
    assert got == {(1, 3), (3, 1)}


@pytest.mark.parametrize("nbytes", [48, 64])
def test_matcher_bit_exact_vs_oracle(ctx, okb, oracle, nbytes):
    rng = np.random.default_rng(41)
    for nA, nB, flips in ((1000, 1000, 20), (37, 513, 40), (400, 7, 10), (129, 128, 64)):
        base = rng.integers(0, 256, (max(nA, nB), nbytes), dtype=np.uint8)
        A = base[:nA].copy()
        B = base[rng.permutation(max(nA, nB))[:nB]].copy()
        # few bit flips => many small, tied distances (ties are the hard part of the semantics)
        for row in B:
            idx = rng.integers(0, nbytes * 8, rng.integers(0, flips))
            for i in idx:
                row[i >> 3] ^= (1 << (i & 7))
        skipA = (rng.random(nA) < 0.05).astype(np.uint8)
        skipB = (rng.random(nB) < 0.05).astype(np.uint8)
        for use_ratio in (False, True):
            g = ctx.hamming_match(A, B, skipA, skipB, threshold=60.0, num_best=4, use_ratio=use_ratio)
            o = oracle.match_hamming(A, B, skipA, skipB, threshold=60.0, num_best=4, use_ratio=use_ratio)
            assert np.array_equal(g["topk"]["index_a"], o["topk"]["index_a"])
            assert np.array_equal(g["topk"]["distance"], o["topk"]["distance"])
            assert np.array_equal(g["pairs"]["index_a"], o["pairs"]["index_a"])
            assert np.array_equal(g["pairs"]["distance"], o["pairs"]["distance"])
            m = okb.matches_from_pairs(g["pairs"], g["topk"], 60.0, use_ratio, 3.0)
            assert [(a, b) for a, b, _ in m] == [tuple(x) for x in o["matches"].tolist()]
        rp, col, dist = ctx.hamming_candidates(A, B, threshold=60.0)
        rp0, col0, dist0 = oracle.hamming_candidates(A, B, threshold=60.0)
        assert np.array_equal(rp, rp0) and np.array_equal(col, col0) and np.array_equal(dist, dist0)


def test_candidates_capacity_error(ctx, okb):
    A = np.zeros((10, 48), np.uint8)
    with pytest.raises(okb.OkbError) as e:
        ctx.hamming_candidates(A, A, threshold=60.0, cap=5)
    assert e.value.code == abi.OKB_ERR_CAPACITY


@pytest.mark.parametrize("radius,maxk,nbytes", [(15.0, 1000, 48), (40.0, 400, 48), (15.0, 1000, 64)])
def test_detect_describe_bit_exact_vs_oracle(ctx, oracle, radius, maxk, nbytes):
    left, right = images.stereo_pair()
    R_CW = np.array([[0.9950, 0.0, -0.0998], [0.0198, 0.9801, 0.1977], [0.0978, -0.1987, 0.9752]])
    for img in (left, right):
        kg, dg = ctx.detect_describe(img, CAM, R_CW, uniformity_radius=radius, max_keypoints=maxk, desc_bytes=nbytes)
        ko, do = oracle.detect_describe(img, CAM, R_CW, uniformity_radius=radius, max_keypoints=maxk, desc_bytes=nbytes)
        assert len(kg) == len(ko) > 100
        for f in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(kg[f], ko[f]), f
        assert np.abs(kg["angle"] - ko["angle"]).max() < 1e-3
        assert np.array_equal(dg, do)


def test_frontend_to_matcher_pipeline(ctx, okb, oracle):
    """cfg-3: detect+describe both views, match, bit-exact match indices vs the oracle."""
    left, right = images.stereo_pair()
    ka, da = ctx.detect_describe(left, CAM, np.eye(3), uniformity_radius=15, max_keypoints=1000)
    kb, db = ctx.detect_describe(right, CAM, np.eye(3), uniformity_radius=15, max_keypoints=1000, cam_slot=1)
    g = ctx.hamming_match(da, db, threshold=60.0)
    o = oracle.match_hamming(da, db, threshold=60.0)
    m = okb.matches_from_pairs(g["pairs"], g["topk"], 60.0)
    assert [(a, b) for a, b, _ in m] == [tuple(x) for x in o["matches"].tolist()]
    assert len(m) > 200
