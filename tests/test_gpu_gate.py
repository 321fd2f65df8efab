"""GPU parity of the geometry-gated matcher (okb_hamming_match_gated: VioKeyframeWindowMatchingAlgorithm::distance +
verifyMatch inside the list kernel) against the oracle: bit-exact top-k lists and per-B winners in both gate modes."""
import numpy as np
import pytest

from gate_scene import gates, make_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(okb):
    c = okb.Context(0, 1)
    yield c
    c.close()


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("use_ratio", [False, True])
def test_gated_match_bit_exact_vs_oracle(ctx, oracle, seed, use_ratio):
    sc = make_scene(seed)
    rng = np.random.default_rng(seed)
    skipA = (rng.random(len(sc["A"])) < 0.05).astype(np.uint8)
    skipB = (rng.random(len(sc["B"])) < 0.05).astype(np.uint8)
    for g in gates(sc):
        got = ctx.hamming_match_gated(sc["A"], sc["B"], g, skipA, skipB, use_ratio=use_ratio)
        ref = oracle.match_hamming_gated(sc["A"], sc["B"], g, skipA, skipB, use_ratio=use_ratio)
        assert np.array_equal(got["topk"]["index_a"], ref["topk"]["index_a"])
        assert np.array_equal(got["topk"]["distance"], ref["topk"]["distance"])
        assert np.array_equal(got["pairs"]["index_a"], ref["pairs"]["index_a"])
        assert np.array_equal(got["pairs"]["distance"], ref["pairs"]["distance"])
    plain = ctx.hamming_match(sc["A"], sc["B"], skipA, skipB, use_ratio=use_ratio)
    assert not np.array_equal(plain["pairs"]["index_a"], got["pairs"]["index_a"])     # the gate changes the result
