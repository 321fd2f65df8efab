"""The reference's only golden answers for the hot path: okvis_matcher/test/testMatcher.cpp:69-155,
ported verbatim as known-answer tests of the CPU oracle (sequential DenseMatcher semantics)."""
import numpy as np


def _dist_matrix(listA, listB):
    return np.abs(np.asarray(listA, np.float64)[:, None] - np.asarray(listB, np.float64)[None, :]).astype(np.float32)


def test_dense_matcher_known_answer(oracle):
    # testMatcher.cpp:69-110: threshold 4.0, skipA(0), expect {1->2, 2->1, 3->3}
    listA = [1.0, 3.0, 2.0, 0.9]
    listB = [18.0, 2.1, 4.0, 1.0]
    res = oracle.match_matrix(_dist_matrix(listA, listB), skipA=[1, 0, 0, 0], threshold=4.0, num_best=4)
    got = {(int(a), int(b)) for a, b in res["matches"]}
    assert got == {(1, 2), (2, 1), (3, 3)}


def test_dense_matcher_distance_ratio_known_answer(oracle):
    # testMatcher.cpp:112-155: DenseMatcher(4, 4, true), ratio threshold 3, expect {1->3, 3->1}
    listA = [8.0, 1.0, 3.0, 2.0, 0.9]
    listB = [18.0, 2.1, 4.0, 1.0, 7.0]
    res = oracle.match_matrix(_dist_matrix(listA, listB), skipA=[1, 0, 0, 0, 0], threshold=4.0, num_best=4,
                              use_ratio=True, ratio_threshold=3.0)
    got = {(int(a), int(b)) for a, b in res["matches"]}
    assert got == {(1, 3), (3, 1)}


def test_tie_rules(oracle):
    """DenseMatcher.hpp(impl):153-179: lower_bound insertion puts a new entry BEFORE equal distances;
    an entry equal to the current worst is rejected; assignbest only displaces on strictly smaller."""
    D = np.array([[5, 5, 5, 5, 5, 5]], np.float32)
    res = oracle.match_matrix(D, threshold=10.0, num_best=4)
    # B=0..3 fill the list (each inserted before its equals), B=4,5 are rejected (not < worst)
    assert [int(x) for x in res["topk"][0]["index_a"]] == [3, 2, 1, 0]
    # two A with identical distance to the same B: the first keeps it, the second takes its next choice
    D = np.array([[1, 3], [1, 2]], np.float32)
    res = oracle.match_matrix(D, threshold=10.0, num_best=4)
    assert [int(x) for x in res["pairs"]["index_a"]] == [0, 1]


def test_hamming_matches_bruteforce(oracle):
    rng = np.random.default_rng(0)
    A = rng.integers(0, 256, (50, 48), dtype=np.uint8)
    B = A[rng.permutation(50)].copy()
    flip = rng.integers(0, 48, 50)
    B[np.arange(50), flip] ^= 0x11
    res = oracle.match_hamming(A, B, threshold=60.0)
    D = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(2)
    for a, b in res["matches"]:
        assert D[a, b] == D[a].min() == 2
    assert len(res["matches"]) == 50
    rp, col, dist = oracle.hamming_candidates(A, B, threshold=60.0)
    assert rp[-1] == (D < 60).sum()
    for a in range(50):
        assert list(col[rp[a]:rp[a + 1]]) == list(np.nonzero(D[a] < 60)[0])
