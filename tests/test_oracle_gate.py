"""CPU: the oracle's restatement of the geometric match gate (oracle/oracle_gate.hpp) behaves like the reference's
verifyMatch on a seeded stereo scene: true correspondences pass, geometrically inconsistent look-alikes are rejected,
and the gated DenseMatcher recovers (almost) only true matches."""
import numpy as np

from gate_scene import gates, make_scene


def test_gate_keeps_true_matches_and_rejects_distractors(oracle):
    sc = make_scene(0)
    g3, g2 = gates(sc)
    plain = oracle.match_hamming(sc["A"], sc["B"])
    wrong_plain = sum(1 for b, p in enumerate(plain["pairs"]) if p["index_a"] >= 0 and p["distance"] < 60 and sc["truth_b"][b] != p["index_a"])
    assert wrong_plain > 20                     # without the gate the look-alikes steal matches
    for g, min_true in ((g3, 0.85), (g2, 0.85)):
        r = oracle.match_hamming_gated(sc["A"], sc["B"], g)
        good = wrong = 0
        for b, p in enumerate(r["pairs"]):
            if p["index_a"] >= 0 and p["distance"] < 60:
                if sc["truth_b"][b] == p["index_a"]:
                    good += 1
                else:
                    wrong += 1
        assert good >= min_true * len(sc["A"]), (good, wrong)
        assert wrong <= 0.03 * len(sc["A"]), (good, wrong)
