#!/usr/bin/env python
"""Golden vectors for the matcher produced by the REFERENCE ITSELF: okvis_matcher compiled unmodified from /root/reference
(oracle/Makefile.ref -> oracle/_ref/libokvis_matcher_ref.so, one matcher thread = the sequential order of the project's
contract).  Writes tests/golden/matcher_reference.npz: descriptor lists, skip flags, parameters and the (A, B, distance)
matches DenseMatcher::match emitted.  The reference tree does not exist on the GPU box; the vectors travel instead.
Usage: python tools/make_golden_matcher.py [--check]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "matcher_reference.npz")
SO = os.path.join(ROOT, "oracle", "_ref", "libokvis_matcher_ref.so")

CASES = [  # nA, nB, bytes, max flipped bits, skip fraction, threshold, numBest, ratio test, ratio threshold
    (300, 280, 48, 30, 0.05, 60.0, 4, False, 3.0),
    (300, 280, 48, 30, 0.05, 60.0, 4, True, 3.0),
    (64, 400, 48, 60, 0.10, 60.0, 4, False, 3.0),
    (400, 37, 48, 12, 0.00, 60.0, 4, True, 1.5),
    (129, 128, 64, 64, 0.05, 80.0, 4, False, 3.0),
    (200, 200, 48, 8, 0.02, 60.0, 1, False, 3.0),
    (200, 200, 48, 8, 0.02, 60.0, 2, True, 3.0),
    (150, 170, 48, 20, 0.00, 60.0, 8, False, 3.0),
    (1, 1, 48, 0, 0.00, 60.0, 4, False, 3.0),
    (50, 60, 48, 400, 0.00, 60.0, 4, False, 3.0),       # almost nothing below the threshold
]


def make_case(i, spec):
    nA, nB, nbytes, flips, skipf, thr, nb, use_ratio, ratio = spec
    rng = np.random.Generator(np.random.PCG64(0x0B200 + 7000 + i))
    base = rng.integers(0, 256, (max(nA, nB), nbytes), dtype=np.uint8)
    A = base[:nA].copy()
    B = base[rng.permutation(max(nA, nB))[:nB]].copy()
    for row in B:      # few bit flips => many small, tied distances
        for j in rng.integers(0, nbytes * 8, rng.integers(0, flips + 1)):
            row[j >> 3] ^= np.uint8(1 << (j & 7))
    skipA = (rng.random(nA) < skipf).astype(np.uint8)
    skipB = (rng.random(nB) < skipf).astype(np.uint8)
    return A, B, skipA, skipB


def reference_matches(lib, A, B, skipA, skipB, thr, nb, use_ratio, ratio):
    D = np.ascontiguousarray(np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(2), np.float32)
    oa, od = np.zeros(len(B), np.int32), np.zeros(len(B), np.float32)
    lib.okr_match(C.c_void_p(D.ctypes.data), len(A), len(B), C.c_void_p(skipA.ctypes.data), C.c_void_p(skipB.ctypes.data), C.c_float(thr), nb,
                  int(use_ratio), C.c_float(ratio), 1, C.c_void_p(oa.ctypes.data), C.c_void_p(od.ctypes.data))
    m = np.array([(int(oa[b]), b, float(od[b])) for b in range(len(B)) if oa[b] >= 0], np.float64).reshape(-1, 3)
    return m


def generate():
    subprocess.run(["make", "-s", "-f", "Makefile.ref"], cwd=os.path.join(ROOT, "oracle"), check=True)
    lib = C.CDLL(SO)
    lib.okr_match.restype = C.c_int
    out = {"n_cases": np.array(len(CASES)), "source": np.array("okvis_matcher (unmodified reference sources, 1 matcher thread) via oracle/Makefile.ref")}
    for i, spec in enumerate(CASES):
        A, B, sA, sB = make_case(i, spec)
        out["A%d" % i], out["B%d" % i], out["skipA%d" % i], out["skipB%d" % i] = A, B, sA, sB
        out["params%d" % i] = np.array([spec[5], spec[6], float(spec[7]), spec[8]])
        out["matches%d" % i] = reference_matches(lib, A, B, sA, sB, spec[5], spec[6], spec[7], spec[8])
    return out


def main():
    out = generate()
    if "--check" in sys.argv:
        old = np.load(OUT)
        for k in out:
            if k != "source":
                assert np.array_equal(out[k], old[k]), k
        print("fixture reproduced")
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT), "bytes;", [len(out["matches%d" % i]) for i in range(len(CASES))], "matches per case")


if __name__ == "__main__":
    main()
