"""Oracle of the dense numeric core of OKVIS' marginalisation (SURVEY 8(f) row 1 -- NEXT TIER: groundwork only, the
device path does not exist yet).  The reference has no golden values for it and Eigen's eigenvector basis is not unique,
so the restatement (oracle/oracle_marg.hpp) is pinned through the quantities that ARE unique: the reduced system of
MarginalizationError::marginalizeOut against an independent numpy Schur complement with pseudo-inverses, and
J^T J / J^T e0 / rank of updateErrorComputation."""
import numpy as np
import pytest


def random_system(rng, n_dense, n_lm, rank_deficient=False):
    n = n_dense + 3 * n_lm
    m = n + 20
    A = rng.normal(size=(m, n))
    A[:, n_dense:] *= 30.0                               # landmark columns on another scale (exercises the preconditioner)
    # landmarks only couple to the dense part (block-diagonal landmark Hessian, as in the estimator)
    H = A.T @ A
    for i in range(n_lm):
        for j in range(n_lm):
            if i != j:
                H[n_dense + 3 * i:n_dense + 3 * i + 3, n_dense + 3 * j:n_dense + 3 * j + 3] = 0.0
    if rank_deficient:
        H[:, 0] = 0.0; H[0, :] = 0.0                    # an unobservable direction
    b = H @ rng.normal(size=n)                           # b in range(H)
    return H, b


def numpy_schur(H, b, idx_b):
    n = H.shape[0]
    idx_a = [i for i in range(n) if i not in set(idx_b)]
    U, W, V = H[np.ix_(idx_a, idx_a)], H[np.ix_(idx_a, idx_b)], H[np.ix_(idx_b, idx_b)]
    Vi = np.linalg.pinv(0.5 * (V + V.T), hermitian=True)
    return U - W @ Vi @ W.T, b[idx_a] - W @ Vi @ b[idx_b]


def test_jacobi_eigensolver(oracle):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 9, 40):
        A = rng.normal(size=(n, n)); A = A + A.T
        w, V = oracle.sym_eig(A)
        assert np.all(np.diff(w) >= -1e-12)
        assert np.allclose(V.T @ V, np.eye(n), atol=1e-12)
        assert np.allclose(V @ np.diag(w) @ V.T, A, atol=1e-10 * max(1.0, np.abs(A).max()))
        assert np.allclose(w, np.linalg.eigvalsh(A), atol=1e-10 * max(1.0, np.abs(A).max()))


def test_landmark_stage_matches_schur_complement(oracle):
    rng = np.random.default_rng(4)
    n_dense, n_lm = 15, 8
    H, b = random_system(rng, n_dense, n_lm)
    # marginalise landmarks 1, 2 (contiguous -> one unified range), 5 and 7
    ranges = [(n_dense + 3, 6), (n_dense + 15, 3), (n_dense + 21, 3)]
    idx_b = [s + k for s, l in ranges for k in range(l)]
    Hr, br = oracle.marginalize_stage(H, b, ranges, landmark_blocks=True)
    He, be = numpy_schur(H, b, idx_b)
    assert Hr.shape == He.shape
    assert np.allclose(Hr, He, rtol=1e-9, atol=1e-9 * np.abs(He).max())
    assert np.allclose(br, be, rtol=1e-9, atol=1e-9 * np.abs(be).max())
    assert np.allclose(Hr, Hr.T, atol=1e-9 * np.abs(Hr).max())


def test_dense_stage_matches_schur_complement_also_rank_deficient(oracle):
    rng = np.random.default_rng(5)
    for deficient in (False, True):
        H, b = random_system(rng, 21, 0, rank_deficient=deficient)
        ranges = [(0, 6), (12, 9)]                       # a pose and a speed/bias block
        idx_b = [s + k for s, l in ranges for k in range(l)]
        Hr, br = oracle.marginalize_stage(H, b, ranges, landmark_blocks=False)
        He, be = numpy_schur(H, b, idx_b)
        assert np.allclose(Hr, He, rtol=1e-8, atol=1e-8 * np.abs(He).max())
        assert np.allclose(br, be, rtol=1e-8, atol=1e-8 * np.abs(be).max())


def test_two_stages_equal_joint_elimination(oracle):
    """marginalizeOut runs the landmark stage, then the dense stage: same result as eliminating everything at once."""
    rng = np.random.default_rng(6)
    n_dense, n_lm = 18, 5
    H, b = random_system(rng, n_dense, n_lm)
    lm_ranges = [(n_dense, 3 * n_lm)]
    H1, b1 = oracle.marginalize_stage(H, b, lm_ranges, landmark_blocks=True)
    H2, b2 = oracle.marginalize_stage(H1, b1, [(0, 6)], landmark_blocks=False)
    He, be = numpy_schur(H, b, list(range(0, 6)) + list(range(n_dense, n_dense + 3 * n_lm)))
    assert np.allclose(H2, He, rtol=1e-8, atol=1e-8 * np.abs(He).max())
    assert np.allclose(b2, be, rtol=1e-8, atol=1e-8 * np.abs(be).max())


@pytest.mark.parametrize("deficient", [False, True])
def test_update_error_computation_invariants(oracle, deficient):
    rng = np.random.default_rng(7)
    H, b = random_system(rng, 24, 0, rank_deficient=deficient)
    J, e0, rank = oracle.marg_update_error_computation(H, b)
    n = H.shape[0]
    assert rank == (n - 1 if deficient else n)
    scale = np.abs(H).max()
    assert np.allclose(J.T @ J, H, atol=1e-9 * scale)                 # H = J^T J
    assert np.allclose(-J.T @ e0, b, atol=1e-8 * np.abs(b).max())     # b0 = -J^T e0 (b in range(H))
    # the prior evaluates to e = e0 + J dchi: its gradient at dchi = 0 is J^T e0 = -b and its Hessian J^T J = H
    dchi = rng.normal(size=n) * 1e-3
    e = e0 + J @ dchi
    assert np.isclose(0.5 * e @ e, 0.5 * e0 @ e0 - b @ dchi + 0.5 * dchi @ H @ dchi, rtol=1e-9)


def test_marginalised_prior_feeds_the_existing_functor(oracle):
    """J, e0 of updateErrorComputation are exactly what okb_window_desc.marg / okb_eval_marginalization consume
    (row (a) M): at the linearisation point the functor returns e0 and J (pose blocks lifted at x0)."""
    rng = np.random.default_rng(8)
    H, b = random_system(rng, 15, 0)                                   # one pose (6) + one speed/bias (9)
    J, e0, _ = oracle.marg_update_error_computation(H, b)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    x0 = np.concatenate([rng.normal(size=3), q, rng.normal(size=9)])
    marg = dict(J=J, e0=e0, x0=x0, block_kind=np.array([0, 1], np.int32), block_idx=np.array([0, 0], np.uint32))
    r, J_eff = oracle.eval_marginalization(marg, x0)
    assert np.allclose(r, e0, atol=1e-12 * max(1.0, np.abs(e0).max()))
    assert np.allclose(J_eff, J, atol=1e-9 * np.abs(J).max())
