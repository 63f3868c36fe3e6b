"""GPU: ccb_are_solve (K5, structure-preserving doubling) against the restatement of util.jl (oracle/are.py, ordered Schur forms).
The two algorithms differ, so the bound is the conditioning of the equation, not summation order: 1e-9 relative on well-posed problems."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(seed, n, m, discrete):
    rng = np.random.default_rng(seed)
    # spectral radius ~1.1 (discrete) / spectrum in a disc of radius ~1 around -0.5 (continuous): a few unstable modes, well conditioned
    A = rng.standard_normal((n, n)) / np.sqrt(n) * 1.1 if discrete else rng.standard_normal((n, n)) / np.sqrt(n) - 0.5 * np.eye(n)
    B = rng.standard_normal((n, m))
    L = rng.standard_normal((n, n)); Q = L @ L.T / n + np.eye(n)
    M = rng.standard_normal((m, m)); R = M @ M.T / m + np.eye(m)
    return A, B, Q, R


def _rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.mark.parametrize("n,m", [(1, 1), (2, 1), (6, 2), (12, 3), (24, 1), (48, 7), (60, 4)])
def test_dare_and_dlqr_match_the_schur_restatement(cc, n, m):
    import are as orc_are
    probs = [_problem(1000 * n + s, n, m, True) for s in range(5)]
    A, B, Q, R = (np.stack(x) for x in zip(*probs))
    X, K, info = cc.util.are_solve(cc.util.DISCRETE, A, B, Q, R, want_X=True, want_K=True, info=True)
    assert (info["status"] == 0).all() and (info["iters"] < 40).all()
    for s in range(5):
        assert _rel(X[s], orc_are.dare(*probs[s])) < 1e-9
        assert _rel(K[s], orc_are.dlqr(*probs[s])) < 1e-9
        assert np.array_equal(X[s], X[s].T)


@pytest.mark.parametrize("n,m", [(1, 1), (2, 1), (6, 2), (12, 3), (24, 1), (48, 7), (60, 4)])
def test_care_and_lqr_match_the_schur_restatement(cc, n, m):
    import are as orc_are
    probs = [_problem(2000 * n + s, n, m, False) for s in range(5)]
    A, B, Q, R = (np.stack(x) for x in zip(*probs))
    X, K, info = cc.util.are_solve(cc.util.CONTINUOUS, A, B, Q, R, want_X=True, want_K=True, info=True)
    assert (info["status"] == 0).all()
    for s in range(5):
        Xo = orc_are.care(*probs[s])
        res = lambda P: np.abs(probs[s][0].T @ P + P @ probs[s][0] - P @ probs[s][1] @ np.linalg.solve(probs[s][3], probs[s][1].T @ P) + probs[s][2]).max()
        # the Hamiltonian Schur form loses digits on the larger random problems: compare where it is itself accurate, and require
        # the device residual to be at least as small as the restatement's everywhere
        assert res(X[s]) <= max(10 * res(Xo), 1e-9 * np.abs(Xo).max())
        if res(Xo) < 1e-10 * np.abs(Xo).max():
            assert _rel(X[s], Xo) < 1e-8
            assert _rel(K[s], orc_are.lqr(*probs[s])) < 1e-8


def test_single_problem_api_and_dt_method(cc):
    import are as orc_are
    A, B, Q, R = _problem(5, 4, 1, False)
    # reference-shaped calls: one problem in, one matrix out
    assert cc.util.dare(A * 0.3, B, Q, R).shape == (4, 4)
    assert cc.util.lqr(A, B, Q, R).shape == (1, 4)
    dt = 0.01
    assert _rel(cc.util.dlqr(A, B, Q, R, dt), orc_are.dlqr(A, B, Q, R, dt)) < 1e-9
    # the linearised cartpole on a cart (minimal coordinates), the textbook use of lqr / dlqr
    g, mc, mp, l = 9.81, 0.5, 1.0, 0.5
    Ac = np.array([[0, 1, 0, 0], [0, 0, -mp * g / mc, 0], [0, 0, 0, 1], [0, 0, (mc + mp) * g / (mc * l), 0]], float)
    Bc = np.array([[0], [1 / mc], [0], [-1 / (mc * l)]], float)
    assert _rel(cc.util.lqr(Ac, Bc, np.eye(4), np.eye(1)), orc_are.lqr(Ac, Bc, np.eye(4), np.eye(1))) < 1e-9
    assert _rel(cc.util.dlqr(Ac, Bc, np.eye(4), np.eye(1), dt), orc_are.dlqr(Ac, Bc, np.eye(4), np.eye(1), dt)) < 1e-8


def test_unstabilisable_problem_is_flagged_and_bad_arguments_raise(cc):
    # unstable mode that the input cannot reach: no stabilising solution
    A = np.diag([1.5, 0.5]); B = np.array([[0.0], [1.0]]); Q = np.eye(2); R = np.eye(1)
    X, info = cc.util.dare(A, B, Q, R, info=True)
    assert info["status"][0] != 0
    with pytest.raises(cc.CcbError):
        cc.util.dare(np.eye(2), np.ones((2, 3)), np.eye(2), np.eye(3))      # m > n is not supported


def test_large_batch_of_cartpole_weights(cc):
    """4096 (Q, R) scalings of one plant in one launch; every closed loop is stable and the batch equals problem-by-problem calls"""
    rng = np.random.default_rng(3)
    A, B, _, _ = _problem(11, 12, 2, True)
    S = 4096
    qs = 10 ** rng.uniform(-1, 1, S); rs = 10 ** rng.uniform(-1, 1, S)
    Q = qs[:, None, None] * np.eye(12); R = rs[:, None, None] * np.eye(2)
    K = cc.util.dlqr(A, B, Q, R)
    assert K.shape == (S, 2, 12)
    rho = np.abs(np.linalg.eigvals(A[None] - B[None] @ K)).max(axis=1)
    assert rho.max() < 1.0
    for s in (0, 17, 4095):
        assert np.array_equal(K[s], cc.util.dlqr(A, B, Q[s], R[s]))
