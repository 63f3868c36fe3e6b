"""CPU: the restatement of util.jl's Riccati utilities (oracle/are.py) against independent checks -- residuals of the equations,
stability of the closed loop, scipy's own ARE solvers, and the scalar closed form."""
import numpy as np
import pytest
import scipy.linalg as sl

import are as orc_are


def _problem(seed, n, m, discrete):
    rng = np.random.default_rng(seed)
    # spectral radius ~1.1 (discrete) / spectrum in a disc of radius ~1 around -0.5 (continuous): a few unstable modes, well conditioned
    A = rng.standard_normal((n, n)) / np.sqrt(n) * 1.1 if discrete else rng.standard_normal((n, n)) / np.sqrt(n) - 0.5 * np.eye(n)
    B = rng.standard_normal((n, m))
    L = rng.standard_normal((n, n)); Q = L @ L.T / n + np.eye(n)
    M = rng.standard_normal((m, m)); R = M @ M.T / m + np.eye(m)
    return A, B, Q, R


@pytest.mark.parametrize("n,m", [(2, 1), (6, 2), (12, 3), (24, 1)])
def test_dare_residual_stability_and_scipy(n, m):
    A, B, Q, R = _problem(n, n, m, True)
    X = orc_are.dare(A, B, Q, R)
    assert np.abs(X - X.T).max() < 1e-10 * np.abs(X).max()
    res = A.T @ X @ A - X - A.T @ X @ B @ np.linalg.solve(R + B.T @ X @ B, B.T @ X @ A) + Q
    assert np.abs(res).max() < 1e-10 * np.abs(X).max()
    K = orc_are.dlqr(A, B, Q, R)
    assert np.abs(np.linalg.eigvals(A - B @ K)).max() < 1.0
    assert np.abs(X - sl.solve_discrete_are(A, B, Q, R)).max() < 1e-9 * np.abs(X).max()


@pytest.mark.parametrize("n,m", [(2, 1), (6, 2), (12, 3), (24, 1)])
def test_care_residual_stability_and_scipy(n, m):
    A, B, Q, R = _problem(100 + n, n, m, False)
    X = orc_are.care(A, B, Q, R)
    res = A.T @ X + X @ A - X @ B @ np.linalg.solve(R, B.T @ X) + Q
    assert np.abs(res).max() < 1e-9 * np.abs(X).max()
    K = orc_are.lqr(A, B, Q, R)
    assert np.linalg.eigvals(A - B @ K).real.max() < 0.0
    assert np.abs(X - sl.solve_continuous_are(A, B, Q, R)).max() < 1e-8 * np.abs(X).max()


def test_scalar_closed_forms():
    # x+ = a x + b u, cost q x^2 + r u^2:  b^2 X^2 + (r - a^2 r - q b^2) X - q r = 0
    a, b, q, r = 1.2, 0.5, 2.0, 0.3
    X = orc_are.dare([[a]], [[b]], [[q]], [[r]])[0, 0]
    c1 = r - a * a * r - q * b * b
    assert abs(X - (-c1 + np.sqrt(c1 * c1 + 4 * b * b * q * r)) / (2 * b * b)) < 1e-12
    # continuous: 2 a X - b^2 X^2 / r + q = 0
    Xc = orc_are.care([[a]], [[b]], [[q]], [[r]])[0, 0]
    assert abs(Xc - r * (a + np.sqrt(a * a + b * b * q / r)) / (b * b)) < 1e-12


def test_dlqr_dt_method_scales_before_solving():
    A, B, Q, R = _problem(7, 4, 1, False)
    dt = 0.01
    K = orc_are.dlqr(A, B, Q, R, dt)
    K2 = orc_are.dlqr(A * dt + np.eye(4), B * dt, Q * dt, R * dt)
    assert np.array_equal(K, K2)
