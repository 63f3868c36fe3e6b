"""CPU: the oracle's restatement of the constrained Riccati recursion (lqr.jl:141-184) against a brute-force solution of the optimal-control
problem it is meant to solve, built WITHOUT the recursion and without eliminating the multipliers:

    min  sum_{k=1}^{N-1} (z_k' Q z_k + u_k' R u_k) + z_N' Q z_N
    s.t. z_{k+1} = A z_k + Bu u_k + Bl lam_k ,   G z_{k+1} = 0          (k = 1 .. N-1)

solved as one dense KKT system in (u_k, lam_k, z_{k+1}).  The first input of the optimum must be u_1 = -Ku[1] z_1, the second u_2 = -Ku[2] z_2,
for any z_1 -- a transcription error in M, b or the P update (lqr.jl:152-170) breaks it."""
import numpy as np
import pytest


def brute_force_inputs(A, Bu, Bl, G, Q, R, N, z1):
    n, mu, ml = A.shape[0], Bu.shape[1], Bl.shape[1]
    per = mu + ml + n                       # (u_k, lam_k, z_{k+1}) for k = 1 .. N-1
    nv = per * (N - 1)
    nc = (n + ml) * (N - 1)
    H = np.zeros((nv, nv)); C = np.zeros((nc, nv)); d = np.zeros(nc)
    for k in range(N - 1):
        o = per * k
        H[o:o + mu, o:o + mu] = 2 * R
        H[o + mu + ml:o + per, o + mu + ml:o + per] = 2 * Q
        r = (n + ml) * k
        # z_{k+1} - Bu u_k - Bl lam_k - A z_k = 0
        C[r:r + n, o + mu + ml:o + per] = np.eye(n)
        C[r:r + n, o:o + mu] = -Bu
        C[r:r + n, o + mu:o + mu + ml] = -Bl
        if k == 0:
            d[r:r + n] = A @ z1
        else:
            C[r:r + n, per * (k - 1) + mu + ml:per * k] = -A
        C[r + n:r + n + ml, o + mu + ml:o + per] = G          # G z_{k+1} = 0
    KKT = np.block([[H, C.T], [C, np.zeros((nc, nc))]])
    sol = np.linalg.solve(KKT, np.concatenate([np.zeros(nv), d]))
    u = [sol[per * k:per * k + mu] for k in range(N - 1)]
    z = [z1] + [sol[per * k + mu + ml:per * (k + 1)] for k in range(N - 1)]
    return u, z


@pytest.mark.parametrize("seed,n,mu,ml,N", [(0, 5, 1, 2, 6), (1, 8, 2, 3, 7), (2, 6, 3, 1, 5)])
def test_recursion_gains_are_the_optimal_feedback(orc, seed, n, mu, ml, N):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n)) * 0.6
    Bu = rng.normal(size=(n, mu)); Bl = rng.normal(size=(n, ml)); G = rng.normal(size=(ml, n))
    L = rng.normal(size=(n, n)); Q = L @ L.T / n + 0.1 * np.eye(n)
    M = rng.normal(size=(mu, mu)); R = M @ M.T / mu + 0.5 * np.eye(mu)
    ref = orc.dlqr(A, Bu, Bl, G, Q, R, N)
    assert ref["kexit"] == 1                                   # no early break at this short horizon
    Ku = np.asarray(ref["Ku"])                                 # [N-1][mu][n], row k-1 = gain of step k
    for trial in range(3):
        z1 = rng.normal(size=n)
        u, z = brute_force_inputs(A, Bu, Bl, G, Q, R, N, z1)
        for k in range(2):                                     # the first two steps of the optimal trajectory
            uk = -Ku[k] @ z[k]
            assert np.abs(uk - u[k]).max() < 1e-9 * max(1.0, np.abs(u[k]).max()), (trial, k)
        assert np.abs(G @ z[1]).max() < 1e-10
