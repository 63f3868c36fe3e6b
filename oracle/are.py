"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's algebraic-Riccati utilities (src/util/util.jl:1-59).

Parity unpinned: the reference ships no test or golden vector for these functions (they are only reachable through the
mis-dispatching branch lqr.jl:30-37); this file restates them line by line on scipy's LAPACK wrappers, which is the same
(generalised) ordered real Schur form Julia's `schur` / `ordschur` call.  Product code never imports this module.
"""
import numpy as np
import scipy.linalg as sl


def dare(A, B, Q, R):
    """util.jl:1-20: ordered QZ of the symplectic pencil F - z E, stable deflating subspace -> X = Z21 / Z11"""
    A = np.asarray(A, float); B = np.asarray(B, float); Q = np.asarray(Q, float); R = np.asarray(R, float)
    n = A.shape[0]
    I = np.eye(n)
    E = np.block([[I, B @ np.linalg.solve(R, B.T)], [np.zeros((n, n)), A.T]])            # util.jl:5-8   (B/R*B' = B R^-1 B')
    F = np.block([[A, np.zeros((n, n))], [-Q, I]])                                       # util.jl:10-13
    # util.jl:15-16  schur(F, E) then ordschur(select = |alpha/beta| < 1)
    _, _, _, _, _, Z = sl.ordqz(F, E, sort=lambda alpha, beta: np.abs(alpha) < np.abs(beta), output="real")
    return np.linalg.solve(Z[:n, :n].T, Z[n:, :n].T).T                                   # util.jl:18   Z21 / Z11


def care(A, B, Q, R):
    """util.jl:22-37: ordered real Schur form of the Hamiltonian matrix, stable invariant subspace -> X = U21 / U11"""
    A = np.asarray(A, float); B = np.asarray(B, float); Q = np.asarray(Q, float); R = np.asarray(R, float)
    n = A.shape[0]
    G = B @ np.linalg.solve(R, B.T)                                                      # util.jl:23
    Zm = np.block([[A, -G], [-Q, -A.T]])                                                 # util.jl:25-26
    _, U, _ = sl.schur(Zm, output="real", sort="lhp")                                    # util.jl:28-29
    U11, U21 = U[:n, :n], U[n:, :n]                                                      # util.jl:32-34
    return np.linalg.solve(U11.T, U21.T).T                                               # util.jl:36


def lqr(A, B, Q, R):
    """util.jl:39-43"""
    P = care(A, B, Q, R)
    return np.linalg.solve(np.asarray(R, float), np.asarray(B, float).T @ P)


def dlqr(A, B, Q, R, dt=None):
    """util.jl:45-49 and the Δt method :51-59"""
    A = np.asarray(A, float); B = np.asarray(B, float); Q = np.asarray(Q, float); R = np.asarray(R, float)
    if dt is not None:
        Q = Q * dt; R = R * dt; A = A * dt + np.eye(A.shape[0]); B = B * dt
    P = dare(A, B, Q, R)
    return np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
