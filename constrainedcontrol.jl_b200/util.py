"""Algebraic-Riccati utilities of the reference (src/util/util.jl:1-59) through the C-ABI (GPU only, batched).

    dare(A, B, Q, R)        util.jl:1-20      care(A, B, Q, R)        util.jl:22-37
    lqr(A, B, Q, R)         util.jl:39-43     dlqr(A, B, Q, R[, Δt])  util.jl:45-49, :51-59

Every argument may carry a leading batch axis ([S][n][n] ...); matrices without one are shared by the whole batch.  A single
problem returns a single matrix, like the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f64, ptr

DISCRETE, CONTINUOUS = 0, 1


def _batched(A, B, Q, R):
    A = f64(A); B = f64(B); Q = f64(Q); R = f64(R)
    n = A.shape[-1]
    if B.ndim == 1:
        B = B.reshape(n, 1)
    m = B.shape[-1]
    if R.ndim < 2:                  # a scalar weight (or one scalar per problem) for a single input
        R = R.reshape(1, 1) if R.size == 1 else R.reshape(-1, 1, 1)
    single = A.ndim == 2 and B.ndim == 2 and Q.ndim == 2 and R.ndim == 2
    S = max(x.shape[0] if x.ndim == 3 else 1 for x in (A, B, Q, R))
    def full(x, r, c):
        assert x.shape[-2:] == (r, c), f"expected trailing shape {(r, c)}, got {x.shape}"
        x = x.reshape(-1, r, c)
        assert x.shape[0] in (1, S), "batch sizes differ"
        return np.ascontiguousarray(np.broadcast_to(x, (S, r, c)))
    return full(A, n, n), full(B, n, m), full(Q, n, n), full(R, m, m), S, n, m, single


def are_solve(kind, A, B, Q, R, want_X=True, want_K=False, device: int = 0, info: bool = False):
    """S problems at once: returns X [S][n][n] and/or K [S][m][n] (+ dict(iters, status) with info=True)"""
    A, B, Q, R, S, n, m, single = _batched(A, B, Q, R)
    X = np.empty((S, n, n)) if want_X else None
    K = np.empty((S, m, n)) if want_K else None
    iters = np.empty(S, np.int32); status = np.empty(S, np.int32)
    check(_lib.lib().ccb_are_solve(C.c_int32(device), C.c_int32(kind), C.c_int32(S), C.c_int32(n), C.c_int32(m), ptr(A), ptr(B), ptr(Q), ptr(R),
                                   ptr(X) if want_X else None, ptr(K) if want_K else None, ptr(iters), ptr(status)))
    out = [x[0] if single else x for x in (X, K) if x is not None]
    if info:
        out.append({"iters": iters, "status": status})
    elif np.any(status != 0):
        # the reference would return garbage or throw from ordschur (util.jl:15-18, :28-36); without info=True the caller cannot see the
        # per-problem flags, so an unsolved problem must not pass silently
        bad = np.flatnonzero(status != 0)
        raise ArithmeticError(f"are_solve: {bad.size} of {S} problems have no stabilising solution within the doubling iteration "
                              f"(first: problem {int(bad[0])}, status {int(status[bad[0]])}); pass info=True to get the per-problem flags")
    return out[0] if len(out) == 1 else tuple(out)


def dare(A, B, Q, R, **kw):
    """util.jl:1-20"""
    return are_solve(DISCRETE, A, B, Q, R, **kw)


def care(A, B, Q, R, **kw):
    """util.jl:22-37"""
    return are_solve(CONTINUOUS, A, B, Q, R, **kw)


def lqr(A, B, Q, R, **kw):
    """util.jl:39-43  K = R \\ B'P"""
    return are_solve(CONTINUOUS, A, B, Q, R, want_X=False, want_K=True, **kw)


def dlqr(A, B, Q, R, dt=None, **kw):
    """util.jl:45-49; with Δt (util.jl:51-59): Q Δt, R Δt, A Δt + I, B Δt first"""
    if dt is not None:
        A = f64(A); n = A.shape[-1]
        A = A * dt + np.eye(n); B = f64(B) * dt; Q = f64(Q) * dt; R = f64(R) * dt
    return are_solve(DISCRETE, A, B, Q, R, want_X=False, want_K=True, **kw)
