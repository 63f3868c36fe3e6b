"""linearsystem / dlqr / LQR and TrackingLQR gain synthesis through the C-ABI (GPU only).

    linearsystem(mechanism, xd, vd, qd, ωd, Fτd, bodyids, eqcids) -> A, Bu, Bλ, G      reference src/control/lqr.jl:63
    dlqr(A, Bu, Bλ, G, Q, R, N)                                                        reference src/control/lqr.jl:141-184
    dlqr(mechanism, xd, vd, qd, ωd, Fτd, eqcids, Q, R, N)                              reference src/control/lqr_tracking.jl:73-122
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, f64, i32, ptr


def _opts(eps=1e-10, newton_iter=100, line_iter=10, lanes=0):
    return _lib.SimOpts(eps, newton_iter, line_iter, lanes, 0)


def _mh(mech):
    from .control import handle
    return handle(mech)


def linearsystem(mech, zd, Fd, eqc_idx, device: int = 0, want_lambda: bool = False, **kw):
    """zd [S][nb][13] (xd,qd,vd,ωd), Fd [S][nu]; returns A [S][n][n], Bu [S][n][nu], Bl [S][n][ml], G [S][ml][n]"""
    mh = _mh(mech)
    eqc_idx = i32(eqc_idx); nu = eqc_idx.size
    zd = f64(zd).reshape(-1, mh.nb, 13); S = zd.shape[0]
    Fd = np.broadcast_to(f64(Fd).reshape(-1, nu), (S, nu)).copy()
    n, ml = mh.n, mh.ml
    A = np.empty((S, n, n)); Bu = np.empty((S, n, nu)); Bl = np.empty((S, n, ml)); G = np.empty((S, ml, n)); lam = np.empty((S, ml))
    o = _opts(**kw)
    check(_lib.lib().ccb_linearize(mh.h, C.c_int32(device), C.c_int32(S), ptr(zd), ptr(Fd), ptr(eqc_idx), C.c_int32(nu), C.byref(o), ptr(A), ptr(Bu),
                                   ptr(Bl), ptr(G), ptr(lam)))
    return (A, Bu, Bl, G, lam) if want_lambda else (A, Bu, Bl, G)


def _weights(W, S, k):
    W = f64(W)
    if W.size == k * k:
        return W.reshape(k, k), 0
    assert W.size == S * k * k, "weights must be one matrix or one per system"
    return W.reshape(S, k, k), k * k


def dlqr(A, Bu, Bl, G, Q, R, N: int, device: int = 0):
    """batched dlqr: A [S][n][n] ... -> dict(Ku [S][N-1][nu][n], kexit [S], iters [S])"""
    A = f64(A)
    n = A.shape[-1]
    A = A.reshape(-1, n, n); S = A.shape[0]
    Bu = f64(Bu).reshape(S, n, -1); nu = Bu.shape[2]
    Bl = f64(Bl).reshape(S, n, -1); ml = Bl.shape[2]
    G = f64(G).reshape(S, ml, n)
    Q, qs = _weights(Q, S, n)
    R, rs = _weights(R, S, nu)
    Ku = np.empty((S, N - 1, nu, n)); kexit = np.empty(S, np.int32); iters = np.empty(S, np.int32)
    check(_lib.lib().ccb_dlqr(C.c_int32(device), C.c_int32(S), C.c_int32(n), C.c_int32(nu), C.c_int32(ml), C.c_int32(N), ptr(A), ptr(Bu), ptr(Bl),
                              ptr(G), ptr(Q), C.c_int64(qs), ptr(R), C.c_int64(rs), ptr(Ku), ptr(kexit), ptr(iters)))
    return dict(Ku=Ku, kexit=kexit, iters=iters)


def lqr_synthesize(mech, zd, Fd, eqc_idx, Q, R, N: int, device: int = 0, **kw):
    mh = _mh(mech)
    eqc_idx = i32(eqc_idx); nu = eqc_idx.size
    zd = f64(zd).reshape(-1, mh.nb, 13); S = zd.shape[0]
    Fd = np.broadcast_to(f64(Fd).reshape(-1, nu), (S, nu)).copy()
    Q, qs = _weights(Q, S, mh.n)
    R, rs = _weights(R, S, nu)
    Ku = np.empty((S, N - 1, nu, mh.n)); kexit = np.empty(S, np.int32); iters = np.empty(S, np.int32)
    o = _opts(**kw)
    check(_lib.lib().ccb_lqr_synthesize(mh.h, C.c_int32(device), C.c_int32(S), ptr(zd), ptr(Fd), ptr(eqc_idx), C.c_int32(nu), ptr(Q), C.c_int64(qs),
                                        ptr(R), C.c_int64(rs), C.c_int32(N), C.byref(o), ptr(Ku), ptr(kexit), ptr(iters)))
    return dict(Ku=Ku, kexit=kexit, iters=iters)


def tracking_synthesize(mech, traj, Ftau, eqc_idx, Q, R, device: int = 0, **kw):
    mh = _mh(mech)
    eqc_idx = i32(eqc_idx); nu = eqc_idx.size
    traj = f64(traj); traj = traj.reshape(-1, traj.shape[-3], mh.nb, 13)
    S, N = traj.shape[0], traj.shape[1]
    Ftau = np.broadcast_to(f64(Ftau).reshape(-1, N, nu), (S, N, nu)).copy()
    Q = f64(Q).reshape(mh.n, mh.n); R = f64(R).reshape(nu, nu)
    Ku = np.empty((S, N - 1, nu, mh.n)); kexit = np.empty(S, np.int32); iters = np.empty(S, np.int32)
    o = _opts(**kw)
    check(_lib.lib().ccb_tracking_synthesize(mh.h, C.c_int32(device), C.c_int32(S), C.c_int32(N), ptr(traj), ptr(Ftau), ptr(eqc_idx), C.c_int32(nu),
                                             ptr(Q), ptr(R), C.byref(o), ptr(Ku), ptr(kexit), ptr(iters)))
    return dict(Ku=Ku, kexit=kexit, iters=iters)


def phase_ms():
    """device time (ms) of the phases of the last synthesis call: dict(linearize, project, sweep)"""
    out = (C.c_double * 4)()
    check(_lib.lib().ccb_synth_phase_ms(out))
    return dict(linearize=out[0], project=out[1], sweep=out[2])


def phase_total_ms():
    """elapsed device time (ms) of the last synthesis call: the separate total of an overlapped call (TrackingLQR pipeline), else the sum of the phases"""
    out = (C.c_double * 4)()
    check(_lib.lib().ccb_synth_phase_ms(out))
    return out[3] if out[3] > 0 else out[0] + out[1] + out[2]


def release_scratch():
    """give the library's persistent device scratch back to the driver (ccb_scratch_release)"""
    check(_lib.lib().ccb_scratch_release())
