"""ctypes wrapper of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (constrainedcontrol.jl_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "orc_math.hpp", "orc_mech.hpp", "orc_dense.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs if os.path.exists(s)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


class OrcCtrl(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("nu", C.c_int32), ("N", C.c_int32), ("pad", C.c_int32),
        ("eqc", C.c_void_p), ("K", C.c_void_p), ("K_bstride", C.c_int64),
        ("zd", C.c_void_p), ("zd_bstride", C.c_int64), ("Fd", C.c_void_p), ("Fd_bstride", C.c_int64),
        ("U", C.c_void_p), ("U_bstride", C.c_int64),
        ("P", C.c_void_p), ("I", C.c_void_p), ("D", C.c_void_p),
        ("goal", C.c_void_p), ("goal_bstride", C.c_int64),
        ("pid_int", C.c_void_p), ("pid_last", C.c_void_p),
    ]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_mech_create.restype = C.c_void_p
        _LIB.orc_dlqr.restype = C.c_int
        _LIB.orc_tracking_dlqr.restype = C.c_int
        _LIB.orc_linearize.restype = C.c_int
        _LIB.orc_solve_both.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


KIND = dict(none=0, table=1, lqr_inf=2, lqr_finite=3, tracking=4, pid=5)


class OracleMech:
    def __init__(self, desc: dict):
        L = lib()
        d = desc
        self.desc = d
        self._keep = [_f64(d["mass"]), _f64(d["inertia"]), _i32(d["eqc_parent"]), _i32(d["eqc_joint_ptr"]), _i32(d["joint_child"]),
                      _i32(d["joint_nT"]), _i32(d["joint_nR"]), _f64(d["joint_axis"]), _f64(d["joint_p1"]), _f64(d["joint_p2"]),
                      _f64(d["joint_qoff"])]
        self.h = C.c_void_p(L.orc_mech_create(C.c_int(d["nb"]), C.c_int(d["ne"]), C.c_int(d["nj"]), C.c_double(d["dt"]), C.c_double(d["g"]),
                                              *[_p(a) for a in self._keep], C.c_double(d.get("reg", 0.0))))
        dims = np.zeros(8, dtype=np.int32)
        L.orc_mech_dims(self.h, _p(dims))
        self.nb, self.ne, self.nj, self.ml, self.nnodes, self.nloop = [int(v) for v in dims[:6]]
        self.n = 12 * self.nb
        self.dt = d["dt"]

    def __del__(self):
        try:
            if self.h:
                lib().orc_mech_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def graph(self) -> dict:
        N = self.nnodes
        eqc_id = np.zeros(self.ne, np.int32); body_id = np.zeros(self.nb, np.int32)
        dfs = np.zeros(N, np.int32); par = np.zeros(N, np.int32); loops = np.zeros(2 * max(self.nloop, 1), np.int32)
        eqc_dim = np.zeros(self.ne, np.int32); eqc_row0 = np.zeros(self.ne, np.int32); act = np.zeros(self.ne, np.int32)
        lib().orc_mech_graph(self.h, _p(eqc_id), _p(body_id), _p(dfs), _p(par), _p(loops), _p(eqc_dim), _p(eqc_row0), _p(act))
        return dict(eqc_id=eqc_id, body_id=body_id, dfs_list=dfs, parents=par, loop_edges=loops[: 2 * self.nloop].reshape(-1, 2),
                    eqc_dim=eqc_dim, eqc_row0=eqc_row0, act_kind=act)

    # ---- low-level pieces for unit tests
    def residual(self, z, s, lam, u):
        f = np.zeros(6 * self.nb + self.ml)
        lib().orc_residual(self.h, _p(_f64(z)), _p(_f64(s)), _p(_f64(lam)), _p(_f64(u)), _p(f))
        return f

    def kkt(self, z, s, lam, u):
        n = 6 * self.nb + self.ml
        M = np.zeros((n, n)); f = np.zeros(n)
        lib().orc_kkt(self.h, _p(_f64(z)), _p(_f64(s)), _p(_f64(lam)), _p(_f64(u)), _p(M), _p(f))
        return M, f

    def kkt_ad(self, z, s, lam, u):
        n = 6 * self.nb + self.ml
        M = np.zeros((n, n))
        lib().orc_kkt_ad(self.h, _p(_f64(z)), _p(_f64(s)), _p(_f64(lam)), _p(_f64(u)), _p(M))
        return M

    def solve_both(self, z, s, lam, u):
        n = 6 * self.nb + self.ml
        a = np.zeros(n); b = np.zeros(n)
        rc = lib().orc_solve_both(self.h, _p(_f64(z)), _p(_f64(s)), _p(_f64(lam)), _p(_f64(u)), _p(a), _p(b))
        return rc, a, b

    # ---- controller struct
    def make_ctrl(self, B, kind="none", eqc=None, N=0, K=None, zd=None, Fd=None, U=None, P=None, I=None, D=None, goal=None,
                  pid_int=None, pid_last=None):
        c = OrcCtrl()
        keep = []

        def arr(a, dt=np.float64):
            if a is None:
                return None, None
            a = np.ascontiguousarray(a, dtype=dt)
            keep.append(a)
            return a, _p(a)

        c.kind = KIND[kind]
        eqc_a, c.eqc = arr(eqc, np.int32)
        c.nu = 0 if eqc_a is None else int(eqc_a.size)
        c.N = int(N)
        nu, n = c.nu, self.n

        def with_stride(a, per_system_shape):
            if a is None:
                return None, 0
            a = np.ascontiguousarray(a, dtype=np.float64)
            per = int(np.prod(per_system_shape))
            if a.size == per:
                stride = 0
            else:
                assert a.size == per * B, (a.shape, per_system_shape, B)
                stride = per
            keep.append(a)
            return _p(a), stride

        if kind in ("lqr_inf", "lqr_finite", "tracking"):
            rows = 1 if kind == "lqr_inf" else N - 1
            c.K, c.K_bstride = with_stride(K, (rows, nu, n))
            if kind == "tracking":
                c.zd, c.zd_bstride = with_stride(zd, (N, self.nb, 13))
                c.Fd, c.Fd_bstride = with_stride(Fd, (N, nu))
            else:
                c.zd, c.zd_bstride = with_stride(zd, (self.nb, 13))
                c.Fd, c.Fd_bstride = with_stride(Fd, (nu,))
        if kind == "table":
            U = np.ascontiguousarray(U, dtype=np.float64)
            steps = U.shape[-2] if U.ndim >= 2 else U.size // max(nu, 1)
            c.U, c.U_bstride = with_stride(U, (steps, nu))
        if kind == "pid":
            _, c.P = arr(P); _, c.I = arr(I); _, c.D = arr(D)
            c.goal, c.goal_bstride = with_stride(goal, (nu,))
            self_int = np.zeros((B, nu)) if pid_int is None else pid_int
            self_last = np.zeros((B, nu)) if pid_last is None else pid_last
            assert self_int.dtype == np.float64 and self_int.flags.c_contiguous
            keep += [self_int, self_last]
            c.pid_int, c.pid_last = _p(self_int), _p(self_last)
        c._keep = keep
        return c

    # ---- simulate!
    def simulate(self, states, steps, ctrl=None, lambdas=None, k0=1, eps=1e-10, newton_iter=100, line_iter=10, solver=1, record=False,
                 nthreads=1, marginal_band=0.0):
        states = np.array(states, dtype=np.float64, copy=True).reshape(-1, self.nb, 13)
        B = states.shape[0]
        lambdas = np.zeros((B, self.ml)) if lambdas is None else np.array(lambdas, dtype=np.float64, copy=True).reshape(B, self.ml)
        storage = np.zeros((steps, B, self.nb, 13)) if record else None
        iters = np.zeros(B, dtype=np.int32)
        marginal = np.zeros(B, dtype=np.int32)
        if marginal_band > 1.0:      # count the Newton iterations whose termination decision is rounding-marginal (oracle.cpp: newton)
            lib().orc_set_marginal(C.c_double(marginal_band), _p(marginal))
        lib().orc_simulate(self.h, C.c_int(B), _p(states), _p(lambdas), C.c_int(k0), C.c_int(steps), C.byref(ctrl) if ctrl is not None else None,
                           C.c_double(eps), C.c_int(newton_iter), C.c_int(line_iter), C.c_int(solver), _p(storage), _p(iters), C.c_int(nthreads))
        if marginal_band > 1.0:
            lib().orc_set_marginal(C.c_double(0.0), None)
        return dict(states=states, lambdas=lambdas, storage=storage, iters=iters, marginal=marginal)

    def count_flops(self, state, lam=None, u=None, eps=1e-10, newton_iter=100, line_iter=10):
        """exact floating-point operation count of ONE literal newton! solve at `state` [nb][13] (oracle.cpp: orc_count_flops)"""
        out = np.zeros(6, dtype=np.int64)
        lam = np.zeros(self.ml) if lam is None else _f64(lam)
        u = np.zeros(self.ne) if u is None else _f64(u)
        lib().orc_count_flops(self.h, _p(_f64(state)), _p(lam), _p(u), C.c_double(eps), C.c_int(newton_iter), C.c_int(line_iter), _p(out))
        add, mul, div, sq, cmp_, it = (int(v) for v in out)
        return dict(add=add, mul=mul, div=div, sqrt=sq, compare=cmp_, iterations=it, flops=add + mul + div + sq,
                    flops_per_iteration=(add + mul + div + sq) / max(it, 1))

    def count_flops_rollout(self, state, steps, ctrl=None, sys=0, lam=None, k0=1, eps=1e-10, newton_iter=100, line_iter=10):
        """exact operation count of the literal algorithm over a closed-loop rollout of ONE system (oracle.cpp: orc_count_flops_rollout);
        returns the counts, the final state (bit-identical to simulate()) and flops per sim step / per Newton iteration"""
        st = np.array(state, dtype=np.float64, copy=True).reshape(self.nb, 13)
        lam = np.zeros(self.ml) if lam is None else np.array(lam, dtype=np.float64, copy=True)
        out = np.zeros(6, dtype=np.int64)
        rc = lib().orc_count_flops_rollout(self.h, _p(st), _p(lam), C.byref(ctrl) if ctrl is not None else None, C.c_int(sys), C.c_int(k0), C.c_int(steps),
                                           C.c_double(eps), C.c_int(newton_iter), C.c_int(line_iter), _p(out))
        if rc != 0:
            raise RuntimeError(f"orc_count_flops_rollout: Newton failure {rc}")
        add, mul, div, sq, cmp_, it = (int(v) for v in out)
        fl = add + mul + div + sq
        return dict(add=add, mul=mul, div=div, sqrt=sq, compare=cmp_, iterations=it, flops=fl, state=st, lam=lam,
                    flops_per_step=fl / steps, flops_per_iteration=fl / max(it, 1), iterations_per_step=it / steps)

    def control(self, ctrl, sys, k, state):
        u = np.zeros(self.ne)
        lib().orc_control(self.h, C.byref(ctrl), C.c_int(sys), C.c_int(k), _p(_f64(state)), _p(u))
        return u

    # ---- linearsystem / dlqr
    def linearize(self, zd, Fd, eqc_idx, lam0=None, eps=1e-10, newton_iter=100, line_iter=10):
        n, ml = self.n, self.ml
        eqc_idx = _i32(eqc_idx); nu = eqc_idx.size
        A = np.zeros((n, n)); Bu = np.zeros((n, nu)); Bl = np.zeros((n, ml)); G = np.zeros((ml, n))
        lam = np.zeros(ml) if lam0 is None else np.array(lam0, dtype=np.float64, copy=True)
        it = lib().orc_linearize(self.h, _p(_f64(zd)), _p(_f64(Fd)), _p(eqc_idx), C.c_int(nu), _p(lam), C.c_double(eps), C.c_int(newton_iter),
                                 C.c_int(line_iter), _p(A), _p(Bu), _p(Bl), _p(G))
        return dict(A=A, Bu=Bu, Bl=Bl, G=G, lam=lam, iters=it)

    def tracking_dlqr(self, traj, Ftau, eqc_idx, Q, R, eps=1e-10, newton_iter=100, line_iter=10, prec=None):
        """prec None: the literal Float64 loop; 0/1/2: orc_tracking_dlqr_mp (Float64 linearisations, recursion in Float64 / x87 / binary128)"""
        traj = _f64(traj).reshape(-1, self.nb, 13)
        N = traj.shape[0]
        eqc_idx = _i32(eqc_idx); nu = eqc_idx.size
        Ku = np.zeros((N - 1, nu, self.n)); iters = C.c_int(0)
        if prec is not None:
            k = lib().orc_tracking_dlqr_mp(C.c_int(prec), self.h, _p(traj), _p(_f64(Ftau)), _p(eqc_idx), C.c_int(nu), _p(_f64(Q)), _p(_f64(R)), C.c_int(N),
                                           C.c_double(eps), C.c_int(newton_iter), C.c_int(line_iter), _p(Ku), C.byref(iters))
            return dict(Ku=Ku, kexit=k, iters=iters.value)
        k = lib().orc_tracking_dlqr(self.h, _p(traj), _p(_f64(Ftau)), _p(eqc_idx), C.c_int(nu), _p(_f64(Q)), _p(_f64(R)), C.c_int(N),
                                    C.c_double(eps), C.c_int(newton_iter), C.c_int(line_iter), _p(Ku), C.byref(iters))
        return dict(Ku=Ku, kexit=k, iters=iters.value)


def dlqr(A, Bu, Bl, G, Q, R, N):
    """dlqr(A,Bu,Bλ,G,Q,R,N)  src/control/lqr.jl:141-184 -> Ku [N-1][mu][n], exit k, iterations, last Pk"""
    A = _f64(A); n = A.shape[0]
    Bu = _f64(Bu).reshape(n, -1); mu = Bu.shape[1]
    Bl = _f64(Bl).reshape(n, -1); ml = Bl.shape[1]
    G = _f64(G).reshape(ml, n)
    Ku = np.zeros((N - 1, mu, n)); P = np.zeros((n, n)); iters = C.c_int(0)
    k = lib().orc_dlqr(_p(A), _p(Bu), _p(Bl), _p(G), _p(_f64(Q)), _p(_f64(R)), C.c_int(n), C.c_int(mu), C.c_int(ml), C.c_int(N), _p(Ku),
                       C.byref(iters), _p(P))
    return dict(Ku=Ku, kexit=k, iters=iters.value, P=P)


def dlqr_mp(A, Bu, Bl, G, Q, R, N, prec=2):
    """dlqr (lqr.jl:141-184) with the recursion carried in prec 0 = Float64, 1 = long double (x87), 2 = binary128; Float64 inputs / outputs.
    Also returns the trace of ||Pk - Pkp1||_F (the early-break quantity of lqr.jl:172) and max |P| at exit."""
    A = _f64(A); n = A.shape[0]
    Bu = _f64(Bu).reshape(n, -1); mu = Bu.shape[1]
    Bl = _f64(Bl).reshape(n, -1); ml = Bl.shape[1]
    G = _f64(G).reshape(ml, n)
    Ku = np.zeros((N - 1, mu, n)); trace = np.zeros(N - 1); iters = C.c_int(0); pmax = C.c_double(0)
    k = lib().orc_dlqr_mp(C.c_int(prec), _p(A), _p(Bu), _p(Bl), _p(G), _p(_f64(Q)), _p(_f64(R)), C.c_int(n), C.c_int(mu), C.c_int(ml), C.c_int(N),
                          _p(Ku), C.byref(iters), _p(trace), C.byref(pmax))
    if k < 0:
        raise RuntimeError("orc_dlqr_mp: precision not built")
    return dict(Ku=Ku, kexit=k, iters=iters.value, trace=trace[:iters.value], Pmax=pmax.value)


def dlqr_batch(A, Bu, Bl, G, Q, R, N, nthreads=1):
    A = _f64(A); B, n = A.shape[0], A.shape[1]
    Bu = _f64(Bu).reshape(B, n, -1); mu = Bu.shape[2]
    Bl = _f64(Bl).reshape(B, n, -1); ml = Bl.shape[2]
    G = _f64(G).reshape(B, ml, n)
    Ku = np.zeros((B, N - 1, mu, n)); kexit = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
    lib().orc_dlqr_batch(C.c_int(B), _p(A), _p(Bu), _p(Bl), _p(G), _p(_f64(Q)), _p(_f64(R)), C.c_int(n), C.c_int(mu), C.c_int(ml), C.c_int(N),
                         _p(Ku), _p(kexit), _p(iters), C.c_int(nthreads))
    return dict(Ku=Ku, kexit=kexit, iters=iters)
