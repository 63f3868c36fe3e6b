"""Host-side mirror of the reference's controller API on top of the C-ABI (include/ccb.h).

Reference surface kept (names, argument meaning, error behaviour):
    LQR(mechanism, bodyids, eqcids, Q, R, horizon; xd, vd, qd, ωd, Fτd)     src/control/lqr.jl:49-66 (+ inner ctor :17-47)
    TrackingLQR(mechanism, storage, Fτ, eqcids, Q, R)                        src/control/lqr_tracking.jl:17-43
    PID(mechanism, eqcids, goals; P, I, D)                                   src/control/pid.jl:16-39
    simulate!(mechanism, tend | storage, controller | control!; record)      examples/lqr_pendulum.jl:43
Every one of them accepts a *batch*: per-system setpoints / weights / goals / initial states, and runs on the
GPU through libccb.so.  No CPU path exists here: without the library or a device these raise `CcbError`.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import CcbError, check, f64, i32, ptr
from .mechanism import Mechanism

CTRL_NONE, CTRL_TABLE, CTRL_LQR_INF, CTRL_LQR_FIN, CTRL_TRACKING, CTRL_PID = range(6)


# ----------------------------------------------------------------------------------------------- handles
class MechHandle:
    """ccb_mech handle (graph orderings + device tables) of one Mechanism"""

    def __init__(self, mech: Mechanism):
        L = _lib.lib()
        d = mech.desc()
        self._keep = {k: (f64(v) if v.dtype.kind == "f" else i32(v)) for k, v in d.items() if isinstance(v, np.ndarray)}
        md = _lib.MechDesc()
        md.nb, md.ne, md.nj, md.id_scheme = d["nb"], d["ne"], d["nj"], d["id_scheme"]
        md.dt, md.g, md.reg = d["dt"], d["g"], d["reg"]
        for k in ("mass", "inertia", "eqc_parent", "eqc_joint_ptr", "joint_child", "joint_nT", "joint_nR", "joint_axis", "joint_p1", "joint_p2",
                  "joint_qoff"):
            setattr(md, k, ptr(self._keep[k]))
        self.h = C.c_void_p()
        check(L.ccb_mech_create(C.byref(md), C.byref(self.h)))
        dims = np.zeros(8, dtype=np.int32)
        check(L.ccb_mech_dims(self.h, ptr(dims)))
        self.nb, self.ne, self.nj, self.ml, self.nnodes, self.nloop, self.n, self.lanes = [int(v) for v in dims]

    def graph(self) -> dict:
        L = _lib.lib()
        out = dict(eqc_id=np.zeros(self.ne, np.int32), body_id=np.zeros(self.nb, np.int32), dfs_list=np.zeros(self.nnodes, np.int32),
                   parents=np.zeros(self.nnodes, np.int32), loop_edges=np.zeros(2 * max(self.nloop, 1), np.int32),
                   eqc_dim=np.zeros(self.ne, np.int32), eqc_row0=np.zeros(self.ne, np.int32), act_kind=np.zeros(self.ne, np.int32),
                   joint_elim=np.zeros(self.nj, np.int32))
        check(L.ccb_mech_graph(self.h, *[ptr(out[k]) for k in ("eqc_id", "body_id", "dfs_list", "parents", "loop_edges", "eqc_dim", "eqc_row0",
                                                               "act_kind", "joint_elim")]))
        out["loop_edges"] = out["loop_edges"][: 2 * self.nloop].reshape(-1, 2)
        return out

    def __del__(self):
        try:
            if self.h:
                _lib.lib().ccb_mech_destroy(self.h)
                self.h = None
        except Exception:
            pass


def handle(mech: Mechanism) -> MechHandle:
    h = getattr(mech, "_ccb_handle", None)
    if h is None:
        h = MechHandle(mech)
        mech._ccb_handle = h
    return h


class Batch:
    """B copies of one mechanism resident on one device (ccb_batch)"""

    def __init__(self, mech: Mechanism, B: int, device: int = 0):
        self.mech, self.B, self.device = mech, int(B), int(device)
        self.mh = handle(mech)
        self.h = C.c_void_p()
        check(_lib.lib().ccb_batch_create(self.mh.h, C.c_int32(self.B), C.c_int32(self.device), C.byref(self.h)))

    def set_state(self, state, lam=None):
        state = f64(state).reshape(self.B, self.mh.nb, 13)
        lam = None if lam is None else f64(lam).reshape(self.B, self.mh.ml)
        check(_lib.lib().ccb_batch_set_state(self.h, ptr(state), ptr(lam)))

    def get_state(self, want_lambda: bool = False):
        st = np.empty((self.B, self.mh.nb, 13))
        lam = np.empty((self.B, self.mh.ml)) if want_lambda else None
        check(_lib.lib().ccb_batch_get_state(self.h, ptr(st), ptr(lam)))
        return (st, lam) if want_lambda else st

    def device_ptrs(self):
        a, b = C.c_void_p(), C.c_void_p()
        check(_lib.lib().ccb_batch_device_ptrs(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stream(self) -> int:
        s = C.c_void_p()
        check(_lib.lib().ccb_batch_stream(self.h, C.byref(s)))
        return s.value or 0

    def sync(self):
        check(_lib.lib().ccb_batch_sync(self.h))

    def simulate(self, ctrl=None, k0: int = 1, steps: int = 1, record: bool = False, eps: float = 1e-10, newton_iter: int = 100,
                 line_iter: int = 10, lanes: int = 0, want_status: bool = True, spw: int = 0):
        opts = _lib.SimOpts(eps, newton_iter, line_iter, lanes, spw)
        storage = np.empty((steps, self.B, self.mh.nb, 13)) if record else None
        status = np.empty(self.B, dtype=np.int32) if want_status else None
        ch = ctrl.device_handle(self.B, self.device) if ctrl is not None else None
        check(_lib.lib().ccb_simulate(self.h, ch, C.c_int32(k0), C.c_int32(steps), C.byref(opts), ptr(storage), ptr(status)))
        return storage, status

    def simulate_device(self, ctrl=None, k0: int = 1, steps: int = 1, eps: float = 1e-10, newton_iter: int = 100, line_iter: int = 10,
                        lanes: int = 0, storage_ptr: int = 0, status_ptr: int = 0, spw: int = 0):
        """asynchronous, device-resident variant (used by bench.py's kernel-only timing)"""
        opts = _lib.SimOpts(eps, newton_iter, line_iter, lanes, spw)
        ch = ctrl.device_handle(self.B, self.device) if ctrl is not None else None
        check(_lib.lib().ccb_simulate_device(self.h, ch, C.c_int32(k0), C.c_int32(steps), C.byref(opts), C.c_void_p(storage_ptr or None),
                                             C.c_void_p(status_ptr or None)))

    def __del__(self):
        try:
            if self.h:
                _lib.lib().ccb_batch_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------- Storage
class Storage:
    """Storage{T,N}: trajectory record, `storage.x[i][k]` body-major then step (lqr_tracking.jl:32-35).
    Backed by one array data[N][B][nb][13]; `x[i][k]` etc. are views for system `sys` (default 0)."""

    def __init__(self, steps: int, nb: int, B: int = 1, data: Optional[np.ndarray] = None):
        self.N, self.nb, self.B = int(steps), int(nb), int(B)
        self.data = np.zeros((self.N, self.B, self.nb, 13)) if data is None else data

    def _field(self, lo, hi, sys=0):
        return [self.data[:, sys, i, lo:hi] for i in range(self.nb)]

    @property
    def x(self):
        return self._field(0, 3)

    @property
    def q(self):
        return self._field(3, 7)

    @property
    def v(self):
        return self._field(7, 10)

    @property
    def w(self):
        return self._field(10, 13)

    def __len__(self):
        return self.N


# ----------------------------------------------------------------------------------------------- controllers
class Controller:
    """`Controller` supertype: objects carrying the parameters of an enumerated device control law."""

    kind = CTRL_NONE

    def __init__(self):
        self._dev = {}

    def _desc(self, B: int):
        raise NotImplementedError

    def device_handle(self, B: int, device: int):
        key = (B, device)
        if key not in self._dev:
            d, keep = self._desc(B)
            d.B = B
            h = C.c_void_p()
            check(_lib.lib().ccb_ctrl_create(handle(self.mech).h, C.byref(d), C.c_int32(device), C.byref(h)))
            self._dev[key] = (h, keep)
        return self._dev[key][0]

    def replicate(self, per_table: int):
        """let `per_table` consecutive systems of the batch share one gain table / desired trajectory (device-resident controllers)"""
        nat = getattr(self, "_native", None)
        if nat is None:
            raise CcbError("replicate() needs a controller synthesised on the device")
        check(_lib.lib().ccb_ctrl_set_replication(nat[0], C.c_int32(per_table)))
        self._per_table = int(per_table)
        return self

    def invalidate(self):
        for h, _ in self._dev.values():
            _lib.lib().ccb_ctrl_destroy(h)
        self._dev = {}

    def __del__(self):
        try:
            self.invalidate()
        except Exception:
            pass


class TableLQR(Controller):
    """finite-horizon LQR (control_lqr!, lqr.jl:89-114) over a device-resident gain table filled by dist.PeerGainTable"""

    kind = CTRL_LQR_FIN

    def __init__(self, table):
        super().__init__()
        self.mech, self.table = table.mech, table
        self._h = {}

    def device_handle(self, B: int, device: int):
        t = self.table
        if B != t.S or device != t.device:
            raise CcbError(f"TableLQR: the table holds {t.S} systems on device {t.device}")
        if device not in self._h:
            h = C.c_void_p()
            check(_lib.lib().ccb_lqr_ctrl_from_table(handle(self.mech).h, C.c_int32(device), C.c_int32(t.S), C.c_int32(t.N), C.c_int32(t.nu), ptr(t.eqc_idx),
                                                    t.ptr, ptr(t.zd_all), ptr(t.Fd_all), C.byref(h)))
            self._h[device] = h
        return self._h[device]

    def invalidate(self):
        for h in self._h.values():
            _lib.lib().ccb_ctrl_destroy(h)
        self._h = {}


class TableTracking(TableLQR):
    """TrackingLQR (control_trackinglqr!, lqr_tracking.jl:46-71) over a device-resident gain table filled by dist.PeerTrackingTable"""

    kind = CTRL_TRACKING

    def device_handle(self, B: int, device: int):
        t = self.table
        per = getattr(self, "_per_table", 1)
        if B != t.S * per or device != t.device:
            raise CcbError(f"TableTracking: the table holds {t.S} trajectories (x {per} systems each) on device {t.device}")
        if device not in self._h:
            h = C.c_void_p()
            check(_lib.lib().ccb_tracking_ctrl_from_table(handle(self.mech).h, C.c_int32(device), C.c_int32(t.S), C.c_int32(t.N), C.c_int32(t.nu), ptr(t.eqc_idx),
                                                         t.ptr, ptr(t.traj), ptr(t.Ftau), C.byref(h)))
            self._h[device] = h
            if per > 1:
                check(_lib.lib().ccb_ctrl_set_replication(h, C.c_int32(per)))
        return self._h[device]

    def replicate(self, per_table: int):
        """let `per_table` consecutive systems of the batch share one trajectory's tables"""
        self._per_table = int(per_table)
        for h in self._h.values():
            check(_lib.lib().ccb_ctrl_set_replication(h, C.c_int32(per_table)))
        return self


def _stride(a: np.ndarray, per: int, B: int) -> int:
    if a.size == per:
        return 0
    if a.size != per * B:
        raise CcbError(f"controller table has {a.size} entries, expected {per} (shared) or {per}*{B} (per system)")
    return per


def _zd_array(nb, xd, vd, qd, wd):
    """(xd,vd,qd,ωd) lists -> [..., nb, 13] with the reference's defaults (lqr.jl:51-54)"""
    def arr(v, k, default):
        if v is None:
            return None
        return f64(v).reshape(-1, nb, k)
    parts = [arr(xd, 3, 0), arr(qd, 4, 0), arr(vd, 3, 0), arr(wd, 3, 0)]
    S = max([p.shape[0] for p in parts if p is not None] + [1])
    zd = np.zeros((S, nb, 13))
    zd[:, :, 3] = 1.0
    for p, (lo, hi) in zip(parts, ((0, 3), (3, 7), (7, 10), (10, 13))):
        if p is not None:
            zd[:, :, lo:hi] = p
    return zd


def blockdiag(mats: Sequence[np.ndarray]) -> np.ndarray:
    """cat(Q..., dims=(1,2))  (lqr.jl:18)"""
    mats = [np.atleast_2d(f64(m)) for m in mats]
    n = sum(m.shape[0] for m in mats)
    out = np.zeros((n, n))
    o = 0
    for m in mats:
        out[o:o + m.shape[0], o:o + m.shape[0]] = m
        o += m.shape[0]
    return out


class LQR(Controller):
    """LQR(mechanism, bodyids, eqcids, Q, R, horizon; xd, vd, qd, ωd, Fτd)   (reference src/control/lqr.jl:17-66).

    Batched: xd/vd/qd/ωd/Fτd may carry a leading batch dimension S (one setpoint per system) and Q/R may be lists
    of length S of per-body block lists; the gains are synthesised on the GPU (ccb_lqr_synthesize).
    `LQR.from_gains` wraps gains that already exist (K [rows][nu][12 nb] or [S][rows][nu][12 nb])."""

    def __init__(self, mechanism: Mechanism, bodyids, eqcids, Q, R, horizon, xd=None, vd=None, qd=None, wd=None, Ftaud=None, device: int = 0,
                 opts: Optional[dict] = None, Q_batch=None, R_batch=None, xtheta_d=None, vomega_d=None):
        """Q_batch [S][12 nb][12 nb] / R_batch [S][nu][nu] (optional): one assembled weight matrix per system instead of the
        shared block lists Q / R (a synthesis batch; they are scaled by Δt like lqr.jl:18-19)"""
        super().__init__()
        mech = mechanism
        nb = mech.nb
        if Q_batch is None and Q is not None and len(Q) > 0 and all(np.ndim(q) == 0 for q in Q):
            # minimal-coordinate constructor LQR(mechanism, controlledids, controlids, Q::Vector{T}, R::Vector{T}, horizon; xθd, vωd, Fτd)
            # (lqr.jl:68-86): `bodyids` are the controlled eqcs (one 1-DoF joint per body), the setpoint is given in joint coordinates
            controlledids, controlids = list(bodyids), list(eqcids)
            xth = np.zeros(len(controlledids)) if xtheta_d is None else f64(xtheta_d).ravel()
            vom = np.zeros(len(controlledids)) if vomega_d is None else f64(vomega_d).ravel()
            Fta = np.zeros(len(controlids)) if Ftaud is None else f64(Ftaud).ravel()
            if not (len(controlledids) == len(Q) == len(xth) == len(vom) == nb):          # lqr.jl:76
                raise AssertionError("Missmatched length for bodies")
            if not (len(controlids) == len(R) == len(Fta)):                               # lqr.jl:77
                raise AssertionError("Missmatched length for constraints")
            from .mechanism import propagate_minimal
            saved = mech.state_vector().copy()
            idx = [mech.eqc_index(i) for i in controlledids]
            for e in idx:
                if not mech.is_1dof(int(e)):
                    raise AssertionError("Only 1 DOF joints are supported")
            propagate_minimal(mech, {e: xth[k] for k, e in enumerate(idx)}, {e: vom[k] for k, e in enumerate(idx)})   # linearsystem(…6 args) lqr.jl:80
            zdm = mech.state_vector().copy()
            mech.set_state_vector(saved)
            Q = [np.eye(12) * float(q) for q in Q]                                        # lqr.jl:82
            R = [np.eye(1) * float(r) for r in R]                                         # lqr.jl:83
            xd, qd, vd, wd = zdm[:, 0:3], zdm[:, 3:7], zdm[:, 7:10], zdm[:, 10:13]
            bodyids, eqcids, Ftaud = [b.id for b in mech.bodies], controlids, Fta
        # lqr.jl:59-60
        if not (len(bodyids) == nb and (Q_batch is not None or len(Q) == nb)):
            raise AssertionError("Missmatched length for bodies")
        if not (R_batch is not None or len(eqcids) == len(R)) or (Ftaud is not None and np.asarray(Ftaud).reshape(-1, len(eqcids)).shape[-1] != len(eqcids)):
            raise AssertionError("Missmatched length for constraints")
        if [mech.body_index(i) for i in bodyids] != list(range(nb)):
            raise CcbError("bodyids must list the mechanism's bodies in order (Δz is indexed by body position, lqr.jl:91)")
        self.mech = mech
        self.eqcids = list(eqcids)
        self.eqc_idx = i32([mech.eqc_index(i) for i in eqcids])
        nu = len(eqcids)
        zd = _zd_array(nb, xd, vd, qd, wd)
        Fd = np.zeros((1, nu)) if Ftaud is None else f64(Ftaud).reshape(-1, nu)
        S = max(zd.shape[0], Fd.shape[0], 1 if Q_batch is None else len(Q_batch), 1 if R_batch is None else len(R_batch))
        zd = np.broadcast_to(zd, (S, nb, 13)).copy()
        Fd = np.broadcast_to(Fd, (S, nu)).copy()
        dt = mech.dt
        Qd = blockdiag(Q) * dt if Q_batch is None else f64(Q_batch) * dt            # lqr.jl:18
        Rd = blockdiag(R) * dt if R_batch is None else f64(R_batch) * dt            # lqr.jl:19
        N = horizon / dt                  # lqr.jl:21-27
        if N < math.inf:
            N = int(math.ceil(horizon / dt))
            Ntemp = N
        else:
            Ntemp = int(math.ceil(10 / dt))
        self.N = N
        # synthesis + controller tables in one C-ABI call: the gain table stays on the device (ccb_lqr_create)
        from .synth import _opts, _weights
        mh = handle(mech)
        Qw, qs = _weights(Qd, S, mh.n)
        Rw, rs = _weights(Rd, S, nu)
        self.kexit = np.empty(S, np.int32); self.riccati_iters = np.empty(S, np.int32); conv = np.empty(S, np.int32)
        o = _opts(**(opts or {}))
        h = C.c_void_p()
        check(_lib.lib().ccb_lqr_create(mh.h, C.c_int32(device), C.c_int32(S), ptr(zd), ptr(Fd), ptr(self.eqc_idx), C.c_int32(nu), ptr(Qw), C.c_int64(qs),
                                        ptr(Rw), C.c_int64(rs), C.c_int32(Ntemp), C.c_int32(1 if N == math.inf else 0), C.byref(o), C.byref(h),
                                        ptr(self.kexit), ptr(self.riccati_iters), ptr(conv)))
        self._native = (h, S, device)
        self.converged = conv.astype(bool)          # lqr.jl:41 (Ku[1] == Ku[2]); only meaningful for horizon = Inf
        self._K = None
        self._rows = 1 if N == math.inf else Ntemp - 1
        self.zd, self.Fd = zd, Fd
        self.kind = CTRL_LQR_INF if N == math.inf else CTRL_LQR_FIN

    @property
    def K(self):
        """gain table [S][rows][nu][12 nb]; fetched from the device on first use"""
        if self._K is None:
            h, S, _ = self._native
            n, nu = 12 * self.mech.nb, len(self.eqcids)
            K = np.empty((S, self._rows, nu, n))
            check(_lib.lib().ccb_ctrl_get_gains(h, C.c_int32(n), ptr(K), C.c_int64(K.size)))
            self._K = K
        return self._K

    @K.setter
    def K(self, v):
        self._K = v

    def device_handle(self, B: int, device: int):
        nat = getattr(self, "_native", None)
        if nat is not None and nat[2] == device and (nat[1] == 1 or nat[1] * getattr(self, "_per_table", 1) == B):
            return nat[0]
        return super().device_handle(B, device)

    def invalidate(self):
        nat = getattr(self, "_native", None)
        if nat is not None:
            _lib.lib().ccb_ctrl_destroy(nat[0])
            self._native = None
        super().invalidate()

    @classmethod
    def from_gains(cls, mechanism: Mechanism, eqcids, K, horizon_steps, xd=None, vd=None, qd=None, wd=None, Ftaud=None):
        self = cls.__new__(cls)
        Controller.__init__(self)
        nb, nu = mechanism.nb, len(eqcids)
        self.mech, self.eqcids = mechanism, list(eqcids)
        self.eqc_idx = i32([mechanism.eqc_index(i) for i in eqcids])
        self.N = horizon_steps
        self._native = None
        rows = 1 if horizon_steps == math.inf else int(horizon_steps) - 1
        self.K = f64(K).reshape(-1, rows, nu, 12 * nb)
        self.zd = _zd_array(nb, xd, vd, qd, wd)
        self.Fd = np.zeros((1, nu)) if Ftaud is None else f64(Ftaud).reshape(-1, nu)
        self.kind = CTRL_LQR_INF if horizon_steps == math.inf else CTRL_LQR_FIN
        return self

    def _desc(self, B):
        nb, nu, n = self.mech.nb, len(self.eqcids), 12 * self.mech.nb
        rows = self.K.shape[1]
        d = _lib.CtrlDesc()
        K, zd, Fd = f64(self.K), f64(self.zd), f64(self.Fd)
        d.kind, d.nu, d.N = self.kind, nu, (0 if self.N == math.inf else int(self.N))
        d.eqc = ptr(self.eqc_idx)
        d.K, d.K_bstride = ptr(K), _stride(K, rows * nu * n, B)
        d.zd, d.zd_bstride = ptr(zd), _stride(zd, nb * 13, B)
        d.Fd, d.Fd_bstride = ptr(Fd), _stride(Fd, nu, B)
        return d, (K, zd, Fd)


class TrackingLQR(Controller):
    """TrackingLQR(mechanism, storage, Fτ, eqcids, Q, R)   (reference src/control/lqr_tracking.jl:17-43)"""

    def __init__(self, mechanism: Mechanism, storage, Ftau, eqcids, Q, R, device: int = 0, opts: Optional[dict] = None):
        super().__init__()
        mech = mechanism
        self.mech, self.eqcids = mech, list(eqcids)
        self.eqc_idx = i32([mech.eqc_index(i) for i in eqcids])
        nu, nb = len(eqcids), mech.nb
        data = storage.data if isinstance(storage, Storage) else f64(storage)
        N = data.shape[0]
        traj = np.ascontiguousarray(np.moveaxis(data.reshape(N, -1, nb, 13), 1, 0))   # [S][N][nb][13]  (lqr_tracking.jl:30-37)
        S = traj.shape[0]
        Ft = f64(Ftau).reshape(-1, N, nu)
        Ft = np.broadcast_to(Ft, (S, N, nu)).copy()
        Qd = blockdiag(Q) * mech.dt       # lqr_tracking.jl:22-23
        Rd = blockdiag(R) * mech.dt
        # synthesis + controller tables in one C-ABI call: gains, desired trajectory and feed-forward stay on the device
        from .synth import _opts
        mh = handle(mech)
        self.kexit = np.empty(S, np.int32); self.riccati_iters = np.empty(S, np.int32)
        o = _opts(**(opts or {}))
        h = C.c_void_p()
        Qd, Rd = f64(Qd).reshape(mh.n, mh.n), f64(Rd).reshape(nu, nu)
        check(_lib.lib().ccb_tracking_create(mh.h, C.c_int32(device), C.c_int32(S), C.c_int32(N), ptr(traj), ptr(Ft), ptr(self.eqc_idx), C.c_int32(nu),
                                             ptr(Qd), ptr(Rd), C.byref(o), C.byref(h), ptr(self.kexit), ptr(self.riccati_iters)))
        self._native = (h, S, device)
        self._K = None
        self.N, self.zd, self.Fd = N, traj, Ft
        self.kind = CTRL_TRACKING

    @property
    def K(self):
        """gain tables [S][N-1][nu][12 nb]; fetched from the device on first use"""
        if self._K is None:
            h, S, _ = self._native
            n, nu = 12 * self.mech.nb, len(self.eqcids)
            K = np.empty((S, self.N - 1, nu, n))
            check(_lib.lib().ccb_ctrl_get_gains(h, C.c_int32(n), ptr(K), C.c_int64(K.size)))
            self._K = K
        return self._K

    @K.setter
    def K(self, v):
        self._K = v

    def device_handle(self, B: int, device: int):
        nat = getattr(self, "_native", None)
        if nat is not None and nat[2] == device and (nat[1] == 1 or nat[1] * getattr(self, "_per_table", 1) == B):
            return nat[0]
        return super().device_handle(B, device)

    def invalidate(self):
        nat = getattr(self, "_native", None)
        if nat is not None:
            _lib.lib().ccb_ctrl_destroy(nat[0])
            self._native = None
        super().invalidate()

    @classmethod
    def from_gains(cls, mechanism: Mechanism, eqcids, K, traj, Ftau):
        self = cls.__new__(cls)
        Controller.__init__(self)
        nb, nu = mechanism.nb, len(eqcids)
        self.mech, self.eqcids = mechanism, list(eqcids)
        self.eqc_idx = i32([mechanism.eqc_index(i) for i in eqcids])
        traj = f64(traj)
        traj = traj.reshape(-1, traj.shape[-3], nb, 13)
        self._native = None
        self.N = traj.shape[1]
        self.zd = traj
        self.Fd = f64(Ftau).reshape(-1, self.N, nu)
        self.K = f64(K).reshape(-1, self.N - 1, nu, 12 * nb)
        self.kind = CTRL_TRACKING
        return self

    def _desc(self, B):
        nb, nu, n, N = self.mech.nb, len(self.eqcids), 12 * self.mech.nb, self.N
        d = _lib.CtrlDesc()
        K, zd, Fd = f64(self.K), f64(self.zd), f64(self.Fd)
        d.kind, d.nu, d.N = CTRL_TRACKING, nu, N
        d.eqc = ptr(self.eqc_idx)
        d.K, d.K_bstride = ptr(K), _stride(K, (N - 1) * nu * n, B)
        d.zd, d.zd_bstride = ptr(zd), _stride(zd, N * nb * 13, B)
        d.Fd, d.Fd_bstride = ptr(Fd), _stride(Fd, N * nu, B)
        return d, (K, zd, Fd)


class PID(Controller):
    """PID(mechanism, eqcids, goals; P, I, D)   (reference src/control/pid.jl:16-39); goals may be [B][N] (one row per system)"""

    kind = CTRL_PID

    def __init__(self, mechanism: Mechanism, eqcids, goals, P=None, I=None, D=None):
        super().__init__()
        mech = mechanism
        scalar = np.isscalar(eqcids)
        eqcids = [eqcids] if scalar else list(eqcids)
        nu = len(eqcids)
        self.mech, self.eqcids = mech, eqcids
        self.eqc_idx = i32([mech.eqc_index(i) for i in eqcids])
        for e in self.eqc_idx:   # pid.jl:20,36
            if not mech.is_1dof(int(e)):
                raise AssertionError("Only 1 DOF joints are supported")
        z = np.zeros(nu)
        self.P, self.I, self.D = (f64(z if v is None else np.atleast_1d(v)) for v in (P, I, D))
        self.goals = f64(np.atleast_1d(goals)).reshape(-1, nu)

    def _desc(self, B):
        nu = len(self.eqcids)
        d = _lib.CtrlDesc()
        goals = f64(self.goals)
        d.kind, d.nu, d.N = CTRL_PID, nu, 0
        d.eqc = ptr(self.eqc_idx)
        d.P, d.I, d.D = ptr(self.P), ptr(self.I), ptr(self.D)
        d.goal, d.goal_bstride = ptr(goals), _stride(goals, nu, B)
        return d, (goals,)

    def pid_state(self, B: int, device: int = 0):
        """(integratederrors, lasterrors) [B][N]  (pid.jl:9-10)"""
        h = self.device_handle(B, device)
        a, b = np.zeros((B, len(self.eqcids))), np.zeros((B, len(self.eqcids)))
        check(_lib.lib().ccb_ctrl_pid_state(h, ptr(a), ptr(b), C.c_int32(0)))
        return a, b


class OpenLoop(Controller):
    """`control!(mechanism, k) = setForce!(mechanism, joint, [U[k]])`   (examples/trackingLQR_triple_cartpole.jl:45-47)"""

    kind = CTRL_TABLE

    def __init__(self, mechanism: Mechanism, eqcids, U):
        super().__init__()
        self.mech, self.eqcids = mechanism, list(eqcids)
        self.eqc_idx = i32([mechanism.eqc_index(i) for i in eqcids])
        nu = len(self.eqcids)
        U = f64(U)
        self.U = U.reshape(-1, U.shape[-2] if U.ndim >= 2 else U.size // nu, nu)   # [S][steps][nu]

    def _desc(self, B):
        nu = len(self.eqcids)
        d = _lib.CtrlDesc()
        U = f64(self.U)
        d.kind, d.nu, d.N = CTRL_TABLE, nu, U.shape[1]
        d.eqc = ptr(self.eqc_idx)
        d.U, d.U_bstride = ptr(U), _stride(U, U.shape[1] * nu, B)
        return d, (U,)


# ----------------------------------------------------------------------------------------------- simulate!
def simulate(mechanism: Mechanism, tend_or_storage, controller: Optional[Controller] = None, record: bool = False, states=None,
             device: int = 0, eps: float = 1e-10, newton_iter: int = 100, line_iter: int = 10, lanes: int = 0, return_status: bool = False):
    """simulate!(mechanism, tend, controller; record) / simulate!(mechanism, storage, control!; record)   (SURVEY A.7).

    `states` [B][nb][13] is the batch of initial continuous states; without it the mechanism's own bodies are
    simulated (B = 1) and their `state` fields are updated in place, exactly like the reference."""
    mech = mechanism
    if isinstance(tend_or_storage, Storage):
        steps = tend_or_storage.N
    else:
        steps = int(math.ceil(tend_or_storage / mech.dt))      # steps = 1:ceil(tend/Δt)
    own = states is None
    st0 = mech.state_vector()[None] if own else f64(states).reshape(-1, mech.nb, 13)
    B = st0.shape[0]
    batch = Batch(mech, B, device)
    batch.set_state(st0)
    storage, status = batch.simulate(controller, 1, steps, record=record, eps=eps, newton_iter=newton_iter, line_iter=line_iter, lanes=lanes)
    final = batch.get_state()
    if own:
        mech.set_state_vector(final[0])
    out = None
    if record:
        if isinstance(tend_or_storage, Storage) and tend_or_storage.B == B:
            tend_or_storage.data[...] = storage
            out = tend_or_storage
        else:
            out = Storage(steps, mech.nb, B, storage)
    if return_status:
        return out, final, status
    return out if record else None
