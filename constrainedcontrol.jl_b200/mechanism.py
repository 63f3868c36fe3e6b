"""Host-side mechanism description (no GPU, no native code needed).

Mirrors the model-building surface of ConstrainedDynamics.jl that the reference's examples use
(reference examples/lqr_cartpole.jl:20-34, examples/lqr_deltabot.jl:25-41, examples/lqr_sawyer.jl:9-14):
``Origin``, ``Box``, ``Body``, ``Revolute``, ``Prismatic``, ``Fixed``, ``FixedOrientation``,
``EqualityConstraint``, ``Mechanism(origin, bodies, eqcs; g, Δt)``, ``setPosition!``, ``setVelocity!``,
``getid``, ``geteqconstraint``.  Python cannot spell ``!`` or ``Δ``: ``setPosition!`` -> ``setPosition``,
``Δt`` -> ``dt``, ``Δx``/``Δq`` -> ``dx``/``dq``.

Everything here is a plain description that is flattened (``Mechanism.desc()``) into the arrays the
C-ABI (`include/ccb.h: ccb_mech_desc`) takes.  Quaternions are scalar-first ``[s, x, y, z]``
(reference src/control/lqr.jl:101-102).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------- quaternion helpers


def qmul(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = a[0] * b[0] - np.dot(a[1:], b[1:])
    v = a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])
    return np.array([s, v[0], v[1], v[2]])


def qconj(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a[0], -a[1], -a[2], -a[3]])


def qrot(q, p):
    """R(q) p"""
    q = np.asarray(q, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    t = 2.0 * np.cross(q[1:], p)
    return p + q[0] * t + np.cross(q[1:], t)


def axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=np.float64)
    n = np.linalg.norm(axis)
    if n == 0.0:
        return np.array([1.0, 0.0, 0.0, 0.0])
    h = 0.5 * angle
    return np.concatenate([[math.cos(h)], math.sin(h) * axis / n])


def RotX(a):
    """Quaternion(RotX(a)) = (cos a/2, sin a/2, 0, 0)  """
    return axis_angle([1.0, 0.0, 0.0], a)


def RotY(a):
    return axis_angle([0.0, 1.0, 0.0], a)


def RotZ(a):
    return axis_angle([0.0, 0.0, 1.0], a)


def quat_from_rpy(r, p, y):
    """URDF fixed-axis roll-pitch-yaw: R = Rz(y) Ry(p) Rx(r)"""
    return qmul(qmul(RotZ(y), RotY(p)), RotX(r))


def quat_to_matrix(q):
    return np.stack([qrot(q, e) for e in np.eye(3)], axis=1)


ONE_Q = np.array([1.0, 0.0, 0.0, 0.0])

# ----------------------------------------------------------------------------- bodies


@dataclass
class State:
    """Continuous state of one body: the fields the reference reads (examples/lqr_sawyer.jl:16-17)."""

    xc: np.ndarray = field(default_factory=lambda: np.zeros(3))
    qc: np.ndarray = field(default_factory=lambda: ONE_Q.copy())
    vc: np.ndarray = field(default_factory=lambda: np.zeros(3))
    wc: np.ndarray = field(default_factory=lambda: np.zeros(3))

    def vector(self):
        return np.concatenate([self.xc, self.qc, self.vc, self.wc])


class Origin:
    id = 0

    def __init__(self):
        self.state = State()
        self.name = "origin"


class Body:
    def __init__(self, m: float, J, name: str = ""):
        self.m = float(m)
        self.J = np.asarray(J, dtype=np.float64).reshape(3, 3).copy()
        self.state = State()
        self.id = -1
        self.name = name


def Box(x: float, y: float, z: float, m: float, name: str = "") -> Body:
    """Box(x,y,z,m): J = m/12 diag(y^2+z^2, x^2+z^2, x^2+y^2)  """
    J = m / 12.0 * np.diag([y * y + z * z, x * x + z * z, x * x + y * y])
    return Body(m, J, name)


# ----------------------------------------------------------------------------- joints


@dataclass
class Joint:
    parent: object
    child: Body
    nT: int
    nR: int
    axis: np.ndarray
    p1: np.ndarray
    p2: np.ndarray
    qoffset: np.ndarray

    @property
    def dim(self):
        return self.nT + self.nR


def _joint(parent, child, nT, nR, axis, p1, p2, qoffset):
    z = np.zeros(3)
    return Joint(parent, child, nT, nR, np.asarray(axis, dtype=np.float64), np.asarray(z if p1 is None else p1, dtype=np.float64),
                 np.asarray(z if p2 is None else p2, dtype=np.float64), np.asarray(ONE_Q if qoffset is None else qoffset, dtype=np.float64))


# Joint prototypes: Revolute = Translational3 + Rotational2(axis) (confirmed by pid.jl:44), ...
def Revolute(parent, child, axis, p1=None, p2=None, qoffset=None):
    return _joint(parent, child, 3, 2, axis, p1, p2, qoffset)


def Prismatic(parent, child, axis, p1=None, p2=None, qoffset=None):
    return _joint(parent, child, 2, 3, axis, p1, p2, qoffset)


def Fixed(parent, child, p1=None, p2=None, qoffset=None):
    return _joint(parent, child, 3, 3, [1.0, 0.0, 0.0], p1, p2, qoffset)


def FixedOrientation(parent, child, qoffset=None):
    return _joint(parent, child, 0, 3, [1.0, 0.0, 0.0], None, None, qoffset)


def Spherical(parent, child, p1=None, p2=None):
    return _joint(parent, child, 3, 0, [1.0, 0.0, 0.0], p1, p2, None)


def Cylindrical(parent, child, axis, p1=None, p2=None, qoffset=None):
    return _joint(parent, child, 2, 2, axis, p1, p2, qoffset)


class EqualityConstraint:
    """One or several joints that share a parent (examples/lqr_deltabot.jl:25 holds three)."""

    def __init__(self, *joints: Joint, name: str = ""):
        assert len(joints) >= 1
        self.constraints = list(joints)
        self.parent = joints[0].parent
        for j in joints:
            assert j.parent is self.parent, "all joints of an EqualityConstraint share the parent"
        self.id = -1
        self.name = name

    def __len__(self):  # ConstrainedDynamics.length(eqc)  (pid.jl:19, lqr_tracking.jl:80)
        return sum(j.dim for j in self.constraints)

    @property
    def parentid(self):
        return None if isinstance(self.parent, Origin) else self.parent.id

    @property
    def childids(self):
        return [j.child.id for j in self.constraints]


# ----------------------------------------------------------------------------- mechanism


class Mechanism:
    """Mechanism(origin, bodies, eqcs; g=-9.81, Δt=0.01)  (examples/lqr_cartpole.jl:32).

    Node ids (SURVEY A.1 decision, documented host option): ``id_scheme='eqcs_first'`` numbers eqcs
    1..Ne then bodies Ne+1..Ne+Nb; ``'bodies_first'`` the other way round.  Gain columns / Δz are always
    indexed by body *position* in ``bodies`` (as lqr.jl:91 `enumerate` does), never by id.
    """

    def __init__(self, origin: Origin, bodies: Sequence[Body], eqcs: Sequence[EqualityConstraint], g: float = -9.81, dt: float = 0.01,
                 id_scheme: str = "eqcs_first", reg: float = 0.0):
        self.origin = origin
        self.bodies: List[Body] = list(bodies)
        self.eqconstraints: List[EqualityConstraint] = list(eqcs)
        self.g = float(g)
        self.dt = float(dt)
        self.reg = float(reg)
        self.id_scheme = id_scheme
        nb, ne = len(self.bodies), len(self.eqconstraints)
        if id_scheme == "eqcs_first":
            for i, e in enumerate(self.eqconstraints):
                e.id = i + 1
            for i, b in enumerate(self.bodies):
                b.id = ne + i + 1
        elif id_scheme == "bodies_first":
            for i, b in enumerate(self.bodies):
                b.id = i + 1
            for i, e in enumerate(self.eqconstraints):
                e.id = nb + i + 1
        else:
            raise ValueError(id_scheme)
        self._bindex = {id(b): i for i, b in enumerate(self.bodies)}
        for e in self.eqconstraints:
            for j in e.constraints:
                assert id(j.child) in self._bindex, "joint child is not a body of this mechanism"
                assert isinstance(j.parent, Origin) or id(j.parent) in self._bindex

    # -- sizes
    @property
    def nb(self):
        return len(self.bodies)

    @property
    def ne(self):
        return len(self.eqconstraints)

    @property
    def ml(self):
        return sum(len(e) for e in self.eqconstraints)

    def body_index(self, body_or_id) -> int:
        if isinstance(body_or_id, Body):
            return self._bindex[id(body_or_id)]
        for i, b in enumerate(self.bodies):
            if b.id == body_or_id:
                return i
        raise KeyError(body_or_id)

    def eqc_index(self, eqc_or_id) -> int:
        if isinstance(eqc_or_id, EqualityConstraint):
            return self.eqconstraints.index(eqc_or_id)
        if isinstance(eqc_or_id, str):
            for i, e in enumerate(self.eqconstraints):
                if e.name == eqc_or_id:
                    return i
            raise KeyError(eqc_or_id)
        for i, e in enumerate(self.eqconstraints):
            if e.id == eqc_or_id:
                return i
        raise KeyError(eqc_or_id)

    def is_1dof(self, eqc_index: int) -> bool:
        """pid.jl:31-37: 6*|unique(childids)| - length(eqc) == 1"""
        e = self.eqconstraints[eqc_index]
        return 6 * len(set(e.childids)) - len(e) == 1 and len(e.constraints) == 1

    # -- flat description for the C-ABI / oracle
    def desc(self) -> dict:
        nb, ne = self.nb, self.ne
        joints = [j for e in self.eqconstraints for j in e.constraints]
        nj = len(joints)
        ptr = np.zeros(ne + 1, dtype=np.int32)
        for i, e in enumerate(self.eqconstraints):
            ptr[i + 1] = ptr[i] + len(e.constraints)
        d = dict(
            nb=nb, ne=ne, nj=nj, dt=self.dt, g=self.g, reg=self.reg,
            id_scheme=0 if self.id_scheme == "eqcs_first" else 1,
            mass=np.array([b.m for b in self.bodies], dtype=np.float64),
            inertia=np.array([b.J.reshape(9) for b in self.bodies], dtype=np.float64).reshape(nb, 9),
            eqc_parent=np.array([-1 if isinstance(e.parent, Origin) else self.body_index(e.parent) for e in self.eqconstraints], dtype=np.int32),
            eqc_joint_ptr=ptr,
            joint_child=np.array([self.body_index(j.child) for j in joints], dtype=np.int32),
            joint_nT=np.array([j.nT for j in joints], dtype=np.int32),
            joint_nR=np.array([j.nR for j in joints], dtype=np.int32),
            joint_axis=np.array([j.axis for j in joints], dtype=np.float64).reshape(nj, 3),
            joint_p1=np.array([j.p1 for j in joints], dtype=np.float64).reshape(nj, 3),
            joint_p2=np.array([j.p2 for j in joints], dtype=np.float64).reshape(nj, 3),
            joint_qoff=np.array([j.qoffset for j in joints], dtype=np.float64).reshape(nj, 4),
        )
        return d

    # -- state <-> array
    def state_vector(self) -> np.ndarray:
        """[nb][13] = (xc, qc, vc, wc) per body"""
        return np.stack([b.state.vector() for b in self.bodies])

    def set_state_vector(self, s):
        s = np.asarray(s, dtype=np.float64).reshape(self.nb, 13)
        for b, r in zip(self.bodies, s):
            b.state.xc, b.state.qc, b.state.vc, b.state.wc = r[0:3].copy(), r[3:7].copy(), r[7:10].copy(), r[10:13].copy()

    def constraint_values(self) -> np.ndarray:
        """position-level joint residuals g(xc, qc) of every joint, concatenated: translational part
        R_a^T (x_b + R_b p2 - x_a) - p1 and rotational part vec(q_a^-1 q_b q_off^-1), each restricted to the constrained
        directions (all 3; the 2 orthogonal to the axis; or the axis).  A host-side diagnostic, not used by any solver."""
        out = []
        for e in self.eqconstraints:
            for j in e.constraints:
                if isinstance(j.parent, Origin):
                    xa, qa = np.zeros(3), ONE_Q
                else:
                    xa, qa = j.parent.state.xc, j.parent.state.qc
                xb, qb = j.child.state.xc, j.child.state.qc
                ax = j.axis / (np.linalg.norm(j.axis) or 1.0)
                gT = qrot(qconj(qa), xb + qrot(qb, j.p2) - xa) - j.p1
                gR = qmul(qmul(qconj(qa), qb), qconj(j.qoffset))[1:4]
                for n, g in ((j.nT, gT), (j.nR, gR)):
                    if n == 3:
                        out.append(g)
                    elif n == 2:
                        out.append(g - np.dot(ax, g) * ax)
                    elif n == 1:
                        out.append(np.array([np.dot(ax, g)]))
        return np.concatenate(out) if out else np.zeros(0)

    def batch_states(self, B: int) -> np.ndarray:
        """[B][nb][13], every system initialised from the bodies' current states"""
        return np.repeat(self.state_vector()[None], B, axis=0).copy()


def getid(x):
    if isinstance(x, (list, tuple)):
        return [getid(v) for v in x]
    return x.id


def geteqconstraint(mech: Mechanism, id_or_name) -> EqualityConstraint:
    return mech.eqconstraints[mech.eqc_index(id_or_name)]


def getbody(mech: Mechanism, id_) -> Body:
    return mech.bodies[mech.body_index(id_)]


# ----------------------------------------------------------------------------- setPosition! / setVelocity!


def setPosition(a, b=None, p1=None, p2=None, dx=None, dq=None, x=None, q=None):
    """setPosition!(a, b; p1, p2, Δx, Δq): q_b = q_a ⊗ Δq ; x_b = x_a + R(q_a)(p1 + Δx) - R(q_b) p2   (SURVEY A.3).
    With a single body: setPosition!(body; x, q)."""
    z = np.zeros(3)
    if b is None:
        a.state.xc = np.asarray(z if x is None else x, dtype=np.float64).copy()
        a.state.qc = np.asarray(ONE_Q if q is None else q, dtype=np.float64).copy()
        return
    p1 = np.asarray(z if p1 is None else p1, dtype=np.float64)
    p2 = np.asarray(z if p2 is None else p2, dtype=np.float64)
    dx = np.asarray(z if dx is None else dx, dtype=np.float64)
    dq = np.asarray(ONE_Q if dq is None else dq, dtype=np.float64)
    qa, xa = a.state.qc, a.state.xc
    qb = qmul(qa, dq)
    b.state.qc = qb
    b.state.xc = xa + qrot(qa, p1 + dx) - qrot(qb, p2)


def setVelocity(body, v=None, w=None):
    z = np.zeros(3)
    body.state.vc = np.asarray(z if v is None else v, dtype=np.float64).copy()
    body.state.wc = np.asarray(z if w is None else w, dtype=np.float64).copy()


def setJointPosition(mech: Mechanism, eqc, coords):
    """setPosition!(mech, eqc, [θ]) (examples/lqr_sawyer.jl:11-14): place the child of a 1-DoF eqc from the parent's
    pose and the joint's minimal coordinate (angle for Revolute, displacement for Prismatic)."""
    e = eqc if isinstance(eqc, EqualityConstraint) else geteqconstraint(mech, eqc)
    assert len(e.constraints) == 1
    j = e.constraints[0]
    th = float(np.atleast_1d(coords)[0])
    a, b = j.parent, j.child
    ax = j.axis / np.linalg.norm(j.axis)
    if j.nT == 3 and j.nR == 2:
        qb = qmul(qmul(a.state.qc, axis_angle(ax, th)), j.qoffset)
        b.state.qc = qb
        b.state.xc = a.state.xc + qrot(a.state.qc, j.p1) - qrot(qb, j.p2)
    elif j.nT == 2 and j.nR == 3:
        qb = qmul(a.state.qc, j.qoffset)
        b.state.qc = qb
        b.state.xc = a.state.xc + qrot(a.state.qc, j.p1 + ax * th) - qrot(qb, j.p2)
    else:
        raise ValueError("setJointPosition needs a Revolute or Prismatic eqc")


def setJointVelocity(mech: Mechanism, eqc, rate):
    """velocity counterpart of setJointPosition: child (vc, wc) from the parent's velocities and the joint rate (θ' or d').
    World angular velocity Ω = R(q) w (w is stored in the body frame):
      Revolute   Ω_b = Ω_a + R_a ax θ' ;  v_b = v_a + Ω_a x (R_a p1) - Ω_b x (R_b p2)
      Prismatic  Ω_b = Ω_a ;             v_b = v_a + Ω_a x (R_a (p1 + ax d)) + R_a ax d' - Ω_b x (R_b p2)"""
    e = eqc if isinstance(eqc, EqualityConstraint) else geteqconstraint(mech, eqc)
    assert len(e.constraints) == 1
    j = e.constraints[0]
    rate = float(np.atleast_1d(rate)[0])
    a, b = j.parent, j.child
    ax = j.axis / np.linalg.norm(j.axis)
    if isinstance(a, Origin):
        xa, qa, va, Wa = np.zeros(3), ONE_Q, np.zeros(3), np.zeros(3)
    else:
        xa, qa, va, Wa = a.state.xc, a.state.qc, a.state.vc, qrot(a.state.qc, a.state.wc)
    xb, qb = b.state.xc, b.state.qc
    r2 = qrot(qb, j.p2)
    if j.nT == 3 and j.nR == 2:
        Wb = Wa + qrot(qa, ax) * rate
        vb = va + np.cross(Wa, qrot(qa, j.p1)) - np.cross(Wb, r2)
    elif j.nT == 2 and j.nR == 3:
        Wb = Wa
        arm = xb + r2 - xa                       # = R_a (p1 + ax d)
        vb = va + np.cross(Wa, arm) + qrot(qa, ax) * rate - np.cross(Wb, r2)
    else:
        raise ValueError("setJointVelocity needs a Revolute or Prismatic eqc")
    b.state.vc = vb
    b.state.wc = qrot(qconj(qb), Wb)


def propagate_minimal(mech: Mechanism, coords: dict, rates: dict = None):
    """forward kinematics of positions AND velocities over a tree of 1-DoF eqcs (eqc index -> coordinate / rate), root first:
    the maximal-coordinate setpoint of the 6-argument `linearsystem` (reference src/control/lqr.jl:80)"""
    rates = rates or {}
    done = set()
    pending = list(range(mech.ne))
    guard = 0
    while pending and guard < 10 * mech.ne + 10:
        guard += 1
        i = pending.pop(0)
        e = mech.eqconstraints[i]
        par = e.parent
        if not isinstance(par, Origin) and id(par) not in done:
            pending.append(i)
            continue
        setJointPosition(mech, e, [coords.get(i, 0.0)])
        setJointVelocity(mech, e, [rates.get(i, 0.0)])
        for j in e.constraints:
            done.add(id(j.child))
    return mech


def propagate_positions(mech: Mechanism, coords: dict):
    """Forward kinematics over a tree: set every 1-DoF eqc named in `coords` (eqc index -> coordinate), root first."""
    done = set()
    pending = list(range(mech.ne))
    guard = 0
    while pending and guard < 10 * mech.ne + 10:
        guard += 1
        i = pending.pop(0)
        e = mech.eqconstraints[i]
        par = e.parent
        if not isinstance(par, Origin) and id(par) not in done:
            pending.append(i)
            continue
        setJointPosition(mech, e, [coords.get(i, 0.0)])
        for j in e.constraints:
            done.add(id(j.child))
