"""Mechanism / controller-weight definitions of the reference's example scripts (the BASELINE configs).

Each builder restates the *parameters* of one reference example (bodies, joints, vertices, initial
pose, Q/R/horizon, setpoints); nothing here computes anything.  Citations are to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

from .mechanism import (Box, EqualityConstraint, FixedOrientation, Mechanism, Origin, Prismatic, Revolute, RotX, setPosition)

EX = np.array([1.0, 0.0, 0.0])
EY = np.array([0.0, 1.0, 0.0])


def pendulum(phi0: float = math.pi - 0.4, g: float = -9.81, dt: float = 0.01):
    """examples/lqr_pendulum.jl:8-40"""
    p2 = np.array([0.0, 0.0, 0.5])
    origin = Origin()
    link1 = Box(0.1, 0.1, 1.0, 1.0, name="link1")
    joint = EqualityConstraint(Revolute(origin, link1, EX, p2=p2), name="joint1")
    mech = Mechanism(origin, [link1], [joint], g=g, dt=dt)
    setPosition(origin, link1, p2=p2, dq=RotX(phi0))
    Q = [np.zeros((12, 12))]
    Q[0][6, 6] = 1000.0
    Q[0][9, 9] = 100.0
    R = [np.ones((1, 1))]
    cfg = dict(Q=Q, R=R, horizon=math.inf, xd=[np.array([0.0, 0.0, 0.5])], qd=[RotX(math.pi)], eqcids=[joint.id])
    return mech, cfg


def cartpole(y0: float = 0.5, phi0: float = 0.2, dt: float = 0.01):
    """examples/lqr_cartpole.jl:8-41"""
    p2 = np.array([0.0, 0.0, 0.5])
    origin = Origin()
    cart = Box(0.1, 0.5, 0.1, 0.5, name="cart")
    pole = Box(0.1, 0.1, 1.0, 1.0, name="pole")
    joint1 = EqualityConstraint(Prismatic(origin, cart, EY), name="joint1")
    joint2 = EqualityConstraint(Revolute(cart, pole, EX, p2=-p2), name="joint2")
    mech = Mechanism(origin, [cart, pole], [joint1, joint2], g=-9.81, dt=dt)
    setPosition(origin, cart, dx=[0.0, y0, 0.0])
    setPosition(cart, pole, p2=-p2, dq=RotX(phi0))
    Q = [np.eye(12) for _ in range(2)]
    R = [np.ones((1, 1))]
    cfg = dict(Q=Q, R=R, horizon=10.0, xd=[np.zeros(3), np.array([0.0, 0.0, 0.5])], qd=None, eqcids=[joint1.id])
    return mech, cfg


def set_cartpole_pose(mech: Mechanism, y0: float, phis):
    """initial pose of a cart + chain of poles (examples/lqr_cartpole_n_pendulum.jl:39-43)"""
    p2 = np.array([0.0, 0.0, 0.5])
    bodies = mech.bodies
    setPosition(mech.origin, bodies[0], dx=[0.0, y0, 0.0])
    phis = np.atleast_1d(phis)
    setPosition(bodies[0], bodies[1], p2=-p2, dq=RotX(phis[0]))
    for i in range(1, len(phis)):
        setPosition(bodies[i], bodies[i + 1], p1=p2, p2=-p2, dq=RotX(phis[i]))


def cartpole_n_pendulum(n: int = 3, y0: float = 0.0, phis=None, dt: float = 0.01):
    """examples/lqr_cartpole_n_pendulum.jl:8-50"""
    p2 = np.array([0.0, 0.0, 0.5])
    origin = Origin()
    cart = Box(0.1, 0.5, 0.1, 0.5, name="cart")
    bodies = [cart] + [Box(0.1, 0.1, 1.0, 1.0, name=f"pole{i + 1}") for i in range(n)]
    eqcs = [EqualityConstraint(Prismatic(origin, cart, EY), name="joint1"),
            EqualityConstraint(Revolute(cart, bodies[1], EX, p2=-p2), name="joint2")]
    for i in range(2, n + 1):
        eqcs.append(EqualityConstraint(Revolute(bodies[i - 1], bodies[i], EX, p1=p2, p2=-p2), name=f"joint{i + 1}"))
    mech = Mechanism(origin, bodies, eqcs, g=-9.81, dt=dt)
    set_cartpole_pose(mech, y0, np.zeros(n) if phis is None else phis)
    Q = [np.eye(12) for _ in range(n + 1)]
    R = [np.ones((1, 1))]
    xd = [np.zeros(3)] + [np.array([0.0, 0.0, i + 0.5]) for i in range(n)]
    cfg = dict(Q=Q, R=R, horizon=10.0, xd=xd, qd=None, eqcids=[eqcs[0].id])
    return mech, cfg


def acrobot(dt: float = 0.01):
    """examples/lqr_acrobot.jl:8-47"""
    p2a = np.array([0.0, 0.0, 0.5])
    p2b = np.array([0.0, 0.0, 1.0])
    origin = Origin()
    link1 = Box(0.1, 0.1, 1.0, 1.0, name="link1")
    link2 = Box(0.1, 0.1, 2.0, 1.0, name="link2")
    joint1 = EqualityConstraint(Revolute(origin, link1, EX, p2=p2a), name="joint1")
    joint2 = EqualityConstraint(Revolute(link1, link2, EX, p1=-p2a, p2=p2b), name="joint2")
    mech = Mechanism(origin, [link1, link2], [joint1, joint2], g=-9.81, dt=dt)
    setPosition(origin, link1, p2=p2a, dq=RotX(math.pi - 0.1))
    setPosition(link1, link2, p1=-p2a, p2=p2b, dq=RotX(0.1))
    Q = [np.zeros((12, 12)) for _ in range(2)]
    Q[0][6, 6] = 4.0; Q[0][9, 9] = 4.0; Q[1][6, 6] = 1.0; Q[1][9, 9] = 1.0
    R = [np.ones((1, 1))]
    cfg = dict(Q=Q, R=R, horizon=10.0, xd=[np.array([0.0, 0.0, 0.5]), np.array([0.0, 0.0, 2.0])], qd=[RotX(math.pi), RotX(math.pi)],
               eqcids=[joint2.id])
    return mech, cfg


def prismatic(dt: float = 0.01):
    """examples/lqr_prismatic.jl:8-30"""
    origin = Origin()
    link1 = Box(0.1, 0.1, 0.1, 1.0, name="link1")
    joint = EqualityConstraint(Prismatic(origin, link1, EX), name="joint1")
    mech = Mechanism(origin, [link1], [joint], g=0.0, dt=dt)
    setPosition(origin, link1, dx=[1.0, 0.0, 0.0])
    cfg = dict(Q=[np.eye(12)], R=[np.ones((1, 1))], horizon=10.0, xd=None, qd=None, eqcids=[joint.id])
    return mech, cfg


def doublependulum(phi1: float = 0.0, phi2: float = 0.0, dt: float = 0.01):
    """examples/pid_doublependulum.jl:6-36"""
    p2 = np.array([0.0, 0.0, 0.5])
    origin = Origin()
    link1 = Box(0.1, 0.1, 1.0, 1.0, name="link1")
    link2 = Box(0.1, 0.1, 1.0, 1.0, name="link2")
    joint1 = EqualityConstraint(Revolute(origin, link1, EX, p2=p2), name="joint1")
    joint2 = EqualityConstraint(Revolute(link1, link2, EX, p1=-p2, p2=p2), name="joint2")
    mech = Mechanism(origin, [link1, link2], [joint1, joint2], dt=dt)
    setPosition(origin, link1, p2=p2, dq=RotX(phi1))
    setPosition(link1, link2, p1=-p2, p2=p2, dq=RotX(phi2))
    cfg = dict(eqcids=[joint1.id, joint2.id], goals=[math.pi / 2, -math.pi / 4], P=[10.0, 10.0], I=[10.0, 10.0], D=[5.0, 5.0])
    return mech, cfg


def triple_cartpole(dt: float = 0.01):
    """examples/trackingLQR_triple_cartpole.jl:11-44 (bodies/joints) and :64-73 (weights)"""
    p2 = np.array([0.0, 0.0, 0.5])
    origin = Origin()
    cart = Box(0.1, 0.5, 0.1, 0.5, name="cart")
    poles = [Box(0.1, 0.1, 1.0, 1.0, name=f"pole{i + 1}") for i in range(3)]
    j1 = EqualityConstraint(Prismatic(origin, cart, EY), name="joint1")
    j2 = EqualityConstraint(Revolute(cart, poles[0], EX, p2=p2), name="joint2")
    j3 = EqualityConstraint(Revolute(poles[0], poles[1], EX, p1=-p2, p2=p2), name="joint3")
    j4 = EqualityConstraint(Revolute(poles[1], poles[2], EX, p1=-p2, p2=p2), name="joint4")
    mech = Mechanism(origin, [cart] + poles, [j1, j2, j3, j4], g=-9.81, dt=dt)
    setPosition(origin, cart, dx=[0.0, 0.0, 0.0])
    setPosition(cart, poles[0], p2=p2, dq=RotX(0.0))
    setPosition(poles[0], poles[1], p1=-p2, p2=p2, dq=RotX(0.0))
    setPosition(poles[1], poles[2], p1=-p2, p2=p2, dq=RotX(0.0))
    Q = [np.zeros((12, 12)) for _ in range(4)]
    Q[0][1, 1] = 10.0; Q[0][4, 4] = 1.0
    for i in (1, 2, 3):
        Q[i][6, 6] = 40.0; Q[i][9, 9] = 1.0
    R = [np.ones((1, 1)) * 0.1]
    cfg = dict(Q=Q, R=R, eqcids=[j1.id])
    return mech, cfg


def deltabot(dt: float = 0.01, reg: float = 1e-10):
    """examples/lqr_deltabot.jl:8-53 (closed chain: 5 bodies, 5 eqcs incl. a 13-row one, 2 loops)"""
    L = 1.0
    pll = np.array([0.0, 0.0, L / 2])
    pul = np.array([0.0, 0.0, L / 4])
    pp = np.array([0.0, 0.0, L / 4 * math.sqrt(2)])
    origin = Origin()
    lowerlegl = Box(0.1, 0.1, L, L, name="lowerlegl")
    lowerlegr = Box(0.1, 0.1, L, L, name="lowerlegr")
    upperlegl = Box(0.1, 0.1, L / 2, L / 2, name="upperlegl")
    upperlegr = Box(0.1, 0.1, L / 2, L / 2, name="upperlegr")
    platform = Box(0.1, 0.1, L / 2 * math.sqrt(2), L / 2 * math.sqrt(2), name="platform")
    floorlr = EqualityConstraint(Revolute(origin, lowerlegl, EX, p2=-pll), Revolute(origin, lowerlegr, EX, p2=-pll),
                                 FixedOrientation(origin, platform, qoffset=RotX(math.pi / 2)), name="floorlr")
    kneel = EqualityConstraint(Revolute(lowerlegl, upperlegl, EX, p1=pll, p2=-pul), name="kneel")
    kneer = EqualityConstraint(Revolute(lowerlegr, upperlegr, EX, p1=pll, p2=-pul), name="kneer")
    platl = EqualityConstraint(Revolute(platform, upperlegl, EX, p2=pul, p1=pp), name="platl")
    platr = EqualityConstraint(Revolute(platform, upperlegr, EX, p2=pul, p1=-pp), name="platr")
    links = [lowerlegl, lowerlegr, upperlegl, upperlegr, platform]
    constraints = [platl, platr, floorlr, kneel, kneer]
    mech = Mechanism(origin, links, constraints, g=-9.81, dt=dt, reg=reg)
    setPosition(origin, lowerlegl, p2=-pll, dq=RotX(math.pi / 4))
    setPosition(origin, lowerlegr, p2=-pll, dq=RotX(-math.pi / 4))
    setPosition(lowerlegl, upperlegl, p1=pll, p2=-pul, dq=RotX(-math.pi / 2))
    setPosition(lowerlegr, upperlegr, p1=pll, p2=-pul, dq=RotX(math.pi / 2))
    setPosition(upperlegl, platform, p1=pul, p2=pp, dq=RotX(3 * math.pi / 4))
    Q = [np.zeros((12, 12)) for _ in range(5)]
    Q[4][1, 1] = 10.0; Q[4][2, 2] = 10.0; Q[4][4, 4] = 1.0; Q[4][5, 5] = 1.0
    R = [np.ones((1, 1)) * 0.1 for _ in range(2)]
    cfg = dict(Q=Q, R=R, horizon=math.inf, xd=[b.state.xc.copy() for b in links], qd=[b.state.qc.copy() for b in links],
               eqcids=[platl.id, platr.id], Ftaud=[6.7879484, -6.7879484])
    return mech, cfg


def deltabot_initial_conditions(s1: int = 101, s2: int = 101, L: float = 1.0):
    """examples/lqr_deltabot.jl:57-145: the grid of platform positions the two legs can reach and the leg angles that realise them.
    Returns (A [nc][2] platform (y, z), anglesl [nc][2], anglesr [nc][2]); 1973 valid positions for the example's 101 x 101 grid."""
    ppz3 = L / 4 * math.sqrt(2)
    coords = []
    for i in range(1, s1 + 1):                                      # :64-86
        ppz = -1.5 * L + 3 * L * (i - 1) / (s1 - 1)
        if ppz < 0:
            continue
        for j in range(1, s2 + 1):
            ppy = -1.5 * L + 3 * L * (j - 1) / (s2 - 1)
            ppr = np.array([0.0, ppy + ppz3, ppz]); ppl = np.array([0.0, ppy - ppz3, ppz])
            if np.linalg.norm(ppr) <= 1.5 * L and np.linalg.norm(ppl) <= 1.5 * L and np.linalg.norm(ppr) >= 0.5 * L and np.linalg.norm(ppl) >= 0.5 * L:
                coords.append((ppy, ppz))
    A = np.array(coords)
    nc = A.shape[0]
    anglesl = np.zeros((nc, 2)); anglesr = np.zeros((nc, 2))
    a, b = L, L / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(nc):                                         # :100-141
            ppr = A[i] + np.array([ppz3, 0.0]); ppl = A[i] - np.array([ppz3, 0.0])
            cl = np.linalg.norm(ppl)
            bl = math.acos((a * a + cl * cl - b * b) / (2 * a * cl)); gl = math.acos((a * a + b * b - cl * cl) / (2 * a * b))
            dl = abs(math.atan(ppl[0] / ppl[1])) if ppl[1] != 0 else math.pi / 2
            cr = np.linalg.norm(ppr)
            br = math.acos((a * a + cr * cr - b * b) / (2 * a * cr)); gr = math.acos((a * a + b * b - cr * cr) / (2 * a * b))
            dr = abs(math.atan(ppr[0] / ppr[1])) if ppr[1] != 0 else math.pi / 2
            if ppl[0] <= 0 and ppl[1] >= 0: anglesl[i] = [dl + bl, -math.pi + gl]
            elif ppl[0] >= 0 and ppl[1] >= 0: anglesl[i] = [-dl + bl, -math.pi + gl]
            elif ppl[0] >= 0 and ppl[1] <= 0: anglesl[i] = [-math.pi + dl - bl, math.pi - gl]
            elif ppl[0] <= 0 and ppl[1] <= 0: anglesl[i] = [-math.pi - dl - bl, math.pi - gl]
            if ppr[0] <= 0 and ppr[1] >= 0: anglesr[i] = [dr - br, math.pi - gr]
            elif ppr[0] >= 0 and ppr[1] >= 0: anglesr[i] = [-dr - br, math.pi - gr]
            elif ppr[0] >= 0 and ppr[1] <= 0: anglesr[i] = [-math.pi + dr + br, -math.pi + gr]
            elif ppr[0] <= 0 and ppr[1] <= 0: anglesr[i] = [-math.pi - dr + br, -math.pi + gr]
    return A, anglesl, anglesr


def deltabot_states(mech: Mechanism, B: int, grid: int = 105):
    """[B][5][13]: the first B valid initial conditions of the example's grid construction (lqr_deltabot.jl:147-152, one `i` per system).
    SURVEY 8(d) C5: a 105 x 105 grid has more than 2048 valid platform positions."""
    L = 1.0
    pll = np.array([0.0, 0.0, L / 2]); pul = np.array([0.0, 0.0, L / 4])
    A, al, ar = deltabot_initial_conditions(grid, grid, L)
    assert A.shape[0] >= B, f"grid {grid} x {grid} has only {A.shape[0]} valid positions"
    lowerlegl, lowerlegr, upperlegl, upperlegr, platform = mech.bodies
    keep = mech.state_vector()
    out = np.zeros((B, mech.nb, 13))
    for i in range(B):
        setPosition(mech.origin, lowerlegl, p2=-pll, dq=RotX(al[i][0]))
        setPosition(mech.origin, lowerlegr, p2=-pll, dq=RotX(ar[i][0]))
        setPosition(lowerlegl, upperlegl, p1=pll, p2=-pul, dq=RotX(al[i][1]))
        setPosition(lowerlegr, upperlegr, p1=pll, p2=-pul, dq=RotX(ar[i][1]))
        setPosition(mech.origin, platform, p1=np.array([0.0, A[i][0], A[i][1]]), dq=RotX(math.pi / 2))
        out[i] = mech.state_vector()
    mech.set_state_vector(keep)
    return out


def sawyer(description_path: str, dt: float = 0.01):
    """examples/lqr_sawyer.jl:8-29: sawyer_arm.urdf, floating=false, g=0; Q = 1000 I12 x 7, R = 1 x 7, horizon 20.
    `description_path` is the URDF itself or the parsed JSON fixture (tests/golden/sawyer_arm.json)."""
    from . import urdf
    if description_path.endswith(".json"):
        mech = urdf.mechanism_from_json(description_path, floating=False, g=0.0, dt=dt)
    else:
        mech = urdf.mechanism_from_urdf(description_path, floating=False, g=0.0, dt=dt)
    Q = [np.eye(12) * 1000.0 for _ in range(mech.nb)]
    R = [np.ones((1, 1)) for _ in range(mech.ne)]
    cfg = dict(Q=Q, R=R, horizon=20.0, xd=[b.state.xc.copy() for b in mech.bodies], qd=[b.state.qc.copy() for b in mech.bodies],
               eqcids=[e.id for e in mech.eqconstraints])
    return mech, cfg
