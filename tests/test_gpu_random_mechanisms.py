"""Property-style parity: randomly generated serial chains (random masses, full inertia tensors, joint axes, vertices, orientation
offsets, joint types) through every rollout kernel and through linearsystem / dlqr, against the CPU oracle."""
import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def random_chain(cc, rng, nj, g=-9.81):
    Body, EqualityConstraint, Mechanism, Origin, Prismatic, Revolute = cc.Body, cc.EqualityConstraint, cc.Mechanism, cc.Origin, cc.Prismatic, cc.Revolute
    origin = Origin()
    bodies, eqcs = [], []
    parent = origin
    for j in range(nj):
        m = rng.uniform(0.3, 2.0)
        Aq = rng.normal(size=(3, 3))
        J = (Aq @ Aq.T) * 0.02 + np.eye(3) * 0.02 * m               # symmetric positive definite, full
        b = Body(m, J, name=f"b{j}")
        axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
        p1 = rng.uniform(-0.3, 0.3, 3); p2 = rng.uniform(-0.3, 0.3, 3)
        qoff = rng.normal(size=4); qoff /= np.linalg.norm(qoff)
        if j > 0 and rng.uniform() < 0.35:
            jt = Prismatic(parent, b, axis, p1=p1, p2=p2, qoffset=qoff)
        else:
            jt = Revolute(parent, b, axis, p1=p1, p2=p2, qoffset=qoff)
        eqcs.append(EqualityConstraint(jt, name=f"j{j}"))
        bodies.append(b)
        parent = b
    mech = Mechanism(origin, bodies, eqcs, g=g, dt=0.01)
    return mech


def random_states(cc, mech, rng, B, amp=0.4, rate=0.5):
    st = np.zeros((B, mech.nb, 13))
    for s in range(B):
        cc.mechanism.propagate_minimal(mech, {i: rng.uniform(-amp, amp) for i in range(mech.ne)}, {i: rng.uniform(-rate, rate) for i in range(mech.ne)})
        st[s] = mech.state_vector()
    return st


@pytest.mark.parametrize("seed,nj", [(1, 1), (2, 2), (3, 3), (4, 4), (8, 5), (5, 6), (6, 7), (7, 9), (9, 12)])
def test_random_chain_rollout_every_kernel(cc, orc, seed, nj):
    rng = np.random.default_rng(seed)
    mech = random_chain(cc, rng, nj)
    B = 7
    st0 = random_states(cc, mech, rng, B)
    assert np.abs(mech.constraint_values()).max() < 1e-12
    om = orc.OracleMech(mech.desc())
    # open-loop inputs on every joint (exercises applyFτ! for rotated axes, force and torque inputs, parent and child sides)
    steps = 50
    U = 0.3 * rng.normal(size=(B, steps, nj))
    eqc = list(range(nj))
    co = om.make_ctrl(B, kind="table", eqc=eqc, U=U)
    r = om.simulate(st0, steps, ctrl=co, record=True, nthreads=7)
    assert np.all(r["iters"] > 0)
    ol = cc.OpenLoop(mech, cc.getid(mech.eqconstraints), U)
    lanes_list = [0, 1, 32] + ([5] if nj <= 6 else []) + [6]
    for lanes in lanes_list:
        sto, fin, status = cc.simulate(mech, steps * mech.dt, ol, record=True, states=st0, lanes=lanes, return_status=True)
        assert np.all(status > 0), lanes
        assert rel_err(sto.data, r["storage"]) < TOL, lanes
        assert rel_err(fin, r["states"]) < TOL, lanes


@pytest.mark.parametrize("seed,nj", [(11, 2), (12, 3), (13, 5)])
def test_random_chain_linearsystem_and_gains(cc, orc, seed, nj):
    rng = np.random.default_rng(seed)
    mech = random_chain(cc, rng, nj, g=-9.81)
    om = orc.OracleMech(mech.desc())
    S = 3
    zd = random_states(cc, mech, rng, S, amp=0.5, rate=0.2)
    Fd = 0.2 * rng.normal(size=(S, nj))
    eqc = list(range(nj))
    A, Bu, Bl, G, lam = cc.linearsystem(mech, zd, Fd, eqc, want_lambda=True)
    n = 12 * nj
    Qd = np.eye(n) * mech.dt; Rd = np.eye(nj) * mech.dt
    for s in range(S):
        ref = om.linearize(zd[s], Fd[s], eqc)
        assert rel_err(A[s], ref["A"]) < TOL and rel_err(Bu[s], ref["Bu"]) < TOL and rel_err(Bl[s], ref["Bl"]) < TOL and rel_err(G[s], ref["G"]) < TOL
    # gains: dlqr on the GPU-linearised systems vs the oracle's dlqr on ITS linearisation (short horizon: no conditioning issues)
    out = cc.lqr_synthesize(mech, zd, Fd, eqc, Qd, Rd, 40)
    for s in range(S):
        ref = om.linearize(zd[s], Fd[s], eqc)
        rd = orc.dlqr(ref["A"], ref["Bu"], ref["Bl"], ref["G"], Qd, Rd, 40)
        assert rel_err(out["Ku"][s], rd["Ku"]) < 1e-9, s


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_random_branching_tree_with_mixed_joint_dimensions(cc, orc, seed):
    # random tree (not a chain) with Revolute (5 rows), Prismatic (5), Fixed (6), Cylindrical (4), Spherical (3) and FixedOrientation-free joints:
    # the general kernel K1 (table-driven, joint blocks of size 3..6) falling under gravity from the zero configuration
    Body, Cylindrical, EqualityConstraint, Fixed, Mechanism, Origin = cc.Body, cc.Cylindrical, cc.EqualityConstraint, cc.Fixed, cc.Mechanism, cc.Origin
    Prismatic, Revolute, Spherical, qmul, qrot = cc.Prismatic, cc.Revolute, cc.Spherical, cc.mechanism.qmul, cc.mechanism.qrot
    rng = np.random.default_rng(seed)
    origin = Origin()
    nb = 7
    bodies, eqcs = [], []
    kinds = ["rev", "pri", "fix", "cyl", "sph", "rev", "rev"]
    rng.shuffle(kinds)
    kinds[0] = "rev"
    for j in range(nb):
        m = rng.uniform(0.3, 2.0)
        Aq = rng.normal(size=(3, 3))
        b = Body(m, (Aq @ Aq.T) * 0.02 + np.eye(3) * 0.02 * m, name=f"b{j}")
        parent = origin if j == 0 else bodies[rng.integers(0, j)]
        axis = rng.normal(size=3); axis /= np.linalg.norm(axis)
        p1 = rng.uniform(-0.3, 0.3, 3); p2 = rng.uniform(-0.3, 0.3, 3)
        qoff = rng.normal(size=4); qoff /= np.linalg.norm(qoff)
        k = kinds[j]
        if k == "rev": jt = Revolute(parent, b, axis, p1=p1, p2=p2, qoffset=qoff)
        elif k == "pri": jt = Prismatic(parent, b, axis, p1=p1, p2=p2, qoffset=qoff)
        elif k == "fix": jt = Fixed(parent, b, p1=p1, p2=p2, qoffset=qoff)
        elif k == "cyl": jt = Cylindrical(parent, b, axis, p1=p1, p2=p2, qoffset=qoff)
        else: jt = Spherical(parent, b, p1=p1, p2=p2)
        # zero configuration of any joint type: q_b = q_a (x) qoff, vertices coincide
        qa = np.array([1.0, 0, 0, 0]) if parent is origin else parent.state.qc
        xa = np.zeros(3) if parent is origin else parent.state.xc
        b.state.qc = qmul(qa, jt.qoffset)
        b.state.xc = xa + qrot(qa, p1) - qrot(b.state.qc, p2)
        bodies.append(b)
        eqcs.append(EqualityConstraint(jt, name=f"j{j}"))
    mech = Mechanism(origin, bodies, eqcs, g=-9.81, dt=0.01)
    assert np.abs(mech.constraint_values()).max() < 1e-12
    g = cc.handle(mech).graph()
    om = orc.OracleMech(mech.desc())
    go = om.graph()
    for key in go:
        assert np.array_equal(g[key], go[key]), key
    B = 4
    st0 = mech.batch_states(B)
    r = om.simulate(st0, 50, record=True, nthreads=4)
    assert np.all(r["iters"] > 0)
    for lanes in (0, 16, 32):
        sto, fin, status = cc.simulate(mech, 0.5, None, record=True, states=st0, lanes=lanes, return_status=True)
        assert np.all(status > 0), lanes
        assert rel_err(sto.data, r["storage"]) < TOL, lanes


@pytest.mark.parametrize("order", [[0, 1, 2], [2, 0, 1], [1, 2, 0], [0, 2, 1]])
def test_constraint_listing_order_does_not_change_the_trajectory(cc, orc, order):
    """the same cart + double pendulum with its equality constraints listed in every order (ids, DFS order and the elimination order change):
    every rollout kernel reproduces the trajectory of the canonical listing"""
    from test_oracle_invariance import _chain
    base = _chain(cc, [0, 1, 2])
    ref = orc.OracleMech(base.desc()).simulate(base.state_vector()[None], 150, record=True)
    mech = _chain(cc, order)
    st0 = mech.state_vector()[None].copy()
    for lanes in (0, 5, 6, 1, 32):
        sto, fin, status = cc.simulate(mech, 1.5, None, record=True, states=st0, lanes=lanes, return_status=True)
        assert np.all(status > 0), lanes
        assert rel_err(sto.data, ref["storage"]) < 1e-9, lanes


@pytest.mark.parametrize("body_order,order", [((2, 0, 1), [0, 1, 2]), ((1, 2, 0), [2, 1, 0])])
def test_body_listing_order_does_not_change_the_motion(cc, orc, body_order, order):
    """bodies (and constraints) listed in another order: the state array is permuted, the motion of each named body is not -- every kernel"""
    from test_oracle_invariance import _chain
    base = _chain(cc, [0, 1, 2])
    ref = orc.OracleMech(base.desc()).simulate(base.state_vector()[None], 150, record=True)["storage"][:, 0]
    ref_by_name = {b.name: ref[:, i] for i, b in enumerate(base.bodies)}
    mech = _chain(cc, order, body_order)
    st0 = mech.state_vector()[None].copy()
    for lanes in (0, 6, 1, 32):
        sto, fin, status = cc.simulate(mech, 1.5, None, record=True, states=st0, lanes=lanes, return_status=True)
        assert np.all(status > 0), lanes
        for i, b in enumerate(mech.bodies):
            assert rel_err(sto.data[:, 0, i], ref_by_name[b.name]) < 1e-9, (lanes, b.name)


def test_lqr_closed_loop_is_invariant_under_listing_orders(cc):
    """LQR about the upright pose of the cart + double pendulum, synthesised and rolled out on mechanisms that list bodies and constraints in
    different orders: gains are permuted consistently (columns by body position, rows by controlled eqc) and the closed-loop motion of each
    named body is the same"""
    import math
    from ccb200 import RotX, setPosition
    from test_oracle_invariance import _chain
    ref = None
    for body_order, order in (((0, 1, 2), [0, 1, 2]), ((2, 0, 1), [1, 2, 0]), ((1, 2, 0), [2, 0, 1])):
        mech = _chain(cc, order, body_order)
        by = {b.name: b for b in mech.bodies}
        p = np.array([0.0, 0.0, 0.5])
        # setpoint: upright
        setPosition(mech.origin, by["cart"], dx=[0.0, 0.0, 0.0])
        setPosition(by["cart"], by["l1"], p2=-p, dq=RotX(math.pi))
        setPosition(by["l1"], by["l2"], p1=p, p2=-p, dq=RotX(0.0))
        xd = [b.state.xc.copy() for b in mech.bodies]; qd = [b.state.qc.copy() for b in mech.bodies]
        j0 = [e for e in mech.eqconstraints if e.name == "j0"][0]
        lqr = cc.LQR(mech, cc.getid(mech.bodies), [j0.id], [np.eye(12) for _ in mech.bodies], [np.ones((1, 1))], 3.0, xd=xd, qd=qd)
        # start: slightly off upright
        setPosition(mech.origin, by["cart"], dx=[0.0, 0.05, 0.0])
        setPosition(by["cart"], by["l1"], p2=-p, dq=RotX(math.pi - 0.02))
        setPosition(by["l1"], by["l2"], p1=p, p2=-p, dq=RotX(0.03))
        st0 = mech.state_vector()[None].copy()
        sto, fin, status = cc.simulate(mech, 2.0, lqr, record=True, states=st0, return_status=True)
        assert np.all(status > 0)
        by_name = {b.name: sto.data[:, 0, i] for i, b in enumerate(mech.bodies)}
        if ref is None:
            ref = by_name
            assert np.abs(by_name["cart"][-1, 1]) < 0.05           # the controller does bring the cart back
        for name in ref:
            assert rel_err(by_name[name], ref[name]) < 1e-9, (body_order, order, name)


def test_two_independent_pendulums_in_one_mechanism(cc, orc):
    """a forest: two pendulums that share nothing but the origin, as ONE mechanism (two components, two origin joints).  Each must move exactly
    like the single pendulum -- exercises the component handling of the graph code and of the kernel choice (this is not a serial chain)"""
    import math
    from ccb200 import Box, EqualityConstraint, Mechanism, Origin, Revolute, RotX, setPosition
    EX = np.array([1.0, 0.0, 0.0])
    p = np.array([0.0, 0.0, 0.5])

    def build(angles):
        origin = Origin()
        bodies, eqcs = [], []
        for i, a in enumerate(angles):
            b = Box(0.1, 0.1, 1.0, 1.0 + 0.5 * i, name=f"b{i}")
            eqcs.append(EqualityConstraint(Revolute(origin, b, EX, p2=-p), name=f"j{i}"))
            bodies.append(b)
        mech = Mechanism(origin, bodies, eqcs, g=-9.81, dt=0.01)
        for b, a in zip(bodies, angles):
            setPosition(origin, b, p2=-p, dq=RotX(a))
        return mech

    both = build([0.7, -1.1])
    singles = [build([0.7]), None]
    # the second pendulum alone (heavier box): build it as the only body of its own mechanism
    origin = Origin()
    b1 = Box(0.1, 0.1, 1.0, 1.5, name="b1")
    m1 = Mechanism(origin, [b1], [EqualityConstraint(Revolute(origin, b1, EX, p2=-p), name="j1")], g=-9.81, dt=0.01)
    setPosition(origin, b1, p2=-p, dq=RotX(-1.1))
    singles[1] = m1
    refs = [orc.OracleMech(m.desc()).simulate(m.state_vector()[None], 200, record=True)["storage"][:, 0, 0] for m in singles]
    ob = orc.OracleMech(both.desc()).simulate(both.state_vector()[None], 200, record=True)
    assert ob["iters"][0] > 0
    for i in range(2):
        assert rel_err(ob["storage"][:, 0, i], refs[i]) < 1e-9
    for lanes in (0, 8, 32):
        sto, fin, status = cc.simulate(both, 2.0, None, record=True, states=both.state_vector()[None].copy(), lanes=lanes, return_status=True)
        assert np.all(status > 0), lanes
        for i in range(2):
            assert rel_err(sto.data[:, 0, i], refs[i]) < 1e-9, (lanes, i)
