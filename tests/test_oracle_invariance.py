"""CPU: invariances the restatement must have.  They hold for the mathematics of the reference whatever internal
conventions ConstrainedDynamics.jl uses, so they pin the oracle independently of those conventions:
  * the gains Ku do not depend on the basis / scaling chosen for the multipliers (G -> T G, B_lambda -> B_lambda S);
  * the states do not depend on the order in which the equality constraints are listed (ids, DFS order, elimination order all change)."""
import math

import numpy as np


def test_gains_are_invariant_under_constraint_basis_and_multiplier_scaling(orc):
    rng = np.random.default_rng(4)
    n, mu, ml, N = 12, 2, 5, 40
    A = rng.normal(size=(n, n)) * 0.3
    Bu = rng.normal(size=(n, mu)); Bl = rng.normal(size=(n, ml)); G = rng.normal(size=(ml, n))
    Q = np.eye(n) * 0.02; R = np.eye(mu) * 0.05
    ref = orc.dlqr(A, Bu, Bl, G, Q, R, N)
    T = rng.normal(size=(ml, ml)) + 3 * np.eye(ml)            # another basis of the constraint rows
    Sc = np.diag(10.0 ** rng.uniform(-2, 2, ml)) @ (rng.normal(size=(ml, ml)) + 3 * np.eye(ml))     # another scaling / mixing of lambda
    out = orc.dlqr(A, Bu, Bl @ Sc, T @ G, Q, R, N)
    assert out["kexit"] == ref["kexit"]
    assert np.abs(np.asarray(out["Ku"]) - np.asarray(ref["Ku"])).max() < 1e-10 * np.abs(np.asarray(ref["Ku"])).max()


def _chain(cc, order, body_order=(0, 1, 2)):
    """cart + two pendulum links; `order` permutes the list of equality constraints handed to Mechanism, `body_order` the list of bodies"""
    from ccb200 import Box, EqualityConstraint, Mechanism, Origin, Prismatic, Revolute, RotX, setPosition
    EX, EY = np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])
    p = np.array([0.0, 0.0, 0.5])
    origin = Origin()
    cart = Box(0.1, 0.5, 0.1, 0.5, name="cart")
    l1 = Box(0.1, 0.1, 1.0, 1.0, name="l1"); l2 = Box(0.1, 0.1, 1.0, 0.7, name="l2")
    j0 = EqualityConstraint(Prismatic(origin, cart, EY), name="j0")
    j1 = EqualityConstraint(Revolute(cart, l1, EX, p2=-p), name="j1")
    j2 = EqualityConstraint(Revolute(l1, l2, EX, p1=p, p2=-p), name="j2")
    eqcs = [j0, j1, j2]
    bodies = [cart, l1, l2]
    mech = Mechanism(origin, [bodies[i] for i in body_order], [eqcs[i] for i in order], g=-9.81, dt=0.01)
    setPosition(origin, cart, dx=[0.0, 0.1, 0.0])
    setPosition(cart, l1, p2=-p, dq=RotX(math.pi - 0.3))
    setPosition(l1, l2, p1=p, p2=-p, dq=RotX(0.4))
    return mech


def test_states_are_invariant_under_the_order_of_the_equality_constraints(cc, orc):
    runs = []
    for order in ([0, 1, 2], [2, 0, 1], [1, 2, 0]):
        mech = _chain(cc, order)
        om = orc.OracleMech(mech.desc())
        r = om.simulate(mech.state_vector()[None], 200, record=True)
        assert r["iters"][0] > 0
        runs.append((om.graph(), r["storage"]))
    assert not np.array_equal(runs[0][0]["dfs_list"], runs[1][0]["dfs_list"])      # the orderings really differ ...
    for g, sto in runs[1:]:
        assert np.abs(sto - runs[0][1]).max() < 1e-9                                # ... the trajectory (a swinging double pendulum on a cart) does not


def test_states_are_invariant_under_the_order_of_the_bodies(cc, orc):
    """listing the bodies in another order permutes the state array (and the columns of A, K) but not the motion of each body"""
    ref = None
    for body_order, order in (((0, 1, 2), [0, 1, 2]), ((2, 0, 1), [0, 1, 2]), ((1, 2, 0), [2, 1, 0])):
        mech = _chain(cc, order, body_order)
        om = orc.OracleMech(mech.desc())
        r = om.simulate(mech.state_vector()[None], 200, record=True)
        assert r["iters"][0] > 0
        by_name = {b.name: r["storage"][:, 0, i] for i, b in enumerate(mech.bodies)}
        if ref is None:
            ref = by_name
        for name in ref:
            assert np.abs(by_name[name] - ref[name]).max() < 1e-9, (body_order, name)
