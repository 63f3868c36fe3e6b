"""The oracle against the numbers the reference owns: its one numeric fixture, the input sequence `U` of
examples/trackingLQR_triple_cartpole.jl:1 (identical copy in test/examples/), committed as tests/golden/triple_cartpole_U.npy, and the
gravity-compensation literal of examples/lqr_deltabot.jl:53 (last test).

`U` is a swing-up of the triple pendulum on a cart that upstream computed on ITS dynamics (ConstrainedDynamics.jl): the example plays it
open loop (`simulate!(mech, storage0, control!, record = true)` with `setForce!(mechanism, joint1, [U[k]])`, :45-53) from the hanging
configuration, and TrackingLQR then stabilises the recorded trajectory.  Nothing about the target is stored next to `U`, but the target is
known: all three poles upright.  An open-loop swing-up of a triple pendulum only arrives there on a model that agrees with the one it was
optimised on, so "the restated dynamics, fed the reference's U, end upright" is a known-answer test of the whole rollout path of the
restatement (Box mass / inertia, joint vertices, gravity, the force convention of `setForce!` on a prismatic joint, the time base k -> U[k],
the integrator and the constraint solve of a 4-body chain) against upstream itself -- the one place where the oracle is checked against
numbers the reference owns.  What it pins is quantified below: a 1 % error in the pole mass, the force scale or the time step, or 0.1 % in
gravity, misses the upright configuration by 4 ... 30 times more than the restatement does.  It is a physics-level pin (about 0.05 % in
those parameters), not a 1e-9 one: bit-level parity with upstream stays unverified."""
import math

import numpy as np
import pytest

HERE = __file__.rsplit("/", 1)[0]
P2 = np.array([0.0, 0.0, 0.5])


def _U():
    return np.load(HERE + "/golden/triple_cartpole_U.npy").reshape(-1, 1)


def _swing_up(cc, orc, mech, U, steps):
    om = orc.OracleMech(mech.desc())
    ctrl = om.make_ctrl(1, kind="table", eqc=[mech.eqc_index(mech.eqconstraints[0].id)], U=U)
    r = om.simulate(mech.state_vector()[None], steps, ctrl=ctrl, record=True)
    assert r["iters"][0] > 0                                   # every newton! converged
    return r["storage"][:, 0]


def _upright_error(s):
    """(cart offset, pole angles relative to upright in (-pi, pi], pole centre heights) of one recorded state [4][13]"""
    ang = 2 * np.arctan2(s[1:, 4], s[1:, 3])                  # rotation about x: q = (cos(a/2), sin(a/2), 0, 0)
    return s[0, 1], np.mod(ang, 2 * math.pi) - math.pi, s[1:, 2]


def test_reference_U_swings_the_triple_cartpole_up(cc, orc):
    mech, _ = cc.examples.triple_cartpole()
    U = _U()
    assert U.shape == (1000, 1) and abs(U[0, 0] + 4.10328755) < 1e-8           # first literal of the reference's vector
    sto = _swing_up(cc, orc, mech, U, 1000)
    assert np.allclose(sto[0, 1:, 2], [-0.5, -1.5, -2.5])                       # hanging at k = 1 (:39-42)
    assert np.abs(sto[:, :, [0]]).max() < 1e-12 and np.abs(sto[:, :, 5:7]).max() < 1e-12      # planar: no x motion, rotation about x only
    y, dphi, z = _upright_error(sto[-1])
    assert np.abs(dphi).max() < 0.01, dphi                                      # all three poles within 0.6 degrees of upright
    assert abs(y) < 0.01                                                        # the cart is back at the origin
    assert np.abs(z - [0.5, 1.5, 2.5]).max() < 1e-4
    # it is a genuine swing-up (the poles pass through large angles and rates), not a quasi-static lift
    assert np.abs(sto[:, 1:, 10]).max() > 4.0
    # the joints held all the way: pole i's lower end sits on pole i-1's upper end / the cart
    for k in (250, 500, 750, 999):
        s = sto[k]
        lower = [s[i, 0:3] + cc.qrot(s[i, 3:7], P2) for i in (1, 2, 3)]        # joint vertex p2 of the child (:19, :33-35)
        upper = [s[0, 0:3]] + [s[i, 0:3] + cc.qrot(s[i, 3:7], -P2) for i in (1, 2)]
        assert np.abs(np.array(lower) - np.array(upper)).max() < 1e-9


def test_the_swing_up_does_not_depend_on_the_step_size(cc, orc):
    """the same continuous input integrated with half the step (every U[k] held for two steps) arrives upright as well: the restated
    integrator converges to the trajectory U was designed for, it does not merely compensate one discretisation error with another"""
    mech, _ = cc.examples.triple_cartpole(dt=0.005)
    sto = _swing_up(cc, orc, mech, np.repeat(_U(), 2, axis=0), 2000)
    y, dphi, z = _upright_error(sto[-1])
    assert np.abs(dphi).max() < 0.01 and abs(y) < 0.01


@pytest.mark.parametrize("what,bound", [("pole mass + 1 %", 0.1), ("force scale + 1 %", 0.1), ("time step + 1 %", 0.15), ("gravity 9.80", 0.025),
                                        ("force scale + 0.1 %", 0.015)])
def test_what_the_fixture_pins(cc, orc, what, bound):
    """discriminating power: each model error below misses the upright configuration by at least `bound` rad (nominal: 0.008 rad)"""
    U = _U()
    dt = 0.0101 if what.startswith("time step") else 0.01
    mech, _ = cc.examples.triple_cartpole(dt=dt)
    if what.startswith("pole mass"):
        for b in mech.bodies[1:]:
            b.m *= 1.01
            b.J = b.J * 1.01
    elif what.startswith("force scale + 1 %"):
        U = U * 1.01
    elif what.startswith("force scale + 0.1 %"):
        U = U * 1.001
    elif what.startswith("gravity"):
        mech.g = -9.80
    sto = _swing_up(cc, orc, mech, U, 1000)
    _, dphi, _ = _upright_error(sto[-1])
    assert np.abs(dphi).max() > bound, (what, dphi)


def _deltabot_velocity_after_one_step(om, st0, eqc, u):
    ctrl = om.make_ctrl(1, kind="table", eqc=eqc, U=np.asarray(u, dtype=float).reshape(1, 2))
    return om.simulate(st0, 1, ctrl=ctrl, eps=1e-13)["states"][0][:, 7:13].ravel()


def test_deltabot_gravity_compensation_constant(cc, orc):
    """the reference's second numeric literal: `Fτd = [[[6.7879484]];[[-6.7879484]]]` (examples/lqr_deltabot.jl:53), the joint torques that hold the
    closed-chain deltabot at the example's pose against gravity (the LQR linearises about that equilibrium).  The restatement's own equilibrium
    torques -- the inputs for which one step from rest leaves every body at rest, found by Newton on the step map -- agree with the literal in all
    eight printed digits; statics of a closed chain, so this pins masses, geometry, joint vertices, g = 9.81 and the sign / axis convention of
    `setForce!` on a revolute joint between two moving bodies at the 1e-8 level."""
    mech, cfg = cc.examples.deltabot()
    om = orc.OracleMech(mech.desc())
    st0 = mech.state_vector()[None]
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    assert np.abs(_deltabot_velocity_after_one_step(om, st0, eqc, [0.0, 0.0])).max() > 0.1          # it falls without the torques
    assert np.abs(_deltabot_velocity_after_one_step(om, st0, eqc, [6.7879484, -6.7879484])).max() < 1e-8   # the reference's literal holds it
    u = np.array([6.0, -6.0])
    for _ in range(5):
        a = _deltabot_velocity_after_one_step(om, st0, eqc, u)
        h = 1e-4
        B = np.stack([(_deltabot_velocity_after_one_step(om, st0, eqc, u + [h, 0.0]) - a) / h,
                      (_deltabot_velocity_after_one_step(om, st0, eqc, u + [0.0, h]) - a) / h], axis=1)
        u = u + np.linalg.lstsq(B, -a, rcond=None)[0]
    assert np.abs(_deltabot_velocity_after_one_step(om, st0, eqc, u)).max() < 1e-11
    assert abs(u[0] + u[1]) < 1e-9                                                                  # mirror symmetry of the two legs
    assert abs(u[0] - 6.7879484) < 1e-7 and int(u[0] * 1e7) == 67879484                             # 6.78794845...: every printed digit
    # the third leg: statics by virtual work, independent of both programs.  Symmetric vertical motion of the platform (legs at a = b = 45 degrees
    # from the vertical, lower leg 1 m / 1 kg, upper leg 0.5 m / 0.5 kg, platform sqrt(2)/2 kg): U = g [(m_l + 2 m_u + m_p) cos a + (m_u / 2 + m_p / 2)
    # cos b], loop closure sin a - sin b / 2 = const gives da = db / 2, the two actuated joint angles change by -+ db: 2 tau = |dU / db|
    m_l, m_u, m_p, g = 1.0, 0.5, math.sqrt(2) / 2, 9.81
    tau = g * math.sin(math.pi / 4) * ((m_l + 2 * m_u + m_p) / 2 + (m_u + m_p) / 2) / 2
    assert abs(tau - 6.7879484) < 6e-8                      # the reference's literal is this number cut after eight digits
    assert abs(u[0] - tau) < 1e-9                           # and the restated step map has exactly this equilibrium
