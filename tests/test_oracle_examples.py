"""The reference's example scripts, end to end through the oracle, do what they are for.

The reference's test suite runs these very scripts for 100 steps and asserts `@test true` (test/example_test.jl:3-19): the only outcome
it records is "did not throw".  The scripts themselves state a purpose -- a setpoint to stabilise, a goal angle to reach -- so the
restatement can be held to more than that: each example, built from the script's literals (constrainedcontrol.jl_b200/examples.py cites the
lines), synthesised and simulated as the script does, must reach ITS OWN setpoint.  A sign or ordering error in a control law (Δz ordering
lqr.jl:92-95, `imag(qd \\ q)` :102, `Fτd - K Δz` :108, the ±2π wrap pid.jl:46-53), in `setForce!`, or in `linearsystem` shows up here as a
closed loop that does not converge.  (examples/lqr_pendulum.jl and pid_pendulum.jl are in tests/test_oracle_control.py, the open-loop
examples/trackingLQR_triple_cartpole.jl in tests/test_oracle_reference_fixture.py.)"""
import math

import numpy as np
import pytest

from helpers import oracle_lqr_gains


def _closed_loop(cc, orc, mech, cfg, steps):
    o = oracle_lqr_gains(cc, orc, mech, cfg)
    kind = "lqr_inf" if o["N"] == math.inf else "lqr_finite"
    kw = {} if kind == "lqr_inf" else dict(N=o["N"])
    ctrl = o["om"].make_ctrl(1, kind=kind, eqc=o["eqc_idx"], K=o["K"], zd=o["zd"], Fd=o["Fd"], **kw)
    st0 = mech.state_vector()[None]
    r = o["om"].simulate(st0, steps, ctrl=ctrl, record=True)
    assert r["iters"][0] > 0
    last, zd = r["storage"][-1, 0], o["zd"]
    pos = np.abs(last[:, 0:3] - zd[:, 0:3]).max()
    rot = max(np.abs(cc.qmul(cc.qconj(zd[i, 3:7]), last[i, 3:7])[1:]).max() for i in range(mech.nb))
    vel = np.abs(last[:, 7:13]).max()
    pos0 = np.abs(st0[0, :, 0:3] - zd[:, 0:3]).max()
    return pos0, pos, rot, vel


@pytest.mark.parametrize("name,steps,tol", [
    ("acrobot", 1000, 1e-5),        # examples/lqr_acrobot.jl: both links inverted, torque on the elbow only, 0.1 rad off at the start
    ("cartpole", 1000, 2e-3),       # examples/lqr_cartpole.jl: cart 0.5 m and pole 0.2 rad off
    ("prismatic", 1000, 1e-3),      # examples/lqr_prismatic.jl (minimal-coordinate weights): slider 1 m off, g = 0
])
def test_lqr_example_reaches_its_setpoint(cc, orc, name, steps, tol):
    mech, cfg = getattr(cc.examples, name)()
    pos0, pos, rot, vel = _closed_loop(cc, orc, mech, cfg, steps)
    assert pos0 > 0.09                                   # the script starts away from its setpoint ...
    assert pos < tol and rot < tol and vel < tol         # ... and the closed loop brings it there (10 s as in the script)


def test_lqr_cartpole_n_pendulum_example(cc, orc):
    # examples/lqr_cartpole_n_pendulum.jl with the test suite's N = 3 links (test/examples/lqr_cartpole_n_pendulum.jl:8), small random-like offsets (:19-21)
    mech, cfg = cc.examples.cartpole_n_pendulum(3, y0=0.3, phis=[0.02, -0.02, 0.01])
    pos0, pos, rot, vel = _closed_loop(cc, orc, mech, cfg, 1000)
    assert pos0 >= 0.3 and pos < 5e-3 and rot < 1e-3 and vel < 5e-3
    # uncontrolled, the same start falls over within the 10 s
    om = orc.OracleMech(mech.desc())
    free = om.simulate(mech.state_vector()[None], 1000)["states"][0]
    assert free[3, 2] < 2.0                               # the top pole's centre (upright: z = 2.5) has come down


def test_lqr_deltabot_example_holds_its_pose(cc, orc):
    # examples/lqr_deltabot.jl: infinite-horizon LQR about the gravity-compensated pose (closed chain, rank-deficient constraints)
    mech, cfg = cc.examples.deltabot()
    pos0, pos, rot, vel = _closed_loop(cc, orc, mech, cfg, 500)
    assert pos0 == 0.0 and pos < 1e-7 and rot < 1e-7 and vel < 1e-10
    # without the feedback (feed-forward only) the pose is an UNSTABLE equilibrium: Newton residuals at the 1e-10 level are enough to leave it
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    st0 = mech.state_vector()[None]
    ol = om.simulate(st0, 500, ctrl=om.make_ctrl(1, kind="table", eqc=eqc, U=np.tile(np.array(cfg["Ftaud"]), (500, 1))))["states"][0]
    assert np.abs(ol[4, 0:3] - st0[0, 4, 0:3]).max() > 100 * pos


def test_pid_doublependulum_example_reaches_its_goals(cc, orc):
    # examples/pid_doublependulum.jl:36-39: goals (pi/2, -pi/4) on the two revolute joints, P = 10, I = 10, D = 5, 10 s
    mech, cfg = cc.examples.doublependulum()
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    goal = np.array([[math.pi / 2, -math.pi / 4]])
    assert np.allclose(goal[0], cfg["goals"])
    ctrl = om.make_ctrl(1, kind="pid", eqc=eqc, P=cfg["P"], I=cfg["I"], D=cfg["D"], goal=goal, pid_int=np.zeros((1, 2)), pid_last=np.zeros((1, 2)))
    last = om.simulate(mech.state_vector()[None], 1000, ctrl=ctrl, record=True)["storage"][-1, 0]
    a = 2 * np.arctan2(last[:, 4], last[:, 3])                                   # absolute link angles about x
    assert abs(a[0] - goal[0, 0]) < 5e-3 and abs((a[1] - a[0]) - goal[0, 1]) < 5e-3     # joint angles = relative angles
    assert np.abs(last[:, 10]).max() < 5e-3                                        # at rest: the integral term holds the links against gravity
