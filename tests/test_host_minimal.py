"""Host logic of the minimal-coordinate LQR constructor (reference src/control/lqr.jl:68-86): forward kinematics of joint
coordinates and rates into the maximal-coordinate setpoint.  CPU only (no compute calls)."""
import math

import numpy as np


def fk_state(cc, mech, coords, rates):
    cc.mechanism.propagate_minimal(mech, dict(enumerate(coords)), dict(enumerate(rates)))
    return mech.state_vector().copy()


def test_velocity_fk_matches_finite_differences(cc):
    # cart + 3 poles: prismatic + 3 revolutes; velocities from the analytic FK vs central differences of the position FK
    mech, _ = cc.examples.cartpole_n_pendulum(3)
    rng = np.random.default_rng(0)
    th = rng.uniform(-1, 1, 4); rate = rng.uniform(-1, 1, 4)
    z = fk_state(cc, mech, th, rate)
    assert np.abs(mech.constraint_values()).max() < 1e-14
    h = 1e-6
    zp = fk_state(cc, mech, th + h * rate, rate); zm = fk_state(cc, mech, th - h * rate, rate)
    v_fd = (zp[:, 0:3] - zm[:, 0:3]) / (2 * h)
    assert np.abs(z[:, 7:10] - v_fd).max() < 1e-8
    # body-frame angular velocity: q' = 1/2 q (x) (0, w)  ->  w = 2 vec(conj(q) (x) q')
    qd = (zp[:, 3:7] - zm[:, 3:7]) / (2 * h)
    for b in range(mech.nb):
        w = 2.0 * cc.mechanism.qmul(cc.mechanism.qconj(z[b, 3:7]), qd[b])[1:4]
        assert np.abs(z[b, 10:13] - w).max() < 1e-8


def test_minimal_ctor_asserts_like_the_reference(cc):
    import pytest
    mech, cfg = cc.examples.prismatic()
    ids = cc.getid(mech.eqconstraints)
    with pytest.raises(AssertionError, match="Missmatched length for bodies"):
        cc.LQR(mech, ids, ids, [1.0, 1.0], [1.0], 10.0)
    with pytest.raises(AssertionError, match="Missmatched length for constraints"):
        cc.LQR(mech, ids, ids, [1.0], [1.0, 2.0], 10.0)
