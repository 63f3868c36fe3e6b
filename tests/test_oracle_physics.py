"""Known-answer / invariant tests that pin the CPU oracle.

The reference's own tests are `@test true` smoke runs, so these replace the missing golden vectors."""
import math

import numpy as np
import pytest


def _unit_rows(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def test_pendulum_small_swing_period(cc, orc):
    # period 2*pi*sqrt(I_pivot/(m g d)) = 1.640 s for Box(0.1,0.1,1,1) pivoted at its end
    mech, _ = cc.examples.pendulum(phi0=0.1)
    om = orc.OracleMech(mech.desc())
    r = om.simulate(mech.state_vector()[None], 500, record=True)
    st = r["storage"][:, 0, 0]
    ang = 2 * np.arctan2(st[:, 4], st[:, 3])
    zc = np.where(np.sign(ang[:-1]) != np.sign(ang[1:]))[0]
    half = np.diff(zc)
    I_pivot = 1.0 / 12 * (0.1 ** 2 + 1.0) + 0.25
    period = 2 * math.pi * math.sqrt(I_pivot / (9.81 * 0.5))
    assert abs(period - 1.640) < 2e-3
    assert np.all(np.abs(half * 0.01 * 2 - period) < 0.02)
    # (2) constraint satisfied, unit quaternion, (3) bounded amplitude (symplectic)
    x, q = st[:, 0:3], st[:, 3:7]
    tip = np.array([x[i] + cc.qrot(q[i], [0, 0, 0.5]) for i in range(len(x))])
    assert np.abs(tip).max() < 1e-10
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 1e-12
    assert abs(ang.max() - 0.1) < 5e-4 and abs(ang.min() + 0.1) < 5e-4


def test_free_body_momentum_conservation(cc, orc):
    # torque-free asymmetric body: world-frame angular momentum R(q) J w and linear momentum are conserved
    origin = cc.Origin()
    J = np.array([[0.3, 0.02, -0.01], [0.02, 0.5, 0.03], [-0.01, 0.03, 0.9]])
    body = cc.Body(2.0, J)
    mech = cc.Mechanism(origin, [body], [], g=0.0, dt=0.01)
    body.state.vc = np.array([0.3, -0.2, 0.1])
    body.state.wc = np.array([1.0, 2.0, -1.5])
    om = orc.OracleMech(mech.desc())
    r = om.simulate(mech.state_vector()[None], 400, record=True)
    st = r["storage"][:, 0, 0]
    # discrete angular momentum of the variational integrator (exactly conserved when tau = 0):
    #   pi = R(q2) (sqrt(4/dt^2 - w2.w2) I + [w2]x) J w2 ; the recorded pair (qc, wc) is exactly (q2, w2) of the previous step
    def pi(s):
        w = s[10:13]
        sq = math.sqrt(4 / 0.01 ** 2 - w @ w)
        return cc.qrot(s[3:7], sq * (J @ w) + np.cross(w, J @ w))
    L = np.array([pi(s) for s in st])
    assert np.abs(L - L[0]).max() < 1e-10 * np.abs(L[0]).max()
    assert np.abs(st[:, 7:10] - st[0, 7:10]).max() < 1e-13
    E = np.array([0.5 * s[10:13] @ J @ s[10:13] for s in st])
    assert np.abs(E - E[0]).max() < 1e-4 * E[0]


def test_cartpole_free_axis_momentum_and_energy(cc, orc):
    # SURVEY 8(c)(3,4): sum m v_y conserved along the free prismatic axis; energy error bounded
    mech, _ = cc.examples.cartpole(y0=0.0, phi0=0.8)
    om = orc.OracleMech(mech.desc())
    r = om.simulate(mech.state_vector()[None], 1500, record=True)
    st = r["storage"][:, 0]
    m = np.array([0.5, 1.0])
    py = (m * st[:, :, 8]).sum(1)
    assert np.abs(py).max() < 1e-9
    Jp = 1.0 / 12 * np.array([0.01 + 1.0, 0.01 + 1.0, 0.02])
    Jc = 0.5 / 12 * np.array([0.25 + 0.01, 0.01 + 0.01, 0.01 + 0.25])
    E = (0.5 * (m * (st[:, :, 7:10] ** 2).sum(2)).sum(1) + 0.5 * (st[:, 0, 10:13] ** 2 * Jc).sum(1) + 0.5 * (st[:, 1, 10:13] ** 2 * Jp).sum(1)
         + 9.81 * (m * st[:, :, 2]).sum(1))
    assert np.abs(E - E[0]).max() < 0.3
    assert abs(E[-1] - E[0]) < 0.3


@pytest.mark.parametrize("name", ["cartpole", "npend3", "acrobot", "deltabot"])
def test_kkt_analytic_matches_ad_and_sparse_matches_dense(cc, orc, name):
    rng = np.random.default_rng(5)
    mech = dict(cartpole=cc.examples.cartpole, npend3=lambda: cc.examples.cartpole_n_pendulum(3), acrobot=cc.examples.acrobot,
                deltabot=lambda: cc.examples.deltabot(reg=0.0))[name]()[0]
    om = orc.OracleMech(mech.desc())
    z = rng.normal(size=(om.nb, 13)); z[:, 3:7] = _unit_rows(z[:, 3:7])
    s = rng.normal(size=(om.nb, 6)); lam = rng.normal(size=om.ml); u = rng.normal(size=om.ne)
    M, f = om.kkt(z, s, lam, u)
    assert np.abs(M - om.kkt_ad(z, s, lam, u)).max() < 1e-12
    assert np.abs(f - om.residual(z, s, lam, u)).max() == 0.0
    if name != "deltabot":   # closed chains: the un-pivoted block sweep is not defined at a random state
        rc, a, b = om.solve_both(z, s, lam, u)
        assert rc == 0
        assert np.abs(a - b).max() < 1e-9 * np.abs(a).max()
        assert np.abs(a - np.linalg.solve(M, f)).max() < 1e-9 * np.abs(a).max()


def test_graph_orderings(cc, orc):
    # restatement-defined orderings: eqcs first, DFS from the smallest id, post-order
    mech, _ = cc.examples.cartpole()
    g = orc.OracleMech(mech.desc()).graph()
    assert g["eqc_id"].tolist() == [1, 2] and g["body_id"].tolist() == [3, 4]
    assert g["dfs_list"].tolist() == [4, 2, 3, 1]
    assert g["parents"].tolist() == [0, 3, 1, 2]
    mech, _ = cc.examples.deltabot()
    g = orc.OracleMech(mech.desc()).graph()
    assert len(g["loop_edges"]) == 2
    assert sorted(g["dfs_list"].tolist()) == list(range(1, 11))


def test_integrator_is_first_order_accurate(cc, orc):
    """the reference's variational integrator is first order (README.md:14 paper): halving the step halves the error of a pendulum swing
    against a fine-step run of the same restatement"""
    import math
    T = 1.0

    def angle_at_T(dt):
        mech, _ = cc.examples.pendulum(phi0=0.5, dt=dt)
        om = orc.OracleMech(mech.desc())
        r = om.simulate(mech.state_vector()[None], int(round(T / dt)), record=False)
        assert r["iters"][0] > 0
        q = r["states"][0, 0, 3:7]
        return 2 * math.atan2(q[1], q[0])

    ref = angle_at_T(0.00025)
    e1, e2 = abs(angle_at_T(0.01) - ref), abs(angle_at_T(0.005) - ref)
    assert 1e-6 < e2 < e1 < 5e-2
    assert 1.6 < e1 / e2 < 2.5, e1 / e2
