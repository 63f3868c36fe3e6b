"""Known-answer tests pinning the oracle's linearsystem / dlqr / control laws."""
import math

import numpy as np
import pytest
import scipy.linalg


def _zd(mech, cfg):
    zd = np.zeros((mech.nb, 13))
    zd[:, 3] = 1.0
    if cfg.get("xd") is not None:
        zd[:, 0:3] = np.array(cfg["xd"])
    if cfg.get("qd") is not None:
        zd[:, 3:7] = np.array(cfg["qd"])
    return zd


def _tangent_step(cc, om, zd, dz, u, lam):
    """one-step map in tangent coordinates about zd, lambda held fixed (independent input): returns z+ state arrays"""
    nb = om.nb
    z = zd.copy()
    for b in range(nb):
        d = dz[12 * b:12 * b + 12]
        z[b, 0:3] += d[0:3]; z[b, 7:10] += d[3:6]; z[b, 10:13] += d[9:12]
        ph = d[6:9]
        z[b, 3:7] = cc.qmul(zd[b, 3:7], np.concatenate([[math.sqrt(1 - ph @ ph)], ph]))
    # solve d(z, s, u, lam) = 0 for s with lam FIXED: Newton on the body equations only
    s = z[:, 7:13].copy()
    for _ in range(30):
        M, f = om.kkt(z, s, lam, u)
        n6 = 6 * nb
        ds = np.linalg.solve(M[:n6, :n6], f[:n6])
        s = s - ds.reshape(nb, 6)
        if np.abs(ds).max() < 1e-14:
            break
    out = np.zeros((nb, 13))
    dt = om.dt
    for b in range(nb):
        out[b, 0:3] = z[b, 0:3] + dt * s[b, 0:3]
        w = s[b, 3:6]
        out[b, 3:7] = cc.qmul(z[b, 3:7], dt / 2 * np.concatenate([[math.sqrt(4 / dt ** 2 - w @ w)], w]))
        out[b, 7:13] = s[b]
    return out


def _tangent_diff(cc, zp, zm, nb):
    d = np.zeros(12 * nb)
    for b in range(nb):
        d[12 * b:12 * b + 3] = zp[b, 0:3] - zm[b, 0:3]
        d[12 * b + 3:12 * b + 6] = zp[b, 7:10] - zm[b, 7:10]
        d[12 * b + 6:12 * b + 9] = cc.qmul(cc.qconj(zm[b, 3:7]), zp[b, 3:7])[1:]
        d[12 * b + 9:12 * b + 12] = zp[b, 10:13] - zm[b, 10:13]
    return d


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "acrobot"])
def test_linearsystem_matches_central_differences(cc, orc, name):
    # A, Bu, Bl against central finite differences of the one-step map in tangent coordinates
    mech, cfg = getattr(cc.examples, name)()
    om = orc.OracleMech(mech.desc())
    zd = _zd(mech, cfg)
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    lin = om.linearize(zd, np.zeros(len(eqc_idx)), eqc_idx)
    lam = lin["lam"]
    n, h = om.n, 1e-6
    u0 = np.zeros(om.ne)
    A_fd = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n); e[j] = h
        zp = _tangent_step(cc, om, zd, e, u0, lam); zm = _tangent_step(cc, om, zd, -e, u0, lam)
        A_fd[:, j] = _tangent_diff(cc, zp, zm, om.nb) / (2 * h)
    assert np.abs(lin["A"] - A_fd).max() < 2e-6 * max(1.0, np.abs(A_fd).max())
    for i, e_i in enumerate(eqc_idx):
        up = u0.copy(); up[e_i] = h
        zp = _tangent_step(cc, om, zd, np.zeros(n), up, lam); zm = _tangent_step(cc, om, zd, np.zeros(n), -up, lam)
        assert np.abs(lin["Bu"][:, i] - _tangent_diff(cc, zp, zm, om.nb) / (2 * h)).max() < 1e-6
    for r in range(om.ml):
        lp = lam.copy(); lp[r] += h; lm = lam.copy(); lm[r] -= h
        zp = _tangent_step(cc, om, zd, np.zeros(n), u0, lp); zm = _tangent_step(cc, om, zd, np.zeros(n), u0, lm)
        assert np.abs(lin["Bl"][:, r] - _tangent_diff(cc, zp, zm, om.nb) / (2 * h)).max() < 1e-6


def test_inverted_pendulum_open_loop_instability(cc, orc):
    # rho(A - Bl (G Bl)^-1 G A) = exp(dt sqrt(m g d / I_pivot)) = 1.0391 -- needs the lambda_d Hessian term
    mech, cfg = cc.examples.pendulum()
    om = orc.OracleMech(mech.desc())
    lin = om.linearize(_zd(mech, cfg), [0.0], [0])
    assert np.allclose(lin["lam"], [0, 0, 9.81, 0, 0], atol=1e-9)
    A, Bl, G = lin["A"], lin["Bl"], lin["G"]
    rho = np.abs(np.linalg.eigvals(A - Bl @ np.linalg.solve(G @ Bl, G @ A))).max()
    I_pivot = 1.0 / 12 * (0.01 + 1.0) + 0.25
    assert abs(rho - math.exp(0.01 * math.sqrt(9.81 * 0.5 / I_pivot))) < 5e-5


def test_dlqr_unconstrained_matches_dare(orc):
    # with Bl, G empty the recursion is the textbook Riccati difference equation
    rng = np.random.default_rng(3)
    n, mu = 6, 2
    A = rng.normal(size=(n, n)) * 0.4
    Bu = rng.normal(size=(n, mu))
    Q = np.eye(n); R = np.eye(mu)
    res = orc.dlqr(A, Bu, np.zeros((n, 0)), np.zeros((0, n)), Q, R, 1000)
    P = scipy.linalg.solve_discrete_are(A, Bu, Q, R)
    K = np.linalg.solve(R + Bu.T @ P @ Bu, Bu.T @ P @ A)
    assert res["iters"] < 999
    assert np.abs(res["Ku"][0] - K).max() < 1e-4
    # back-fill (lqr.jl:179-181): all rows before the break index equal the last computed one
    assert np.array_equal(res["Ku"][0], res["Ku"][res["kexit"] - 1])


def test_dlqr_constraint_consistency(orc):
    # G * Abar = 0 for the closed loop of every iterate
    rng = np.random.default_rng(4)
    n, mu, ml = 8, 2, 3
    A = rng.normal(size=(n, n)) * 0.3; Bu = rng.normal(size=(n, mu)); Bl = rng.normal(size=(n, ml)); G = rng.normal(size=(ml, n))
    Q = np.eye(n); R = np.eye(mu)
    res = orc.dlqr(A, Bu, Bl, G, Q, R, 5)
    # recompute K_lambda for the last iterate from the definition and check G Abar = 0
    P = Q.copy()
    D = Bu - Bl @ np.linalg.solve(G @ Bl, G @ Bu)
    M = np.block([[R + D.T @ P @ Bu, D.T @ P @ Bl], [G @ Bu, G @ Bl]])
    Kk = np.linalg.solve(M, np.vstack([D.T @ P, G]) @ A)
    assert np.abs(res["Ku"][3] - Kk[:mu]).max() < 1e-10 * np.abs(Kk).max()
    Abar = A - Bu @ Kk[:mu] - Bl @ Kk[mu:]
    assert np.abs(G @ Abar).max() < 1e-12


def test_lqr_pendulum_end_to_end_known_answers(cc, orc):
    # 451 Riccati iterations, K values, rho(Abar), closed loop from pi-0.4 reaches the setpoint
    mech, cfg = cc.examples.pendulum()
    om = orc.OracleMech(mech.desc())
    zd = _zd(mech, cfg)
    lin = om.linearize(zd, [0.0], [0])
    Q = cfg["Q"][0] * mech.dt; R = cfg["R"][0] * mech.dt
    res = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Q, R, 1000)
    assert res["iters"] == 451
    K = res["Ku"][0]
    assert np.array_equal(res["Ku"][0], res["Ku"][1])      # lqr.jl:41 convergence criterion
    ref = np.array([0, -351.97, 0, 0, -13.65, 0, -313.93, 0, 0, 2.298, 0, 0])
    assert np.abs(K[0] - ref).max() < 1e-2
    ctrl = om.make_ctrl(1, kind="lqr_inf", eqc=[0], K=K[None], zd=zd, Fd=[0.0])
    u0 = om.control(ctrl, 0, 1, mech.state_vector())
    assert abs(u0[0] - 6.164) < 2e-3
    r = om.simulate(mech.state_vector()[None], 1000, ctrl=ctrl, record=True)
    st = r["storage"][:, 0, 0]
    th = 2 * np.arctan2(st[:, 4], st[:, 3])
    assert abs(abs(th[100] - math.pi) - 9.1e-2) < 5e-3
    assert abs(abs(th[300] - math.pi) - 3.5e-3) < 5e-4
    assert abs(th[999] - math.pi) < 1e-7


def test_pid_pendulum_reaches_goal(cc, orc):
    # examples/pid_pendulum.jl: P=10, I=10, D=5 drives the revolute angle to pi/2
    mech, _ = cc.examples.pendulum(phi0=0.0)
    om = orc.OracleMech(mech.desc())
    ctrl = om.make_ctrl(1, kind="pid", eqc=[0], P=[10.0], I=[10.0], D=[5.0], goal=[math.pi / 2])
    r = om.simulate(mech.state_vector()[None], 1000, ctrl=ctrl, record=True)
    q = r["states"][0, 0, 3:7]
    assert abs(2 * math.atan2(q[1], q[0]) - math.pi / 2) < 0.05


def test_newton_count_is_rounding_marginal_at_default_eps(cc, orc):
    """The per-system Newton iteration count is NOT a bit-exact contract at eps = 1e-10: the oracle's own two linear solvers (dense LU and
    block-sparse LDU on the SAME residual / Jacobian code, states equal to 1e-14) return different counts for some cartpole systems, because the
    second Newton step of a nearly linear time step is amplified rounding noise of size ~eps.  At eps = 1e-7 the counts are equal.  This is the
    CPU half of tests/test_gpu_fullsize.py::test_newton_iteration_counts_against_the_oracle."""
    from helpers import cartpole_states, oracle_lqr_gains
    mech, cfg = cc.examples.cartpole()
    B, steps = 96, 600
    st0 = cartpole_states(cc, mech, B, seed=1)
    o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=20.0)
    out = {}
    for eps in (1e-10, 1e-7):
        res = []
        for solver in (0, 1):
            co = o["om"].make_ctrl(B, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
            res.append(o["om"].simulate(st0, steps, ctrl=co, nthreads=8, solver=solver, eps=eps, marginal_band=3.0 if solver == 1 else 0.0))
        assert np.abs(res[0]["states"] - res[1]["states"]).max() < 1e-12
        out[eps] = (np.abs(res[0]["iters"] - res[1]["iters"]), res[1]["marginal"])
    assert not out[1e-7][0].any()                                # away from the noise floor the count is well defined
    d, marg = out[1e-10]
    assert d.any() and d.max() <= 2 and np.all(d <= marg)        # at the default eps it is not, exactly where the oracle flags marginal decisions


def test_counting_scalar_rollout_is_bit_identical_and_pins_the_flop_count(cc, orc):
    """"exact count from the oracle instrumented with a counting scalar type".  The Newton solve of the oracle is templated on its
    scalar; instantiated with orc::Count it performs the same operations in the same order (final states and iteration counts BIT-identical to
    the Float64 instantiation) while tallying them.  bench.py uses this count as the algorithmic FP64 work of a sim step."""
    from helpers import cartpole_states, oracle_lqr_gains
    mech, cfg = cc.examples.cartpole()
    B, steps = 4, 300
    st0 = cartpole_states(cc, mech, B, seed=1)
    o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=20.0)
    co = o["om"].make_ctrl(B, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    r = o["om"].simulate(st0, steps, ctrl=co)
    for s in range(B):
        c = o["om"].count_flops_rollout(st0[s], steps, ctrl=co, sys=s)
        assert np.array_equal(c["state"], r["states"][s]) and c["iterations"] == r["iters"][s]
        # cartpole (2 bodies, 2 joints of 5 rows): 16.5 - 19 kflop per literal Newton iteration (residual x2, analytic blocks, mixed-node LDU, substitution)
        assert 16.0e3 < c["flops_per_iteration"] < 19.5e3, c["flops_per_iteration"]
        assert c["mul"] > c["add"] > 100 * c["div"] > 0 and c["sqrt"] > 0
    one = o["om"].count_flops(st0[0])
    assert one["iterations"] >= 1 and one["flops"] == one["add"] + one["mul"] + one["div"] + one["sqrt"]
