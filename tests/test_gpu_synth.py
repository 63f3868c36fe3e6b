"""GPU parity tests of linearsystem (K2) and the constrained Riccati recursion (K3) against the CPU oracle.
Tolerance 1e-9 relative on A, Bu, Bλ, G and on the gains (north_star)."""
import math

import numpy as np
import pytest

from helpers import cartpole_states, oracle_lqr_gains, rel_err, zd_of

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "acrobot", "triple_cartpole", "npend8"])
def test_linearsystem_matches_oracle(cc, orc, name):
    mech, cfg = cc.examples.cartpole_n_pendulum(8) if name == "npend8" else getattr(cc.examples, name)()
    om = orc.OracleMech(mech.desc())
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    rng = np.random.default_rng(7)
    S = 3
    zd = np.repeat(zd_of(mech, cfg)[None], S, axis=0)
    Fd = np.zeros((S, len(eqc_idx)))
    # second setpoint: non-zero desired input and velocities (tracking-style linearisation point)
    Fd[1] = rng.normal(size=len(eqc_idx))
    zd[2, :, 7:13] = 0.05 * rng.normal(size=(mech.nb, 6))
    A, Bu, Bl, G, lam = cc.linearsystem(mech, zd, Fd, eqc_idx, want_lambda=True)
    for s in range(S):
        ref = om.linearize(zd[s], Fd[s], eqc_idx)
        assert rel_err(lam[s], ref["lam"]) < 1e-9
        assert rel_err(A[s], ref["A"]) < TOL, (name, s)
        assert rel_err(Bu[s], ref["Bu"]) < TOL
        assert rel_err(Bl[s], ref["Bl"]) < TOL
        assert rel_err(G[s], ref["G"]) < TOL


def test_dlqr_matches_oracle_random_and_unconstrained(cc, orc):
    rng = np.random.default_rng(11)
    S, n, nu, ml, N = 5, 10, 2, 3, 60
    A = rng.normal(size=(S, n, n)) * 0.3
    Bu = rng.normal(size=(S, n, nu)); Bl = rng.normal(size=(S, n, ml)); G = rng.normal(size=(S, ml, n))
    Q = np.eye(n) * 0.01; R = np.eye(nu) * 0.02
    out = cc.dlqr(A, Bu, Bl, G, Q, R, N)
    for s in range(S):
        ref = orc.dlqr(A[s], Bu[s], Bl[s], G[s], Q, R, N)
        assert out["kexit"][s] == ref["kexit"] and out["iters"][s] == ref["iters"]
        assert rel_err(out["Ku"][s], ref["Ku"]) < TOL
    out = cc.dlqr(A, Bu, np.zeros((S, n, 0)), np.zeros((S, 0, n)), Q, R, N)
    for s in range(S):
        ref = orc.dlqr(A[s], Bu[s], np.zeros((n, 0)), np.zeros((0, n)), Q, R, N)
        assert rel_err(out["Ku"][s], ref["Ku"]) < TOL


@pytest.mark.parametrize("n,nu,ml,N", [(132, 3, 20, 25), (204, 1, 85, 12), (116, 8, 0, 20)])
def test_dlqr_large_state_blocked_tensor_core_sweep(cc, orc, n, nu, ml, N):
    """n > 108 runs on K4b (P, Abar in the global workspace, blocked DMMA products): same gains and exit index as the oracle"""
    rng = np.random.default_rng(n)
    S = 3
    A = rng.normal(size=(S, n, n)) * (0.9 / np.sqrt(n))
    Bu = rng.normal(size=(S, n, nu)); G = rng.normal(size=(S, ml, n))
    Bl = np.swapaxes(G, 1, 2) + 0.1 * rng.normal(size=(S, n, ml))          # G Bl ~ G G': a well conditioned projection
    L = rng.normal(size=(n, n)); Q = (L @ L.T / n + np.eye(n)) * 0.01; R = np.eye(nu) * 0.02
    out = cc.dlqr(A, Bu, Bl, G, Q, R, N)
    for s in range(S):
        ref = orc.dlqr(A[s], Bu[s], Bl[s], G[s], Q, R, N)
        assert out["kexit"][s] == ref["kexit"] and out["iters"][s] == ref["iters"]
        assert rel_err(out["Ku"][s], ref["Ku"]) < TOL


def test_lqr_pendulum_reference_example(cc, orc):
    # examples/lqr_pendulum.jl end to end through the reference API: LQR(...) then simulate!(mech, 10, lqr, record=true)
    mech, cfg = cc.examples.pendulum()
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cc.getid(mech.eqconstraints), cfg["Q"], cfg["R"], math.inf, xd=cfg["xd"], qd=cfg["qd"])
    o = oracle_lqr_gains(cc, orc, mech, cfg)
    assert lqr.riccati_iters[0] == 451 and bool(lqr.converged[0])      # known answer and lqr.jl:41
    assert rel_err(lqr.K[0], o["K"]) < TOL
    ref = np.array([0, -351.97, 0, 0, -13.65, 0, -313.93, 0, 0, 2.298, 0, 0])
    assert np.abs(lqr.K[0, 0, 0] - ref).max() < 1e-2
    storage = cc.simulate(mech, 10.0, lqr, record=True)
    th = 2 * np.arctan2(storage.q[0][:, 1], storage.q[0][:, 0])
    assert abs(th[-1] - math.pi) < 1e-7


def test_lqr_cartpole_batched_weights(cc, orc):
    # synthesis batch with distinct (Q-scale, R-scale) per system (SURVEY 8(d) C3 style), finite horizon
    mech, cfg = cc.examples.cartpole()
    rng = np.random.default_rng(2)
    S = 6
    qs = np.exp(rng.uniform(np.log(0.1), np.log(10), S)); rs = np.exp(rng.uniform(np.log(0.1), np.log(10), S))
    Qd = np.stack([cc.blockdiag(cfg["Q"]) * q * mech.dt for q in qs]); Rd = np.stack([cc.blockdiag(cfg["R"]) * r * mech.dt for r in rs])
    zd = zd_of(mech, cfg)
    N = 400
    out = cc.lqr_synthesize(mech, np.repeat(zd[None], S, 0), np.zeros((S, 1)), [0], Qd, Rd, N)
    om = orc.OracleMech(mech.desc())
    lin = om.linearize(zd, [0.0], [0])
    for s in range(S):
        ref = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Qd[s], Rd[s], N)
        assert out["kexit"][s] == ref["kexit"]
        assert rel_err(out["Ku"][s], ref["Ku"]) < TOL


def test_tracking_lqr_triple_cartpole(cc, orc):
    # examples/trackingLQR_triple_cartpole.jl: open-loop trajectory from U, then TrackingLQR about it (short horizon here)
    mech, cfg = cc.examples.triple_cartpole()
    N = 120
    U = np.load(__file__.replace("test_gpu_synth.py", "golden/triple_cartpole_U.npy"))[:N].reshape(N, 1)
    ol = cc.OpenLoop(mech, cfg["eqcids"], U)
    storage0 = cc.Storage(N, mech.nb)
    st_init = mech.state_vector().copy()
    cc.simulate(mech, storage0, ol, record=True)
    tl = cc.TrackingLQR(mech, storage0, U.reshape(1, N, 1), cfg["eqcids"], cfg["Q"], cfg["R"])
    om = orc.OracleMech(mech.desc())
    ref = om.tracking_dlqr(storage0.data[:, 0], U, [0], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt)
    assert tl.kexit[0] == ref["kexit"]
    assert rel_err(tl.K[0], ref["Ku"]) < 1e-9
    # closed loop about the trajectory from a perturbed start, GPU vs oracle with the same gains
    mech.set_state_vector(st_init)
    st0 = mech.state_vector()[None].copy()
    st0[0, 0, 1] += 0.01
    ctrl_o = om.make_ctrl(1, kind="tracking", eqc=[0], N=N, K=tl.K[0], zd=storage0.data[:, 0], Fd=U)
    r = om.simulate(st0, N, ctrl=ctrl_o, record=True)
    out, final, _ = cc.simulate(mech, N * mech.dt, tl, record=True, states=st0, return_status=True)
    assert rel_err(out.data, r["storage"]) < 1e-9


def test_lqr_per_system_setpoints_device_resident(cc, orc):
    # LQR(...) with one setpoint per system: gains synthesised and kept on the device (ccb_lqr_create), rollout vs the oracle
    mech, cfg = cc.examples.cartpole()
    S = 4
    ys = np.array([-0.3, 0.0, 0.2, 0.45])
    xd = np.zeros((S, 2, 3)); xd[:, 0, 1] = ys; xd[:, 1, 1] = ys; xd[:, 1, 2] = 0.5
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 2.0, xd=xd)
    assert lqr.K.shape == (S, 199, 1, 24)
    om = orc.OracleMech(mech.desc())
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
    st0 = cartpole_states(cc, mech, S, seed=5)
    for s in range(S):
        lin = om.linearize(lqr.zd[s], [0.0], [0])
        ref = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Qd, Rd, 200)
        assert rel_err(lqr.K[s], ref["Ku"]) < TOL
        ctrl_o = om.make_ctrl(1, kind="lqr_finite", eqc=[0], N=200, K=ref["Ku"], zd=lqr.zd[s], Fd=np.zeros(1))
        r = om.simulate(st0[s:s + 1], 250, ctrl=ctrl_o, record=True)
        if s == 0:
            refs = np.zeros((250, S, 2, 13))
        refs[:, s] = r["storage"][:, 0]
    storage, final, status = cc.simulate(mech, 2.5, lqr, record=True, states=st0, return_status=True)
    assert np.all(status > 0)
    assert rel_err(storage.data, refs) < TOL


def test_lqr_minimal_coordinate_ctor_prismatic_example(cc, orc):
    # examples/lqr_prismatic.jl: LQR(mech, getid.(constraints), getid.(constraints), ones(1), ones(1), 10.) then simulate!(mech, 10., lqr)
    mech, cfg = cc.examples.prismatic()
    ids = cc.getid(mech.eqconstraints)
    lqr = cc.LQR(mech, ids, ids, np.ones(1), np.ones(1), 10.0)        # setpoint: joint coordinate 0 -> link at the origin (lqr.jl:70-71,80)
    assert np.allclose(lqr.zd[0, 0, 0:3], 0.0) and lqr.N == 1000
    om = orc.OracleMech(mech.desc())
    lin = om.linearize(lqr.zd[0], [0.0], [0])
    ref = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], np.eye(12) * mech.dt, np.eye(1) * mech.dt, 1000)     # Q = I12 * 1, R = 1 (lqr.jl:82-83)
    assert rel_err(lqr.K[0], ref["Ku"]) < TOL
    st0 = mech.state_vector()[None].copy()                              # link displaced by 1 m along the joint axis
    co = om.make_ctrl(1, kind="lqr_finite", eqc=[0], N=1000, K=ref["Ku"], zd=lqr.zd[0], Fd=np.zeros(1))
    r = om.simulate(st0, 1000, ctrl=co, record=True)
    storage = cc.simulate(mech, 10.0, lqr, record=True)
    assert rel_err(storage.data, r["storage"]) < TOL
    assert abs(mech.bodies[0].state.xc[0]) < 1e-2                       # regulated to the origin
    # a non-zero joint setpoint on the cartpole: cart at y = 0.3, pole upright
    mech2, cfg2 = cc.examples.cartpole()
    ids2 = cc.getid(mech2.eqconstraints)
    lqr2 = cc.LQR(mech2, ids2, [ids2[0]], np.ones(2), np.ones(1), 3.0, xtheta_d=[0.3, 0.0])
    assert np.allclose(lqr2.zd[0, 0, 0:3], [0, 0.3, 0]) and np.allclose(lqr2.zd[0, 1, 0:3], [0, 0.3, 0.5])
    lin2 = orc.OracleMech(mech2.desc()).linearize(lqr2.zd[0], [0.0], [0])
    ref2 = orc.dlqr(lin2["A"], lin2["Bu"], lin2["Bl"], lin2["G"], np.eye(24) * mech2.dt, np.eye(1) * mech2.dt, 300)
    assert rel_err(lqr2.K[0], ref2["Ku"]) < TOL


def test_peer_gain_table_single_rank_matches_synthesis_and_drives_a_rollout(cc, orc):
    """ccb_lqr_synthesize_peers with one table (no process group): the sweep writes the gain rows straight into the device-resident
    table; bit-identical to ccb_lqr_synthesize, and a controller built on the table reproduces the LQR rollout"""
    mech, cfg = cc.examples.cartpole()
    S, N = 6, 120
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    zd = np.repeat(zd_of(mech, cfg)[None], S, 0)
    ys = np.linspace(-0.3, 0.3, S); zd[:, :, 1] += ys[:, None]
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
    ref = cc.lqr_synthesize(mech, zd, np.zeros((S, 1)), eqc_idx, Qd, Rd, N)
    tab = cc.dist.PeerGainTable(mech, zd, np.zeros((S, 1)), eqc_idx, Qd, Rd, N)
    assert np.array_equal(tab.read(), ref["Ku"]) and np.array_equal(tab.kexit, ref["kexit"])
    st0 = cartpole_states(cc, mech, S, seed=9)
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], ref["Ku"], N, xd=zd[:, :, 0:3])
    a = cc.simulate(mech, 1.5, lqr, record=False, states=st0)
    b = cc.simulate(mech, 1.5, tab.controller(), record=False, states=st0)
    assert np.array_equal(a[1] if isinstance(a, tuple) else a, b[1] if isinstance(b, tuple) else b)
    # a shape outside the tensor-core sweep (n = 12 nb with nb = 9 -> 108 is K4; the double pendulum's n = 24 is too: use CCB_NO_DMMA elsewhere) --
    # here: the large-n path copies the finished shard instead
    tab.free()


def test_tracking_synthesis_into_a_device_table(cc, orc):
    """ccb_tracking_synthesize_peers with one table: the gain rows of the TrackingLQR sweep (pipelined over step blocks at this horizon) land in
    the device-resident table at the given system offset, bit-identical to ccb_tracking_synthesize"""
    import ctypes as C
    from ccb200 import _lib
    mech, cfg = cc.examples.triple_cartpole()
    N = 300
    U = np.load(__file__.replace("test_gpu_synth.py", "golden/triple_cartpole_U.npy"))[:N].reshape(N, 1)
    ol = cc.OpenLoop(mech, cfg["eqcids"], U)
    sto = cc.Storage(N, mech.nb)
    cc.simulate(mech, sto, ol, record=True)
    tl = cc.TrackingLQR(mech, sto, U.reshape(1, N, 1), cfg["eqcids"], cfg["Q"], cfg["R"])
    Kref = tl.K[0]
    n, nu = 12 * mech.nb, 1
    per = (N - 1) * nu * n
    L = _lib.lib()
    tab = C.c_void_p()
    _lib.check(L.ccb_gains_alloc(C.c_int32(0), C.c_int64(3 * per), C.byref(tab), None))
    arr = (C.c_void_p * 1)(tab.value)
    traj = np.ascontiguousarray(sto.data[:, 0]).reshape(1, N, mech.nb, 13)
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
    eqc = _lib.i32([mech.eqc_index(i) for i in cfg["eqcids"]])
    kx = np.zeros(1, np.int32); it = np.zeros(1, np.int32)
    Ft = np.ascontiguousarray(U.reshape(1, N, 1))
    _lib.check(L.ccb_tracking_synthesize_peers(cc.handle(mech).h, C.c_int32(0), C.c_int32(1), C.c_int32(N), _lib.ptr(traj), _lib.ptr(Ft), _lib.ptr(eqc),
                                              C.c_int32(nu), _lib.ptr(Qd), _lib.ptr(Rd), None, C.c_int32(1), arr, C.c_int64(2), _lib.ptr(kx), _lib.ptr(it)))
    out = np.empty((N - 1, nu, n))
    _lib.check(L.ccb_gains_read(C.c_int32(0), tab, C.c_int64(2 * per), _lib.ptr(out), C.c_int64(per)))
    _lib.check(L.ccb_gains_free(tab))
    assert kx[0] == tl.kexit[0]
    assert np.array_equal(out, Kref)


@pytest.mark.parametrize("n", [10, 24, 132])
def test_gains_invariant_under_constraint_basis_and_multiplier_scaling_on_the_device(cc, n):
    """SURVEY 8(c) item 9 on the product (K3 / K4 / K4b): Ku does not depend on the basis of the constraint rows or the scaling of the multipliers"""
    rng = np.random.default_rng(n)
    mu, ml, N = 2, 5, 30
    A = rng.normal(size=(n, n)) * (0.8 / np.sqrt(n))
    Bu = rng.normal(size=(n, mu)); Bl = rng.normal(size=(n, ml)); G = rng.normal(size=(ml, n))
    Q = np.eye(n) * 0.02; R = np.eye(mu) * 0.05
    T = rng.normal(size=(ml, ml)) + 3 * np.eye(ml)
    Sc = np.diag(10.0 ** rng.uniform(-2, 2, ml)) @ (rng.normal(size=(ml, ml)) + 3 * np.eye(ml))
    a = cc.dlqr(A[None], Bu[None], Bl[None], G[None], Q, R, N)
    b = cc.dlqr(A[None], Bu[None], (Bl @ Sc)[None], (T @ G)[None], Q, R, N)
    assert a["kexit"][0] == b["kexit"][0]
    assert rel_err(b["Ku"][0], a["Ku"][0]) < 1e-9


def test_peer_tracking_table_matches_tracking_lqr(cc):
    """dist.PeerTrackingTable (single rank here): the TrackingLQR gain tables of 3 reference trajectories in a device-resident table, and a
    controller over it (2 perturbed systems per trajectory) -- bit-identical to TrackingLQR + replicate"""
    mech, cfg = cc.examples.triple_cartpole()
    N, S, PER = 200, 3, 2
    U0 = np.load(__file__.replace("test_gpu_synth.py", "golden/triple_cartpole_U.npy"))[:N]
    U = np.stack([U0 * sc for sc in (1.0, 0.8, 1.1)]).reshape(S, N, 1)
    st0 = mech.batch_states(S)
    ol = cc.OpenLoop(mech, cfg["eqcids"], U)
    sto, fin, status = cc.simulate(mech, N * mech.dt, ol, record=True, states=st0, return_status=True)
    assert np.all(status > 0)
    tl = cc.TrackingLQR(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"])
    tab = cc.dist.PeerTrackingTable(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"])
    assert np.array_equal(tab.read(), tl.K) and np.array_equal(tab.kexit, tl.kexit)
    stp = np.repeat(st0, PER, axis=0); stp[:, 0, 1] += np.linspace(-0.01, 0.01, S * PER)
    a = cc.simulate(mech, N * mech.dt, tl.replicate(PER), record=False, states=stp)
    b = cc.simulate(mech, N * mech.dt, tab.controller().replicate(PER), record=False, states=stp)
    fa = a[1] if isinstance(a, tuple) else a; fb = b[1] if isinstance(b, tuple) else b
    assert np.array_equal(fa, fb)
    tab.free()
