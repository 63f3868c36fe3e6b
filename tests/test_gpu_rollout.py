"""GPU parity tests of the fused rollout kernel (K1) through the C-ABI against the CPU oracle.
Tolerance: 1e-9 relative on states (north_star), integer orderings bit-exact."""
import math

import numpy as np
import pytest

from helpers import cartpole_states, oracle_lqr_gains, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_graph_orderings_bit_exact(cc, orc):
    for name in ["pendulum", "cartpole", "acrobot", "deltabot", "triple_cartpole"]:
        mech, _ = getattr(cc.examples, name)()
        g = cc.handle(mech).graph()
        go = orc.OracleMech(mech.desc()).graph()
        for k in go:
            assert np.array_equal(g[k], go[k]), (name, k)


def test_pendulum_lqr_inf_c1(cc, orc):
    # BASELINE config 0: examples/lqr_pendulum.jl, 1000-step closed loop, record
    mech, cfg = cc.examples.pendulum()
    o = oracle_lqr_gains(cc, orc, mech, cfg)
    st0 = mech.state_vector()[None]
    ctrl_o = o["om"].make_ctrl(1, kind="lqr_inf", eqc=o["eqc_idx"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    ref = o["om"].simulate(st0, 1000, ctrl=ctrl_o, record=True)
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], math.inf, xd=cfg["xd"], qd=cfg["qd"])
    storage = cc.simulate(mech, 10.0, lqr, record=True)
    assert rel_err(storage.data, ref["storage"]) < TOL
    assert rel_err(mech.state_vector(), ref["states"][0]) < TOL
    th = 2 * np.arctan2(storage.q[0][:, 1], storage.q[0][:, 0])
    assert abs(th[-1] - math.pi) < 1e-7


@pytest.mark.parametrize("lanes", [0, 5, 1, 8, 32])   # auto (K1r), K1r, K1c, K1 with 8 / 32 lanes
def test_cartpole_lqr_finite_batch(cc, orc, lanes):
    mech, cfg = cc.examples.cartpole()
    o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=3.0)
    B, steps = 96, 400   # runs past the horizon: no input for k >= N (lqr.jl:106)
    st0 = cartpole_states(cc, mech, B)
    ctrl_o = o["om"].make_ctrl(B, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    ref = o["om"].simulate(st0, steps, ctrl=ctrl_o, record=True, nthreads=8)
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"])
    storage, final, status = cc.simulate(mech, steps * mech.dt, lqr, record=True, states=st0, lanes=lanes, return_status=True)
    assert np.all(status > 0)
    assert rel_err(storage.data, ref["storage"]) < TOL
    assert rel_err(final, ref["states"]) < TOL


def test_uncontrolled_chain_short_horizon(cc, orc):
    # chaotic without control: pin a short horizon only (SURVEY 7 "chaotic sensitivity")
    mech, _ = cc.examples.cartpole_n_pendulum(5)
    B = 16
    st0 = cartpole_states(cc, mech, B, seed=2, phimax=0.5)
    om = orc.OracleMech(mech.desc())
    ref = om.simulate(st0, 60, record=True)
    storage, final, status = cc.simulate(mech, 0.6, None, record=True, states=st0, return_status=True)
    assert rel_err(storage.data, ref["storage"]) < TOL
    assert rel_err(final, ref["states"]) < TOL


def test_acrobot_lqr(cc, orc):
    mech, cfg = cc.examples.acrobot()
    o = oracle_lqr_gains(cc, orc, mech, cfg)
    st0 = mech.state_vector()[None]
    ctrl_o = o["om"].make_ctrl(1, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    ref = o["om"].simulate(st0, 300, ctrl=ctrl_o, record=True)
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"], qd=cfg["qd"])
    storage, final, _ = cc.simulate(mech, 3.0, lqr, record=True, states=st0, return_status=True)
    assert rel_err(storage.data, ref["storage"]) < 1e-9   # (measured 8e-14 over the 300 steps)
    assert rel_err(final, ref["states"]) < 1e-9


def test_pid_doublependulum_sweep(cc, orc):
    # examples/pid_doublependulum.jl with per-system goals; integratederrors persist across calls (pid.jl:76)
    mech, cfg = cc.examples.doublependulum()
    B = 16
    g1, g2 = np.meshgrid(np.linspace(-2.5, 2.5, 4), np.linspace(-2.5, 2.5, 4))
    goals = np.stack([g1.ravel(), g2.ravel()], axis=1)
    st0 = mech.batch_states(B)
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    integ = np.zeros((B, 2)); last = np.zeros((B, 2))
    ctrl_o = om.make_ctrl(B, kind="pid", eqc=eqc, P=cfg["P"], I=cfg["I"], D=cfg["D"], goal=goals, pid_int=integ, pid_last=last)
    ref = om.simulate(st0, 200, ctrl=ctrl_o, record=True)
    pid = cc.PID(mech, cfg["eqcids"], goals, P=cfg["P"], I=cfg["I"], D=cfg["D"])
    storage, final, _ = cc.simulate(mech, 2.0, pid, record=True, states=st0, return_status=True)
    assert rel_err(storage.data, ref["storage"]) < TOL
    gi, gl = pid.pid_state(B)
    assert rel_err(gi, integ) < TOL and rel_err(gl, last) < TOL
    # a 1-DoF check like pid.jl:20
    mech2, _ = cc.examples.deltabot()
    with pytest.raises(AssertionError):
        cc.PID(mech2, [mech2.eqconstraints[2].id], [0.0])


def test_open_loop_table_and_recording(cc, orc):
    # examples/trackingLQR_triple_cartpole.jl:45-53: open-loop U[k] on the cart
    mech, cfg = cc.examples.triple_cartpole()
    U = np.load(__file__.replace("test_gpu_rollout.py", "golden/triple_cartpole_U.npy"))[:150]
    om = orc.OracleMech(mech.desc())
    st0 = mech.state_vector()[None]
    ctrl_o = om.make_ctrl(1, kind="table", eqc=[0], U=U.reshape(-1, 1))
    ref = om.simulate(st0, 150, ctrl=ctrl_o, record=True)
    ol = cc.OpenLoop(mech, cfg["eqcids"], U.reshape(-1, 1))
    storage = cc.Storage(150, mech.nb)
    out = cc.simulate(mech, storage, ol, record=True)
    assert out is storage
    assert rel_err(storage.data, ref["storage"]) < TOL
    assert np.array_equal(storage.x[0][0], st0[0, 0, 0:3])   # Storage[1] is the initial continuous state


@pytest.mark.parametrize("lanes", [0, 6, 1, 16])   # auto (K1r: 4 joints = 20 lanes), K1m, K1c, general kernel K1
def test_reference_U_swings_the_triple_cartpole_up(cc, orc, lanes):
    """the reference's only numeric fixture played open loop for its full 1000 steps (examples/trackingLQR_triple_cartpole.jl:45-53): the CUDA path
    ends with all three poles upright like the oracle does (tests/test_oracle_reference_fixture.py: the known answer that checks the restated
    dynamics against the dynamics upstream optimised U on) and agrees with the oracle along the whole swing-up"""
    mech, cfg = cc.examples.triple_cartpole()
    U = np.load(__file__.replace("test_gpu_rollout.py", "golden/triple_cartpole_U.npy")).reshape(1000, 1)
    om = orc.OracleMech(mech.desc())
    st0 = mech.state_vector()[None]
    ref = om.simulate(st0, 1000, ctrl=om.make_ctrl(1, kind="table", eqc=[0], U=U), record=True)
    ol = cc.OpenLoop(mech, cfg["eqcids"], U)
    storage, final, status = cc.simulate(mech, 10.0, ol, record=True, states=st0, lanes=lanes, return_status=True)
    assert status[0] > 0
    last = storage.data[-1, 0]                                  # Storage[1000]: the state at the start of the last step
    ang = 2 * np.arctan2(last[1:, 4], last[1:, 3])
    assert np.abs(np.mod(ang, 2 * math.pi) - math.pi).max() < 0.01 and abs(last[0, 1]) < 0.01
    assert np.abs(last[1:, 2] - [0.5, 1.5, 2.5]).max() < 1e-4
    assert rel_err(storage.data, ref["storage"]) < TOL
    assert rel_err(final, ref["states"]) < TOL


def test_shared_goal_pid_built_for_one_system_runs_a_batch(cc, orc):
    """ADVICE r1 (ccb_api.cu): the Julia shim's BatchPID with goals (nu,1) creates the controller with B = 1 / goal_bstride = 0 and then
    simulates a larger batch.  The per-system integrator state must be re-sized to the batch (it used to be indexed out of bounds); a
    controller with per-system tables laid out for another batch size must be refused."""
    import ctypes as C
    from ccb200 import _lib
    mech, cfg = cc.examples.doublependulum()
    B = 37
    rng = np.random.default_rng(4)
    st0 = mech.batch_states(B)
    st0[:, :, 10] = rng.uniform(-0.5, 0.5, (B, mech.nb))          # distinct initial angular velocities
    goal = np.array([[0.7, -0.4]])
    pid = cc.PID(mech, cfg["eqcids"], goal, P=cfg["P"], I=cfg["I"], D=cfg["D"])
    h1 = pid.device_handle(1, 0)                                   # built for ONE system: d.B = 1, goal_bstride = 0
    batch = cc.Batch(mech, B, 0)
    batch.set_state(st0)
    opts = _lib.SimOpts(1e-10, 100, 10, 0, 0)
    status = np.zeros(B, np.int32)
    _lib.check(_lib.lib().ccb_simulate(batch.h, h1, C.c_int32(1), C.c_int32(150), C.byref(opts), None, _lib.ptr(status)))
    final = batch.get_state()
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    integ = np.zeros((B, 2)); last = np.zeros((B, 2))
    ctrl_o = om.make_ctrl(B, kind="pid", eqc=eqc, P=cfg["P"], I=cfg["I"], D=cfg["D"], goal=np.repeat(goal, B, axis=0), pid_int=integ, pid_last=last)
    ref = om.simulate(st0, 150, ctrl=ctrl_o)
    assert np.all(status > 0)
    assert rel_err(final, ref["states"]) < TOL
    gi, gl = np.zeros((B, 2)), np.zeros((B, 2))
    _lib.check(_lib.lib().ccb_ctrl_pid_state(h1, _lib.ptr(gi), _lib.ptr(gl), C.c_int32(0)))       # now sized for the batch
    assert rel_err(gi, integ) < TOL and rel_err(gl, last) < TOL
    # per-system goals laid out for 5 systems on a batch of 37: refused
    pid5 = cc.PID(mech, cfg["eqcids"], np.tile(goal, (5, 1)) + np.arange(5)[:, None] * 0.01, P=cfg["P"], I=cfg["I"], D=cfg["D"])
    h5 = pid5.device_handle(5, 0)
    rc = _lib.lib().ccb_simulate(batch.h, h5, C.c_int32(1), C.c_int32(1), C.byref(opts), None, None)
    assert rc != 0 and b"another batch size" in _lib.lib().ccb_last_error()
