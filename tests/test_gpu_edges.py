"""Edge cases of the C-ABI rollout / synthesis entry points on the GPU: empty and ragged batches, split rollouts (k0), error
returns (the reference's @assert messages), per-system tables."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import cartpole_states, oracle_lqr_gains, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def cartpole_setup(cc, orc):
    mech, cfg = cc.examples.cartpole()
    o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=3.0)
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"])
    return mech, cfg, o, lqr


@pytest.mark.parametrize("B", [1, 2, 3, 4, 7])
def test_ragged_batches_fill_partial_warps(cc, orc, cartpole_setup, B):
    # K1r packs 3 cartpoles per warp: batches that leave lanes or whole system slots idle
    mech, cfg, o, lqr = cartpole_setup
    st0 = cartpole_states(cc, mech, 8, seed=9)[:B]
    co = o["om"].make_ctrl(B, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    r = o["om"].simulate(st0, 120, ctrl=co, record=True)
    sto, fin, status = cc.simulate(mech, 1.2, lqr, record=True, states=st0, return_status=True)
    assert np.all(status > 0) and rel_err(sto.data, r["storage"]) < TOL and rel_err(fin, r["states"]) < TOL


def test_zero_steps_and_split_rollout(cc, orc, cartpole_setup):
    mech, cfg, o, lqr = cartpole_setup
    B = 6
    st0 = cartpole_states(cc, mech, B, seed=4)
    batch = cc.Batch(mech, B)
    batch.set_state(st0)
    sto, status = batch.simulate(lqr, 1, 0, record=True)           # simulate!(mech, 0, ...): nothing happens
    assert sto.shape[0] == 0 and np.array_equal(batch.get_state(), st0) and np.all(status == 0)
    # one 200-step rollout == 120 steps followed by 80 steps with k0 = 121 (gain row K[k], lambda warm start and state carry over)
    batch.simulate(lqr, 1, 200)
    full = batch.get_state(want_lambda=True)
    batch.set_state(st0)
    batch.simulate(lqr, 1, 120)
    batch.simulate(lqr, 121, 80)
    split = batch.get_state(want_lambda=True)
    assert np.array_equal(full[0], split[0]) and np.array_equal(full[1], split[1])


def test_storage_semantics(cc, cartpole_setup):
    # Storage records (xc,qc,vc,wc) at the START of each step: entry 0 is the initial state, the final state is not in it
    mech, cfg, o, lqr = cartpole_setup
    st0 = cartpole_states(cc, mech, 5, seed=6)
    sto, fin, _ = cc.simulate(mech, 0.5, lqr, record=True, states=st0, return_status=True)
    assert np.array_equal(sto.data[0], st0)
    sto2, fin2, _ = cc.simulate(mech, 0.51, lqr, record=True, states=st0, return_status=True)     # ceil(tend / dt) = 51 steps
    assert sto2.data.shape[0] == 51 and np.array_equal(sto2.data[50], fin)


def test_per_system_gain_tables(cc, orc, cartpole_setup):
    # K_bstride != 0: every system its own table (here: the shared table scaled per system) through ccb_ctrl_create
    mech, cfg, o, lqr = cartpole_setup
    B = 5
    scale = np.linspace(0.8, 1.2, B)
    Kb = o["K"][None] * scale[:, None, None, None]
    st0 = cartpole_states(cc, mech, B, seed=8)
    ctrl = cc.LQR.from_gains(mech, cfg["eqcids"], Kb, o["N"], xd=cfg["xd"])
    sto, fin, status = cc.simulate(mech, 1.0, ctrl, record=True, states=st0, return_status=True)
    for s in range(B):
        co = o["om"].make_ctrl(1, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=Kb[s], zd=o["zd"], Fd=o["Fd"])
        r = o["om"].simulate(st0[s:s + 1], 100, ctrl=co, record=True)
        assert rel_err(sto.data[:, s], r["storage"][:, 0]) < TOL, s


def test_error_returns(cc, cartpole_setup):
    mech, cfg, o, lqr = cartpole_setup
    L = cc._lib.lib()
    # kernel selection that does not apply to the mechanism
    sw, cfg_s = cc.examples.sawyer(__file__.replace("test_gpu_edges.py", "golden/sawyer_arm.json"))
    with pytest.raises(cc.CcbError, match="row-per-lane"):
        cc.simulate(sw, 0.02, None, states=sw.batch_states(2), lanes=5)
    mech_d, _ = cc.examples.deltabot()
    with pytest.raises(cc.CcbError, match="serial chain"):
        cc.simulate(mech_d, 0.02, None, states=mech_d.batch_states(2), lanes=6)
    with pytest.raises(cc.CcbError, match="lanes per system"):
        cc.simulate(mech, 0.02, None, states=mech.batch_states(2), lanes=3)
    # controller laid out for another batch size
    Kb = np.repeat(o["K"][None], 4, axis=0)
    ctrl4 = cc.LQR.from_gains(mech, cfg["eqcids"], Kb, o["N"], xd=cfg["xd"])
    with pytest.raises(cc.CcbError):
        cc.simulate(mech, 0.02, ctrl4, states=mech.batch_states(6))
    # the reference's assertion messages
    with pytest.raises(AssertionError, match="Missmatched length for bodies"):
        cc.LQR(mech, cc.getid(mech.bodies)[:1], cfg["eqcids"], cfg["Q"], cfg["R"], 1.0)
    with pytest.raises(AssertionError, match="Missmatched length for constraints"):
        cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"] + cfg["R"], 1.0)
    # NULL / bad arguments straight at the ABI: error code + message, never a crash
    assert L.ccb_simulate(None, None, C.c_int32(1), C.c_int32(1), None, None, None) == -1
    assert b"NULL" in L.ccb_last_error() or b"bad" in L.ccb_last_error()
    assert L.ccb_dlqr(C.c_int32(0), C.c_int32(0), C.c_int32(4), C.c_int32(1), C.c_int32(0), C.c_int32(5), None, None, None, None, None, C.c_int64(0), None,
                      C.c_int64(0), None, None, None) == -1
    # more controlled inputs than the Riccati kernels support
    n, nu = 12, 9
    A = np.eye(n)[None]; Bu = np.ones((1, n, nu)); Q = np.eye(n); R = np.eye(nu)
    with pytest.raises(cc.CcbError, match="8 controlled"):
        cc.dlqr(A, Bu, np.zeros((1, n, 0)), np.zeros((1, 0, n)), Q, R, 5)
    # no such device
    with pytest.raises(cc.CcbError, match="CUDA device"):
        cc.lqr_synthesize(mech, o["zd"][None], np.zeros((1, 1)), [0], np.eye(24), np.eye(1), 5, device=99)


def test_peer_gain_tables_two_ranks(cc):
    """2 GPUs: the Riccati sweep of each rank stores its gain rows into both ranks' tables (CUDA IPC, NVLink); tools/peer_check.py asserts
    bit-identity with the NCCL all-gather and with a local synthesis.  Skipped on single-GPU boxes."""
    import subprocess, sys, os
    if cc.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(root, "tools", "peer_check.py"), "cartpole", "64", "300"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("table == gathered: True") == 2
    # a shard of more than one wave of sweep CTAs (200 Sawyer syntheses per rank): the group-pipelined copy-engine exchange
    cmd = cmd[:-4] + [os.path.join(root, "tools", "peer_check.py"), "sawyer", "400", "120"]
    cmd[cmd.index("29541")] = "29542"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("table == gathered: True") == 2 and r.stdout.count("peer-written slice == local synthesis: True") == 2
    # TrackingLQR tables of 8 Sawyer trajectories sharded over the two ranks (in-kernel peer stores), refilled once
    cmd = cmd[:-4] + [os.path.join(root, "tools", "peer_tracking_check.py"), "8", "300"]
    cmd[cmd.index("29542")] = "29543"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("tables bit-identical across ranks: True") == 2


def test_record_pipeline_is_bit_identical_to_one_launch(cc, monkeypatch):
    """ccb_simulate with a host trajectory buffer cuts the rollout into chunks (device double buffer drained by a copy stream while the next chunk
    computes).  The kernels resume from the stored state / multipliers / PID integrators / step index, so the recorded trajectory, the final
    state and the per-system Newton counts (accumulated over the launches) equal those of ONE launch bit for bit -- LQR (finite horizon: the
    `k < N` gate crosses a chunk boundary) and PID (persistent integrator)."""
    mech, cfg = cc.examples.cartpole()
    B, steps = 300, 260
    st0 = cartpole_states(cc, mech, B, seed=5)
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 2.0, xd=cfg["xd"])      # N = 200 < steps
    mech2, cfg2 = cc.examples.doublependulum()
    goals = np.stack([np.linspace(-1, 1, 64), np.linspace(1, -1, 64)], axis=1)
    out = {}
    for mode in ("one", "chunks"):
        if mode == "one":
            monkeypatch.setenv("CCB_RECORD_NO_PIPELINE", "1")
        else:
            monkeypatch.delenv("CCB_RECORD_NO_PIPELINE")
            monkeypatch.setenv("CCB_RECORD_CHUNK_MB", "1")       # 1 MB chunks: 16 steps of the cartpole batch, 78 of the pendulum batch
        sto, fin, status = cc.simulate(mech, steps * mech.dt, lqr, record=True, states=st0, return_status=True)
        pid = cc.PID(mech2, cfg2["eqcids"], goals, P=cfg2["P"], I=cfg2["I"], D=cfg2["D"])
        sto2, fin2, status2 = cc.simulate(mech2, 4.0, pid, record=True, states=mech2.batch_states(64), return_status=True)
        out[mode] = (sto.data.copy(), fin, status, sto2.data.copy(), fin2, status2, pid.pid_state(64))
    a, b = out["one"], out["chunks"]
    for x, y in zip(a[:6], b[:6]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[6][0], b[6][0]) and np.array_equal(a[6][1], b[6][1])
    assert np.all(a[2] > 0) and np.all(a[5] != 0)


def test_persistent_time_sliced_launch_is_bit_identical(cc, monkeypatch):
    """K1r's persistent launch (a batch of 1184 < W <= 1.35 x 1184 warps: two warps per SM sub-partition work through (chunk of steps, triple of
    systems) items on a static schedule) must give the bits of the plain launch: states, multipliers and accumulated Newton counts, for LQR (finite horizon ending
    inside the rollout), open-loop tables and the uncontrolled mechanism.  B = 3700 cartpoles -> 1234 warps."""
    mech, cfg = cc.examples.cartpole()
    B, steps = 3700, 272                                         # chunks of 17 steps, the last one ragged
    st0 = cartpole_states(cc, mech, B, seed=11)
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 2.0, xd=cfg["xd"])       # N = 200 < steps
    U = 0.3 * np.sin(np.arange(steps)[:, None] * 0.05)
    out = {}
    for mode in ("plain", "persistent"):
        if mode == "plain":
            monkeypatch.setenv("CCB_K1R_NO_PERSIST", "1")
        else:
            monkeypatch.delenv("CCB_K1R_NO_PERSIST")
        res = []
        for ctrl in (lqr, cc.OpenLoop(mech, cfg["eqcids"], U), None):
            batch = cc.Batch(mech, B)
            batch.set_state(st0)
            n0 = cc._lib.lib().ccb_launch_count()
            _, status = batch.simulate(ctrl, 1, steps)
            fin, lam = batch.get_state(want_lambda=True)
            res.append((fin, lam, status))
        out[mode] = res
    for a, b in zip(out["plain"], out["persistent"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.all(a[2] > 0)


@pytest.mark.parametrize("B", [3552, 3553, 4794, 4795])
def test_launch_selection_edges_against_the_oracle(cc, orc, B):
    """the batch sizes at which K1r's launcher changes route on a 148-SM device -- 3552 cartpoles = 1184 warps (whole batch on the 255-register
    instantiation), 3553 (time-sliced, ONE extra triple), 4794 (time-sliced, the last batch inside the 1.35x window: 414 extra triples),
    4795 (plain launch) -- each against the oracle: states, multiplier-free failure flags and Newton totals of every system"""
    mech, cfg = cc.examples.cartpole()
    steps = 260                                                   # >= 256: the persistent routes are eligible; chunks of 33 steps, the last ragged
    o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=2.0)          # N = 200 < steps: the k < N gate closes inside a chunk
    st0 = cartpole_states(cc, mech, B, seed=13)
    ctrl_o = o["om"].make_ctrl(B, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    ref = o["om"].simulate(st0, steps, ctrl=ctrl_o, nthreads=8, eps=1e-7)
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"])
    batch = cc.Batch(mech, B)
    batch.set_state(st0)
    _, status = batch.simulate(lqr, 1, steps, eps=1e-7)
    assert np.all(status > 0) and np.all(ref["iters"] > 0)
    assert rel_err(batch.get_state(), ref["states"]) < 1e-9
    assert np.array_equal(status, ref["iters"])                   # at eps = 1e-7 the iteration count is a well-defined integer (DESIGN "Integer outputs")


def test_whole_batch_launch_on_the_spill_free_instantiation_is_bit_identical(cc, monkeypatch):
    """batches of at most two warps per scheduler run K1r's 255-register instantiation in one chunk (no time slicing): the bits of the plain
    168-register kernel for LQR, open-loop tables, no controller and PID (whose integrator state lives on the device)"""
    mech, cfg = cc.examples.cartpole()
    B, steps = 901, 300
    st0 = cartpole_states(cc, mech, B, seed=12)
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 2.0, xd=cfg["xd"])
    U = 0.3 * np.sin(np.arange(steps)[:, None] * 0.05)
    dp, dcfg = cc.examples.doublependulum()
    Bp = 77
    goals = np.stack([np.linspace(-2.0, 2.0, Bp), np.linspace(1.5, -1.5, Bp)], axis=1)
    out = {}
    for mode in ("plain", "whole"):
        if mode == "plain":
            monkeypatch.setenv("CCB_K1R_NO_PERSIST", "1")
        else:
            monkeypatch.delenv("CCB_K1R_NO_PERSIST")
        res = []
        for ctrl in (lqr, cc.OpenLoop(mech, cfg["eqcids"], U), None):
            batch = cc.Batch(mech, B)
            batch.set_state(st0)
            _, status = batch.simulate(ctrl, 1, steps)
            fin, lam = batch.get_state(want_lambda=True)
            res.append((fin, lam, status))
        pid = cc.PID(dp, dcfg["eqcids"], goals, P=dcfg["P"], I=dcfg["I"], D=dcfg["D"])
        batch = cc.Batch(dp, Bp)
        batch.set_state(dp.batch_states(Bp))
        _, status = batch.simulate(pid, 1, steps)
        _, status2 = batch.simulate(pid, steps + 1, 40)               # a short second call (plain kernel in both modes) continues from the PID state
        fin, lam = batch.get_state(want_lambda=True)
        gi, gl = pid.pid_state(Bp)
        res.append((fin, lam, status, status2, gi, gl))
        out[mode] = res
    for a, b in zip(out["plain"], out["whole"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.all(a[2] > 0)


@pytest.mark.parametrize("B", [700, 1300])
def test_persistent_launch_of_the_long_chain_kernel_is_bit_identical(cc, monkeypatch, B):
    """K1m's persistent time-sliced launch (a batch beyond one wave of the 255-register variant: 592 CTAs simulate B Sawyer arms in chunks of
    steps; B = 700: one own system per CTA + 108 rotating extras, B = 1300: two own systems + 116 extras) gives the bits of the plain launch."""
    import os
    mech, cfg = cc.examples.sawyer(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sawyer_arm.json"))
    steps = 272
    rng = np.random.default_rng(17)
    st0 = np.zeros((B, mech.nb, 13))
    for s in range(B):
        cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)})
        st0[s] = mech.state_vector()
    cc.urdf.zero_configuration(mech)
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 2.0, xd=cfg["xd"], qd=cfg["qd"])     # N = 200 < steps
    out = {}
    for mode in ("plain", "persistent"):
        if mode == "plain":                                   # the same 255-register variant, launched plainly (two or three waves)
            monkeypatch.setenv("CCB_K1M_NO_PERSIST", "1")
            monkeypatch.setenv("CCB_K1M_REGS", "255")          # (the register variants differ from each other at the 1e-14 level: ptxas fuses
        else:                                                  # multiplies and adds differently under different caps)
            monkeypatch.delenv("CCB_K1M_NO_PERSIST")
            monkeypatch.delenv("CCB_K1M_REGS")
        res = []
        for ctrl in (lqr, None):
            batch = cc.Batch(mech, B)
            batch.set_state(st0)
            _, status = batch.simulate(ctrl, 1, steps)
            fin, lam = batch.get_state(want_lambda=True)
            res.append((fin, lam, status))
        out[mode] = res
    for a, b in zip(out["plain"], out["persistent"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.all(a[2] > 0)


def test_gains_backfill_completes_rows_below_the_break_index(cc):
    """ccb_gains_backfill (deferred back-fill of the multi-GPU exchange): rows 0 .. kexit-2 of every system become copies of row kexit-1
    (lqr.jl:179-181), everything else is untouched; break indices outside 1 .. N-1 are refused"""
    import ctypes as C
    import torch
    from ccb200 import _lib
    L = _lib.lib()
    for nu, n in ((2, 12), (7, 84), (1, 3)):          # even rows (16-byte kernel) and an odd row length (copy path)
        S, N = 5, 9
        row = nu * n
        ref = np.arange(S * (N - 1) * row, dtype=np.float64).reshape(S, N - 1, row)
        t = torch.from_numpy(ref.copy()).cuda()
        kexit = np.array([1, 3, 8, 2, 5], np.int32)
        _lib.check(L.ccb_gains_backfill(C.c_int32(0), C.c_void_p(t.data_ptr()), C.c_int64(S), C.c_int32(N), C.c_int32(nu), C.c_int32(n), _lib.ptr(kexit)))
        out = t.cpu().numpy()
        for s in range(S):
            k = int(kexit[s])
            assert np.array_equal(out[s, k - 1:], ref[s, k - 1:])
            for r in range(k - 1):
                assert np.array_equal(out[s, r], ref[s, k - 1])
    bad = np.array([1, 0, 8, 2, 5], np.int32)
    with pytest.raises(_lib.CcbError):
        _lib.check(L.ccb_gains_backfill(C.c_int32(0), C.c_void_p(t.data_ptr()), C.c_int64(5), C.c_int32(9), C.c_int32(1), C.c_int32(3), _lib.ptr(bad)))
