"""One small invocation of ONE kernel family for ncu (round 2: summaries of every kernel a BASELINE config runs).
    python tools/profile_case.py <case>       cases: k1r k1r_record k1m_sawyer k1m_sawyer128 k1m_npend8 k1m_npend32 k1_deltabot tracking sawyer_lqr are k3 k4b pid
"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import ccb200 as cc
from helpers import cartpole_states, zd_of
GOLD = os.path.join(ROOT, "tests", "golden")
case = sys.argv[1]


def lqr_for(mech, cfg, horizon):
    return cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], horizon, xd=cfg["xd"], qd=cfg.get("qd"))


def chain_states(mech, B, seed, hanging):
    rng = np.random.default_rng(seed); n = mech.nb - 1
    out = np.zeros((B, mech.nb, 13))
    for s in range(B):
        ph = rng.uniform(-0.05, 0.05, size=n)
        if hanging: ph[0] += math.pi
        cc.examples.set_cartpole_pose(mech, rng.uniform(-0.5, 0.5), ph); out[s] = mech.state_vector()
    return out


def sawyer_states(mech, B):
    rng = np.random.default_rng(3); st = np.zeros((B, mech.nb, 13))
    for s in range(B):
        cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)}); st[s] = mech.state_vector()
    cc.urdf.zero_configuration(mech)
    return st


if case in ("k1r", "k1r_record"):
    mech, cfg = cc.examples.cartpole()
    lqr = lqr_for(mech, cfg, 20.0)
    st = cartpole_states(cc, mech, 4096, seed=1)
    sto, fin, status = cc.simulate(mech, (2000 if case == "k1r" else 256) * mech.dt, lqr, record=case == "k1r_record", states=st, return_status=True)
elif case in ("k1m_sawyer", "k1m_sawyer128"):
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    lqr = lqr_for(mech, cfg, 20.0)
    B = 1024 if case == "k1m_sawyer" else 128
    _, _, status = cc.simulate(mech, 2.0, lqr, states=sawyer_states(mech, B), return_status=True)
elif case in ("k1m_npend8", "k1m_npend32"):
    n = 8 if case.endswith("8") else 32
    mech, cfg = cc.examples.cartpole_n_pendulum(n)
    _, _, status = cc.simulate(mech, 1.0 if n == 8 else 0.4, None, states=chain_states(mech, 1024, 3, True), return_status=True)
elif case == "k1_deltabot":
    mech, cfg = cc.examples.deltabot()
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], math.inf, xd=cfg["xd"], qd=cfg["qd"], Ftaud=cfg["Ftaud"])
    _, _, status = cc.simulate(mech, 1.0, lqr, states=cc.examples.deltabot_states(mech, 2048), return_status=True)
elif case == "pid":
    mech, cfg = cc.examples.doublependulum()
    g1, g2 = np.meshgrid(np.linspace(-2.5, 2.5, 32), np.linspace(-2.5, 2.5, 32))
    pid = cc.PID(mech, cfg["eqcids"], np.stack([g1.ravel(), g2.ravel()], axis=1), P=cfg["P"], I=cfg["I"], D=cfg["D"])
    _, _, status = cc.simulate(mech, 3.0, pid, states=mech.batch_states(1024), return_status=True)
elif case == "tracking":      # linearize_kernel + riccati_project_mech_kernel + riccati_sweep_dmma_kernel (time varying), 16 x 400 knots
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    NT, N = 16, 400
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, (NT, 7)) * np.array([0.05, 0.05, 0.02, 0.02, 0.005, 0.005, 0.0005]); f = rng.integers(1, 4, (NT, 7))
    U = a[:, None, :] * np.sin(2 * np.pi * f[:, None, :] * np.arange(N)[None, :, None] / 2000.0)
    cc.urdf.zero_configuration(mech)
    sto, _, status = cc.simulate(mech, N * mech.dt, cc.OpenLoop(mech, cfg["eqcids"], U), record=True, states=mech.batch_states(NT), return_status=True)
    tl = cc.TrackingLQR(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"])
    print("phase ms", cc.synth.phase_ms())
elif case == "sawyer_lqr":    # projection + time-invariant K4 sweep, 1024 setpoints, N = 200
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    S = 1024
    zd = np.repeat(zd_of(mech, cfg)[None], S, 0)
    for _ in range(2):
        out = cc.lqr_synthesize(mech, zd, np.zeros((S, 7)), [mech.eqc_index(i) for i in cfg["eqcids"]], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, 200)
        print("phase ms", cc.synth.phase_ms())
elif case == "are":
    rng = np.random.default_rng(5); n, m, S = 24, 1, 4096
    A = rng.standard_normal((S, n, n)) / np.sqrt(n) * 1.05
    K = cc.util.are_solve(cc.util.DISCRETE, A, rng.standard_normal((S, n, m)), np.eye(n), np.eye(m), want_X=False, want_K=True)
elif case == "k3":
    os.environ["CCB_NO_DMMA"] = "1"
    mech, cfg = cc.examples.cartpole()
    S = 1024
    zd = np.repeat(zd_of(mech, cfg)[None], S, 0)
    out = cc.lqr_synthesize(mech, zd, np.zeros((S, 1)), [mech.eqc_index(i) for i in cfg["eqcids"]], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, 300)
elif case == "k4b":
    mech, cfg = cc.examples.cartpole_n_pendulum(16)
    S = 148
    zd = np.repeat(zd_of(mech, cfg)[None], S, 0)
    out = cc.lqr_synthesize(mech, zd, np.zeros((S, 1)), [mech.eqc_index(i) for i in cfg["eqcids"]], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, 12)
else:
    raise SystemExit("unknown case")
print(case, "ok")
