"""Parity (small) + throughput (full size) sweep over the BASELINE configs C1..C5; writes JSON lines to stdout.
Not the bench: a measurement helper whose output is archived under profiles/."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import ccb200 as cc, orc
from helpers import rel_err, zd_of

GOLD = os.path.join(ROOT, "tests", "golden")
only = sys.argv[1:] or None


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        t0 = time.time(); out = fn(); best = min(best, time.time() - t0)
    return best, out


def chain_states(mech, B, seed, amp):
    """amp None: the example's own distribution phi_i ~ U(0,1)/3^n (examples/lqr_cartpole_n_pendulum.jl:21)"""
    rng = np.random.default_rng(seed)
    out = np.zeros((B, mech.nb, 13))
    n = mech.nb - 1
    for s in range(B):
        phis = rng.uniform(0, 1, size=n) / 3 ** n if amp is None else rng.uniform(-amp, amp, size=n)
        cc.examples.set_cartpole_pose(mech, rng.uniform(-0.5, 0.5), phis)
        out[s] = mech.state_vector()
    return out


def run_lqr_config(tag, mech, cfg, st_fn, B, steps, horizon, parity_B=8, parity_steps=150, synth_S=None):
    om = orc.OracleMech(mech.desc())
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    zd = zd_of(mech, cfg)
    Fd = np.array(cfg.get("Ftaud", np.zeros(len(eqc_idx))), dtype=float)
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
    N = int(math.ceil(horizon / mech.dt)) if horizon != math.inf else int(math.ceil(10 / mech.dt))
    # synthesis: GPU vs oracle
    t_syn, out = timed(lambda: cc.lqr_synthesize(mech, zd[None], Fd[None], eqc_idx, Qd, Rd, N), reps=1)
    t0 = time.time(); lin = om.linearize(zd, Fd, eqc_idx); ref = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Qd, Rd, N); t_cpu = time.time() - t0
    rec = dict(config=tag, nb=mech.nb, ml=mech.ml, n=12 * mech.nb, N=N, riccati_iters=int(out["iters"][0]), riccati_iters_oracle=ref["iters"],
               gain_rel_err=rel_err(out["Ku"][0], ref["Ku"]), synth_1_gpu_s=t_syn, synth_1_cpu_s=t_cpu)
    K = out["Ku"][0]
    if horizon == math.inf:
        K = K[:1]; Nh = math.inf
    else:
        Nh = N
    lqr = cc.LQR.from_gains(mech, cfg["eqcids"], K, Nh, xd=zd[:, 0:3], qd=zd[:, 3:7], Ftaud=Fd)
    # rollout parity on a small batch
    st = st_fn(parity_B)
    kind = "lqr_inf" if horizon == math.inf else "lqr_finite"
    co = om.make_ctrl(parity_B, kind=kind, eqc=eqc_idx, N=0 if horizon == math.inf else N, K=K, zd=zd, Fd=Fd)
    r = om.simulate(st, parity_steps, ctrl=co, record=True, nthreads=8)
    sto, fin, status = cc.simulate(mech, parity_steps * mech.dt, lqr, record=True, states=st, return_status=True)
    rec.update(rollout_rel_err=rel_err(sto.data, r["storage"]), newton_ok=bool(np.all(status > 0)))
    # throughput at full size
    stB = st_fn(B)
    batch = cc.Batch(mech, B)
    def go():
        batch.set_state(stB)
        return batch.simulate(lqr, 1, steps)
    t, (_, status) = timed(go)
    rec.update(B=B, steps=steps, rollout_s=t, sim_steps_per_s=B * steps / t, newton_per_step=float(np.abs(status).mean() / steps),
               failed=int((status <= 0).sum()), lanes=cc.handle(mech).lanes)
    # CPU oracle rate on a sample
    nb_s = 16
    t0 = time.time(); om.simulate(stB[:nb_s], 100, ctrl=om.make_ctrl(nb_s, kind=kind, eqc=eqc_idx, N=0 if horizon == math.inf else N, K=K, zd=zd, Fd=Fd), nthreads=8)
    rec["cpu_sim_steps_per_s_8thr"] = nb_s * 100 / (time.time() - t0)
    if synth_S:
        rng = np.random.default_rng(2)
        qs = np.exp(rng.uniform(np.log(0.1), np.log(10), synth_S)); rs = np.exp(rng.uniform(np.log(0.1), np.log(10), synth_S))
        Qb = Qd[None] * qs[:, None, None]; Rb = Rd[None] * rs[:, None, None]
        t, o2 = timed(lambda: cc.lqr_synthesize(mech, np.repeat(zd[None], synth_S, 0), np.repeat(Fd[None], synth_S, 0), eqc_idx, Qb, Rb, N), reps=1)
        ph = cc.synth.phase_ms()
        rec.update(synth_S=synth_S, synth_batch_s=t, syntheses_per_s=synth_S / t, synth_mean_iters=float(o2["iters"].mean()), phase_ms=ph,
                   syntheses_per_s_device=synth_S / (sum(ph.values()) * 1e-3))
    print(json.dumps(rec), flush=True)


def want(tag):
    return only is None or tag in only


if want("C1"):
    mech, cfg = cc.examples.pendulum()
    run_lqr_config("C1 lqr_pendulum", mech, cfg, lambda B: mech.batch_states(B), 4096, 1000, math.inf, synth_S=1024)
if want("C2"):
    mech, cfg = cc.examples.cartpole()
    run_lqr_config("C2 lqr_cartpole", mech, cfg, lambda B: chain_states(mech, B, 1, 0.2), 4096, 2000, 20.0, synth_S=4096)
if want("C3a"):
    mech, cfg = cc.examples.cartpole_n_pendulum(8)
    run_lqr_config("C3 n_pendulum n=8", mech, cfg, lambda B: chain_states(mech, B, 2, None), 1024, 1000, 10.0, parity_steps=100, synth_S=64)
if want("C3b"):
    mech, cfg = cc.examples.cartpole_n_pendulum(32)
    run_lqr_config("C3 n_pendulum n=32", mech, cfg, lambda B: chain_states(mech, B, 2, None), 1024, 200, 10.0, parity_B=2, parity_steps=40, synth_S=None)
if want("C4"):
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    def st_fn(B):
        rng = np.random.default_rng(3)
        out = np.zeros((B, mech.nb, 13))
        for s in range(B):
            cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)})
            out[s] = mech.state_vector()
        cc.urdf.zero_configuration(mech)
        return out
    run_lqr_config("C4 lqr_sawyer", mech, cfg, st_fn, 1024, 2000, 20.0, parity_steps=100, synth_S=64)
if want("C5a"):
    # lqr_deltabot: closed chain (2 loops, rank-deficient constraint set); gravity-compensating feed-forward, horizon Inf
    mech, cfg = cc.examples.deltabot()
    def st_fn(B):          # the example's own grid of reachable platform positions (lqr_deltabot.jl:57-152), first B of a 105 x 105 grid
        return cc.examples.deltabot_states(mech, B) if B > 8 else mech.batch_states(B)
    try:
        run_lqr_config("C5 lqr_deltabot", mech, cfg, st_fn, 2048, 1000, math.inf, parity_B=4, parity_steps=100, synth_S=64)
    except Exception as e:
        print(json.dumps(dict(config="C5 lqr_deltabot", error=repr(e))), flush=True)
if want("C5b"):
    mech, cfg = cc.examples.doublependulum()
    n = 32
    g1, g2 = np.meshgrid(np.linspace(-math.pi, math.pi, n, endpoint=False), np.linspace(-math.pi, math.pi, n, endpoint=False))
    goals = np.stack([g1.ravel(), g2.ravel()], axis=1)
    B = goals.shape[0]
    st0 = mech.batch_states(B)
    batch = cc.Batch(mech, B)
    def go():
        pid = cc.PID(mech, cfg["eqcids"], goals, P=cfg["P"], I=cfg["I"], D=cfg["D"])   # fresh controller: integratederrors persist otherwise (pid.jl:76)
        batch.set_state(st0)
        return batch.simulate(pid, 1, 1000)
    t, (_, status) = timed(go)
    print(json.dumps(dict(config="C5 pid_doublependulum 32x32 goal sweep", B=B, steps=1000, rollout_s=t, sim_steps_per_s=B * 1000 / t,
                          newton_per_step=float(np.abs(status).mean() / 1000), flagged=int((status <= 0).sum()))), flush=True)

def run_uncontrolled(tag, n, B, steps):
    """hanging n-pendulum swinging freely: the tree solve at scale without the (ill-conditioned) upright LQR"""
    mech, cfg = cc.examples.cartpole_n_pendulum(n)
    rng = np.random.default_rng(3)
    st = np.zeros((B, mech.nb, 13))
    for s_ in range(B):
        ph = rng.uniform(-0.05, 0.05, size=n); ph[0] += math.pi
        cc.examples.set_cartpole_pose(mech, rng.uniform(-0.5, 0.5), ph); st[s_] = mech.state_vector()
    om = orc.OracleMech(mech.desc())
    r = om.simulate(st[:4], 40, record=True, nthreads=4)
    sto, fin, status = cc.simulate(mech, 0.4, None, record=True, states=st[:4], return_status=True)
    batch = cc.Batch(mech, B)
    def go():
        batch.set_state(st)
        return batch.simulate(None, 1, steps)
    t, (_, status) = timed(go)
    t0 = time.time(); om.simulate(st[:16], 50, nthreads=8); tc = time.time() - t0
    print(json.dumps(dict(config=tag, nb=mech.nb, ml=mech.ml, B=B, steps=steps, rollout_s=t, sim_steps_per_s=B * steps / t,
                          newton_per_step=float(np.abs(status).mean() / steps), failed=int((status <= 0).sum()),
                          rollout_rel_err=rel_err(sto.data, r["storage"]), cpu_sim_steps_per_s_8thr=16 * 50 / tc)), flush=True)


if want("C3c"):
    run_uncontrolled("C3 n_pendulum n=8 hanging chain, uncontrolled", 8, 1024, 1000)
if want("C3d"):
    run_uncontrolled("C3 n_pendulum n=32 hanging chain, uncontrolled", 32, 1024, 200)
