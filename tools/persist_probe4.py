"""K1r: when does the 255-register instantiation pay?  2000-step closed loops, one process; modes: plain (168 registers), whole (255 registers,
one chunk; batches of at most two warps per scheduler), sliced with a given window (percent of two-warps-per-scheduler)."""
import os, sys, math
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch, ccb200 as cc
from helpers import cartpole_states

KEYS = ("CCB_K1R_FORCE_PERSIST", "CCB_K1R_NO_PERSIST", "CCB_K1R_CHUNKS", "CCB_K1R_WINDOW", "CCB_K1R_WHOLE")


def run(tag, mech, ctrl, st, steps, **env):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    B = st.shape[0]
    batch = cc.Batch(mech, B)
    stream = torch.cuda.ExternalStream(batch.stream())
    ts = []
    for rep in range(4):
        batch.set_state(st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); batch.simulate_device(ctrl, 1, steps); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    fin = batch.get_state()
    del batch
    print("%-16s B %5d %-40s ms %.3f" % (tag, B, env, min(ts[1:])), flush=True)
    return fin


mech, cfg = cc.examples.cartpole()
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"])
st_all = cartpole_states(cc, mech, 5328, seed=1)
for B in (512, 1024, 2048, 3072, 3552):
    a = run("cartpole", mech, lqr, st_all[:B], 2000, CCB_K1R_NO_PERSIST=1)
    b = run("cartpole", mech, lqr, st_all[:B], 2000, CCB_K1R_WHOLE=1)
    print("   bit-identical:", np.array_equal(a, b), flush=True)
for B in (4096, 4400, 4608, 4800, 5328):
    a = run("cartpole", mech, lqr, st_all[:B], 2000, CCB_K1R_NO_PERSIST=1)
    b = run("cartpole", mech, lqr, st_all[:B], 2000, CCB_K1R_WINDOW=150)
    print("   bit-identical:", np.array_equal(a, b), flush=True)

# C1 pendulum (one joint: six systems per warp), infinite-horizon LQR
pm, pcfg = cc.examples.pendulum()
plqr = cc.LQR(pm, cc.getid(pm.bodies), pcfg["eqcids"], pcfg["Q"], pcfg["R"], math.inf, xd=pcfg["xd"], qd=pcfg["qd"])
rng = np.random.default_rng(3)
pst = np.zeros((8192, 1, 13))
for s in range(8192):
    m2, _ = cc.examples.pendulum(phi0=math.pi + rng.uniform(-0.4, 0.4))
    pst[s] = m2.state_vector()
for B in (4096, 7104, 8192):
    a = run("pendulum", pm, plqr, pst[:B], 1000, CCB_K1R_NO_PERSIST=1)
    b = run("pendulum", pm, plqr, pst[:B], 1000, CCB_K1R_WHOLE=1, CCB_K1R_WINDOW=150)
    print("   bit-identical:", np.array_equal(a, b), flush=True)

# triple cartpole (four joints: one system per warp), open-loop table
tm, tcfg = cc.examples.triple_cartpole()
U = np.load(ROOT + "/tests/golden/triple_cartpole_U.npy").reshape(1000, 1)
ol = cc.OpenLoop(tm, tcfg["eqcids"], U)
tst = np.tile(tm.state_vector()[None], (1500, 1, 1))
tst[:, :, 1] += np.linspace(-0.2, 0.2, 1500)[:, None]
for B in (592, 1184, 1366, 1500):
    a = run("triple_cartpole", tm, ol, tst[:B], 1000, CCB_K1R_NO_PERSIST=1)
    b = run("triple_cartpole", tm, ol, tst[:B], 1000, CCB_K1R_WHOLE=1, CCB_K1R_WINDOW=150)
    print("   bit-identical:", np.array_equal(a, b), flush=True)
