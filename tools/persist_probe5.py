"""K1r time-sliced launch: extra item after the whole own chunk (CCB_K1R_STAGGER=0) against the staggered cut (default), and the plain launch;
2000 cartpole steps; bit-identity of states and Newton totals across the three."""
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch, ccb200 as cc
from helpers import cartpole_states

KEYS = ("CCB_K1R_NO_PERSIST", "CCB_K1R_STAGGER", "CCB_K1R_CHUNKS")
mech, cfg = cc.examples.cartpole()
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"])
st_all = cartpole_states(cc, mech, 4794, seed=1)


def run(B, steps=2000, **env):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    batch = cc.Batch(mech, B)
    stream = torch.cuda.ExternalStream(batch.stream())
    ts = []
    for rep in range(5):
        batch.set_state(st_all[:B]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); batch.simulate_device(lqr, 1, steps); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    batch.set_state(st_all[:B])
    _, status = batch.simulate(lqr, 1, steps)
    fin, lam = batch.get_state(want_lambda=True)
    print("B %5d %-32s ms best %.3f median %.3f" % (B, env, min(ts[1:]), float(np.median(ts[1:]))), flush=True)
    return fin, lam, status


for B in (4096, 3700, 4400, 4794):
    ref = run(B, CCB_K1R_NO_PERSIST=1)
    for env in (dict(CCB_K1R_STAGGER=0), dict(), dict(CCB_K1R_CHUNKS=16), dict(CCB_K1R_CHUNKS=64)):
        out = run(B, **env)
        print("      bit-identical to the plain launch (states, multipliers, Newton totals):", all(np.array_equal(a, b) for a, b in zip(ref, out)), flush=True)
# ragged: 300 steps (chunks of 34, the last one short), 257 steps
for steps in (300, 257):
    ref = run(4096, steps, CCB_K1R_NO_PERSIST=1)
    out = run(4096, steps)
    print("      steps", steps, "bit-identical:", all(np.array_equal(a, b) for a, b in zip(ref, out)), flush=True)
