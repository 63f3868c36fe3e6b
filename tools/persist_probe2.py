"""K1r persistent time-sliced launch: where does the time go?  One process, several (batch, env) settings on 2000 cartpole steps:
T2 of the persistent instantiation itself (3552 systems = exactly two warps per scheduler, no extra triples: forced persistent vs plain),
the bench batch (4096) with 16 / 32 / 64 chunks, and 3 x 1184 + extras in between."""
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch, ccb200 as cc
from helpers import cartpole_states

steps = 2000
mech, cfg = cc.examples.cartpole()
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"])
KEYS = ("CCB_K1R_FORCE_PERSIST", "CCB_K1R_NO_PERSIST", "CCB_K1R_CHUNKS")


def run(B, **env):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    st = cartpole_states(cc, mech, B, seed=1)
    batch = cc.Batch(mech, B)
    stream = torch.cuda.ExternalStream(batch.stream())
    ts = []
    for rep in range(4):
        batch.set_state(st); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); batch.simulate_device(lqr, 1, steps); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("B", B, env, "ms best %.3f  all %s" % (min(ts[1:]), " ".join("%.2f" % t for t in ts)), flush=True)
    del batch


run(3552, CCB_K1R_NO_PERSIST=1)
run(3552, CCB_K1R_FORCE_PERSIST=1)
run(3552, CCB_K1R_FORCE_PERSIST=1, CCB_K1R_CHUNKS=1)
run(3552, CCB_K1R_FORCE_PERSIST=1, CCB_K1R_CHUNKS=8)
run(4096, CCB_K1R_NO_PERSIST=1)
run(4096)
run(4096, CCB_K1R_CHUNKS=16)
run(4096, CCB_K1R_CHUNKS=64)
run(4096, CCB_K1R_CHUNKS=125)
run(3700)
run(3700, CCB_K1R_NO_PERSIST=1)
