"""one K1r library variant (CCB_LIB): T2 of its persistent instantiation (3552 systems, forced, one chunk) and the bench batch (4096, default schedule)"""
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch, ccb200 as cc
from helpers import cartpole_states

steps = 2000
mech, cfg = cc.examples.cartpole()
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"])
KEYS = ("CCB_K1R_FORCE_PERSIST", "CCB_K1R_NO_PERSIST", "CCB_K1R_CHUNKS")
st_all = cartpole_states(cc, mech, 4096, seed=1)


def run(B, **env):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    batch = cc.Batch(mech, B)
    stream = torch.cuda.ExternalStream(batch.stream())
    ts = []
    for rep in range(4):
        batch.set_state(st_all[:B]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); batch.simulate_device(lqr, 1, steps); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    del batch
    return min(ts[1:])


name = os.path.basename(os.environ.get("CCB_LIB", "current"))
print("VARIANT %s T2_persist %.3f T2_plain %.3f B4096 %.3f" % (name, run(3552, CCB_K1R_FORCE_PERSIST=1, CCB_K1R_CHUNKS=1), run(3552, CCB_K1R_NO_PERSIST=1), run(4096)), flush=True)
