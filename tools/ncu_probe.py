"""small, short invocation of the rollout kernel for ncu (one launch, 4096 systems x 100 steps)"""
import sys, math, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/oracle'); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc, orc
from helpers import oracle_lqr_gains, cartpole_states
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
mech, cfg = cc.examples.cartpole()
o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=20.0)
st0 = np.tile(cartpole_states(cc, mech, 4096), (B // 4096 + 1, 1, 1))[:B].copy()
lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"])
batch = cc.Batch(mech, B)
batch.set_state(st0)
_, status = batch.simulate(lqr, 1, steps, lanes=lanes)
print("ok", status.mean() / steps)
