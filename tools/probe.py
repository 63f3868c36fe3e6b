"""quick timing probe of the headline rollout config (not the bench): python tools_probe.py "B,steps,lanes[,spw];..." """
import sys, time, math, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/oracle'); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc, orc
from helpers import oracle_lqr_gains, cartpole_states
spec = sys.argv[1] if len(sys.argv) > 1 else "4096,2000,0"
mech, cfg = cc.examples.cartpole()
o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=20.0)
lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"])
cache = {}
for item in spec.split(";"):
    f = [int(x) for x in item.split(",")]
    B, steps, lanes = f[0], f[1], f[2]
    spw = f[3] if len(f) > 3 else 0
    if B not in cache:
        base = cartpole_states(cc, mech, min(B, 4096))
        cache[B] = np.tile(base, (B // base.shape[0] + 1, 1, 1))[:B].copy()
    st0 = cache[B]
    batch = cc.Batch(mech, B)
    best = 1e9
    for rep in range(3):
        batch.set_state(st0)
        t0 = time.time()
        batch.simulate_device(lqr, 1, steps, lanes=lanes, spw=spw)
        batch.sync()
        best = min(best, time.time() - t0)
    print(f"{os.environ.get('CCB_LIB','default')[-12:]} lanes={lanes} spw={spw} B={B} steps={steps}: {best*1e3:.1f} ms  {B*steps/best/1e6:.1f} M steps/s", flush=True)
