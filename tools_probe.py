"""quick timing probe of the headline rollout config (not the bench)"""
import sys, time, math
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, ccb200 as cc, orc
from helpers import oracle_lqr_gains, cartpole_states
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
mech, cfg = cc.examples.cartpole()
o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=20.0)
st0 = cartpole_states(cc, mech, B)
lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg["xd"])
for lanes, spw in ((5, 0), (1, 16), (8, 0)):
    batch = cc.Batch(mech, B)
    for rep in range(3):
        batch.set_state(st0)
        t0 = time.time()
        _, status = batch.simulate(lqr, 1, steps, lanes=lanes, spw=spw)
        dt = time.time() - t0
    print(f"lanes={lanes} spw={spw} B={B} steps={steps}: {dt*1e3:.1f} ms  {B*steps/dt/1e6:.1f} M steps/s  newton/step={status.mean()/steps:.2f}", flush=True)
t0 = time.time()
ref = o["om"].simulate(st0[:64], 200, ctrl=o["om"].make_ctrl(64, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"]), nthreads=8)
dt = time.time() - t0
print(f"oracle 8 threads: {64*200/dt/1e6:.3f} M steps/s, newton/step={ref['iters'].mean()/200:.2f}")
