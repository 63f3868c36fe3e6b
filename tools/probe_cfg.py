"""rollout timing probe for a named config: python tools_probe_cfg.py sawyer|npend8|npend32 B steps "lanes,lanes,..." """
import sys, os, time, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc
from helpers import zd_of
name, B, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lanes_list = [int(x) for x in sys.argv[4].split(",")]
rng = np.random.default_rng(3)
if name == "sawyer":
    mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
    st = np.zeros((B, mech.nb, 13))
    for s in range(B):
        cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)}); st[s] = mech.state_vector()
    cc.urdf.zero_configuration(mech)
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"], qd=cfg["qd"])
else:
    n = int(name[5:])
    mech, cfg = cc.examples.cartpole_n_pendulum(n)
    st = np.zeros((B, mech.nb, 13))
    for s in range(B):
        ph = rng.uniform(-0.05, 0.05, size=n); ph[0] += math.pi      # hanging chain: a stable swing, every Newton solve converges
        cc.examples.set_cartpole_pose(mech, rng.uniform(-0.5, 0.5), ph); st[s] = mech.state_vector()
    lqr = None      # uncontrolled swing (the upright LQR of long chains is ill-conditioned in Float64)
batch = cc.Batch(mech, B)
for lanes in lanes_list:
    best = 1e9
    for rep in range(2):
        batch.set_state(st)
        t0 = time.time(); _, status = batch.simulate(lqr, 1, steps, lanes=lanes); best = min(best, time.time() - t0)
    print(f"{name} B={B} steps={steps} lanes={lanes}: {best*1e3:.1f} ms  {B*steps/best/1e6:.2f} M steps/s  iters/step {np.abs(status).mean()/steps:.2f} failed {(status<=0).sum()}", flush=True)
