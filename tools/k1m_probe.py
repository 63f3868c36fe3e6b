"""K1m timing probe: Sawyer LQR rollouts, B systems x 2000 steps, register variant from CCB_K1M_REGS"""
import sys, os
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, ROOT + "/tests"]
import numpy as np, torch, ccb200 as cc
B = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"], qd=cfg["qd"])
rng = np.random.default_rng(3); st = np.zeros((B, mech.nb, 13))
for s in range(B):
    cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)}); st[s] = mech.state_vector()
batch = cc.Batch(mech, B); stream = torch.cuda.ExternalStream(batch.stream())
best = 1e9
for rep in range(3):
    batch.set_state(st); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); batch.simulate_device(lqr, 1, steps); e1.record(stream); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"B {B} steps {steps} regs {os.environ.get('CCB_K1M_REGS')} ms {best:.2f} -> {B * steps / best / 1e3:.2f} M steps/s")
