"""C4 at scale: 64 Sawyer reference trajectories x N steps -> TrackingLQR synthesis -> 16 perturbed rollouts each"""
import sys, os, time, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
PER = int(sys.argv[3]) if len(sys.argv) > 3 else 16
AMP = float(sys.argv[4]) if len(sys.argv) > 4 else 0.005   # the example's Q = 1000 I gains saturate the integrator for larger offsets (also in the oracle)
mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
rng = np.random.default_rng(3)
a = rng.uniform(0, 1, (NT, 7)) * np.array([0.05, 0.05, 0.02, 0.02, 0.005, 0.005, 0.0005]); f = rng.integers(1, 4, (NT, 7))
# torque amplitudes scaled to the link inertias: the frictionless arm in zero gravity must stay well below |w| = 2/dt
k = np.arange(N)[None, :, None]
U = a[:, None, :] * np.sin(2 * np.pi * f[:, None, :] * k / 2000.0)     # [NT][N][7]
cc.urdf.zero_configuration(mech)
st0 = mech.batch_states(NT)
t0 = time.time()
ol = cc.OpenLoop(mech, cfg["eqcids"], U)
sto, fin, status = cc.simulate(mech, N * mech.dt, ol, record=True, states=st0, return_status=True)
t1 = time.time()
print(f"reference trajectories: {NT} x {N} steps open loop, recorded: {t1 - t0:.3f} s, newton ok {np.all(status > 0)}", flush=True)
t0 = time.time()
tl = cc.TrackingLQR(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"])
t1 = time.time()
ph = cc.synth.phase_ms()
print(f"TrackingLQR synthesis: {NT} trajectories x {N - 1} linearisations+Riccati steps: {t1 - t0:.3f} s e2e; device phases {ph}, elapsed {cc.synth.phase_total_ms():.1f} ms; kexit {tl.kexit[:4]}", flush=True)
B = NT * PER
stp = np.repeat(st0, PER, axis=0)
pert = np.zeros((B, 7))
for s in range(B):
    cc.urdf.zero_configuration(mech, {i: rng.uniform(-AMP, AMP) for i in range(7)})
    stp[s] = mech.state_vector()
trk = tl.replicate(PER)      # 16 perturbed starts per reference trajectory share its (device-resident) tables
batch = cc.Batch(mech, B)
for rep in range(2):
    batch.set_state(stp)
    t0 = time.time(); _, status = batch.simulate(trk, 1, N); t1 = time.time()
fin = batch.get_state()
err = np.abs(fin[:, :, 0:3] - np.repeat(sto.data[-1], PER, axis=0)[:, :, 0:3]).max()
print(f"tracking rollouts: {B} x {N}: {t1 - t0:.3f} s -> {B * N / (t1 - t0) / 1e6:.2f} M sim steps/s, newton ok {np.all(status > 0)}, iters/step {np.abs(status).mean() / N:.2f}, "
      f"final |x - x_ref| max {err:.2e}", flush=True)
