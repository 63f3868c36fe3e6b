import sys, os, time
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc
NT, N = 64, 2000
mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
rng = np.random.default_rng(3)
a = rng.uniform(0, 1, (NT, 7)) * np.array([0.05, 0.05, 0.02, 0.02, 0.005, 0.005, 0.0005]); f = rng.integers(1, 4, (NT, 7))
U = a[:, None, :] * np.sin(2 * np.pi * f[:, None, :] * np.arange(N)[None, :, None] / 2000.0)
cc.urdf.zero_configuration(mech)
sto, fin, status = cc.simulate(mech, N * mech.dt, cc.OpenLoop(mech, cfg["eqcids"], U), record=True, states=mech.batch_states(NT), return_status=True)
for env in ("1", "0", "0+dense"):
    os.environ["CCB_TRACK_PIPELINE"] = env[0]
    if "dense" in env: os.environ["CCB_PROJECT_DENSE"] = "1"
    for rep in range(2):
        tl = cc.TrackingLQR(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"]); tl.invalidate()
    print("pipeline", env, cc.synth.phase_ms(), "total", cc.synth.phase_total_ms(), flush=True)
