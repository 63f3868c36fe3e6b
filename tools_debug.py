import sys, math, os
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/oracle'); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc, orc
from helpers import rel_err
mech, cfg = cc.examples.doublependulum()
n = 32
g1, g2 = np.meshgrid(np.linspace(-math.pi, math.pi, n, endpoint=False), np.linspace(-math.pi, math.pi, n, endpoint=False))
goals = np.stack([g1.ravel(), g2.ravel()], axis=1)
B = goals.shape[0]
st0 = mech.batch_states(B)
for lanes in (0, 1, 8):
    pid = cc.PID(mech, cfg["eqcids"], goals, P=cfg["P"], I=cfg["I"], D=cfg["D"])
    sto, fin, status = cc.simulate(mech, 10.0, pid, record=False, states=st0, return_status=True, lanes=lanes)
    bad = np.where(status <= 0)[0]
    print("lanes", lanes, "failed", len(bad), bad[:10], "iters/step", np.abs(status).mean() / 1000)
om = orc.OracleMech(mech.desc())
eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
sub = np.concatenate([bad[:8], np.arange(0, B, 128)])
integ = np.zeros((len(sub), 2)); last = np.zeros((len(sub), 2))
co = om.make_ctrl(len(sub), kind="pid", eqc=eqc, P=cfg["P"], I=cfg["I"], D=cfg["D"], goal=goals[sub], pid_int=integ, pid_last=last)
r = om.simulate(st0[sub], 1000, ctrl=co, record=True, nthreads=8)
print("oracle iters", r["iters"])
pid2 = cc.PID(mech, cfg["eqcids"], goals[sub], P=cfg["P"], I=cfg["I"], D=cfg["D"])
sto2, fin2, st2 = cc.simulate(mech, 10.0, pid2, record=True, states=st0[sub], return_status=True)
print("gpu status", st2)
for i in range(len(sub)):
    e = np.abs(sto2.data[:, i] - r["storage"][:, i]).max(axis=(1, 2))
    first = np.argmax(e > 1e-8) if (e > 1e-8).any() else -1
    print(i, sub[i], goals[sub[i]], "max err", e.max(), "first step >1e-8:", first, "max |w|", np.abs(r["storage"][:, i, :, 10:13]).max())
