"""small invocation of every kernel family for compute-sanitizer (memcheck / racecheck): tiny batches, few steps"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc
# K1r + K2 + K4 (staged, nu = 1): cartpole LQR end to end
mech, cfg = cc.examples.cartpole()
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 0.3, xd=cfg["xd"])
st = mech.batch_states(5)
for lanes in (0, 1, 8):
    cc.simulate(mech, 0.05, lqr, record=True, states=st, lanes=lanes)
# K1m + K4 (nu = 7): Sawyer
mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 0.1, xd=cfg["xd"], qd=cfg["qd"])
cc.simulate(mech, 0.04, lqr, record=True, states=mech.batch_states(3))
# K1 general with loops: deltabot, PID on the double pendulum
mech, cfg = cc.examples.deltabot()
cc.simulate(mech, 0.03, None, record=False, states=mech.batch_states(3))
mech, cfg = cc.examples.doublependulum()
pid = cc.PID(mech, cfg["eqcids"], cfg["goals"], P=cfg["P"], I=cfg["I"], D=cfg["D"])
cc.simulate(mech, 0.03, pid, states=mech.batch_states(4))
print("sanitize case done")
