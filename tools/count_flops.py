"""Exact FP64 operation counts of the literal algorithm (oracle instrumented with a counting scalar) along closed-loop rollouts of
the BASELINE configs -> profiles/flops_r2.json.  bench.py reads that file (it may not execute the oracle outside its cpu_baseline leg).
    python tools/count_flops.py
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import ccb200 as cc, orc
from helpers import cartpole_states, oracle_lqr_gains
orc.build()
GOLD = os.path.join(ROOT, "tests", "golden")
out = {"convention": "add/sub/mul/div/sqrt = 1 flop each on the oracle's scalar (an FMA is 2); literal algorithm = initial residual + per Newton iteration "
                     "(setentries residual and analytic KKT blocks, mixed-node block LDU in dfs_list order, substitution, line-search residuals); "
                     "control-law flops (2 nu n per step) not counted", "configs": {}}


def run(name, mech, cfg, st0, steps, horizon, note):
    o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=horizon)
    B = st0.shape[0]
    kind = "lqr_inf" if horizon == math.inf else "lqr_finite"
    co = o["om"].make_ctrl(B, kind=kind, eqc=o["eqc_idx"], N=0 if horizon == math.inf else o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
    tot = {k: 0 for k in ("add", "mul", "div", "sqrt", "compare", "iterations")}
    for s in range(B):
        c = o["om"].count_flops_rollout(st0[s], steps, ctrl=co, sys=s)
        for k in tot:
            tot[k] += c[k]
    fl = tot["add"] + tot["mul"] + tot["div"] + tot["sqrt"]
    rec = dict(tot, systems=B, steps=steps, flops=fl, flops_per_sim_step=fl / (B * steps), flops_per_newton_iteration=fl / tot["iterations"],
               newton_iterations_per_step=tot["iterations"] / (B * steps), nb=mech.nb, m_lambda=mech.ml, workload=note)
    out["configs"][name] = rec
    print(name, json.dumps(rec))


mech, cfg = cc.examples.cartpole()
run("C2_cartpole", mech, cfg, cartpole_states(cc, mech, 16, seed=1), 2000, 20.0, "bench workload: first 16 of the 4096 seeded systems x 2000 steps, finite-horizon LQR")
mech, cfg = cc.examples.pendulum()
run("C1_pendulum", mech, cfg, mech.batch_states(1), 1000, math.inf, "examples/lqr_pendulum.jl, 1000 steps")
mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
rng = np.random.default_rng(3)
st0 = np.zeros((4, mech.nb, 13))
for s in range(4):
    cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)}); st0[s] = mech.state_vector()
cc.urdf.zero_configuration(mech)
run("C4_sawyer", mech, cfg, st0, 500, 20.0, "lqr_sawyer, 4 systems x 500 steps from +-0.01 rad")
json.dump(out, open(os.path.join(ROOT, "profiles", "flops_r2.json"), "w"), indent=1)
