"""device-time probe of the synthesis kernels: python tools_synth_probe.py "name,S[,N];..." ; CCB_NO_DMMA=1 selects K3"""
import sys, os, math, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, ccb200 as cc
from helpers import zd_of
spec = sys.argv[1] if len(sys.argv) > 1 else "cartpole,4096"
for item in spec.split(";"):
    f = item.split(","); name, S = f[0], int(f[1])
    if name == "sawyer": mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
    elif name.startswith("npend"): mech, cfg = cc.examples.cartpole_n_pendulum(int(name[5:]))
    else: mech, cfg = getattr(cc.examples, name)()
    hz = cfg.get("horizon", 10.0)
    N = int(f[2]) if len(f) > 2 else (int(math.ceil(hz / mech.dt)) if hz != math.inf else 1000)
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    zd = zd_of(mech, cfg); Fd = np.array(cfg.get("Ftaud", np.zeros(len(eqc_idx))), dtype=float)
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
    rng = np.random.default_rng(2)
    qs = np.exp(rng.uniform(np.log(0.1), np.log(10), S)); rs = np.exp(rng.uniform(np.log(0.1), np.log(10), S))
    Qb = Qd[None] * qs[:, None, None]; Rb = Rd[None] * rs[:, None, None]
    t0 = time.time()
    o = cc.lqr_synthesize(mech, np.repeat(zd[None], S, 0), np.repeat(Fd[None], S, 0), eqc_idx, Qb, Rb, N)
    wall = time.time() - t0
    ph = cc.synth.phase_ms()
    n = 12 * mech.nb; nu = len(eqc_idx)
    iters = float(o["iters"].mean())
    T = (n + 7) // 8                 # executed flops: P Abar in full, the symmetric P+ as (T/2+1) of T tile columns per strip
    flops = S * iters * (2.0 * n ** 3 * (1.0 + (T // 2 + 1) / T) + 8.0 * nu * n * n)
    print(f"{name} n={n} nu={nu} S={S} N={N} iters={iters:.0f}: lin {ph['linearize']:.2f} ms, project {ph['project']:.2f} ms, sweep {ph['sweep']:.2f} ms "
          f"-> {S / (sum(ph.values()) * 1e-3):.0f} synth/s device, {S / wall:.0f} e2e; sweep {flops / (ph['sweep'] * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)
