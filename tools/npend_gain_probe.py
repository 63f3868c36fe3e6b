"""Gain parity of the n-pendulum LQR synthesis against the oracle at several horizons (conditioning probe)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import ccb200 as cc
import orc                      # test infrastructure: this tool is a checker, not product
from helpers import zd_of


def main():
    npend = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    mech, cfg = cc.examples.cartpole_n_pendulum(npend)
    om = orc.OracleMech(mech.desc())
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    orc.build()
    zd = zd_of(mech, cfg)
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
    lin = om.linearize(zd, [0.0], eqc_idx)
    for N in ([int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else (2, 5, 20, 80, 300)):
        out = cc.lqr_synthesize(mech, zd[None], np.zeros((1, 1)), eqc_idx, Qd, Rd, N)
        ref = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Qd, Rd, N)
        e = np.abs(out["Ku"][0] - ref["Ku"]).max() / np.abs(ref["Ku"]).max()
        print(f"npend{npend} N={N}: kexit {out['kexit'][0]} vs {ref['kexit']}, rel err {e:.3e}, |K| {np.abs(ref['Ku']).max():.3e}")
    # error of the gain of every 20th step of the longest horizon (index k-1 holds the gain of step k; the recursion runs k = N-1 .. 1)
    Kg, Kr = out["Ku"][0].reshape(N - 1, -1), np.asarray(ref["Ku"]).reshape(N - 1, -1)
    for k in range(N - 2, -1, -20):
        print(f"   k={k + 1}: rel err {np.abs(Kg[k] - Kr[k]).max() / np.abs(Kr[k]).max():.3e}  |K| {np.abs(Kr[k]).max():.3e}")


if __name__ == "__main__":
    main()
