"""Randomised parity sweep of the constrained Riccati recursion (GPU `dlqr` through the C-ABI vs the CPU oracle).

    python tests/fuzz_synth.py [ncases] [seed] > profiles/fuzz_synth_r2.jsonl

Random state / input / constraint dimensions across the kernel-selection boundaries (K3 FMA sweep, K4 DMMA sweep with its staged
variants and the scalar-input path, K4b blocked sweep for n > 108), random horizons including the shortest ones (N = 2: one gain row) and
random batch sizes.  One JSON line per case: worst element-relative gain error (tests/helpers.gain_err), whether the break index and the
iteration count equal the oracle's.  The oracle is the checker (test infrastructure, not the product path).
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
import ccb200 as cc
import orc
from helpers import gain_err, rel_err


def run(ncases=60, seed=5, emit=lambda line: print(line, flush=True)):
    rng = np.random.default_rng(seed)
    worst, nbad = 0.0, 0
    for case in range(ncases):
        n = int(rng.choice([1, 2, 3, 7, 8, 9, 12, 16, 23, 24, 25, 36, 47, 48, 64, 71, 72, 84, 85, 96, 104, 108, 109, 116, 132, 160, 204]))
        nu = int(rng.integers(1, min(8, n) + 1))
        ml = int(rng.integers(0, max(1, min(n - nu, n // 2) + 1)))
        N = int(rng.choice([2, 3, 4, 17, 60, 150]))
        if n > 108:
            N = min(N, 17)
        S = int(rng.choice([1, 2, 5, 33, 149])) if n <= 48 else int(rng.choice([1, 2, 5]))
        A = rng.normal(size=(S, n, n)) * (0.9 / np.sqrt(n))
        Bu = rng.normal(size=(S, n, nu)); G = rng.normal(size=(S, ml, n))
        Bl = np.swapaxes(G, 1, 2) + 0.1 * rng.normal(size=(S, n, ml))          # G Bl ~ G G': a well conditioned projection
        L = rng.normal(size=(n, n)); Q = (L @ L.T / n + np.eye(n)) * 0.01
        Lr = rng.normal(size=(nu, nu)); R = (Lr @ Lr.T / nu + np.eye(nu)) * 0.02
        try:
            out = cc.dlqr(A, Bu, Bl, G, Q, R, N)
            e, eg, k_eq, it_eq = 0.0, 0.0, True, True
            for s in range(S):
                ref = orc.dlqr(A[s], Bu[s], Bl[s], G[s], Q, R, N)
                e = max(e, gain_err(out["Ku"][s], ref["Ku"]))
                eg = max(eg, rel_err(out["Ku"][s], ref["Ku"]))
                k_eq &= bool(out["kexit"][s] == ref["kexit"])
                it_eq &= bool(out["iters"][s] == ref["iters"])
            err = None
        except Exception as ex:
            e = eg = None; k_eq = it_eq = None; err = f"{type(ex).__name__}: {ex}"[:160]
        emit(json.dumps(dict(case=case, n=n, nu=nu, ml=ml, N=N, S=S, gain_err_elementwise=e, gain_err_global=eg, kexit_equal=k_eq, iters_equal=it_eq, error=err)))
        if err is None:
            worst = max(worst, e)
            if not (e <= 1e-9) or not k_eq or not it_eq:
                nbad += 1
        else:
            nbad += 1
    summary = dict(summary=True, cases=ncases, worst_gain_err=worst, cases_failing=nbad)
    emit(json.dumps(summary))
    return summary


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 5)
