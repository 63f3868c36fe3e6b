"""One-off randomized parity sweep (GPU vs oracle) beyond the seeds of the test-suite: dlqr over many (n, nu, ml) shapes (K3 / K4 / K4b),
algebraic Riccati equations (K5), random serial chains on every rollout kernel.  Prints the worst relative error per family."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import ccb200 as cc, orc
import are as orc_are
from helpers import rel_err
import test_gpu_random_mechanisms as trm

orc.build()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 123)

# ---- dlqr shapes
worst = {}
t0 = time.time()
shapes = [(int(n), int(rng.integers(1, 9)), int(rng.integers(0, max(1, n // 3)))) for n in rng.choice(np.arange(8, 209, 4), size=28, replace=False)]
shapes += [(10, 2, 3), (14, 1, 4), (22, 3, 5), (108, 8, 30), (112, 1, 20), (136, 4, 40), (200, 2, 50)]
for n, nu, ml in shapes:
    N = 16
    S = 2
    A = rng.normal(size=(S, n, n)) * (0.9 / np.sqrt(n))
    Bu = rng.normal(size=(S, n, nu)); G = rng.normal(size=(S, ml, n))
    Bl = np.swapaxes(G, 1, 2) + 0.1 * rng.normal(size=(S, n, ml))
    L = rng.normal(size=(n, n)); Q = (L @ L.T / n + np.eye(n)) * 0.01; R = np.eye(nu) * 0.02
    out = cc.dlqr(A, Bu, Bl, G, Q, R, N)
    for s in range(S):
        ref = orc.dlqr(A[s], Bu[s], Bl[s], G[s], Q, R, N)
        e = rel_err(out["Ku"][s], ref["Ku"])
        fam = "K4b" if n > 108 else ("K4" if n % 4 == 0 else "K3")
        worst[fam] = max(worst.get(fam, 0.0), e)
        assert out["kexit"][s] == ref["kexit"], (n, nu, ml)
        assert e < 1e-9, (n, nu, ml, e)
print("dlqr", {k: f"{v:.1e}" for k, v in worst.items()}, f"{len(shapes)} shapes, {time.time() - t0:.1f} s", flush=True)

# ---- ARE
w = 0.0
for n, m in [(3, 1), (5, 2), (9, 3), (17, 4), (33, 5), (57, 8), (64, 3), (80, 2)]:
    A = rng.standard_normal((4, n, n)) / np.sqrt(n) * 1.1; B = rng.standard_normal((4, n, m))
    Lq = rng.standard_normal((n, n)); Q = Lq @ Lq.T / n + np.eye(n); R = np.eye(m) * 0.5
    X, info = cc.util.dare(A, B, Q, R, info=True)
    assert (info["status"] == 0).all()
    for s in range(4):
        w = max(w, rel_err(X[s], orc_are.dare(A[s], B[s], Q, R)))
print(f"dare worst {w:.1e}", flush=True)
assert w < 1e-8

# ---- random chains, every kernel (the test body with more seeds)
class _CC: pass
t0 = time.time()
cnt = 0
for seed in range(40, 52):
    nj = int(rng.integers(1, 13))
    trm.test_random_chain_rollout_every_kernel.__wrapped__(cc, orc, seed, nj) if hasattr(trm.test_random_chain_rollout_every_kernel, "__wrapped__") else trm.test_random_chain_rollout_every_kernel(cc, orc, seed, nj)
    cnt += 1
print(f"random chains: {cnt} mechanisms on every kernel ok, {time.time() - t0:.1f} s", flush=True)
print("STRESS OK")
