"""Multi-precision evidence for the ill-conditioned upright n-pendulum LQR (BASELINE config C3).

    python tools/quad_gains.py <links> <N> [--no-quad]

Runs the oracle's literal dlqr (lqr.jl:141-184) on the Float64 linearisation of examples/lqr_cartpole_n_pendulum.jl with the recursion carried
in Float64, x87 extended and IEEE binary128 (oracle/oracle.cpp: orc_dlqr_mp), then
  * writes tests/golden/npend_quad_n<links>_N<N>.npz: selected rows of the binary128 gains (the "true" gains of the Float64 problem data),
    the Float64-oracle-vs-binary128 error of those rows (the rounding the recursion itself amplifies: the floor ANY Float64 implementation,
    the reference included, sits on), max |P|, iteration counts;
  * simulates the closed loop (oracle, Float64) with the binary128 gains and with the Float64 gains from the example's initial conditions and
    appends one JSON line to profiles/npend_quad_report.jsonl saying whether the loop is stabilised.
TEST INFRASTRUCTURE (oracle side): the product never reads these files; tests/test_oracle_quad.py (CPU) and tests/test_gpu_quad.py (GPU) do.
"""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import ccb200 as cc, orc
from helpers import zd_of, gain_err

links, N = int(sys.argv[1]), int(sys.argv[2])
top = 1 if "--no-quad" in sys.argv else 2
orc.build()
mech, cfg = cc.examples.cartpole_n_pendulum(links)
om = orc.OracleMech(mech.desc())
zd = zd_of(mech, cfg); eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
lin = om.linearize(zd, np.zeros(len(eqc)), eqc)
Q = cc.blockdiag(cfg["Q"]) * mech.dt; R = cc.blockdiag(cfg["R"]) * mech.dt
res, secs = {}, {}
for prec in range(top + 1):
    t0 = time.time()
    res[prec] = orc.dlqr_mp(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Q, R, N, prec=prec)
    secs[prec] = time.time() - t0
    print(f"links {links} N {N} prec {prec}: {secs[prec]:.1f} s, iters {res[prec]['iters']}, max|P| {res[prec]['Pmax']:.3e}", flush=True)
ref = res[top]
rows = sorted(set([0, (N - 1) // 4, (N - 1) // 2, 3 * (N - 1) // 4, N - 2]))
errs = {p: dict(glob=float(np.abs(res[p]["Ku"] - ref["Ku"]).max() / max(1.0, np.abs(ref["Ku"]).max())),
                elem=float(gain_err(res[p]["Ku"], ref["Ku"]))) for p in range(top)}
# spread of the literal Float64 recursion under 1-ulp perturbations of its inputs: 16 realisations with every non-zero entry of A, Bu, Bλ, G
# multiplied by (1 + u 2^-52), u in {-1, 0, 1}.  The recursion amplifies such a perturbation exactly as it amplifies its own rounding, so the
# per-row maximum over the realisations is the band inside which ANY Float64 implementation of dlqr lands (tests/test_gpu_fullsize.py holds the
# CUDA recursion to it).
pert_rows = np.zeros((0, len(rows)))
if links <= 8:
    prng = np.random.default_rng(7)
    def jitter(M):
        return M * (1.0 + prng.integers(-1, 2, size=M.shape) * 2.0 ** -52)
    errs_p = []
    for t in range(16):
        rp = orc.dlqr_mp(jitter(lin["A"]), jitter(lin["Bu"]), jitter(lin["Bl"]), jitter(lin["G"]), Q, R, N, prec=0)
        errs_p.append([gain_err(rp["Ku"][r], ref["Ku"][r]) for r in rows])
    pert_rows = np.array(errs_p)
    print("1-ulp input perturbations, per-row max element-relative error of 16 Float64 realisations:", pert_rows.max(axis=0), flush=True)
np.savez(os.path.join(ROOT, "tests", "golden", f"npend_quad_n{links}_N{N}.npz"), links=links, N=N, rows=np.array(rows), precision=top,
         f64_perturbed_rows=pert_rows,
         Ku_ref_rows=ref["Ku"][rows], Ku_f64_rows=res[0]["Ku"][rows], iters_ref=ref["iters"], iters_f64=res[0]["iters"], Pmax=ref["Pmax"],
         f64_vs_ref_glob=errs[0]["glob"], f64_vs_ref_elem=errs[0]["elem"],
         f64_vs_ref_elem_rows=np.array([gain_err(res[0]["Ku"][r], ref["Ku"][r]) for r in rows]))

# closed loop from the example's initial conditions (phi_i ~ U(0,1)/3^n, y ~ U(-0.5,0.5); examples/lqr_cartpole_n_pendulum.jl:21-22)
rng = np.random.default_rng(2)
B = 8
st = np.zeros((B, mech.nb, 13))
for s in range(B):
    cc.examples.set_cartpole_pose(mech, rng.uniform(-0.5, 0.5), rng.uniform(0, 1, size=links) / 3 ** links)
    st[s] = mech.state_vector()
loop = {}
for name, K in (("ref_gains", ref["Ku"]), ("f64_gains", res[0]["Ku"])):
    co = om.make_ctrl(B, kind="lqr_finite", eqc=eqc, N=N, K=K, zd=zd, Fd=np.zeros(len(eqc)))
    r = om.simulate(st, N - 1, ctrl=co, nthreads=8)
    dev = np.abs(r["states"][:, :, 0:3] - zd[None, :, 0:3])
    ok = np.isfinite(dev).all(axis=(1, 2)) & (dev[:, 1:, :].max(axis=(1, 2)) < 0.5)
    loop[name] = dict(upright_after_horizon=int(ok.sum()), of=B, newton_failed=int((r["iters"] <= 0).sum()),
                      max_pole_deviation_m=float(np.nanmax(dev[:, 1:, :])) if np.isfinite(dev).any() else None)
rec = dict(links=links, n=12 * mech.nb, N=N, reference_precision=["f64", "x87 extended", "binary128"][top], seconds=secs,
           iters={p: res[p]["iters"] for p in res}, Pmax=ref["Pmax"], K1_max=float(np.abs(ref["Ku"][0]).max()),
           f64_vs_ref=errs[0], x87_vs_ref=errs.get(1), closed_loop=loop)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "npend_quad_report.jsonl"), "a") as f:
    f.write(json.dumps(rec) + "\n")
print(json.dumps(rec))
