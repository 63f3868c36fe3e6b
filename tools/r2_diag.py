"""round-2 diagnostics on the GPU: full-size parity numbers behind tests/test_gpu_fullsize.py (prints JSON lines)"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import ccb200 as cc, orc
from helpers import cartpole_states, gain_err, oracle_lqr_gains, rel_err, zd_of
orc.build()
GOLD = os.path.join(ROOT, "tests", "golden")
which = sys.argv[1:] or ["sawyer", "tracking", "counts", "npend"]

def J(**kw):
    print(json.dumps(kw, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)), flush=True)

if "sawyer" in which:
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    om = orc.OracleMech(mech.desc()); eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]; zd = zd_of(mech, cfg)
    Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt; N = 2000
    out = cc.lqr_synthesize(mech, zd[None], np.zeros((1, 7)), eqc, Qd, Rd, N)
    lin = om.linearize(zd, np.zeros(7), eqc)
    f64 = orc.dlqr_mp(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Qd, Rd, N, prec=0)
    x87 = orc.dlqr_mp(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Qd, Rd, N, prec=1)
    it = [int(out["iters"][0]), f64["iters"], x87["iters"]]
    lo = min(it); kc = N - lo
    K = out["Ku"][0]
    J(test="sawyer_N2000", iters_gpu_f64_x87=it, kexit_gpu=int(out["kexit"][0]),
      gain_err_common_vs_f64=gain_err(K[kc - 1:], f64["Ku"][kc - 1:]), gain_err_common_vs_x87=gain_err(K[kc - 1:], x87["Ku"][kc - 1:]),
      f64_vs_x87_common=gain_err(f64["Ku"][kc - 1:], x87["Ku"][kc - 1:]),
      gain_err_all_vs_f64=gain_err(K, f64["Ku"]), gain_err_all_vs_x87=gain_err(K, x87["Ku"]), rel_all_vs_f64=rel_err(K, f64["Ku"]), rel_all_vs_x87=rel_err(K, x87["Ku"]),
      per_row_err_vs_x87=[gain_err(K[r], x87["Ku"][r]) for r in (0, 200, 400, 420, 1000, 1500, 1998)],
      per_row_f64_vs_x87=[gain_err(f64["Ku"][r], x87["Ku"][r]) for r in (0, 200, 400, 420, 1000, 1500, 1998)])

if "tracking" in which:
    mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
    om = orc.OracleMech(mech.desc()); eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    N = 2000
    rng = np.random.default_rng(3)
    a = rng.uniform(0, 1, 7) * np.array([0.05, 0.05, 0.02, 0.02, 0.005, 0.005, 0.0005]); f = rng.integers(1, 4, 7)
    U = a[None, :] * np.sin(2 * np.pi * f[None, :] * np.arange(N)[:, None] / 2000.0)
    cc.urdf.zero_configuration(mech)
    st0 = mech.state_vector()[None].copy()
    sto, _, status = cc.simulate(mech, N * mech.dt, cc.OpenLoop(mech, cfg["eqcids"], U), record=True, states=st0, return_status=True)
    ro = om.simulate(st0, N, ctrl=om.make_ctrl(1, kind="table", eqc=eqc, U=U), record=True)
    tl = cc.TrackingLQR(mech, sto, U.reshape(1, N, 7), cfg["eqcids"], cfg["Q"], cfg["R"])
    ref = om.tracking_dlqr(sto.data[:, 0], U, eqc, cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt)
    x87 = om.tracking_dlqr(sto.data[:, 0], U, eqc, cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, prec=1)
    K = tl.K[0]
    J(test="tracking_N2000_vs_x87", gpu_vs_x87=gain_err(K, x87["Ku"]), f64_vs_x87=gain_err(ref["Ku"], x87["Ku"]), kexit_x87=x87["kexit"])
    J(test="tracking_N2000", openloop_rel=rel_err(sto.data, ro["storage"]), iters_gpu=int(status[0]), iters_orc=int(ro["iters"][0]), kexit_gpu=int(tl.kexit[0]), kexit_orc=ref["kexit"],
      gain_err=gain_err(K, ref["Ku"]), rel=rel_err(K, ref["Ku"]), per_row=[gain_err(K[r], ref["Ku"][r]) for r in (0, 500, 1000, 1500, 1990, 1998)],
      Kmax=float(np.abs(ref["Ku"]).max()))

if "counts" in which:
    def cases():
        mech, cfg = cc.examples.cartpole()
        yield "C2_cartpole", mech, cfg, cartpole_states(cc, mech, 96, seed=1), 2000, 20.0, (0, 1, 8)
        mech, cfg = cc.examples.pendulum()
        rng = np.random.default_rng(11); st0 = mech.batch_states(48)
        for s in range(48):
            cc.mechanism.setPosition(mech.origin, mech.bodies[0], p2=np.array([0.0, 0.0, 0.5]), dq=cc.mechanism.RotX(math.pi + rng.uniform(-0.6, 0.6)))
            st0[s] = mech.state_vector()
        yield "C1_pendulum", mech, cfg, st0, 1000, math.inf, (0, 1, 8)
        mech, cfg = cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json"))
        rng = np.random.default_rng(3); st0 = np.zeros((24, mech.nb, 13))
        for s in range(24):
            cc.urdf.zero_configuration(mech, {i: rng.uniform(-0.01, 0.01) for i in range(7)}); st0[s] = mech.state_vector()
        cc.urdf.zero_configuration(mech)
        yield "C4_sawyer", mech, cfg, st0, 600, 20.0, (0, 1)
    for name, mech, cfg, st0, steps, horizon, lanes_list in cases():
        B = st0.shape[0]
        o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=horizon)
        kind = "lqr_inf" if horizon == math.inf else "lqr_finite"
        lqr = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=o["zd"][:, 0:3], qd=o["zd"][:, 3:7], Ftaud=o["Fd"])
        for eps in (1e-10, 1e-7):
            co = o["om"].make_ctrl(B, kind=kind, eqc=o["eqc_idx"], N=0 if horizon == math.inf else o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
            r = o["om"].simulate(st0, steps, ctrl=co, nthreads=8, eps=eps, marginal_band=3.0)
            for lanes in lanes_list:
                batch = cc.Batch(mech, B); batch.set_state(st0)
                _, status = batch.simulate(lqr, 1, steps, lanes=lanes, eps=eps)
                d = status - r["iters"]
                J(test="counts", config=name, eps=eps, lanes=lanes, state_rel=rel_err(batch.get_state(), r["states"]), systems_differing=int((d != 0).sum()), of=B,
                  max_abs_diff=int(np.abs(d).max()), sum_diff=int(d.sum()), total_iters=int(r["iters"].sum()), marginal_total=int(r["marginal"].sum()),
                  covered=bool(np.all(np.abs(d) <= r["marginal"])))

if "npend" in which:
    for links, N in ((8, 300), (8, 1000)):
        g = np.load(os.path.join(GOLD, f"npend_quad_n{links}_N{N}.npz"))
        mech, cfg = cc.examples.cartpole_n_pendulum(links)
        eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]; zd = zd_of(mech, cfg)
        for env in ("", "1"):
            if env:
                os.environ["CCB_NO_DMMA"] = "1"
            else:
                os.environ.pop("CCB_NO_DMMA", None)
            out = cc.lqr_synthesize(mech, zd[None], np.zeros((1, 1)), eqc, cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, N)
            K = out["Ku"][0][g["rows"]]
            J(test="npend", links=links, N=N, kernel="K3" if env else "K4", rows=g["rows"], gpu_vs_ref_rows=[gain_err(K[i], g["Ku_ref_rows"][i]) for i in range(len(g["rows"]))],
              f64_vs_ref_rows=g["f64_vs_ref_elem_rows"], gpu_vs_ref_glob=rel_err(K, g["Ku_ref_rows"]), f64_vs_ref_glob=rel_err(g["Ku_f64_rows"], g["Ku_ref_rows"]))
        os.environ.pop("CCB_NO_DMMA", None)
