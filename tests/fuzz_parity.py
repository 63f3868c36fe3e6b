"""Randomised parity sweep of the rollout path (GPU through the C-ABI vs the CPU oracle).

    python tests/fuzz_parity.py [ncases] [seed] > profiles/fuzz_r2.jsonl

Every case draws a mechanism, a control law, a batch size around the kernels' quantisation edges (warp fill, persistent-launch
thresholds), a step count, a kernel variant (`lanes`), recording on/off and whether the rollout is issued as one call or as two
consecutive calls (resume from the stored state / multipliers / step index).  It prints one JSON line per case: max relative error of
the final states (and of the recorded trajectory), and whether the per-system newton! iteration totals equal the oracle's.  The oracle
is the checker here (test infrastructure, not the product path).
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
import ccb200 as cc
import orc
from helpers import cartpole_states, oracle_lqr_gains, rel_err

GOLD = os.path.join(ROOT, "tests", "golden")


def perturbed_states(mech, B, scale, seed):
    """initial states around the mechanism's own pose: small random body velocities (the integrator pulls them onto the constraints)"""
    r = np.random.default_rng(seed)
    st = np.repeat(mech.state_vector()[None], B, axis=0)
    st[:, :, 7:13] += scale * r.standard_normal((B, mech.nb, 6))
    return st


MECHS = {
    "pendulum": lambda: cc.examples.pendulum(),
    "cartpole": lambda: cc.examples.cartpole(),
    "acrobot": lambda: cc.examples.acrobot(),
    "doublependulum": lambda: cc.examples.doublependulum(),
    "triple_cartpole": lambda: cc.examples.triple_cartpole(),
    "npend5": lambda: cc.examples.cartpole_n_pendulum(5),
    "npend8": lambda: cc.examples.cartpole_n_pendulum(8),
    "sawyer": lambda: cc.examples.sawyer(os.path.join(GOLD, "sawyer_arm.json")),
    "deltabot": lambda: cc.examples.deltabot(),
}
BATCHES = [1, 2, 3, 4, 5, 6, 7, 31, 32, 33, 95, 96, 97, 191, 257]


def run(ncases=60, seed=7, emit=lambda line: print(line, flush=True)):
    rng = np.random.default_rng(seed)
    worst = 0.0
    nbad = nerr = 0
    for case in range(ncases):
        name = list(MECHS)[int(rng.integers(len(MECHS)))]
        mech, cfg = MECHS[name]()
        B = int(BATCHES[int(rng.integers(len(BATCHES)))])
        steps = int(rng.choice([1, 2, 17, 64, 120, 255, 256, 257, 300]))
        if name in ("sawyer", "npend8", "deltabot"):
            B, steps = min(B, 33), min(steps, 120)
        law = str(rng.choice(["none", "lqr_inf", "lqr_finite"])) if "xd" in cfg and "horizon" in cfg else "none"
        if name == "doublependulum":
            law = str(rng.choice(["none", "pid"]))
        if name in ("npend5", "npend8"):
            law = "none"                       # uncontrolled chains: short horizons only (chaotic)
            steps = min(steps, 64)
        lanes = int(rng.choice([0, 0, 0, 1, 8, 32])) if name not in ("deltabot",) else int(rng.choice([0, 8, 32]))
        record = bool(rng.integers(2))
        split = bool(rng.integers(2)) and steps >= 2
        if name in ("cartpole", "npend5", "npend8"):
            st0 = cartpole_states(cc, mech, B, seed=case, phimax=0.15)
        elif name == "deltabot":
            st0 = cc.examples.deltabot_states(mech, B)
        else:
            st0 = perturbed_states(mech, B, 0.05, case)
        om = orc.OracleMech(mech.desc())
        ctrl_o, ctrl = None, None
        if law in ("lqr_inf", "lqr_finite"):
            o = oracle_lqr_gains(cc, orc, mech, cfg, horizon=(math.inf if law == "lqr_inf" else max(2, steps // 2) * mech.dt))
            if law == "lqr_inf":
                ctrl_o = om.make_ctrl(B, kind="lqr_inf", eqc=o["eqc_idx"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
                ctrl = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], math.inf, xd=cfg.get("xd"), qd=cfg.get("qd"), Ftaud=cfg.get("Ftaud"))
            else:
                ctrl_o = om.make_ctrl(B, kind="lqr_finite", eqc=o["eqc_idx"], N=o["N"], K=o["K"], zd=o["zd"], Fd=o["Fd"])
                ctrl = cc.LQR.from_gains(mech, cfg["eqcids"], o["K"], o["N"], xd=cfg.get("xd"), qd=cfg.get("qd"), Ftaud=cfg.get("Ftaud"))
        if law == "pid":
            goals = rng.uniform(-2.5, 2.5, size=(B, 2))
            eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
            integ = np.zeros((B, 2)); last = np.zeros((B, 2))
            ctrl_o = om.make_ctrl(B, kind="pid", eqc=eqc, P=cfg["P"], I=cfg["I"], D=cfg["D"], goal=goals, pid_int=integ, pid_last=last)
            ctrl = cc.PID(mech, cfg["eqcids"], goals, P=cfg["P"], I=cfg["I"], D=cfg["D"])
        ref = om.simulate(st0.copy(), steps, ctrl=ctrl_o, record=record, nthreads=(1 if law == "pid" else 8))
        batch = cc.Batch(mech, B)
        batch.set_state(st0)
        sto = None
        try:
            if split:
                s1 = steps // 2
                a, st_a = batch.simulate(ctrl, 1, s1, lanes=lanes, record=record)
                b, st_b = batch.simulate(ctrl, 1 + s1, steps - s1, lanes=lanes, record=record)
                status = np.where((st_a > 0) & (st_b > 0), 1, -1) * (np.abs(st_a) + np.abs(st_b))      # negative = some step failed (ccb.h)
                if record:
                    sto = np.concatenate([a, b], axis=0)
            else:
                sto, status = batch.simulate(ctrl, 1, steps, lanes=lanes, record=record)
            final = batch.get_state()
            ok = ref["iters"] > 0                       # systems whose every step converged in the oracle; failures are compared by flag
            fail_eq = bool(np.array_equal(status > 0, ok))
            e_fin = rel_err(final[ok], ref["states"][ok]) if ok.any() else 0.0
            e_sto = (rel_err(sto[:, ok], ref["storage"][:, ok]) if ok.any() else 0.0) if record else None
            it_eq = bool(np.array_equal(status[ok], ref["iters"][ok]))
            it_diff = int(np.abs(status[ok] - ref["iters"][ok]).sum())
            err = None
        except Exception as ex:      # an ineligible variant (e.g. lanes=1 on a non-chain) is reported, not hidden
            e_fin = e_sto = None; it_eq = None; it_diff = None; fail_eq = None; ok = np.zeros(0, bool); err = f"{type(ex).__name__}: {ex}"[:160]
        rec = dict(case=case, mech=name, law=law, B=B, steps=steps, lanes=lanes, record=record, split=split, err_final=e_fin, err_storage=e_sto,
                   newton_counts_equal=it_eq, newton_count_diff=it_diff, oracle_failed_systems=int((~ok).sum()), failure_flags_equal=fail_eq, error=err)
        emit(json.dumps(rec))
        if err is None:
            m = max(e_fin, e_sto or 0.0)
            worst = max(worst, m)
            if not (m <= 1e-9) or not fail_eq:
                nbad += 1
        else:
            nerr += 1
    summary = dict(summary=True, cases=ncases, worst_rel_err=worst, cases_above_1e9=nbad, errors=nerr)
    emit(json.dumps(summary))
    return summary


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 60, int(sys.argv[2]) if len(sys.argv) > 2 else 7)
