"""shared helpers for the parity tests (oracle = checker only)"""
import math

import numpy as np


def zd_of(mech, cfg):
    zd = np.zeros((mech.nb, 13))
    zd[:, 3] = 1.0
    if cfg.get("xd") is not None:
        zd[:, 0:3] = np.array(cfg["xd"])
    if cfg.get("qd") is not None:
        zd[:, 3:7] = np.array(cfg["qd"])
    return zd


def oracle_lqr_gains(cc, orc, mech, cfg, horizon=None):
    """LQR constructor logic (lqr.jl:17-47) evaluated with the oracle: returns K [rows][nu][n], N, zd, Fd, eqc_idx"""
    om = orc.OracleMech(mech.desc())
    zd = zd_of(mech, cfg)
    eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
    Fd = np.array(cfg.get("Ftaud", np.zeros(len(eqc_idx))), dtype=float)
    lin = om.linearize(zd, Fd, eqc_idx)
    Q = cc.blockdiag(cfg["Q"]) * mech.dt
    R = cc.blockdiag(cfg["R"]) * mech.dt
    horizon = cfg["horizon"] if horizon is None else horizon
    if horizon == math.inf:
        N, Ntemp = math.inf, int(math.ceil(10 / mech.dt))
    else:
        N = Ntemp = int(math.ceil(horizon / mech.dt))
    res = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Q, R, Ntemp)
    K = res["Ku"][:1] if N == math.inf else res["Ku"]
    return dict(K=K, N=N, zd=zd, Fd=Fd, eqc_idx=eqc_idx, om=om, lin=lin, res=res)


def _log(kind, val):
    """CCB_ERRLOG=<file>: every error a test evaluates is appended with its call site (tools: which bounds are tight, which are loose)"""
    import os, sys
    path = os.environ.get("CCB_ERRLOG")
    if path:
        f = sys._getframe(2)
        with open(path, "a") as fh:
            fh.write(f"{os.path.basename(f.f_code.co_filename)}:{f.f_lineno} {f.f_code.co_name} {kind} {val:.3e}\n")


def rel_err(a, b):
    a = np.asarray(a); b = np.asarray(b)
    e = float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))
    _log("rel", e)
    return e


def gain_err(a, b, floor=1e-3):
    """ELEMENT-relative error of a gain table against its reference: max_ij |a_ij - b_ij| / max(|b_ij|, floor * s_row),
    s_row = the largest |b| of the entry's own gain row (last axis).  A global scale lets the small entries of a row whose large entries are ~350
    be wrong by 3.5e-7 absolute; here an entry may only be off by 1e-9 of ITS OWN magnitude, down to entries 1000x smaller than the row's largest
    (below that an entry is rounding residue of the row's large entries -- structurally zero columns come out as 1e-13 in every implementation)."""
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), floor * np.abs(b).max(axis=-1, keepdims=True))
    scale = np.where(scale == 0.0, 1.0, scale)
    e = float((np.abs(a - b) / scale).max())
    _log("elem", e)
    return e


def cartpole_states(cc, mech, B, seed=1, ymax=0.5, phimax=0.2):
    """y0 ~ U(-0.5,0.5), theta0 ~ U(-0.2,0.2), zero velocity"""
    rng = np.random.default_rng(seed)
    out = np.zeros((B, mech.nb, 13))
    for s in range(B):
        cc.examples.set_cartpole_pose(mech, rng.uniform(-ymax, ymax), rng.uniform(-phimax, phimax, size=mech.nb - 1))
        out[s] = mech.state_vector()
    return out
