"""Loader for golden vectors dumped by the REAL reference (tools/dump_golden.jl, needs a Julia toolchain -- absent from the build and GPU images).

When tests/golden/julia/manifest.txt exists, every case is rebuilt here from the same literals, and
  * the oracle (CPU, always)  and  * the CUDA path (-m gpu)
are compared with the reference's own A, Bu, Bλ, G, gains, recorded trajectories (1e-9 relative) and graph orderings (bit-exact).
Without the fixtures every test SKIPS and parity with the upstream implementation stays unpinned: the oracle is then
anchored only by the known-answer / invariance tests of tests/test_oracle_*.py.
"""
import math
import os

import numpy as np
import pytest

from helpers import gain_err, rel_err, zd_of

HERE = os.path.dirname(os.path.abspath(__file__))
JDIR = os.path.join(HERE, "golden", "julia")
MANIFEST = os.path.join(JDIR, "manifest.txt")
TOL = 1e-9

pytestmark = pytest.mark.skipif(not os.path.exists(MANIFEST), reason="no golden vectors of the Julia reference (run tools/dump_golden.jl)")


def load_case(case):
    """arrays of one case; Julia wrote column-major, so shapes are reversed: (13, Nb, N) -> [N][Nb][13], (NK, nu, rows) -> [rows][nu][NK]"""
    out = {}
    for line in open(MANIFEST):
        c, name, dt, shape = line.split()
        if c != case:
            continue
        dims = [int(v) for v in shape.split("x")] if shape else []
        raw = np.fromfile(os.path.join(JDIR, f"{c}__{name}.{dt}"), dtype=np.float64 if dt == "f64" else np.int64)
        out[name] = raw.reshape(dims[::-1])        # C-order view of the column-major data = transposed axes
    return out


def parse_manifest(text):
    return sorted({ln.split()[0] for ln in text.splitlines() if ln.strip()})


def _lqr_cases(cc):
    golden = os.path.join(HERE, "golden", "sawyer_arm.json")
    def c1():
        return cc.examples.pendulum() + (math.inf, 1000)
    def c2(h):
        mech, cfg = cc.examples.cartpole(0.5, 0.2)
        return mech, cfg, h, int(h * 100)
    def c3():
        mech, cfg = cc.examples.cartpole_n_pendulum(3, 0.2, [0.02, 0.01, 0.005])
        return mech, cfg, 10.0, 1000
    def c4():
        mech, cfg = cc.examples.sawyer(golden)
        cc.urdf.zero_configuration(mech, {i: 0.01 for i in range(7)})
        return mech, cfg, 20.0, 200
    return {"c1_pendulum": c1, "c2_cartpole_h10": lambda: c2(10.0), "c2_cartpole_h20": lambda: c2(20.0), "c3_npendulum3": c3, "c4_sawyer": c4}


def _graph_check(cc, mech, g):
    """integer orderings, bit-exact: ids of eqcs / bodies and the DFS list of the mixed-node graph"""
    mine = cc.handle(mech).graph()
    assert np.array_equal(mine["eqc_id"], g["eqc_ids"].ravel())
    assert np.array_equal(mine["body_id"], g["body_ids"].ravel())
    assert np.array_equal(mine["dfs_list"], g["dfs_list"].ravel())


CASES = ["c1_pendulum", "c2_cartpole_h10", "c2_cartpole_h20", "c3_npendulum3", "c4_sawyer"]


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_lqr(cc, orc, case):
    g = load_case(case)
    if not g:
        pytest.skip(f"{case} not in the manifest")
    mech, cfg, horizon, steps = _lqr_cases(cc)[case]()
    _graph_check(cc, mech, g)
    st0 = mech.state_vector()[None].copy()
    assert rel_err(st0[0], g["state0"]) < 1e-12
    om = orc.OracleMech(mech.desc())
    zd = zd_of(mech, cfg); eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    lin = om.linearize(zd, np.zeros(len(eqc)), eqc)
    # Julia dumped (n, n) column-major: the reversed-shape view is the transpose
    for mine, ref in ((lin["A"], g["A"].T), (lin["Bu"], g["Bu"].T), (lin["G"], g["G"].T)):
        assert rel_err(mine, ref) < TOL
    # Bλ / λ scaling is a convention of the constraint rows (invariant for gains and states, tests/test_oracle_invariance.py): compare the
    # projector Bλ (G Bλ)^-1 G, which is what the recursion uses (lqr.jl:151)
    proj = lambda Bl, G: Bl @ np.linalg.solve(G @ Bl, G)
    assert rel_err(proj(lin["Bl"], lin["G"]), proj(g["Bl"].T, g["G"].T)) < 1e-8
    N = int(math.ceil((10.0 if horizon == math.inf else horizon) / mech.dt))
    res = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, N)
    K = res["Ku"][:1] if horizon == math.inf else res["Ku"]
    assert K.shape == g["K"].shape
    assert gain_err(K, g["K"]) < TOL
    kind = "lqr_inf" if horizon == math.inf else "lqr_finite"
    co = om.make_ctrl(1, kind=kind, eqc=eqc, N=0 if horizon == math.inf else N, K=g["K"], zd=zd, Fd=np.zeros(len(eqc)))
    r = om.simulate(st0, steps, ctrl=co, record=True)
    assert rel_err(r["storage"][:, 0], g["storage"][:steps]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_reference_lqr(cc, case):
    g = load_case(case)
    if not g:
        pytest.skip(f"{case} not in the manifest")
    mech, cfg, horizon, steps = _lqr_cases(cc)[case]()
    st0 = mech.state_vector()[None].copy()
    lqr = cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], horizon, xd=cfg["xd"], qd=cfg.get("qd"))
    assert gain_err(lqr.K[0], g["K"]) < TOL
    sto, fin, status = cc.simulate(mech, steps * mech.dt, lqr, record=True, states=st0, return_status=True)
    assert np.all(status > 0)
    assert rel_err(sto.data[:, 0], g["storage"][:steps]) < TOL


def test_oracle_matches_reference_pid(cc, orc):
    g = load_case("c5_pid_doublependulum")
    if not g:
        pytest.skip("case not in the manifest")
    mech, cfg = cc.examples.doublependulum()
    _graph_check(cc, mech, g)
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    co = om.make_ctrl(1, kind="pid", eqc=eqc, P=cfg["P"], I=cfg["I"], D=cfg["D"], goal=np.array([cfg["goals"]]))
    r = om.simulate(mech.state_vector()[None], 1000, ctrl=co, record=True)
    assert rel_err(r["storage"][:, 0], g["storage"]) < TOL


@pytest.mark.gpu
def test_gpu_matches_reference_pid(cc):
    g = load_case("c5_pid_doublependulum")
    if not g:
        pytest.skip("case not in the manifest")
    mech, cfg = cc.examples.doublependulum()
    pid = cc.PID(mech, cfg["eqcids"], [cfg["goals"]], P=cfg["P"], I=cfg["I"], D=cfg["D"])
    sto, fin, status = cc.simulate(mech, 10.0, pid, record=True, states=mech.state_vector()[None], return_status=True)
    assert rel_err(sto.data[:, 0], g["storage"]) < TOL


def _tracking_inputs(cc):
    mech, cfg = cc.examples.triple_cartpole()
    U = np.load(os.path.join(HERE, "golden", "triple_cartpole_U.npy")).reshape(1000, 1)
    return mech, cfg, U


def test_oracle_matches_reference_tracking(cc, orc):
    g = load_case("tracking_triple_cartpole")
    if not g:
        pytest.skip("case not in the manifest")
    mech, cfg, U = _tracking_inputs(cc)
    _graph_check(cc, mech, g)
    om = orc.OracleMech(mech.desc())
    r0 = om.simulate(mech.state_vector()[None], 1000, ctrl=om.make_ctrl(1, kind="table", eqc=[0], U=U), record=True)
    assert rel_err(r0["storage"][:, 0], g["storage0"]) < TOL
    ref = om.tracking_dlqr(g["storage0"], U, [0], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt)
    assert gain_err(ref["Ku"], g["K"]) < TOL
    co = om.make_ctrl(1, kind="tracking", eqc=[0], N=1000, K=g["K"], zd=g["storage0"], Fd=U)
    r = om.simulate(g["state0_tracked"][None], 1000, ctrl=co, record=True)
    assert rel_err(r["storage"][:, 0], g["storage_tracked"]) < TOL


@pytest.mark.gpu
def test_gpu_matches_reference_tracking(cc):
    g = load_case("tracking_triple_cartpole")
    if not g:
        pytest.skip("case not in the manifest")
    mech, cfg, U = _tracking_inputs(cc)
    storage0 = cc.Storage(1000, mech.nb)
    cc.simulate(mech, storage0, cc.OpenLoop(mech, cfg["eqcids"], U), record=True)
    assert rel_err(storage0.data[:, 0], g["storage0"]) < TOL
    tl = cc.TrackingLQR(mech, storage0, U.reshape(1, 1000, 1), cfg["eqcids"], cfg["Q"], cfg["R"])
    assert gain_err(tl.K[0], g["K"]) < TOL
    out, fin, status = cc.simulate(mech, 10.0, tl, record=True, states=g["state0_tracked"][None], return_status=True)
    assert rel_err(out.data[:, 0], g["storage_tracked"]) < TOL
