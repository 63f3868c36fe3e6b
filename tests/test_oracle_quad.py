"""How much of a gain difference is the recursion's own rounding?  (replaces the unsupported "meaningless in Float64" claim.)

tests/golden/npend_quad_n<links>_N<N>.npz (tools/quad_gains.py) hold rows of the gains of the upright n-pendulum LQR
(examples/lqr_cartpole_n_pendulum.jl) computed by the oracle's literal dlqr with the recursion carried in IEEE binary128 on the SAME Float64
linearisation, and the error of the Float64 oracle against them.  Facts pinned here (CPU):
  * n = 8, N = 300: the literal Float64 recursion is already 1.3e-9 (global) / 1.4e-7 (element-relative) away from the binary128 result;
  * n = 8, N = 1000 (the example's horizon): 1.4e-6 / 6e-6 -- no Float64 implementation, the Julia reference included, can be closer to the
    exact gains of its own problem data than this, so two Float64 implementations agreeing to 1e-9 there would be a coincidence of rounding;
  * n = 16 and n = 32: max |P| reaches 4e21 resp. > 1e50: beyond Float64's 16 digits -- the Float64 gains share no digit with the binary128 ones
    (and for n = 32 binary128's 34 digits do not resolve P either).
tests/test_gpu_quad.py holds the CUDA path to "not worse than 2x the literal Float64 recursion" on the same rows.
"""
import os

import numpy as np
import pytest

from helpers import gain_err, zd_of

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(links, N):
    p = os.path.join(GOLD, f"npend_quad_n{links}_N{N}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{os.path.basename(p)} not generated (tools/quad_gains.py {links} {N})")
    return np.load(p)


def oracle_f64_rows(cc, orc, links, N, rows):
    mech, cfg = cc.examples.cartpole_n_pendulum(links)
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    lin = om.linearize(zd_of(mech, cfg), np.zeros(len(eqc)), eqc)
    res = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], cc.blockdiag(cfg["Q"]) * mech.dt, cc.blockdiag(cfg["R"]) * mech.dt, N)
    return res["Ku"][rows], lin


@pytest.mark.parametrize("N,lo,hi", [(300, 1e-8, 1e-6), (1000, 1e-6, 1e-4)])
def test_float64_recursion_floor_8_pendulum(cc, orc, N, lo, hi):
    g = load(8, N)
    rows = g["rows"]
    K64, lin = oracle_f64_rows(cc, orc, 8, N, rows)
    assert np.array_equal(K64, g["Ku_f64_rows"])                 # the fixture's Float64 rows are the oracle's, bit for bit
    e = gain_err(K64, g["Ku_ref_rows"])
    assert lo < e < hi, e                                        # rounding amplified by the recursion: far above 1e-9
    assert float(g["Pmax"]) > 1e11


def test_multiprecision_dlqr_agrees_where_the_problem_is_well_conditioned(cc, orc):
    """cartpole (P ~ 1e3): Float64, x87 and binary128 recursions agree to 1e-12 -- the wide instantiations are the same algorithm"""
    mech, cfg = cc.examples.cartpole()
    om = orc.OracleMech(mech.desc())
    eqc = [mech.eqc_index(i) for i in cfg["eqcids"]]
    lin = om.linearize(zd_of(mech, cfg), np.zeros(len(eqc)), eqc)
    Q = cc.blockdiag(cfg["Q"]) * mech.dt; R = cc.blockdiag(cfg["R"]) * mech.dt
    ref = orc.dlqr(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Q, R, 400)
    r0 = orc.dlqr_mp(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Q, R, 400, prec=0)
    assert np.array_equal(r0["Ku"], ref["Ku"]) and r0["kexit"] == ref["kexit"] and r0["iters"] == ref["iters"]   # prec 0 IS orc_dlqr
    for prec in (1, 2):
        r = orc.dlqr_mp(lin["A"], lin["Bu"], lin["Bl"], lin["G"], Q, R, 400, prec=prec)
        assert gain_err(r["Ku"], ref["Ku"]) < 1e-11
        assert r["iters"] == ref["iters"]
    assert len(r0["trace"]) == r0["iters"] and np.all(np.diff(r0["trace"][200:]) < 0)    # ||Pk - Pk+1|| decays monotonically once the transient is over


@pytest.mark.parametrize("links", [16, 32])
def test_float64_cannot_resolve_long_chains(links):
    g = load(links, 1000)
    assert float(g["Pmax"]) > 1e20                               # 16 digits cannot carry a cost-to-go spanning > 20 decades
    assert float(g["f64_vs_ref_glob"]) > 1e-3                    # the literal Float64 gains share (almost) no digits with the wide ones
