"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/ccb.h declares,
validates descriptions like the reference's @asserts, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_exports_every_declared_symbol(cc):
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "ccb.h")).read()
    declared = set(re.findall(r"\b(ccb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(cc._lib.SYMBOLS)
    lib = cc._lib.lib()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.ccb_version() >= 100


def test_graph_orderings_match_oracle_bit_exact(cc, orc):
    for name in ["pendulum", "cartpole", "acrobot", "prismatic", "doublependulum", "triple_cartpole", "deltabot"]:
        mech, _ = getattr(cc.examples, name)()
        g = cc.handle(mech).graph()
        go = orc.OracleMech(mech.desc()).graph()
        for k in go:
            assert np.array_equal(g[k], go[k]), (name, k)
    mech, _ = cc.examples.cartpole_n_pendulum(32)
    g = cc.handle(mech).graph()
    assert g["joint_elim"].tolist() == list(range(32, -1, -1))     # leaf first along the chain
    assert cc.handle(mech).ml == 165 and cc.handle(mech).n == 396


def test_bad_descriptions_are_rejected(cc):
    mech, _ = cc.examples.cartpole()
    d = mech.desc()
    md = cc._lib.MechDesc()
    md.nb, md.ne, md.nj, md.dt = 0, 1, 1, 0.01
    h = C.c_void_p()
    assert cc._lib.lib().ccb_mech_create(C.byref(md), C.byref(h)) == -1
    assert b"bad sizes" in cc._lib.lib().ccb_last_error()
    assert d["nb"] == 2


def test_no_cpu_fallback(cc):
    if cc.device_count() > 0:
        pytest.skip("a GPU is present")
    mech, _ = cc.examples.cartpole()
    with pytest.raises(cc.CcbError, match="no CPU fallback"):
        cc.Batch(mech, 4)
    with pytest.raises(cc.CcbError):
        cc.simulate(mech, 0.1, None)


def test_controller_argument_checks(cc):
    # lqr.jl:59-60 length asserts; pid.jl:20,36 1-DoF assert -- raised before anything touches the device
    mech, cfg = cc.examples.cartpole()
    with pytest.raises(AssertionError, match="Missmatched length for bodies"):
        cc.LQR(mech, cc.getid(mech.bodies)[:1], cfg["eqcids"], cfg["Q"], cfg["R"], 10.0)
    with pytest.raises(AssertionError, match="Missmatched length for constraints"):
        cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"] * 2, 10.0)
    mech2, _ = cc.examples.deltabot()
    with pytest.raises(AssertionError, match="Only 1 DOF"):
        cc.PID(mech2, [mech2.eqconstraints[2].id], [0.0])
