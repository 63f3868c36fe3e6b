"""Static checks of the compiled rollout kernels (cuobjdump on the in-tree objects; no GPU needed).

Round 2 lost 11 % of the headline to a code-generation accident that no runtime test could see: a lane-0 spin wait in K1r's persistent work
loop made ptxas stop treating the warp as converged, and every later `shfl.sync` was compiled with a `BRA.DIV`-guarded slow path (twice the
shuffles, six times the register moves).  The results were bit-identical, only slower.  These tests pin the fix: the
rollout kernels that exchange data by warp shuffles contain no divergence-guarded shuffle, and the persistent instantiation of K1r is the same
code as the plain one."""
import os
import re
import shutil
import subprocess

import pytest

CSRC = os.path.join(os.path.dirname(__file__), "..", "constrainedcontrol.jl_b200", "csrc")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _functions(obj):
    path = os.path.join(CSRC, obj)
    if not os.path.exists(path) or not os.path.exists(CUOBJDUMP):
        pytest.skip("object file or cuobjdump missing (run __graft_entry__.build())")
    out = subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = dict(shfl=0, bradiv=0, instr=0, ldl=0)
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            f = funcs[name]
            f["instr"] += 1
            f["shfl"] += " SHFL" in line
            f["bradiv"] += "BRA.DIV" in line
            f["ldl"] += " LDL" in line
    return funcs


def test_k1r_has_no_divergence_guarded_shuffles_and_persistent_equals_plain_code():
    funcs = {k: v for k, v in _functions("ccb_rollout_rows.o").items() if "rollout_rows_kernel" in k}
    assert len(funcs) == 18                                   # six chain lengths x (LQR-only, every law, persistent)
    for name, f in funcs.items():
        assert f["bradiv"] == 0, name
        assert f["shfl"] > 200
    for nj in range(1, 7):
        plain = next(v for k, v in funcs.items() if f"ILi{nj}ELb1ELb0E" in k)
        pers = next(v for k, v in funcs.items() if f"ILi{nj}ELb1ELb1E" in k)
        assert pers["shfl"] == plain["shfl"], nj             # a doubled count is the signature of the BRA.DIV slow paths
        assert pers["instr"] < 1.15 * plain["instr"], nj     # work loop and flag handling only
        assert pers["ldl"] <= 40 and pers["ldl"] < plain["ldl"] / 2, nj     # 255 registers: the few spills left sit outside the Newton loop


@pytest.mark.parametrize("obj", ["ccb_rollout.o", "ccb_rollout_chain.o", "ccb_rollout_mw.o", "ccb_linearize.o", "ccb_riccati_project.o"])
def test_other_hot_kernels_have_no_divergence_guards(obj):
    for name, f in _functions(obj).items():
        assert f["bradiv"] == 0, name
