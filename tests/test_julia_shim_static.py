"""Static consistency of the two host bindings with include/ccb.h.

There is no julia toolchain in the build image, so julia/ConstrainedControlB200.jl cannot be executed here.  What CAN be
checked without running it is everything a `ccall` gets wrong silently: the symbol, the arity, the class of every argument
(pointer / 32-bit / 64-bit / double -- a Julia `Int32` where the ABI takes `int64_t` corrupts the call frame without an
error), the return type, and the memory layout of the two description structs the shim mirrors.  The same layout check is
applied to the ctypes structures of _lib.py (the binding every GPU test goes through)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.join(os.path.dirname(__file__), "..")
HDR = os.path.join(ROOT, "include", "ccb.h")
SHIM = os.path.join(ROOT, "constrainedcontrol.jl_b200", "julia", "ConstrainedControlB200.jl")


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def _c_class(decl):
    """class of a C parameter / field declaration"""
    d = decl.strip()
    if "*" in d or "[" in d:
        return "ptr"
    t = d.split()
    t = [w for w in t if w not in ("const", "unsigned")]
    base = t[0]
    return {"int32_t": "i32", "int": "i32", "int64_t": "i64", "double": "f64"}[base]


def _jl_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Float64": "f64", "Cdouble": "f64"}[t]


def c_prototypes():
    src = _strip_c_comments(open(HDR).read())
    protos = {}
    for ret, name, args in re.findall(r"\b(const char\*|int64_t|int)\s+(ccb_[a-z_0-9]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        args = " ".join(args.split())
        params = [] if args in ("", "void") else [_c_class(a) for a in args.split(",")]
        protos[name] = ({"const char*": "ptr", "int64_t": "i64", "int": "i32"}[ret], params)
    return protos


def c_struct(name):
    src = _strip_c_comments(open(HDR).read())
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), src, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        fname = re.search(r"([A-Za-z_0-9]+)$", decl).group(1)
        fields.append((fname, _c_class(decl)))
    return fields


def _split_top(s):
    """split at top-level commas (braces and parentheses nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def _call_args(src, start):
    """arguments of the call whose opening parenthesis is at src[start]"""
    depth, i = 0, start
    while True:
        if src[i] in "({[":
            depth += 1
        elif src[i] in ")}]":
            depth -= 1
            if depth == 0:
                return _split_top(src[start + 1:i])
        i += 1


def jl_ccalls():
    src = re.sub(r"#[^\n]*", "", open(SHIM).read())     # no '#' inside the shim's string literals that matter here
    calls = []
    for m in re.finditer(r"ccall\(", src):
        parts = _call_args(src, m.end() - 1)
        sym = re.match(r"\(:(ccb_[a-z_0-9]+),\s*libccb\)", parts[0]).group(1)
        ret = parts[1]
        types = _split_top(parts[2].strip()[1:-1])
        calls.append((sym, ret, [t for t in types if t], parts[3:]))
    return calls


def jl_struct(name):
    src = re.sub(r"#[^\n]*", "", open(SHIM).read())
    body = re.search(r"\nstruct %s\n(.*?)\nend" % name, src, flags=re.S).group(1)
    fields = []
    for decl in re.split(r"[;\n]", body):
        decl = decl.strip()
        if decl:
            fname, ftype = decl.split("::")
            fields.append((fname.strip(), _jl_class(ftype)))
    return fields


def test_header_parses_every_declared_function(cc):
    protos = c_prototypes()
    assert set(protos) == set(cc._lib.SYMBOLS)


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = jl_ccalls()
    assert len(calls) >= 25
    for sym, ret, types, values in calls:
        assert sym in protos, f"{sym} is not declared in include/ccb.h"
        cret, cparams = protos[sym]
        assert len(types) == len(cparams), f"{sym}: {len(types)} ccall argument types, the header declares {len(cparams)}"
        assert len(values) == len(types), f"{sym}: {len(values)} values for {len(types)} argument types"
        for i, (jt, ct) in enumerate(zip(types, cparams)):
            assert _jl_class(jt) == ct, f"{sym}: argument {i} is {jt} in the shim, {ct} in the header"
        assert (ret == "Cstring") == (sym == "ccb_last_error")
        if sym != "ccb_last_error":
            assert _jl_class(ret) == cret, f"{sym}: return type {ret}"


def test_shim_covers_the_reference_facing_entry_points():
    used = {c[0] for c in jl_ccalls()}
    # constructors, simulate!, the util.jl functions and the multi-GPU exchange: what a user of the reference touches
    for sym in ["ccb_mech_create", "ccb_batch_create", "ccb_batch_set_state", "ccb_batch_get_state", "ccb_ctrl_create", "ccb_simulate",
                "ccb_lqr_synthesize", "ccb_lqr_create", "ccb_tracking_create", "ccb_ctrl_set_replication", "ccb_are_solve",
                "ccb_gains_alloc", "ccb_gains_open", "ccb_lqr_synthesize_peers", "ccb_gains_defer_backfill", "ccb_gains_backfill",
                "ccb_mech_destroy", "ccb_batch_destroy", "ccb_ctrl_destroy", "ccb_gains_close", "ccb_gains_free"]:
        assert sym in used, sym


@pytest.mark.parametrize("cname,jname", [("ccb_mech_desc", "MechDesc"), ("ccb_ctrl_desc", "CtrlDesc")])
def test_description_structs_have_the_headers_layout(cc, cname, jname):
    want = c_struct(cname)
    assert jl_struct(jname) == want                                    # names, order and classes: isbits Julia structs use the C layout
    ct = getattr(cc._lib, jname)
    cls = {C.c_int32: "i32", C.c_int64: "i64", C.c_double: "f64", C.c_void_p: "ptr"}
    assert [(n, cls[t]) for n, t in ct._fields_] == want
    # natural alignment, no packing surprises: offsets follow from the classes
    size = {"i32": 4, "i64": 8, "f64": 8, "ptr": 8}
    off = 0
    for (n, k) in want:
        off = (off + size[k] - 1) // size[k] * size[k]
        assert getattr(ct, n).offset == off, n
        off += size[k]
    assert C.sizeof(ct) == (off + 7) // 8 * 8


def test_sim_opts_layout(cc):
    want = c_struct("ccb_sim_opts")
    cls = {C.c_int32: "i32", C.c_double: "f64"}
    assert [(n, cls[t]) for n, t in cc._lib.SimOpts._fields_] == want


def test_struct_constructor_calls_pass_one_value_per_field():
    src = re.sub(r"#[^\n]*", "", open(SHIM).read())
    for jname, cname in [("MechDesc", "ccb_mech_desc"), ("CtrlDesc", "ccb_ctrl_desc")]:
        nfields = len(c_struct(cname))
        sites = [m for m in re.finditer(r"Ref\(%s\(" % jname, src)]
        assert sites, jname
        for m in sites:
            assert len(_call_args(src, m.end() - 1)) == nfields, f"{jname}(...) at offset {m.start()}"


def test_controller_kind_literals_match_the_enum():
    hdr = _strip_c_comments(open(HDR).read())
    enum = dict((k, int(v)) for k, v in re.findall(r"(CCB_CTRL_[A-Z_]+) = (\d)", hdr))
    src = open(SHIM).read()
    assert enum["CCB_CTRL_PID"] == 5 and "Ref(CtrlDesc(5, nu" in src                                  # BatchPID
    assert enum["CCB_CTRL_LQR_INF"] == 2 and enum["CCB_CTRL_LQR_FIN"] == 3 and "lqr.N == Inf ? 2 : 3" in src
