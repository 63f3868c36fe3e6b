"""ctypes binding of the C-ABI shared library csrc/libccb.so (include/ccb.h).

There is NO fallback: if the library is missing (not built) `lib()` raises, and on a machine without a CUDA
device every compute entry point returns CCB_E_CUDA which is raised as `CcbError`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("CCB_LIB", os.path.join(CSRC, "libccb.so"))   # CCB_LIB: kernel-variant experiments only
_LIB = None


class CcbError(RuntimeError):
    pass


class MechDesc(C.Structure):
    _fields_ = [
        ("nb", C.c_int32), ("ne", C.c_int32), ("nj", C.c_int32), ("id_scheme", C.c_int32),
        ("dt", C.c_double), ("g", C.c_double), ("reg", C.c_double),
        ("mass", C.c_void_p), ("inertia", C.c_void_p), ("eqc_parent", C.c_void_p), ("eqc_joint_ptr", C.c_void_p),
        ("joint_child", C.c_void_p), ("joint_nT", C.c_void_p), ("joint_nR", C.c_void_p), ("joint_axis", C.c_void_p),
        ("joint_p1", C.c_void_p), ("joint_p2", C.c_void_p), ("joint_qoff", C.c_void_p),
    ]


class CtrlDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("nu", C.c_int32), ("N", C.c_int32), ("B", C.c_int32),
        ("eqc", C.c_void_p), ("K", C.c_void_p), ("K_bstride", C.c_int64), ("zd", C.c_void_p), ("zd_bstride", C.c_int64),
        ("Fd", C.c_void_p), ("Fd_bstride", C.c_int64), ("U", C.c_void_p), ("U_bstride", C.c_int64),
        ("P", C.c_void_p), ("I", C.c_void_p), ("D", C.c_void_p), ("goal", C.c_void_p), ("goal_bstride", C.c_int64),
    ]


class SimOpts(C.Structure):
    _fields_ = [("eps", C.c_double), ("newton_iter", C.c_int32), ("line_iter", C.c_int32), ("lanes", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/ccb.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ccb_last_error", "ccb_version", "ccb_device_count", "ccb_mech_create", "ccb_mech_destroy", "ccb_mech_dims", "ccb_mech_graph",
    "ccb_batch_create", "ccb_batch_destroy", "ccb_batch_set_state", "ccb_batch_get_state", "ccb_batch_device_ptrs", "ccb_ctrl_create",
    "ccb_ctrl_destroy", "ccb_ctrl_pid_state", "ccb_simulate", "ccb_simulate_device", "ccb_batch_sync", "ccb_batch_stream", "ccb_launch_count",
    "ccb_linearize", "ccb_dlqr", "ccb_lqr_synthesize", "ccb_tracking_synthesize", "ccb_synth_phase_ms", "ccb_lqr_create", "ccb_ctrl_get_gains", "ccb_tracking_create", "ccb_ctrl_set_replication",
    "ccb_are_solve", "ccb_scratch_release", "ccb_gains_alloc", "ccb_gains_open", "ccb_gains_close", "ccb_gains_free", "ccb_gains_read", "ccb_gains_defer_backfill", "ccb_gains_backfill",
    "ccb_lqr_synthesize_peers", "ccb_tracking_synthesize_peers", "ccb_lqr_ctrl_from_table", "ccb_tracking_ctrl_from_table",
]


def build(verbose: bool = False) -> str:
    """compile csrc/*.cu for sm_100a into csrc/libccb.so (nvcc cross-compiles without a GPU)"""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-j", str(min(8, os.cpu_count() or 1)), "-C", CSRC, "libccb.so"], stdout=out)
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise CcbError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the B200 hot path)")
        L = C.CDLL(SO_PATH)
        L.ccb_last_error.restype = C.c_char_p
        L.ccb_launch_count.restype = C.c_int64
        _LIB = L
    return _LIB


def check(rc: int):
    if rc != 0:
        msg = lib().ccb_last_error()
        raise CcbError(f"ccb error {rc}: {msg.decode() if msg else ''}")


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count() -> int:
    return int(lib().ccb_device_count())
