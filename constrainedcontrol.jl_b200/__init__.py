"""constrainedcontrol.jl_b200 -- B200-native batched closed-loop rollouts and LQR synthesis behind the
ConstrainedControl.jl controller API (LQR / TrackingLQR / PID / simulate!).

The directory name contains a dot, so import it through the root-level loader: ``import ccb200``.
Host-side description modules import without the native library; every compute entry point goes through
the C-ABI shared library (csrc/libccb.so) and fails loudly when it is missing (no CPU fallback).
"""
from .mechanism import *  # noqa: F401,F403
from . import mechanism, examples, _lib, control, synth, urdf, dist, util  # noqa: F401
from ._lib import CcbError, device_count  # noqa: F401
from .control import LQR, TrackingLQR, PID, OpenLoop, Controller, Storage, Batch, simulate, handle, blockdiag  # noqa: F401
from .synth import linearsystem, dlqr, lqr_synthesize, tracking_synthesize  # noqa: F401
