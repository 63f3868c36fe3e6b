"""Loader for the package directory `constrainedcontrol.jl_b200/` (its name is not a valid Python identifier).

    import ccb200 as cc
    mech, cfg = cc.examples.cartpole()
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(_HERE, "constrainedcontrol.jl_b200")
_NAME = "constrainedcontrol_jl_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
pkg = sys.modules[_NAME]
globals().update({k: v for k, v in vars(pkg).items() if not k.startswith("__")})
