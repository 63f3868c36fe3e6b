"""Regenerates tests/golden/triple_cartpole_U.npy from the reference's only numeric fixture:
the open-loop input `U` on line 1 of examples/trackingLQR_triple_cartpole.jl (1000 doubles).
Run in the build container (needs /root/reference); the .npy travels to the GPU box."""
import re
import numpy as np

line = open("/root/reference/examples/trackingLQR_triple_cartpole.jl").readline()
vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line.split("=", 1)[1])]
assert len(vals) == 1000
np.save(__file__.replace("make_golden.py", "triple_cartpole_U.npy"), np.array(vals))
