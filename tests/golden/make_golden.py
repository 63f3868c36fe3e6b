"""Regenerates tests/golden/triple_cartpole_U.npy from the reference's only numeric fixture:
the open-loop input `U` on line 1 of examples/trackingLQR_triple_cartpole.jl (1000 doubles).
Run in the build container (needs /root/reference); the .npy travels to the GPU box."""
import re
import numpy as np

line = open("/root/reference/examples/trackingLQR_triple_cartpole.jl").readline()
vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", line.split("=", 1)[1])]
assert len(vals) == 1000
np.save(__file__.replace("make_golden.py", "triple_cartpole_U.npy"), np.array(vals))
# the same numbers one per line for tools/dump_golden.jl (the Julia side reads no .npy)
open(__file__.replace("make_golden.py", "triple_cartpole_U.txt"), "w").write("\n".join(repr(float(v)) for v in vals) + "\n")

# parsed parameters of the Sawyer arm (reference examples/examples_files/sawyer_arm.urdf): masses, COMs, inertia tensors,
# joint origins/axes -- numbers only, so that the GPU box (which has no /root/reference) can build the mechanism.
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import ccb200 as cc
desc = cc.urdf.parse_urdf("/root/reference/examples/examples_files/sawyer_arm.urdf")
json.dump(desc, open(__file__.replace("make_golden.py", "sawyer_arm.json"), "w"), indent=1)

# the branching Sawyer description with fixed joints (reference examples/examples_files/sawyer_simple.urdf): SURVEY 8(f) n3
desc = cc.urdf.parse_urdf("/root/reference/examples/examples_files/sawyer_simple.urdf")
json.dump(desc, open(__file__.replace("make_golden.py", "sawyer_simple.json"), "w"), indent=1)

# the full Sawyer description (reference examples/examples_files/sawyer.urdf: 24 links, 23 joints, massless camera / wrist frames): SURVEY 8(f) n3
desc = cc.urdf.parse_urdf("/root/reference/examples/examples_files/sawyer.urdf")
json.dump(desc, open(__file__.replace("make_golden.py", "sawyer_full.json"), "w"), indent=1)
