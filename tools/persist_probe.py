import sys, os, time
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path[:0]=[ROOT, ROOT+"/tests"]
import numpy as np, torch, ccb200 as cc
from helpers import cartpole_states
B=int(sys.argv[1]); steps=2000
mech,cfg=cc.examples.cartpole()
lqr=cc.LQR(mech, cc.getid(mech.bodies), cfg["eqcids"], cfg["Q"], cfg["R"], 20.0, xd=cfg["xd"])
st=cartpole_states(cc,mech,B,seed=1)
batch=cc.Batch(mech,B); stream=torch.cuda.ExternalStream(batch.stream())
for rep in range(3):
    batch.set_state(st); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(stream); batch.simulate_device(lqr,1,steps); e1.record(stream); e1.synchronize()
print("B",B,"chunks",os.environ.get("CCB_K1R_CHUNKS"),"force",os.environ.get("CCB_K1R_FORCE_PERSIST"),"nopersist",os.environ.get("CCB_K1R_NO_PERSIST"),"ms",e0.elapsed_time(e1))
