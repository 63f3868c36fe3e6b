"""2+ GPU check of the one collective of the path: all-gather of gain tables synthesised on different ranks.
torchrun --nproc-per-node N tools_nccl_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
import numpy as np, torch, torch.distributed as dist
import ccb200 as cc
from helpers import zd_of
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mech, cfg = cc.examples.cartpole()
S, N = 64, 400
rng = np.random.default_rng(2)
qs = np.exp(rng.uniform(np.log(0.1), np.log(10), S)); rs = np.exp(rng.uniform(np.log(0.1), np.log(10), S))
zd = zd_of(mech, cfg); Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
def synth(lo, hi):
    return cc.lqr_synthesize(mech, np.repeat(zd[None], hi - lo, 0), np.zeros((hi - lo, 1)), [0], Qd[None] * qs[lo:hi, None, None], Rd[None] * rs[lo:hi, None, None], N, device=lr)["Ku"]
t0 = time.time()
K = cc.dist.sharded_synthesis(synth, S, device=f"cuda:{lr}")
dt = time.time() - t0
full = synth(0, S) if rank == 0 else None
if rank == 0:
    print(f"world {world}: all-gathered gain tables {K.shape}, equal to the single-rank synthesis: {np.array_equal(K, full)}, {dt:.3f} s")
dist.barrier(); dist.destroy_process_group()
