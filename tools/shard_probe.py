"""N-GPU probe of the sharded Sawyer synthesis: per-run times of the fused peer-store route with and without the spread of flush phases /
peer order (CCB_PEER_NO_SPREAD), and of synthesise + ncclAllGather.   torchrun --nproc-per-node N tools/shard_probe.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch, torch.distributed as dist
import ccb200 as cc
import bench
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
mech, cfg = cc.examples.sawyer(os.path.join(ROOT, "tests", "golden", "sawyer_arm.json"))
S, N = 1024, 2000
eqc_idx = np.ascontiguousarray([mech.eqc_index(i) for i in cfg["eqcids"]], dtype=np.int32)
zd = bench.sawyer_setpoints(cc, mech, S); Fd = np.zeros((S, len(eqc_idx)))
Qd = np.ascontiguousarray(cc.blockdiag(cfg["Q"]) * mech.dt); Rd = np.ascontiguousarray(cc.blockdiag(cfg["R"]) * mech.dt)

def timed(fn):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.time()
    fn()
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([time.time() - t0], dtype=torch.float64, device=f"cuda:{dev}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e3

tab = cc.dist.PeerGainTable(mech, zd, Fd, eqc_idx, Qd, Rd, N, device=dev)
res = {"spread": [], "nospread": []}
for rep in range(reps):
    for mode in ("spread", "nospread"):
        if mode == "nospread": os.environ["CCB_PEER_NO_SPREAD"] = "1"
        else: os.environ.pop("CCB_PEER_NO_SPREAD", None)
        res[mode].append(round(timed(tab.refill), 2))
        syn = tab.timing.get("synthesis_s", 0) * 1e3
        if rank == 0: print(mode, rep, res[mode][-1], "ms; rank0 synthesis", round(syn, 2), flush=True)
os.environ.pop("CCB_PEER_NO_SPREAD", None)
ref = tab.read(0, 2).copy()
if rank == 0: print("fused spread", res["spread"], "\nfused nospread", res["nospread"], flush=True)
tab.free()
