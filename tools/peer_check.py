"""2+ GPUs (torchrun): LQR syntheses sharded over the ranks with the gain rows stored into every rank's table by the Riccati sweep itself
(dist.PeerGainTable, CUDA IPC + NVLink stores) -- checked bit-for-bit against a full local synthesis and timed against synthesise-then-all-gather.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/peer_check.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch, torch.distributed as dist
import ccb200 as cc
from helpers import zd_of

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
name = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 512
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json") if name == "sawyer" else cc.examples.cartpole()
eqc_idx = [mech.eqc_index(i) for i in cfg["eqcids"]]
nu = len(eqc_idx)
zd = np.repeat(zd_of(mech, cfg)[None], S, 0)
rng = np.random.default_rng(2)
qs = np.exp(rng.uniform(np.log(0.1), np.log(10), S)); rs = np.exp(rng.uniform(np.log(0.1), np.log(10), S))
Qd = cc.blockdiag(cfg["Q"]) * mech.dt; Rd = cc.blockdiag(cfg["R"]) * mech.dt
Qb = Qd[None] * qs[:, None, None]; Rb = Rd[None] * rs[:, None, None]
Fd = np.zeros((S, nu))
lo, hi = cc.dist.shard_range(S, rank, world)

def fused():
    return cc.dist.PeerGainTable(mech, zd, Fd, eqc_idx, Qd, Rd, N, device=local, Q_batch=Qb, R_batch=Rb)

def gathered():     # the baseline: synthesise the shard, all-gather the tables through NCCL (device tensors), keep them on the device
    o = cc.lqr_synthesize(mech, zd[lo:hi], Fd[lo:hi], eqc_idx, Qb[lo:hi], Rb[lo:hi], N, device=local)
    K = torch.from_numpy(o["Ku"]).to(f"cuda:{local}")
    if world == 1:
        return K
    out = torch.empty((S,) + K.shape[1:], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_gather_into_tensor(out, K)       # equal shards in this check
    torch.cuda.synchronize()
    return out

tab = fused(); tab.free()                      # warm-up (module load, IPC, arena)
if world > 1: dist.barrier()
t0 = time.time(); tab = fused(); t_f = time.time() - t0
if world > 1: dist.barrier()
_ = gathered()
if world > 1: dist.barrier()
t0 = time.time(); Kg = gathered(); t_g = time.time() - t0
full = tab.read()
ok_gather = np.array_equal(full, Kg.cpu().numpy())
# against a purely local synthesis of a slice owned by ANOTHER rank (bit-identical: same kernel, same inputs)
other = (rank + 1) % world
olo, ohi = cc.dist.shard_range(S, other, world)
olo2 = olo; ohi2 = min(ohi, olo + 8)
loc = cc.lqr_synthesize(mech, zd[olo2:ohi2], Fd[olo2:ohi2], eqc_idx, Qb[olo2:ohi2], Rb[olo2:ohi2], N, device=local)
ok_local = np.array_equal(full[olo2:ohi2], loc["Ku"])
print(f"rank {rank}/{world} {name} S={S} N={N}: fused-into-sweep {t_f*1e3:.1f} ms, synthesise + NCCL all-gather {t_g*1e3:.1f} ms; "
      f"table == gathered: {ok_gather}; peer-written slice == local synthesis: {ok_local}; split {({k: round(v * 1e3, 1) for k, v in tab.timing.items()})} ms", flush=True)
assert ok_gather and ok_local
if world > 1:
    dist.barrier(); dist.destroy_process_group()
