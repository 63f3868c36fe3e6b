"""2+ GPUs (torchrun): TrackingLQR tables of several Sawyer reference trajectories sharded over the ranks, every rank's Riccati sweep pushing its gain
rows into all tables (dist.PeerTrackingTable) -- checked bit-for-bit against a full local TrackingLQR synthesis on every rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/peer_tracking_check.py [NT] [N]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch, torch.distributed as dist
import ccb200 as cc

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
NT = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 600
mech, cfg = cc.examples.sawyer(ROOT + "/tests/golden/sawyer_arm.json")
rng = np.random.default_rng(3)
amp = np.array([0.05, 0.05, 0.02, 0.02, 0.005, 0.005, 0.0005])
a = rng.uniform(0, 1, size=(NT, 7)) * amp; f = rng.integers(1, 4, size=(NT, 7)).astype(float)
k = np.arange(N)[None, :, None]
U = a[:, None, :] * np.sin(2 * np.pi * f[:, None, :] * k / 2000.0)
cc.urdf.zero_configuration(mech)
st0 = mech.batch_states(NT)
ol = cc.OpenLoop(mech, cfg["eqcids"], U)
sto, fin, status = cc.simulate(mech, N * mech.dt, ol, record=True, states=st0, return_status=True, device=local)
assert np.all(status > 0)
tab = cc.dist.PeerTrackingTable(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"], device=local); tab.free()       # warm-up
if world > 1: dist.barrier()
t0 = time.time()
tab = cc.dist.PeerTrackingTable(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"], device=local)
t_f = time.time() - t0
tab.refill()                    # synthesise again into the mapped tables (allocation + IPC mappings kept)
full = tab.read()
tl = cc.TrackingLQR(mech, sto, U, cfg["eqcids"], cfg["Q"], cfg["R"], device=local)          # everything locally
ok_k = np.array_equal(tab.kexit, tl.kexit)
# (1) the exchange itself: every rank must hold the same bits (each slice was written by one rank's kernel into all tables)
h = int(np.bitwise_xor.reduce(full.view(np.uint64).ravel()))
hs = [None] * world
if world > 1:
    dist.all_gather_object(hs, h)
else:
    hs = [h]
ok_x = all(x == hs[0] for x in hs)
# (2) against a local synthesis of ALL trajectories: the batches differ (4 vs 8 trajectories per launch pick different register-cap variants of
# the knot solve), so agreement is to rounding, not to the bit
rel = np.abs(full - tl.K).max() / np.abs(tl.K).max()
print(f"rank {rank}/{world}: {NT} Sawyer trajectories x {N} steps, shard {tab.shard}: exchange {t_f*1e3:.1f} ms, split { {k_: round(v*1e3,1) for k_, v in tab.timing.items()} } ms; "
      f"tables bit-identical across ranks: {ok_x}; vs local TrackingLQR synthesis of all trajectories: rel {rel:.1e}; exit indices equal: {ok_k}", flush=True)
assert ok_x and ok_k and rel < 1e-9
if world > 1:
    dist.barrier(); dist.destroy_process_group()
