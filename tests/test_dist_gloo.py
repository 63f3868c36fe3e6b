"""world_size-2 gloo tests (CPU) of the multi-rank plumbing: batch sharding and the all-gather of gain tables.
The data path itself has no collective."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ccb200 as cc
    lo, hi = cc.dist.shard_range(total, rank, world)

    def synth(lo, hi):   # stand-in for a per-shard gain synthesis: table s is filled with the value s
        return np.stack([np.full((3, 2, 24), float(s)) for s in range(lo, hi)]) if hi > lo else np.zeros((0, 3, 2, 24))

    K = cc.dist.sharded_synthesis(synth, total)
    ok = K.shape == (total, 3, 2, 24) and all(np.all(K[s] == s) for s in range(total))
    q.put((rank, lo, hi, bool(ok)))
    dist.destroy_process_group()


def _worker_counters(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ccb200 as cc

    class Tab:      # the counters of a peer gain table: every rank has filled its own slice only
        pass
    t = Tab()
    lo, hi = cc.dist.shard_range(total, rank, world)
    t.kexit = np.zeros(total, np.int32); t.iters = np.zeros(total, np.int32)
    t.kexit[lo:hi] = np.arange(lo, hi) + 1; t.iters[lo:hi] = 100 + np.arange(lo, hi)
    cc.dist._share_counters(t, dist, 0)
    ok = np.array_equal(t.kexit, np.arange(total) + 1) and np.array_equal(t.iters, 100 + np.arange(total)) and t.kexit.dtype == np.int32
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


class _StubLib:
    """stands in for libccb.so in the protocol test: table handles are fake, nothing is computed"""

    def __init__(self, log):
        self.log = log

    def ccb_gains_alloc(self, device, count, pptr, hbuf):
        pptr._obj.value = 0x1000
        self.log.append("alloc")
        return 0

    def ccb_gains_open(self, device, hb, pptr):
        pptr._obj.value = 0x2000
        self.log.append("open")
        return 0

    def ccb_gains_close(self, p):
        self.log.append("close")
        return 0

    def ccb_gains_free(self, p):
        self.log.append("free")
        return 0


def _worker_peer_failure(rank, world, port, fail_rank, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ccb200 as cc
    log = []
    cc._lib._LIB = _StubLib(log)

    class Tab:
        pass
    t = Tab()
    t.kexit = np.zeros(4, np.int32); t.iters = np.zeros(4, np.int32)

    def run_shard(tables):
        log.append(f"shard:{len(tables)}")
        if rank == fail_rank:
            raise RuntimeError("synthesis failed on this rank")

    try:
        cc.dist._peer_fill(t, 0, 128, run_shard)
        cc.dist._peer_fill(t, 0, 128, lambda tables: log.append(f"refill:{len(tables)}"))      # re-fill: no second allocation / mapping
        outcome = "ok"
        cc.dist._peer_unmap(t)
    except cc.dist.PeerProtocolError as e:
        outcome = "local" if "synthesis failed" in str(e) else "peer"
    q.put((rank, outcome, list(log), t.ptr is None))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_peer_protocol_failure_reaches_every_rank_world2(fail_rank):
    """a rank that fails inside the peer-table protocol must not leave the others at the barrier: every rank raises,
    the mappings are closed and the local table is freed only after the barrier"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_peer_failure, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    if fail_rank < 0:
        assert [r[1] for r in res] == ["ok", "ok"]
        assert all(r[2] == ["alloc", "open", "shard:2", "refill:2", "close"] and not r[3] for r in res)
    else:
        assert [r[1] for r in res] == ["peer", "local"]
        assert all(r[2] == ["alloc", "open", "shard:2", "close", "free"] and r[3] for r in res)


def test_peer_table_counters_are_shared_world2():
    """host half of the fused multi-GPU gain exchange (dist.PeerGainTable): exit indices and iteration counts of every shard reach every rank"""
    world, total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_counters, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [True, True]


@pytest.mark.parametrize("total", [7, 8])
def test_shard_and_allgather_gains_world2(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[3] for r in res] == [True, True]
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total     # contiguous cover of the batch


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import ccb200 as cc
    for total in (1, 5, 4096, 1023):
        for world in (1, 2, 4, 8):
            cuts = [cc.dist.shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1


def test_peer_tables_share_refill_but_not_fill():
    """PeerTrackingTable borrows refill / read / free from PeerGainTable; each class brings its own _fill (the LQR table completes the rows below
    the break index locally after the exchange, the tracking recursion has no break)"""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import ccb200 as cc
    g, t = cc.dist.PeerGainTable, cc.dist.PeerTrackingTable
    assert t.refill is g.refill and t.read is g.read and t.free is g.free
    assert callable(g._fill) and callable(t._fill) and t._fill is not g._fill
