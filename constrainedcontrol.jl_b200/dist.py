"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path shards by *systems*: every rank owns a contiguous slice of the batch and no collective runs on the step
path.  The only exchange is the all-gather of gain tables synthesised on different ranks when every rank needs all of
them (TrackingLQR / per-system LQR gains) -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int):
    """contiguous split of `total` systems: returns (lo, hi) of this rank (sizes differ by at most one)"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_gains(K_local: np.ndarray, total: int, device=None) -> np.ndarray:
    """K_local [S_local, ...] on every rank -> [total, ...] on every rank (order = rank order = batch order)"""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return K_local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    per = int(np.prod(K_local.shape[1:]))
    smax = max(hi - lo for lo, hi in sizes)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros((smax, per), dtype=torch.float64, device=dev)
    buf[: K_local.shape[0]] = torch.from_numpy(np.ascontiguousarray(K_local).reshape(K_local.shape[0], per)).to(dev)
    out = torch.empty((world, smax, per), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out.view(world * smax, per), buf) if dist.get_backend() == "nccl" else dist.all_gather(list(out.unbind(0)), buf)
    parts = [out[r, : hi - lo].cpu().numpy() for r, (lo, hi) in enumerate(sizes)]
    return np.concatenate(parts, axis=0).reshape((total,) + K_local.shape[1:])


def sharded_synthesis(synth_fn, total: int, device=None):
    """run `synth_fn(lo, hi) -> K [hi-lo, ...]` on this rank's shard and all-gather the gain tables"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        lo, hi = shard_range(total, dist.get_rank(), dist.get_world_size())
    else:
        lo, hi = 0, total
    return allgather_gains(synth_fn(lo, hi), total, device)


def _share_counters(tab, dist, device):
    """exit indices / iteration counts of every shard on every rank (each rank filled only its own slice: a sum gathers them)"""
    import torch
    dev = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(np.stack([tab.kexit, tab.iters]).astype(np.int64)).to(dev)
    dist.all_reduce(t)
    t = t.cpu().numpy()
    tab.kexit, tab.iters = t[0].astype(np.int32), t[1].astype(np.int32)


class PeerProtocolError(RuntimeError):
    """a rank failed inside the peer-table protocol; every rank raises (none is left waiting at a barrier)"""


def _agree(ok: bool, device: int) -> bool:
    """all-reduced success flag + barrier: every rank learns whether ANY rank failed, and every rank's stores have landed"""
    import torch
    import torch.distributed as dist
    dev = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_backend() == "nccl":
        torch.cuda.synchronize(device)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    return bool(int(flag.item()))


def _peer_open(tab, device: int, count: int):
    """steps 1-2 of the protocol: allocate my table (`count` doubles), publish its CUDA IPC handle, map the peers' tables.  The mappings stay
    open for the table's lifetime (tab._opened), so that a table can be RE-FILLED (TrackingLQR about a new trajectory, new weights) without
    paying allocation + IPC mapping again -- 0.4 s for 9.6 GB tables on 2 B200s, three times the synthesis itself.
    Returns the local error (or None) and whether every peer is ready; raises nothing."""
    import ctypes as C
    from . import _lib
    from ._lib import check
    import torch.distributed as dist
    L = _lib.lib()
    world, rank = _world()
    err = None
    tab.ptr = C.c_void_p()
    tab._opened = []
    hbuf = (C.c_ubyte * 64)()
    try:
        check(L.ccb_gains_alloc(C.c_int32(device), C.c_int64(count), C.byref(tab.ptr), hbuf))
    except Exception as e:          # noqa: BLE001 -- carried to the agreement
        err = e
        tab.ptr = None
    tab._tables = [tab.ptr]
    peers_ok = True
    if world > 1:
        import torch
        dev = f"cuda:{device}" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor(list(bytes(hbuf)) + [0 if err else 1], dtype=torch.uint8, device=dev)
        allh = [torch.empty(65, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(allh, mine)
        allh = [h.cpu().tolist() for h in allh]
        peers_ok = all(h[64] == 1 for h in allh)
        if peers_ok:
            try:
                for r in range(world):
                    if r == rank:
                        continue
                    hb = (C.c_ubyte * 64)(*allh[r][:64])
                    q = C.c_void_p()
                    check(L.ccb_gains_open(C.c_int32(device), hb, C.byref(q)))
                    tab._opened.append(q); tab._tables.append(q)
            except Exception as e:  # noqa: BLE001
                err = e
    return err, peers_ok


def _peer_unmap(tab):
    from . import _lib
    for q in getattr(tab, "_opened", []) or []:
        _lib.lib().ccb_gains_close(q)
    tab._opened = []
    if getattr(tab, "ptr", None):
        tab._tables = [tab.ptr]


def _peer_fill(tab, device: int, count: int, run_shard):
    """The peer-table protocol shared by PeerGainTable and PeerTrackingTable:

        1. every rank allocates its full table (`count` doubles) and publishes the CUDA IPC handle (+ an ok byte);
        2. the peers' tables are mapped (both once per table: _peer_open; a table that is already open skips them);
        3. `run_shard(tables)` synthesises this rank's shard, its Riccati kernel storing the gain rows into every table;
        4. an all-reduced ok flag and a barrier (every rank's stores into my table have landed).

    A failure on one rank (allocation, mapping, a synthesis error) is carried to step 4 instead of raised on the spot: the collective
    calls are executed by every rank in the same order whatever happened locally, the opened mappings are closed, the local table is
    freed only AFTER the barrier (peers may still be storing into it), and then every rank raises PeerProtocolError."""
    import time
    from . import _lib
    world, rank = _world()
    t_0 = time.time()
    err, peers_ok = (None, True)
    if not getattr(tab, "ptr", None):
        err, peers_ok = _peer_open(tab, device, count)
    t_1 = time.time()
    t_2 = t_1
    all_ok = False
    try:
        if err is None and peers_ok:
            try:
                run_shard(tab._tables)
            except Exception as e:  # noqa: BLE001
                err = e
        t_2 = time.time()
        all_ok = err is None and peers_ok
        if world > 1:
            all_ok = _agree(all_ok, device)
    finally:
        if not all_ok:
            _peer_unmap(tab)
    if not all_ok:
        if tab.ptr:
            _lib.lib().ccb_gains_free(tab.ptr); tab.ptr = None
        raise PeerProtocolError(f"peer gain-table exchange failed on rank {rank}: {err!r}" if err is not None
                                else "peer gain-table exchange failed on another rank") from err
    if world > 1:
        import torch.distributed as dist
        _share_counters(tab, dist, device)
    # wall-clock split of the exchange: table allocation + handle exchange + IPC mapping (first fill only) | synthesis with the pushes | agreement + barrier
    tab.timing = {"alloc_map_s": t_1 - t_0, "synthesis_s": t_2 - t_1, "barrier_s": time.time() - t_2}


def _world():
    import torch.distributed as dist
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    return (dist.get_world_size(), dist.get_rank()) if active else (1, 0)


class PeerGainTable:
    """A device-resident table of finite-horizon LQR gains [S_total][N-1][nu][n] that the GPUs of a node fill TOGETHER: every rank
    synthesises its shard of the systems and its Riccati sweep stores each gain row into every rank's table while it runs (peer
    stores over NVLink through CUDA IPC mappings) -- the all-gather of SURVEY 8(e) fused into the kernel (ccb_lqr_synthesize_peers).
    Without an initialised process group (or with one rank) it degenerates to a local synthesis into the rank's own table."""

    def __init__(self, mech, zd_all, Fd_all, eqc_idx, Qd, Rd, N: int, device: int = 0, Q_batch=None, R_batch=None):
        import ctypes as C
        from . import _lib
        from ._lib import check, f64, i32, ptr
        from .control import handle
        mh = handle(mech)
        self.mech, self.device, self.N = mech, device, int(N)
        self._ctrl, self.ptr = None, None
        eqc_idx = i32(eqc_idx); nu = eqc_idx.size
        zd_all = f64(zd_all).reshape(-1, mh.nb, 13); S = zd_all.shape[0]
        Fd_all = np.ascontiguousarray(np.broadcast_to(f64(Fd_all).reshape(-1, nu), (S, nu)))
        self.S, self.nu, self.n = S, nu, mh.n
        self.eqc_idx, self.zd_all, self.Fd_all = eqc_idx, zd_all, Fd_all
        self.per = (self.N - 1) * nu * mh.n
        L = _lib.lib()
        world, rank = _world()
        lo, hi = shard_range(S, rank, world)
        self.shard = (lo, hi)
        Qd = f64(Qd); Rd = f64(Rd)
        if Q_batch is not None:
            Qs, qbs = np.ascontiguousarray(f64(Q_batch)[lo:hi]), mh.n * mh.n
        else:
            Qs, qbs = Qd, 0
        if R_batch is not None:
            Rs, rbs = np.ascontiguousarray(f64(R_batch)[lo:hi]), nu * nu
        else:
            Rs, rbs = Rd, 0
        self.kexit = np.zeros(S, np.int32); self.iters = np.zeros(S, np.int32)

        def run_shard(tables):      # my shard, written into every table as the sweep produces it
            if hi <= lo:
                return
            arr = (C.c_void_p * len(tables))(*[t.value for t in tables])
            kx = np.empty(hi - lo, np.int32); it = np.empty(hi - lo, np.int32)
            zl = np.ascontiguousarray(zd_all[lo:hi]); fl = np.ascontiguousarray(Fd_all[lo:hi])
            check(L.ccb_lqr_synthesize_peers(mh.h, C.c_int32(device), C.c_int32(hi - lo), ptr(zl), ptr(fl), ptr(eqc_idx), C.c_int32(nu), ptr(Qs), C.c_int64(qbs),
                                             ptr(Rs), C.c_int64(rbs), C.c_int32(self.N), None, C.c_int32(len(tables)), arr, C.c_int64(lo), ptr(kx), ptr(it)))
            self.kexit[lo:hi] = kx; self.iters[lo:hi] = it

        self._run_shard = run_shard
        self._fill()

    def _fill(self):
        """one exchange: the sweeps push the rows they computed into every table (_peer_fill), then -- the break indices of all systems being
        known everywhere -- every rank replicates the rows below each break index in its OWN table (lqr.jl:179-181 back-fill; a fifth of a
        Sawyer table at N = 2000 that therefore never crosses NVLink).  A second agreement closes the fill: nobody starts the next exchange
        (whose stores would land in a table still being completed) before every rank is done, and a failure reaches every rank."""
        import ctypes as C
        import time
        from . import _lib
        from ._lib import check, ptr
        L = _lib.lib()
        world, rank = _world()
        L.ccb_gains_defer_backfill(C.c_int32(1 if world > 1 else 0))
        try:
            _peer_fill(self, self.device, self.S * self.per, self._run_shard)
        finally:
            L.ccb_gains_defer_backfill(C.c_int32(0))
        if world > 1:
            t0 = time.time()
            err = None
            try:
                check(L.ccb_gains_backfill(C.c_int32(self.device), self.ptr, C.c_int64(self.S), C.c_int32(self.N), C.c_int32(self.nu), C.c_int32(self.n),
                                           ptr(self.kexit)))
            except Exception as e:  # noqa: BLE001 -- carried to the agreement
                err = e
            if not _agree(err is None, self.device):
                _peer_unmap(self)
                if self.ptr:
                    L.ccb_gains_free(self.ptr); self.ptr = None
                raise PeerProtocolError(f"gain-table back-fill failed on rank {rank}: {err!r}" if err is not None
                                        else "gain-table back-fill failed on another rank") from err
            self.timing["backfill_s"] = time.time() - t0

    def refill(self):
        """synthesise again into the SAME tables (allocation and IPC mappings are kept): the steady-state cost of the exchange.  Collective."""
        if self._ctrl is not None:
            self._ctrl.invalidate(); self._ctrl = None
        self.kexit[:] = 0; self.iters[:] = 0
        self._fill()
        return self

    def read(self, s0: int = 0, count: int = None) -> np.ndarray:
        """gain tables of systems s0 .. s0+count-1 as a host array [count][N-1][nu][n]"""
        import ctypes as C
        from . import _lib
        from ._lib import check, ptr
        count = self.S - s0 if count is None else count
        out = np.empty((count, self.N - 1, self.nu, self.n))
        check(_lib.lib().ccb_gains_read(C.c_int32(self.device), self.ptr, C.c_int64(s0 * self.per), ptr(out), C.c_int64(count * self.per)))
        return out

    def controller(self):
        """finite-horizon LQR controller (control_lqr!, lqr.jl:89-114) over the table, one gain table per system"""
        from .control import TableLQR
        if self._ctrl is None:
            self._ctrl = TableLQR(self)
        return self._ctrl

    def free(self):
        from . import _lib
        if getattr(self, "_ctrl", None) is not None:
            self._ctrl.invalidate(); self._ctrl = None
        _peer_unmap(self)            # local: the last fill ended with a barrier, nobody stores into a peer's table after it
        if getattr(self, "ptr", None):
            _lib.lib().ccb_gains_free(self.ptr); self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PeerTrackingTable:
    """TrackingLQR gain tables [S_total][N-1][nu][n] of S_total reference trajectories (64 Sawyer tables of 9.4 MB that all
    8 GPUs need) filled by the ranks together: every rank linearises and sweeps its shard of the trajectories and its Riccati kernel
    pushes the gain rows into every rank's table (ccb_tracking_synthesize_peers).  Same protocol as PeerGainTable (_peer_fill)."""

    def __init__(self, mech, storage, Ftau, eqcids, Q, R, device: int = 0):
        import ctypes as C
        from . import _lib
        from ._lib import check, f64, i32, ptr
        from .control import blockdiag, handle
        mh = handle(mech)
        self.mech, self.device = mech, device
        self._ctrl, self.ptr = None, None
        traj = np.ascontiguousarray(np.swapaxes(f64(storage.data), 0, 1))          # [S][N][nb][13]
        S, N = traj.shape[0], traj.shape[1]
        eqc_idx = i32([mech.eqc_index(i) for i in eqcids]); nu = eqc_idx.size
        Ftau = np.ascontiguousarray(np.broadcast_to(f64(Ftau).reshape(-1, N, nu), (S, N, nu)))
        Qd = f64(blockdiag(Q) * mech.dt); Rd = f64(blockdiag(R) * mech.dt)       # lqr_tracking.jl:22-23
        self.S, self.N, self.nu, self.n = S, N, nu, mh.n
        self.eqc_idx, self.traj, self.Ftau = eqc_idx, traj, Ftau
        self.per = (N - 1) * nu * mh.n
        L = _lib.lib()
        world, rank = _world()
        lo, hi = shard_range(S, rank, world)
        self.shard = (lo, hi)
        self.kexit = np.zeros(S, np.int32); self.iters = np.zeros(S, np.int32)

        def run_shard(tables):
            if hi <= lo:
                return
            arr = (C.c_void_p * len(tables))(*[t.value for t in tables])
            kx = np.empty(hi - lo, np.int32); it = np.empty(hi - lo, np.int32)
            tl = np.ascontiguousarray(traj[lo:hi]); fl = np.ascontiguousarray(Ftau[lo:hi])
            check(L.ccb_tracking_synthesize_peers(mh.h, C.c_int32(device), C.c_int32(hi - lo), C.c_int32(N), ptr(tl), ptr(fl), ptr(eqc_idx), C.c_int32(nu),
                                                  ptr(Qd), ptr(Rd), None, C.c_int32(len(tables)), arr, C.c_int64(lo), ptr(kx), ptr(it)))
            self.kexit[lo:hi] = kx; self.iters[lo:hi] = it

        self._run_shard = run_shard
        self._fill()

    def _fill(self):
        # the time-varying recursion has no early break (lqr_tracking.jl:87-115): every row is computed and pushed, nothing to back-fill
        _peer_fill(self, self.device, self.S * self.per, self._run_shard)

    refill = PeerGainTable.refill
    read = PeerGainTable.read
    free = PeerGainTable.free
    __del__ = PeerGainTable.__del__

    def controller(self):
        """TrackingLQR controller (control_trackinglqr!, lqr_tracking.jl:46-71) over the table, one table per reference trajectory"""
        from .control import TableTracking
        if self._ctrl is None:
            self._ctrl = TableTracking(self)
        return self._ctrl
