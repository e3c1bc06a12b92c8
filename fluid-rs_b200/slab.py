"""z-slab decomposition across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink).  Every rank holds the same dense node
grid geometry and owns the particles whose cell floor(pos.z) lies in its slab.  Per substep:

    phase 0  sort + clear + p2g 1     -> the two MASS planes of each interface become complete
    phase 1  p2g 2                    -> the two NODE planes of each interface become complete
    phase 2  update + g2p             -> particles that left the slab go to the neighbour

How the shared planes become complete:
  * peer-memory halo (default on GPUs): every rank maps its neighbours' grids through CUDA IPC and the tile
    kernels add the deposits that fall into the shared planes into the neighbour's copy as well
    (red.global.add over NVLink inside the kernels' flush) — no plane exchange, no accumulate pass, only a
    neighbour barrier after the phase;
  * plane exchange (FLUID_B200_SLAB_P2P=0, and the CPU protocol test): both ranks swap their partial
    planes over the process group and add.

The reference's own analogue is the +-1 block halo ring (`p_rect` vs `a_rect`, 3d:84-86) in which
particles deposit but are not advanced, and the per-block mailboxes `swap_mul` (3d:327-358).

The exchange protocol (`SlabDriver`) only needs an engine object with
    phase(i), planes(side, kind) -> (own, recv) tensors, accumulate(side, kind),
    migrants() -> (lower, upper) tensors of packed records, append(tensor)
so it runs unchanged over NCCL with the CUDA engine and over gloo with a CPU mock (tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

MIG_WORDS = 17     # packed migrant record: 16 f32 + id bits
TILE_Z = 4         # slab faces sit on tile boundaries (multiples of 4 cells from the grid origin)


def plan_slabs(fill_lo: float, fill_hi: float, origin_z: int, size_z: int, world: int):
    """Split the grid's z range [origin_z, origin_z + size_z) into `world` slabs (world cells).
    Interior faces divide the filled range [fill_lo, fill_hi) evenly, rounded to multiples of 4 cells
    from the grid origin; the first and last slab extend to the ends of the grid."""
    faces = [origin_z]
    for r in range(1, world):
        z = fill_lo + (fill_hi - fill_lo) * r / world
        rel = int(round((z - origin_z) / TILE_Z)) * TILE_Z
        rel = min(max(rel, faces[-1] - origin_z + TILE_Z), size_z - TILE_Z * (world - r))
        faces.append(origin_z + rel)
    faces.append(origin_z + size_z)
    return [(faces[r], faces[r + 1]) for r in range(world)]


class SlabDriver:
    """The per-substep exchange protocol, independent of device and backend."""

    def __init__(self, engine, rank: int, world: int, dist=None, device="cpu", p2p: bool = False):
        self.e = engine
        self.rank, self.world = rank, world
        self.dist = dist
        self.device = device
        self.nb = [rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None]
        self.migrated_out = 0
        self.migrated_in = 0
        self.p2p = p2p          # peer-memory halo: the kernels deposit into the neighbours' planes themselves
        self._token = None

    # -- neighbour barrier (peer-memory halo) ---------------------------------------------------
    def neighbour_barrier(self):
        """A token each way with every neighbour: when it completes on this rank's stream, the neighbour's
        phase kernel — and with it its deposits into this rank's planes — has finished."""
        import torch
        dist = self.dist
        if self._token is None:
            self._token = [torch.zeros(1, dtype=torch.float32, device=self.device) for _ in range(4)]
        ops = []
        for side in (0, 1):
            if self.nb[side] is None:
                continue
            ops.append(dist.P2POp(dist.isend, self._token[2 * side], self.nb[side]))
            ops.append(dist.P2POp(dist.irecv, self._token[2 * side + 1], self.nb[side]))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    # -- halo planes -----------------------------------------------------------------------------
    def exchange_planes(self, kind: int):
        """Both ranks of an interface hold partial sums of its two node planes: swap and add."""
        dist = self.dist
        ops, sides = [], []
        for side in (0, 1):
            if self.nb[side] is None:
                continue
            own, recv = self.e.planes(side, kind)
            ops.append(dist.P2POp(dist.isend, own, self.nb[side]))
            ops.append(dist.P2POp(dist.irecv, recv, self.nb[side]))
            sides.append(side)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for side in sides:
            self.e.accumulate(side, kind)

    # -- migration -------------------------------------------------------------------------------
    def migrate(self):
        import torch
        dist = self.dist
        out = self.e.migrants()                      # [lower, upper] tensors (n * 17 words), may be empty
        n_out = [int(t.numel()) // MIG_WORDS for t in out]
        # 1. counts
        send_cnt = [torch.tensor([n_out[s]], dtype=torch.int64, device=self.device) for s in (0, 1)]
        recv_cnt = [torch.zeros(1, dtype=torch.int64, device=self.device) for _ in (0, 1)]
        ops = []
        for side in (0, 1):
            if self.nb[side] is None:
                continue
            ops.append(dist.P2POp(dist.isend, send_cnt[side], self.nb[side]))
            ops.append(dist.P2POp(dist.irecv, recv_cnt[side], self.nb[side]))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        n_in = [int(recv_cnt[s].item()) if self.nb[s] is not None else 0 for s in (0, 1)]
        # 2. records
        bufs = [self.e.recv_buffer(side, n_in[side]) if n_in[side] else None for side in (0, 1)]
        ops = []
        for side in (0, 1):
            if self.nb[side] is None:
                continue
            if n_out[side]:
                ops.append(dist.P2POp(dist.isend, out[side], self.nb[side]))
            if n_in[side]:
                ops.append(dist.P2POp(dist.irecv, bufs[side], self.nb[side]))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for side in (0, 1):
            if n_in[side]:
                self.e.append(bufs[side])
        self.migrated_out += sum(n_out)
        self.migrated_in += sum(n_in)

    def migrate_fixed(self):
        """Migration with one synchronisation: every neighbour pair swaps a fixed-size message (a header
        record with the count, then up to `first_records` records) straight away; only then are the counts
        read (engine.migrants_end, the substep's one host sync).  A count above `first_records` is known to
        both sides and the rest follows in a second message."""
        dist = self.dist
        X = self.e.first_records
        send = self.e.migrants_begin()                  # [lower, upper] full buffers (or None)
        recv = [self.e.recv_buffer(side, X + 1) if self.nb[side] is not None else None for side in (0, 1)]
        ops = []
        for side in (0, 1):
            if self.nb[side] is None:
                continue
            ops.append(dist.P2POp(dist.isend, send[side][: (X + 1) * MIG_WORDS], self.nb[side]))
            ops.append(dist.P2POp(dist.irecv, recv[side][: (X + 1) * MIG_WORDS], self.nb[side]))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        n_out, n_in = self.e.migrants_end(recv)
        ops = []
        for side in (0, 1):
            if self.nb[side] is None:
                continue
            if n_out[side] > X:
                ops.append(dist.P2POp(dist.isend, send[side][(X + 1) * MIG_WORDS:(n_out[side] + 1) * MIG_WORDS], self.nb[side]))
            if n_in[side] > X:
                recv[side] = self.e.recv_buffer(side, n_in[side] + 1, keep=(X + 1) * MIG_WORDS)
                ops.append(dist.P2POp(dist.irecv, recv[side][(X + 1) * MIG_WORDS:(n_in[side] + 1) * MIG_WORDS], self.nb[side]))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for side in (0, 1):
            if self.nb[side] is not None and n_in[side]:
                self.e.append(recv[side][MIG_WORDS:(n_in[side] + 1) * MIG_WORDS])
        self.migrated_out += sum(n_out)
        self.migrated_in += sum(n_in)

    def migrate_peer(self):
        """Peer-memory migration: phase 2 wrote the leavers' records into the neighbours' receive buffers,
        migrants_begin adds the count headers, a peer barrier makes them visible; the one host
        synchronisation of the substep reads the counts."""
        self.e.migrants_begin()
        self.e.peer_barrier()
        n_out, n_in = self.e.migrants_end([None, None])
        for side in (0, 1):
            if self.nb[side] is not None and n_in[side]:
                self.e.append_received(side, n_in[side])
        self.migrated_out += sum(n_out)
        self.migrated_in += sum(n_in)

    def substep(self, mouse=None):
        if self.p2p:
            # nothing here goes through the process group: deposits, records and barriers are stores over NVLink
            self.e.phase(0, None)
            self.e.peer_barrier()
            self.e.phase(1, None)
            self.e.peer_barrier()
            self.e.phase(2, mouse)
            self.migrate_peer()      # its barrier also orders the next substep's deposits behind this one's reads
            return
        self.e.phase(0, None)
        self.exchange_planes(0)
        self.e.phase(1, None)
        self.exchange_planes(1)
        self.e.phase(2, mouse)
        # always a send/recv pair with each neighbour: it also orders the next substep
        if hasattr(self.e, "migrants_begin"):
            self.migrate_fixed()
        else:
            self.migrate()


class _DevArray:
    """Zero-copy view of device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(int(ptr), False), version=2)


class CudaSlabEngine:
    """Adapter: the C ABI's fluid_slab_* calls behind the SlabDriver's engine interface."""

    def __init__(self, pkg, sim):
        self.pkg, self.sim = pkg, sim
        self.L = pkg.lib()
        self._recv = [None, None]
        self._send_cap = None
        self._send_views = None
        self.peer_memory = False      # set by SlabSimulation once the neighbours' buffers are mapped

    def _chk(self, st):
        if st != 0:
            raise self.pkg.FluidError(st, self.L.fluid_last_error().decode())

    def phase(self, i, mouse):
        mp = None
        if mouse is not None:
            m = np.ascontiguousarray(mouse, dtype=np.float32)
            mp = m.ctypes.data_as(C.POINTER(C.c_float))
        self._chk(self.L.fluid_slab_phase(self.sim._h, i, mp))

    def planes(self, side, kind):
        import torch
        own, recv, n = C.c_void_p(), C.c_void_p(), C.c_int64()
        self._chk(self.L.fluid_slab_planes(self.sim._h, side, kind, C.byref(own), C.byref(recv), C.byref(n)))
        words = n.value * (4 if kind == 1 else 1)
        return (torch.as_tensor(_DevArray(own.value, words), device="cuda"),
                torch.as_tensor(_DevArray(recv.value, words), device="cuda"))

    def accumulate(self, side, kind):
        self._chk(self.L.fluid_slab_accumulate(self.sim._h, side, kind))

    def migrants(self):
        import torch
        lo, hi, nlo, nhi = C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        self._chk(self.L.fluid_slab_migrants(self.sim._h, C.byref(lo), C.byref(nlo), C.byref(hi), C.byref(nhi)))
        out = []
        for ptr, n in ((lo, nlo.value), (hi, nhi.value)):
            if n and ptr.value:
                out.append(torch.as_tensor(_DevArray(ptr.value, n * MIG_WORDS), device="cuda"))
            else:
                out.append(torch.empty(0, dtype=torch.float32, device="cuda"))
        return out

    def recv_buffer(self, side, n, keep: int = 0):
        import torch
        need = n * MIG_WORDS
        if self._recv[side] is None or self._recv[side].numel() < need:
            old = self._recv[side]
            self._recv[side] = torch.empty(max(need, 1 << 16), dtype=torch.float32, device="cuda")
            if keep and old is not None:
                self._recv[side][:keep] = old[:keep]
        return self._recv[side][:need]

    first_records = 4096      # records a neighbour pair exchanges before the counts are known

    def migrants_begin(self):
        import torch
        lo, hi = C.c_void_p(), C.c_void_p()
        self._chk(self.L.fluid_slab_migrants_begin(self.sim._h, C.byref(lo), C.byref(hi)))
        if self.peer_memory:
            return None
        if self._send_cap is None:
            self._send_cap = 1 << 18       # fluid_slab_set's migrant buffer (records), + the header record
        words = (self._send_cap + 1) * MIG_WORDS
        key = (lo.value, hi.value)
        if self._send_views is None or self._send_views[0] != key:      # the buffers only move with fluid_slab_set
            self._send_views = (key, [torch.as_tensor(_DevArray(p.value, words), device="cuda") if p.value else None
                                      for p in (lo, hi)])
        return self._send_views[1]

    def peer_barrier(self):
        self._chk(self.L.fluid_slab_peer_barrier(self.sim._h))

    def append_received(self, side, n):
        self._chk(self.L.fluid_slab_append_received(self.sim._h, side, n))

    def migrants_end(self, recv):
        n_out, n_in = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        ptrs = [C.c_void_p(t.data_ptr()) if t is not None else None for t in recv]
        self._chk(self.L.fluid_slab_migrants_end(self.sim._h, ptrs[0], ptrs[1], n_out, n_in))
        return [int(n_out[0]), int(n_out[1])], [int(n_in[0]), int(n_in[1])]

    def append(self, t):
        self._chk(self.L.fluid_slab_append(self.sim._h, C.c_void_p(t.data_ptr()), t.numel() // MIG_WORDS))


class SlabSimulation:
    """`Simulation` spread over the ranks of a torch.distributed process group, one z-slab each."""

    def __init__(self, pkg, cfg, rect_min, rect_max, fill_lo_z, fill_hi_z, rank, world, dist, device: int,
                 reserve: int = 0):
        import torch
        self.pkg, self.rank, self.world = pkg, rank, world
        self.sim = pkg.Simulation.new(cfg, device=device)
        self.stream = torch.cuda.Stream(device=device)
        torch.cuda.set_stream(self.stream)
        self.sim.set_stream(self.stream.cuda_stream)
        self.sim.set_rect(rect_min, rect_max)
        r = self.sim.rects()
        self.slabs = plan_slabs(fill_lo_z, fill_hi_z, int(r["origin"][2]), int(r["size"][2]), world)
        self.z_lo, self.z_hi = self.slabs[rank]
        L = pkg.lib()
        if reserve:
            st = L.fluid_reserve(self.sim._h, int(reserve))
            if st != 0:
                raise pkg.FluidError(st, L.fluid_last_error().decode())
        st = L.fluid_slab_set(self.sim._h, self.z_lo, self.z_hi, int(rank > 0), int(rank < world - 1))
        if st != 0:
            raise pkg.FluidError(st, L.fluid_last_error().decode())
        import os
        p2p = world > 1 and os.environ.get("FLUID_B200_SLAB_P2P", "1") != "0"
        if p2p:
            p2p = self._map_neighbours(L, dist, rank, world, device)
        self.p2p = p2p
        engine = CudaSlabEngine(pkg, self.sim)
        engine.peer_memory = p2p
        self.driver = SlabDriver(engine, rank, world, dist, device=f"cuda:{device}", p2p=p2p)
        self.iterations = int(self.sim.config.iterations)

    def _map_neighbours(self, L, dist, rank, world, device) -> bool:
        """Peer-memory halo set-up: all-gather the CUDA IPC handles of every rank's arrays and map the two
        neighbours'.  Falls back to plane exchanges (on every rank) if any mapping fails."""
        import torch
        IPC_BYTES = 448
        mine = (C.c_ubyte * IPC_BYTES)()
        ok = L.fluid_slab_ipc_export(self.sim._h, C.cast(mine, C.c_void_p)) == 0
        t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=f"cuda:{device}")
        allh = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allh, t)
        for side, nb in ((0, rank - 1), (1, rank + 1)):
            if ok and 0 <= nb < world:
                buf = (C.c_ubyte * IPC_BYTES)(*allh[nb].cpu().tolist())
                ok = L.fluid_slab_ipc_import(self.sim._h, side, C.cast(buf, C.c_void_p)) == 0
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=f"cuda:{device}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # also the barrier between the wipes and the first substep
        if int(flag.item()) == 0:
            for side in (0, 1):
                L.fluid_slab_ipc_import(self.sim._h, side, None)
            return False
        return True

    def owns(self, records: np.ndarray) -> np.ndarray:
        cz = np.floor(records[:, 2]).astype(np.int64)
        return (cz >= self.z_lo) & (cz < self.z_hi)

    def add_particles(self, records, ids=None):
        """Adds the records that fall into this rank's slab (callers may pass the whole scene)."""
        records = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, 16)
        keep = self.owns(records)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)[keep]
        if keep.any():
            self.sim.add_particles(records[keep], ids)
        return int(keep.sum())

    def substeps(self, n, mouse=None):
        for _ in range(n):
            self.driver.substep(mouse)

    def step(self, mouse=None):
        self.substeps(self.iterations, mouse)

    def frame_counts(self, viewport, cols: int = 80, rows: int = 40):
        """`draw`'s 80x40 binning (3d:469-486) over all slabs: every rank bins its own particles on the
        device, the bin counts are summed over the process group."""
        import torch
        counts = self.sim.frame_counts(viewport=viewport, cols=cols, rows=rows)
        t = torch.from_numpy(counts).to(f"cuda:{torch.cuda.current_device()}")
        self.driver.dist.all_reduce(t)
        return t.cpu().numpy()

    def close(self):
        self.sim.close()


def parity_check(pkg, dist, rank: int, world: int, device: int, substeps: int = 60, tol: float = 5e-4,
                 deterministic: bool = False) -> dict:
    """N-rank z-slab run against the single-GPU run of the same sloshing scene (81,920 particles, every
    rank crosses its faces).  Collective: every rank of `dist` must call it.  Rank 0 returns the verdict
    {ok, ranks, particles, substeps, max_dpos, max_dvel, tol, migrated_out, migrated_in, ids_once, halo};
    the other ranks return {ok} only.  Bars: every id present exactly once over the ranks, particles sent ==
    particles received (> 0), and |dpos|, |dvel| < tol against the one-GPU run (the node sums of the shared
    planes are float reductions in a different order, so the last bits differ and 60 substeps of sloshing amplify
    them: measured 4e-5 .. 1.1e-4 from run to run).
    deterministic=True runs both sides in the deterministic mode (64-bit fixed-point node sums, planes exchanged
    and added as integers): the bar becomes BIT-FOR-BIT equality of every field of every particle."""
    import os
    import torch
    old_env = os.environ.get("FLUID_B200_DETERMINISTIC")
    if deterministic:
        os.environ["FLUID_B200_DETERMINISTIC"] = "1"      # read by fluid_create
        tol = 0.0
    sc = pkg.scenes.dam_break_3d(40, 32, 64)
    rec = sc.records()
    rng = np.random.default_rng(3)
    rec[:, 3:6] = rng.normal(0, 0.4, (sc.n, 3)).astype(np.float32)
    rec[:, 5] += 0.8          # a sloshing start along z, so that particles really cross the slab faces
    ids = np.arange(sc.n, dtype=np.int32)
    sim = SlabSimulation(pkg, sc.cfg, sc.rect_min, sc.rect_max, float(sc.fill_lo[2]), float(sc.fill_hi[2]),
                         rank, world, dist, device, reserve=sc.n)
    mine = sim.add_particles(rec, ids)
    sim.substeps(substeps)
    out, oid = sim.sim.read_particles()
    gathered = [None] * world
    dist.gather_object((out, oid, mine, sim.driver.migrated_out, sim.driver.migrated_in, sim.slabs[rank]),
                       gathered if rank == 0 else None, dst=0)
    res = {"ok": True}
    if rank == 0:
        one = pkg.Simulation.new(sc.cfg, device=device)
        one.add_particles(rec, ids)
        one.set_rect(sc.rect_min, sc.rect_max)
        one.set_resident_max(0)
        one.substeps(substeps)
        ref, rid = one.read_particles(sort_by_id=True)
        one.close()
        allrec = np.concatenate([g[0] for g in gathered])
        allid = np.concatenate([g[1] for g in gathered])
        o = np.argsort(allid, kind="stable")
        allrec, allid = allrec[o], allid[o]
        sent, got = sum(g[3] for g in gathered), sum(g[4] for g in gathered)
        ids_once = bool(np.array_equal(allid, rid))
        dp = float(np.abs(allrec[:, :3] - ref[:, :3]).max()) if ids_once else float("inf")
        dv = float(np.abs(allrec[:, 3:6] - ref[:, 3:6]).max()) if ids_once else float("inf")
        bitwise = bool(ids_once and np.array_equal(allrec.view(np.uint32), ref.view(np.uint32)))
        close = bitwise if deterministic else (dp < tol and dv < tol)
        res = {"ok": bool(ids_once and close and sent > 0 and sent == got), "deterministic": deterministic, "bitwise_equal": bitwise,
               "ranks": world, "particles": int(sc.n), "substeps": int(substeps), "max_dpos": dp, "max_dvel": dv,
               "tol": tol, "migrated_out": int(sent), "migrated_in": int(got), "ids_once": ids_once,
               "halo": "peer memory (P2P deposits)" if sim.p2p else "plane exchange",
               "slabs": [list(g[5]) for g in gathered], "start": [int(g[2]) for g in gathered],
               "end": [int(len(g[1])) for g in gathered]}
    sim.close()
    if deterministic:
        if old_env is None:
            os.environ.pop("FLUID_B200_DETERMINISTIC", None)
        else:
            os.environ["FLUID_B200_DETERMINISTIC"] = old_env
    flag = torch.tensor([1 if res["ok"] else 0], device=f"cuda:{device}")
    dist.broadcast(flag, 0)
    if rank != 0:
        res = {"ok": bool(flag.item() == 1)}
    return res
