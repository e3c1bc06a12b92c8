"""World-size-2 gloo test of the z-slab exchange protocol (fluid-rs_b200/slab.py) on CPU.
The CUDA engine is replaced by a numpy mock with the same interface, so what is tested is the
host logic: slab planning, symmetric halo-plane accumulation, migrant counts and routing."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class MockEngine:
    """Particles drift along z with a fixed velocity; the 'grid' is two planes per interface holding
    the number of particles within one cell of the face (a stand-in for partial node sums)."""

    def __init__(self, rank, world, slabs, rec):
        self.rank, self.world, self.slabs = rank, world, slabs
        self.lo, self.hi = slabs[rank]
        self.rec = rec                       # (n, 17) float32: packed migrant layout, pos z at col 2, vz at col 5
        self.planes_own = {}
        self.planes_recv = {}
        self.log = []

    def phase(self, i, mouse):
        if i == 0:
            for side, face in ((0, self.lo), (1, self.hi)):
                z = self.rec[:, 2]
                near = np.float32(((z >= face - 1) & (z < face + 1)).sum())
                for kind, width in ((0, 8), (1, 32)):
                    self.planes_own[(side, kind)] = torch.full((width,), float(near), dtype=torch.float32)
                    self.planes_recv[(side, kind)] = torch.zeros(width, dtype=torch.float32)
        if i == 2:
            self.rec[:, 2] += self.rec[:, 5]

    def planes(self, side, kind):
        return self.planes_own[(side, kind)], self.planes_recv[(side, kind)]

    def accumulate(self, side, kind):
        self.planes_own[(side, kind)] += self.planes_recv[(side, kind)]
        self.log.append((side, kind, float(self.planes_own[(side, kind)][0])))

    def migrants(self):
        cz = np.floor(self.rec[:, 2])
        lo = cz < self.lo if self.rank > 0 else np.zeros(len(cz), bool)
        hi = cz >= self.hi if self.rank < self.world - 1 else np.zeros(len(cz), bool)
        out = [torch.from_numpy(self.rec[m].reshape(-1).copy()) for m in (lo, hi)]
        self.rec = self.rec[~(lo | hi)]
        return out

    def recv_buffer(self, side, n):
        return torch.empty(n * 17, dtype=torch.float32)

    def append(self, t):
        self.rec = np.concatenate([self.rec, t.numpy().reshape(-1, 17)])


class FixedMessageEngine(MockEngine):
    """The same mock behind the one-synchronisation migration interface (migrants_begin / migrants_end):
    every neighbour pair first swaps a fixed-size message — a header record with the count, then up to
    `first_records` records — and only then learns the counts; the rest follows in a second message."""

    first_records = 8          # small, so that the second message is exercised

    def __init__(self, *a):
        super().__init__(*a)
        self._recv = [None, None]
        self._n_out = [0, 0]

    def migrants_begin(self):
        out = self.migrants()
        bufs = []
        for side, t in enumerate(out):
            n = t.numel() // 17
            self._n_out[side] = n
            buf = torch.zeros((max(n, self.first_records) + 1) * 17, dtype=torch.float32)
            buf[:1].view(torch.int32)[0] = n          # header: the count as an int32 in the first word
            buf[17:17 + n * 17] = t
            has_nb = (side == 0 and self.rank > 0) or (side == 1 and self.rank < self.world - 1)
            bufs.append(buf if has_nb else None)
        return bufs

    def recv_buffer(self, side, n, keep: int = 0):
        need = n * 17
        if self._recv[side] is None or self._recv[side].numel() < need:
            old = self._recv[side]
            self._recv[side] = torch.zeros(need, dtype=torch.float32)
            if keep and old is not None:
                self._recv[side][:keep] = old[:keep]
        return self._recv[side][:need]

    def migrants_end(self, recv):
        n_in = [int(t[:1].view(torch.int32)[0]) if t is not None else 0 for t in recv]
        return list(self._n_out), n_in


class PeerMemoryEngine(MockEngine):
    """The mock behind the peer-memory interface the CUDA engine offers once the neighbours' buffers are
    mapped: phase 2 'writes' the leavers into the neighbour's receive buffer (here: a message sent at once),
    peer_barrier() orders it, migrants_end([None, None]) reads the counts of this rank's own receive buffers
    and append_received() joins them.  Exercises SlabDriver's communication-library-free substep path."""

    def __init__(self, *a):
        super().__init__(*a)
        self._n_out = [0, 0]
        self._inbox = [None, None]
        self.barriers = 0

    def migrants_begin(self):
        out = self.migrants()
        reqs, recv_cnt = [], [torch.zeros(1, dtype=torch.int64) for _ in (0, 1)]
        nbs = [self.rank - 1 if self.rank > 0 else None, self.rank + 1 if self.rank < self.world - 1 else None]
        for side, t in enumerate(out):
            self._n_out[side] = t.numel() // 17
            if nbs[side] is None:
                continue
            reqs.append(dist.isend(torch.tensor([self._n_out[side]], dtype=torch.int64), nbs[side], tag=10 + side))
            reqs.append(dist.irecv(recv_cnt[side], nbs[side], tag=10 + (1 - side)))
        for r in reqs:
            r.wait()
        reqs = []
        for side, t in enumerate(out):
            if nbs[side] is None:
                self._inbox[side] = torch.zeros(0)
                continue
            self._inbox[side] = torch.zeros(int(recv_cnt[side]) * 17, dtype=torch.float32)
            if t.numel():
                reqs.append(dist.isend(t, nbs[side], tag=20 + side))
            if self._inbox[side].numel():
                reqs.append(dist.irecv(self._inbox[side], nbs[side], tag=20 + (1 - side)))
        for r in reqs:
            r.wait()
        return None

    def peer_barrier(self):
        self.barriers += 1
        dist.barrier()

    def migrants_end(self, recv):
        assert recv == [None, None]
        return list(self._n_out), [t.numel() // 17 for t in self._inbox]

    def append_received(self, side, n):
        assert n == self._inbox[side].numel() // 17
        self.append(self._inbox[side])


def _worker(rank, world, port, result_dir, engine="MockEngine"):
    sys.path.insert(0, str(ROOT))
    import fluidpkg
    slab = fluidpkg.load().slab
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    slabs = slab.plan_slabs(3.0, 67.0, -16, 96, world)
    rng = np.random.default_rng(7)          # same stream on both ranks: one global scene
    n = 4000
    rec = np.zeros((n, 17), dtype=np.float32)
    rec[:, :3] = rng.uniform(3, 67, (n, 3))
    rec[:, 5] = rng.normal(0, 0.6, n)
    rec[:, 16] = np.arange(n, dtype=np.int32).view(np.float32)
    lo, hi = slabs[rank]
    mine = (np.floor(rec[:, 2]) >= lo) & (np.floor(rec[:, 2]) < hi)
    eng = globals()[engine](rank, world, slabs, rec[mine].copy())
    drv = slab.SlabDriver(eng, rank, world, dist, device="cpu", p2p=engine == "PeerMemoryEngine")
    for _ in range(12):
        drv.substep()
    cz = np.floor(eng.rec[:, 2])
    inside = bool(((cz >= lo) | (rank == 0)).all() and ((cz < hi) | (rank == world - 1)).all())
    ids = eng.rec[:, 16].copy().view(np.int32)
    np.savez(os.path.join(result_dir, f"r{rank}.npz"), ids=ids, inside=inside, out=drv.migrated_out,
             inn=drv.migrated_in, log=np.array(eng.log, dtype=np.float64), z=eng.rec[:, 2],
             z0=rec[:, 2], vz=rec[:, 5])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("engine", ["MockEngine", "FixedMessageEngine", "PeerMemoryEngine"])
def test_slab_protocol_world2_gloo(tmp_path, engine):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), engine), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    ids = np.concatenate([x["ids"] for x in r])
    assert sorted(ids.tolist()) == list(range(4000))          # every particle exactly once
    assert all(bool(x["inside"]) for x in r)                  # and in the slab that owns its cell
    assert int(r[0]["out"]) + int(r[1]["out"]) == int(r[0]["inn"]) + int(r[1]["inn"]) > 0
    if engine != "PeerMemoryEngine":     # (with peer memory the kernels deposit into both copies: no plane exchange)
        # halo planes: rank 0's upper interface and rank 1's lower interface hold the same sums
        l0 = r[0]["log"][r[0]["log"][:, 0] == 1][:, 2]
        l1 = r[1]["log"][r[1]["log"][:, 0] == 0][:, 2]
        np.testing.assert_array_equal(l0, l1)
    else:
        assert all(len(x["log"]) == 0 for x in r)
    # positions advanced exactly as in a single-process run
    z_expect = r[0]["z0"] + 12 * r[0]["vz"]
    for x in r:
        np.testing.assert_allclose(x["z"], z_expect[x["ids"]], rtol=0, atol=1e-4)


def test_plan_slabs_faces_on_tile_boundaries():
    sys.path.insert(0, str(ROOT))
    import fluidpkg
    slab = fluidpkg.load().slab
    for world in (1, 2, 4, 8):
        s = slab.plan_slabs(3.0, 515.0, -16, 560, world)
        assert s[0][0] == -16 and s[-1][1] == 544 and len(s) == world
        for (a, b), (c, d) in zip(s[:-1], s[1:]):
            assert b == c and (b + 16) % 4 == 0 and b > a
        thick = [b - a for a, b in s[1:-1]]
        assert not thick or max(thick) - min(thick) <= 4
