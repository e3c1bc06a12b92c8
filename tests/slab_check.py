"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/slab_check.py
The z-slab run must reproduce the single-GPU run of the same scene: every particle present exactly
once, positions / velocities equal up to summation-order rounding, and (since halo sums are a+b on
both sides) no drift in the counts."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fluidpkg  # noqa: E402


def main():
    substeps = int(os.environ.get("SLAB_CHECK_SUBSTEPS", "60"))
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = fluidpkg.load()
    slab = pkg.slab
    sc = pkg.scenes.dam_break_3d(40, 32, 64)
    # a sloshing start so that particles really cross the slab faces
    rec = sc.records()
    rng = np.random.default_rng(3)
    rec[:, 3:6] = rng.normal(0, 0.4, (sc.n, 3)).astype(np.float32)
    rec[:, 5] += 0.8
    ids = np.arange(sc.n, dtype=np.int32)

    sim = slab.SlabSimulation(pkg, sc.cfg, sc.rect_min, sc.rect_max, float(sc.fill_lo[2]), float(sc.fill_hi[2]),
                              rank, world, dist, local, reserve=sc.n)
    mine = sim.add_particles(rec, ids)
    sim.substeps(substeps)
    out, oid = sim.sim.read_particles()
    counts = sim.sim.particle_counts()
    gathered = [None] * world
    dist.gather_object((out, oid, mine, counts, sim.driver.migrated_out, sim.driver.migrated_in, sim.slabs[rank]),
                       gathered if rank == 0 else None, dst=0)
    ok = True
    if rank == 0:
        one = pkg.Simulation.new(sc.cfg, device=local)
        one.add_particles(rec, ids)
        one.set_rect(sc.rect_min, sc.rect_max)
        one.substeps(substeps)
        ref, rid = one.read_particles(sort_by_id=True)
        allrec = np.concatenate([g[0] for g in gathered])
        allid = np.concatenate([g[1] for g in gathered])
        o = np.argsort(allid, kind="stable")
        allrec, allid = allrec[o], allid[o]
        moved = sum(g[4] for g in gathered)
        print("halo:", "peer memory (P2P deposits)" if sim.p2p else "plane exchange")
        print("slabs", [g[6] for g in gathered], "start", [g[2] for g in gathered],
              "end", [len(g[1]) for g in gathered], "migrated out/in", moved, sum(g[5] for g in gathered))
        ok &= np.array_equal(allid, rid)
        if ok:
            dp = np.abs(allrec[:, :3] - ref[:, :3]).max()
            dv = np.abs(allrec[:, 3:6] - ref[:, 3:6]).max()
            print(f"max |dpos| {dp:.3e}  max |dvel| {dv:.3e} after {substeps} substeps")
            ok &= bool(dp < 5e-3 and dv < 5e-3)
        else:
            print("id sets differ:", len(allid), len(rid), len(set(allid.tolist())))
        ok &= moved > 0 and moved == sum(g[5] for g in gathered)
        print("SLAB CHECK", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
