"""Multi-GPU parity check, run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/slab_check.py
The z-slab run must reproduce the single-GPU run of the same scene: every particle present exactly
once, positions / velocities equal up to summation-order rounding, particles sent == received.  The check
itself is `slab.parity_check` (also run by bench.py outside its timed region for every N > 1)."""
import json
import os
import sys
from pathlib import Path

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
    res = pkg.slab.parity_check(pkg, dist, rank, world, local, substeps=substeps)
    # the same scene in the deterministic mode: the N-rank run must equal the one-GPU run BIT FOR BIT
    det = pkg.slab.parity_check(pkg, dist, rank, world, local, substeps=substeps, deterministic=True)
    if rank == 0:
        print(json.dumps(res))
        print(json.dumps(det))
        print("SLAB CHECK", "OK" if res["ok"] and det["ok"] else "FAILED")
    res["ok"] = res["ok"] and det["ok"]
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
