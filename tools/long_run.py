"""Long single-GPU run of the bench scene with aggregate checks (GPU box only):
particle count constant, every position finite and inside the clip box, kinetic energy bounded; and the
throughput of every step against simulated time (the rate falls as the dam spreads over more, emptier tiles).
usage: python tools/long_run.py [16M|1M] [steps]"""
import sys
import time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fluidpkg


def main():
    pkg = fluidpkg.load()
    scenes = pkg.scenes
    which = sys.argv[1] if len(sys.argv) > 1 else "16M"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    sc = scenes.dam_break_for_gpus(1) if which == "16M" else scenes.dam_break_1m()
    sim = pkg.Simulation.new(sc.cfg)
    sim.set_rect(sc.rect_min, sc.rect_max)
    chunk = 1 << 21
    for s in range(0, sc.n, chunk):
        sim.add_particles(sc.records(s, min(chunk, sc.n - s)))
    lo, hi = np.asarray(sc.cfg["clip_min"][:3]), np.asarray(sc.cfg["clip_max"][:3])
    iters = sc.cfg["iterations"]
    for k in range(steps):
        sim.synchronize()
        t0 = time.perf_counter()
        sim.step()
        sim.synchronize()
        dt = time.perf_counter() - t0
        tiles = sim.debug_tiles()
        tiles = tiles[tiles[:, 2] > 0]
        print(f"step {k + 1}: {sc.n * iters / dt / 1e9:.3f} G updates/s  ({dt * 1e3 / iters:.3f} ms per substep)  "
              f"active tiles {len(tiles)}  particles per tile {tiles[:, 2].mean():.1f}", flush=True)
        if (k + 1) % 10 == 0 or k + 1 == steps:
            c = sim.particle_counts()
            rec, ids = sim.read_particles()
            pos, vel, m = rec[:, :3], rec[:, 3:6], rec[:, -1]
            ok = bool(np.isfinite(rec).all() and (pos >= lo - 1e-3).all() and (pos <= hi + 1e-3).all())
            ke = float(0.5 * (m * (vel.astype(np.float64) ** 2).sum(axis=1)).sum())
            front = float(pos[:, 0].max())
            print(f"step {k + 1}: counts {c} unique ids {len(np.unique(ids))} finite+inside {ok} "
                  f"kinetic energy {ke:.4e} front x {front:.1f} max |v| {float(np.abs(vel).max()):.3f}", flush=True)
            assert ok and c["active"] == sc.n and len(np.unique(ids)) == sc.n
    sim.close()
    print("LONG RUN OK")


if __name__ == "__main__":
    main()
