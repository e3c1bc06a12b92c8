"""Window-fill statistics of the tiled path on the bench scene (GPU box only).
fill = particles / (32 * windows): the share of lanes that carry a particle in the windowed kernels."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import fluidpkg

def main():
    pkg = fluidpkg.load()
    scenes = pkg.scenes
    which = sys.argv[1] if len(sys.argv) > 1 else "16M"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    sc = scenes.dam_break_for_gpus(1) if which == "16M" else scenes.dam_break_3d(128, 64, 128)
    sim = pkg.Simulation.new(sc.cfg)
    sim.set_rect(sc.rect_min, sc.rect_max)
    chunk = 1 << 21
    for s in range(0, sc.n, chunk):
        sim.add_particles(sc.records(s, min(chunk, sc.n - s)))
    for k in range(steps + 1):
        sim.neighbour_table()
        t = sim.debug_tiles()
        t = t[t[:, 2] > 0]
        n, w, ovf = (t[:, j].astype(np.int64) for j in (2, 3, 4))    # rows {tile, first, N, W, overflow}
        need = (n + 31) // 32
        quarters = ((n - ovf) // w + 7) // 8 * (w - (n - ovf) % w) + ((n - ovf) // w + 1 + 7) // 8 * ((n - ovf) % w)
        print(f"after {k} steps: tiles {len(t)} particles {n.sum()} mean N {n.mean():.1f} "
              f"fill {(n - ovf).sum() / (32.0 * w.sum()):.3f} overflow {ovf.sum() / n.sum():.4f} "
              f"lanes per active quarter warp {(n - ovf).sum() / (8.0 * quarters.sum()):.3f} "
              f"mean W {w.mean():.2f} mean ceil(N/32) {need.mean():.2f}", flush=True)
        if k < steps:
            sim.step()
    sim.close()

if __name__ == "__main__":
    main()
