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
        n, w = t[:, 2].astype(np.int64), t[:, 3].astype(np.int64)
        need = (n + 31) // 32
        print(f"after {k} steps: tiles {len(t)} particles {n.sum()} mean N {n.mean():.1f} "
              f"fill {n.sum() / (32.0 * w.sum()):.3f} fill_if_no_column_rule {n.sum() / (32.0 * need.sum()):.3f} "
              f"mean W {w.mean():.2f} mean ceil(N/32) {need.mean():.2f} W>need in {np.mean(w > need):.3f}", flush=True)
        hist = [(lim, float(np.mean(n < lim)), float(n[n < lim].sum() / n.sum())) for lim in (16, 32, 64, 128, 192)]
        print("    tiles / particles below N: " + "  ".join(f"<{lim}: {a:.3f}/{b:.3f}" for lim, a, b in hist), flush=True)
        if k < steps:
            sim.step()
    sim.close()

if __name__ == "__main__":
    main()
