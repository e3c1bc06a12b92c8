"""Generates tests/golden/*.npz from the CPU oracle.  These are SELF-goldens: a regression guard
against accidental changes to oracle/oracle.cpp, not reference-derived vectors (the reference
has none and cannot be compiled here — DESIGN.md).  Run: python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import fluidpkg  # noqa: E402
from oracle import oracle  # noqa: E402

scenes = fluidpkg.load().scenes
out = Path(__file__).parent

for name, sc in (("oracle_3d_default_256_s31", scenes.default_3d(256)),
                 ("oracle_2d_default_256_s31", scenes.default_2d(256))):
    sim = oracle.OracleSim(sc.cfg)
    rec0 = sc.records()
    sim.add_particles(rec0)
    sim.set_rect(sc.rect_min, sc.rect_max)
    # taps after the first substep's p2g_2, then the full step
    for ph in range(5):
        sim.phase(ph)
        if ph == 2:
            taps = sim.read(which=1, debug=True)
    sim.substeps(sc.cfg["iterations"] - 1)
    rec, ids = sim.read()
    o = np.argsort(ids)
    t = np.argsort(taps["ids"])
    np.savez_compressed(out / f"{name}.npz", records0=rec0, records=rec[o],
                        density1=taps["density"][t], pressure1=taps["pressure"][t],
                        cell0=taps["cell"][t], key0=taps["key"][t])
    print(name, rec.shape)
