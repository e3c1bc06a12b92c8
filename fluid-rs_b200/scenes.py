"""Synthetic scenes for the parity tests and the bench.

The reference seeds its scene from `rand::rng()` (OS-seeded, 3d:524 / 2d:501), so there is no
reproducible reference scene.  Positions here come from a counter-based generator
(splitmix64(seed xor index) -> 24-bit uniform in [0,1)), seed 20260101, so any index range can
be regenerated on any rank and the same buffer is fed to the oracle and to the GPU engine.
Velocities and the affine matrix start at zero, mass = 1 (3d:532-534).
"""
from __future__ import annotations

import numpy as np

SEED = 20260101

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def uniform01(seed: int, start: int, count: int) -> np.ndarray:
    """count float32 values in [0,1): element k uses counter start+k."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    z = splitmix64(idx ^ np.uint64(seed))
    return ((z >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / 16777216.0)


def rec_floats(dim: int) -> int:
    return 2 * dim + dim * dim + 1


def default_config(dim: int) -> dict:
    """`Config::default()` of the 2D / 3D binary (2d:17-33 / 3d:17-33) as a plain dict."""
    return dict(
        dim=dim,
        dt=0.032 if dim == 2 else 0.066,
        iterations=int(1.0 / 0.032),
        grid_res=32 if dim == 2 else 16,
        gravity=[0.0, 0.3, 0.0],
        rest_density=4.0 if dim == 2 else 1.0,
        dynamic_viscosity=0.1,
        eos_stiffness=10.0,
        eos_power=4.0,
        mouse_radius=10.0,
        clip_min=[0.0, 0.0, 0.0],
        clip_max=[64.0, 64.0, 64.0],
        boundary_damp_dist=3.0,
        pressure_clamp=-0.0 if dim == 2 else -0.1,
    )


def box_records(dim: int, lo, hi, n: int, seed: int = SEED, start: int = 0,
                count: int | None = None) -> np.ndarray:
    """Records [start, start+count) of an n-particle uniform fill of the box [lo, hi)."""
    count = n - start if count is None else count
    u = uniform01(seed, start * dim, count * dim).reshape(count, dim)
    lo = np.asarray(lo, dtype=np.float32)[:dim]
    hi = np.asarray(hi, dtype=np.float32)[:dim]
    rec = np.zeros((count, rec_floats(dim)), dtype=np.float32)
    rec[:, :dim] = lo + u * (hi - lo)
    rec[:, -1] = 1.0
    return rec


class Scene:
    """A named workload: config dict, active rect, particle count and a record generator."""

    def __init__(self, name, cfg, rect_min, rect_max, n, fill_lo, fill_hi, seed=SEED):
        self.name = name
        self.cfg = cfg
        self.dim = cfg["dim"]
        self.rect_min = np.asarray(rect_min, dtype=np.float32)
        self.rect_max = np.asarray(rect_max, dtype=np.float32)
        self.n = int(n)
        self.fill_lo = np.asarray(fill_lo, dtype=np.float32)
        self.fill_hi = np.asarray(fill_hi, dtype=np.float32)
        self.seed = seed

    def records(self, start: int = 0, count: int | None = None) -> np.ndarray:
        return box_records(self.dim, self.fill_lo, self.fill_hi, self.n, self.seed, start, count)

    def describe(self) -> dict:
        return dict(workload=self.name, dim=self.dim, particles=self.n,
                    box=[float(x) for x in self.rect_max], fill_lo=[float(x) for x in self.fill_lo],
                    fill_hi=[float(x) for x in self.fill_hi],
                    gravity=[float(g) for g in self.cfg["gravity"][: self.dim]])


def default_2d(n: int = 4096) -> Scene:
    """BASELINE config 1: the 2D binary's scene (2d:498-513)."""
    cfg = default_config(2)
    return Scene("2d_default", cfg, [0, 0], [64, 64], n, [16, 16], [48, 48])


def default_3d(n: int = 4096) -> Scene:
    """BASELINE config 2: the 3D binary's scene (3d:521-537)."""
    cfg = default_config(3)
    return Scene("3d_default", cfg, [0, 0, 0], [64, 64, 64], n, [16, 16, 16], [32, 32, 32])


def dam_break_3d(cx: int, cy: int, cz: int, name: str | None = None) -> Scene:
    """3D dam break: a cx*cy*cz-cell column, one particle per cell on average, resting on the
    +y wall (gravity is +y, 3d:23) in the corner x,z = 3.  Box = (3cx, cy+32, cz+6).  Gravity is
    rescaled to 4.8/cy so the hydrostatic load rho0*g*h equals the default scene's
    (h=16, g=0.3); with g=0.3 a deep column is outside the reference's stable range
    (SURVEY.md section 8d)."""
    cfg = default_config(3)
    bx, by, bz = 3 * cx, cy + 32, cz + 6
    cfg["clip_max"] = [float(bx), float(by), float(bz)]
    cfg["gravity"] = [0.0, 4.8 / cy, 0.0]
    n = cx * cy * cz
    lo = [3.0, by - 3.0 - cy, 3.0]
    hi = [3.0 + cx, by - 3.0, 3.0 + cz]
    return Scene(name or f"dam_break_3d_{cx}x{cy}x{cz}", cfg, [0, 0, 0], [bx, by, bz], n, lo, hi)


def dam_break_2d(cells: int, name: str | None = None) -> Scene:
    """2D dam break: a cells x cells column at the 2D scene's density (4 particles per cell, 2d:24,502-511) resting
    on the +y wall of a 3x wider box; gravity rescaled to the default scene's hydrostatic load (32 cells at 0.3)."""
    cfg = default_config(2)
    bx, by = 3.0 * cells, cells + 32.0
    cfg["clip_max"] = [bx, by, 64.0]
    cfg["gravity"] = [0.0, 0.3 * 32.0 / cells, 0.0]
    n = cells * cells * 4
    return Scene(name or f"dam_break_2d_{cells}x{cells}", cfg, [0, 0], [bx, by], n, [3.0, by - 3.0 - cells], [3.0 + cells, by - 3.0])


def dam_break_2d_1m() -> Scene:   # the 2D tiled path at scale (not a BASELINE config): 2^20 particles
    return dam_break_2d(512, "dam_break_2d_1M")


def dam_break_1m() -> Scene:      # BASELINE config 3
    return dam_break_3d(128, 64, 128, "dam_break_3d_1M")


def dam_break_16m() -> Scene:     # BASELINE config 4
    return dam_break_3d(256, 256, 256, "dam_break_3d_16M")


def dam_break_128m() -> Scene:    # BASELINE config 5
    return dam_break_3d(512, 512, 512, "dam_break_3d_128M")


def dam_break_for_gpus(n_gpus: int) -> Scene:
    """Weak-scaling family, 2^24 particles per GPU: 1 GPU = config 4, 8 GPUs = config 5."""
    dims = {1: (256, 256, 256), 2: (256, 256, 512), 4: (512, 256, 512), 8: (512, 512, 512)}
    if n_gpus not in dims:
        raise ValueError("n_gpus must be 1, 2, 4 or 8")
    cx, cy, cz = dims[n_gpus]
    names = {1: "dam_break_3d_16M", 8: "dam_break_3d_128M"}
    return dam_break_3d(cx, cy, cz, names.get(n_gpus))
