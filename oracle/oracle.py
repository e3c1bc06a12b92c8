"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle.cpp header): a C++ restatement of fluid-rs's `Simulation`
(src/3d_multi.rs:50-408, src/2d_multi.rs:50-385), checked only against closed forms and
conservation laws because the reference has no tests and cannot be compiled here.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"


class OrcConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("dt", C.c_float),
        ("iterations", C.c_int32),
        ("grid_res", C.c_int32),
        ("gravity", C.c_float * 3),
        ("rest_density", C.c_float),
        ("dynamic_viscosity", C.c_float),
        ("eos_stiffness", C.c_float),
        ("eos_power", C.c_float),
        ("mouse_radius", C.c_float),
        ("clip_min", C.c_float * 3),
        ("clip_max", C.c_float * 3),
        ("boundary_damp_dist", C.c_float),
        ("pressure_clamp", C.c_float),
    ]


def build(force: bool = False) -> Path:
    """Compile liboracle.so with the committed recipe (oracle/Makefile)."""
    src = _HERE / "oracle.cpp"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        L = C.CDLL(str(_LIB_PATH))
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        L.orc_config_default.argtypes = [C.c_int32, C.POINTER(OrcConfig)]
        L.orc_create.argtypes = [C.POINTER(OrcConfig)]
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_rect.argtypes = [C.c_void_p, fp, fp]
        L.orc_add_particles.argtypes = [C.c_void_p, fp, ip, C.c_int64]
        L.orc_substeps.argtypes = [C.c_void_p, C.c_int32, fp]
        L.orc_step.argtypes = [C.c_void_p, fp]
        L.orc_phase.argtypes = [C.c_void_p, C.c_int32, fp]
        L.orc_count.argtypes = [C.c_void_p, C.c_int32]
        L.orc_count.restype = C.c_int64
        L.orc_read.argtypes = [C.c_void_p, C.c_int32, fp, ip, fp, fp, ip, ip]
        L.orc_read.restype = C.c_int64
        L.orc_rects.argtypes = [C.c_void_p, ip, ip, ip, ip, ip, ip]
        L.orc_read_grid.argtypes = [C.c_void_p, fp]
        L.orc_read_grid.restype = C.c_int64
        L.orc_touched_count.argtypes = [C.c_void_p]
        L.orc_touched_count.restype = C.c_int64
        L.orc_phase_seconds.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.orc_key_from_pos.argtypes = [fp, C.c_int64, C.c_int32, C.c_int32, ip, ip]
        L.orc_quadratic_weights.argtypes = [C.c_float, fp]
        _lib = L
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


PHASES = ("clear", "p2g 1", "p2g 2", "update", "g2p")


def default_config(dim: int) -> OrcConfig:
    c = OrcConfig()
    lib().orc_config_default(dim, C.byref(c))
    return c


def config_from_dict(d: dict) -> OrcConfig:
    """Build an OrcConfig from the plain-dict form shared with the product binding."""
    c = default_config(int(d["dim"]))
    for k, v in d.items():
        if k in ("gravity", "clip_min", "clip_max"):
            arr = getattr(c, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(c, k, v)
    return c


class OracleSim:
    """Mirror of the reference `Simulation` API over the C++ restatement."""

    def __init__(self, cfg):
        if isinstance(cfg, dict):
            cfg = config_from_dict(cfg)
        self.cfg = cfg
        self.dim = int(cfg.dim)
        self.rec_floats = 2 * self.dim + self.dim * self.dim + 1
        self._h = lib().orc_create(C.byref(cfg))
        if not self._h:
            raise ValueError("orc_create failed")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_rect(self, mn, mx):
        mn = np.ascontiguousarray(mn, dtype=np.float32)
        mx = np.ascontiguousarray(mx, dtype=np.float32)
        lib().orc_set_rect(self._h, _fp(mn), _fp(mx))

    def add_particles(self, records, ids=None):
        records = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, self.rec_floats)
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
        lib().orc_add_particles(self._h, _fp(records), _ip(ids), records.shape[0])

    def _mouse(self, mouse):
        if mouse is None:
            return None, None
        m = np.ascontiguousarray(mouse, dtype=np.float32)
        return m, _fp(m)

    def step(self, mouse=None):
        keep, p = self._mouse(mouse)
        lib().orc_step(self._h, p)

    def substeps(self, n, mouse=None):
        keep, p = self._mouse(mouse)
        lib().orc_substeps(self._h, int(n), p)

    def phase(self, ph, mouse=None):
        keep, p = self._mouse(mouse)
        lib().orc_phase(self._h, int(ph), p)

    def count(self, which=0):
        return int(lib().orc_count(self._h, which))

    def read(self, which=0, debug=False):
        """which=0: iter_particle order (a_rect); which=1: every p_rect particle."""
        n = self.count(which)
        rec = np.empty((n, self.rec_floats), dtype=np.float32)
        ids = np.empty(n, dtype=np.int32)
        if debug:
            den = np.empty(n, dtype=np.float32)
            prs = np.empty(n, dtype=np.float32)
            cell = np.empty((n, self.dim), dtype=np.int32)
            key = np.empty((n, self.dim), dtype=np.int32)
            lib().orc_read(self._h, which, _fp(rec), _ip(ids), _fp(den), _fp(prs), _ip(cell),
                           _ip(key))
            return dict(records=rec, ids=ids, density=den, pressure=prs, cell=cell, key=key)
        lib().orc_read(self._h, which, _fp(rec), _ip(ids), None, None, None, None)
        return rec, ids

    def rects(self):
        out = [np.zeros(3, dtype=np.int32) for _ in range(6)]
        lib().orc_rects(self._h, *[_ip(a) for a in out])
        names = ("a_lo", "a_hi", "p_lo", "p_hi", "origin", "size")
        return {k: v[: self.dim].copy() for k, v in zip(names, out)}

    def read_grid(self):
        n = int(lib().orc_read_grid(self._h, None))
        out = np.empty((n, self.dim + 1), dtype=np.float32)
        lib().orc_read_grid(self._h, _fp(out))
        return out

    def touched_count(self):
        return int(lib().orc_touched_count(self._h))

    def phase_seconds(self):
        out = (C.c_double * 5)()
        lib().orc_phase_seconds(self._h, out)
        return dict(zip(PHASES, list(out)))


def key_from_pos(pos, grid_res):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    dim = pos.shape[-1]
    key = np.empty(pos.shape, dtype=np.int32)
    cell = np.empty(pos.shape, dtype=np.int32)
    lib().orc_key_from_pos(_fp(pos), pos.size // dim, dim, int(grid_res), _ip(key), _ip(cell))
    return key, cell


def quadratic_weights(c):
    out = np.empty(3, dtype=np.float32)
    lib().orc_quadratic_weights(C.c_float(c), _fp(out))
    return out
