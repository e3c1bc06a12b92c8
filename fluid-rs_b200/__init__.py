"""fluid_rs_b200 — host-side mirror of fluid-rs's `Simulation` API over the C ABI.

The product is `csrc/libfluid_b200.so` (hand-written sm_100a CUDA behind
`include/fluid_b200.h`).  This module is the thin host layer a Python caller uses; it mirrors
the reference's names and argument meaning (src/3d_multi.rs:63-134,383-387):

    Simulation.new(config) / set_rect(min, max) / add_particle(p) / step(mouse_pos) /
    iter_particle() / config.dt / debug_elapseds

There is NO CPU fallback: loading fails loudly if the CUDA library is missing, and
`Simulation.new` raises if no CUDA device is present.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import scenes  # noqa: F401  (re-export)
from . import slab  # noqa: F401  (z-slab multi-GPU driver)

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libfluid_b200.so"
HEADER_PATH = _HERE.parent / "include" / "fluid_b200.h"

FLUID_OK = 0
STATUS_NAMES = {
    0: "FLUID_OK", 1: "FLUID_ERR_INVALID_ARG", 2: "FLUID_ERR_CUDA", 3: "FLUID_ERR_NO_DEVICE",
    4: "FLUID_ERR_OUT_OF_MEMORY", 5: "FLUID_ERR_STATE", 6: "FLUID_ERR_TOO_SMALL",
}
PHASE_LABELS = ("clear", "p2g 1", "p2g 2", "update", "g2p")


class FluidError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class FluidConfig(C.Structure):
    """`struct fluid_config` == the reference's `struct Config` (3d:3-15)."""
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


_lib = None

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_vpp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); also the list the symbol test checks against the header
SIGNATURES = {
    "fluid_abi_version": (C.c_int32, []),
    "fluid_last_error": (C.c_char_p, []),
    "fluid_config_default": (C.c_int, [C.c_int32, C.POINTER(FluidConfig)]),
    "fluid_create": (C.c_int, [C.POINTER(FluidConfig), C.c_int32, _vpp]),
    "fluid_destroy": (C.c_int, [C.c_void_p]),
    "fluid_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fluid_synchronize": (C.c_int, [C.c_void_p]),
    "fluid_set_rect": (C.c_int, [C.c_void_p, _fp, _fp]),
    "fluid_get_rects": (C.c_int, [C.c_void_p, _ip, _ip, _ip, _ip, _ip, _ip]),
    "fluid_add_particles": (C.c_int, [C.c_void_p, _fp, _ip, C.c_int64]),
    "fluid_add_particles_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "fluid_clear_particles": (C.c_int, [C.c_void_p]),
    "fluid_step": (C.c_int, [C.c_void_p, _fp]),
    "fluid_substeps": (C.c_int, [C.c_void_p, C.c_int32, _fp]),
    "fluid_set_deterministic": (C.c_int, [C.c_void_p, C.c_int32]),
    "fluid_set_resident_max": (C.c_int, [C.c_void_p, C.c_int64]),
    "fluid_set_sparse": (C.c_int, [C.c_void_p, C.c_int64]),
    "fluid_memory_stats": (C.c_int, [C.c_void_p, _i64p]),
    "fluid_particle_count": (C.c_int, [C.c_void_p, _i64p]),
    "fluid_particle_counts": (C.c_int, [C.c_void_p, _i64p]),
    "fluid_slot_count": (C.c_int, [C.c_void_p, _i64p]),
    "fluid_read_particles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _i64p]),
    "fluid_get_dt": (C.c_int, [C.c_void_p, _fp]),
    "fluid_get_phase_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "fluid_phase_label": (C.c_char_p, [C.c_int32]),
    "fluid_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "fluid_profile_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), _i64p]),
    "fluid_debug_substep": (C.c_int, [C.c_void_p, _fp, C.c_int64, _ip, _ip, _ip, _fp, _fp, _i64p]),
    "fluid_debug_neighbour_table": (C.c_int, [C.c_void_p, C.c_int64, _ip, _ip, _i64p]),
    "fluid_render_frame": (C.c_int, [C.c_void_p, _fp, C.c_int32, C.c_int32, _ip]),
    "fluid_frame_char": (C.c_char, [C.c_int32]),
    "fluid_debug_tiles": (C.c_int, [C.c_void_p, C.c_int64, _ip, _i64p]),
    "fluid_debug_windows": (C.c_int, [C.c_void_p, C.c_int64, _ip, _i64p]),
    "fluid_read_grid": (C.c_int, [C.c_void_p, _fp, C.c_int64, _i64p]),
    "fluid_launch_count": (C.c_int, [C.c_void_p, _i64p]),
    "fluid_slab_set": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "fluid_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "fluid_slab_planes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, _vpp, _vpp, _i64p]),
    "fluid_slab_phase": (C.c_int, [C.c_void_p, C.c_int32, _fp]),
    "fluid_slab_accumulate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32]),
    "fluid_slab_migrants": (C.c_int, [C.c_void_p, _vpp, _i64p, _vpp, _i64p]),
    "fluid_slab_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "fluid_slab_migrants_begin": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "fluid_slab_migrants_end": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fluid_slab_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "fluid_slab_peer_barrier": (C.c_int, [C.c_void_p]),
    "fluid_slab_append_received": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64]),
    "fluid_slab_ipc_import": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
}


def lib():
    """Load csrc/libfluid_b200.so.  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is None:
        import os
        path = Path(os.environ.get("FLUID_B200_LIB", str(LIB_PATH)))   # A/B builds of the same CUDA library
        if not path.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This engine has no CPU or eager fallback.")
        L = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(status: int):
    if status != FLUID_OK:
        raise FluidError(status, lib().fluid_last_error().decode())


def config_from_dict(d: dict) -> FluidConfig:
    c = FluidConfig()
    _check(lib().fluid_config_default(int(d["dim"]), C.byref(c)))
    for k, v in d.items():
        if k in ("gravity", "clip_min", "clip_max"):
            arr = getattr(c, k)
            for i, x in enumerate(v):
                arr[i] = float(x)
        else:
            setattr(c, k, v)
    return c


def default_config(dim: int) -> FluidConfig:
    """`Config::default()` (3d:17-33 / 2d:17-33)."""
    c = FluidConfig()
    _check(lib().fluid_config_default(dim, C.byref(c)))
    return c


def _as_fp(a):
    return a.ctypes.data_as(_fp)


class Simulation:
    """Mirror of `struct Simulation` (3d:50-134,383-387) with device-resident state."""

    def __init__(self, config, device: int = 0):
        if isinstance(config, dict):
            config = config_from_dict(config)
        self.config = config
        self.dim = int(config.dim)
        self.rec_floats = 2 * self.dim + self.dim * self.dim + 1
        h = C.c_void_p()
        _check(lib().fluid_create(C.byref(config), device, C.byref(h)))
        self._h = h

    # `Simulation::new(config)` (3d:64)
    @classmethod
    def new(cls, config, device: int = 0) -> "Simulation":
        return cls(config, device)

    def close(self):
        if getattr(self, "_h", None):
            lib().fluid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- reference API -----------------------------------------------------------------
    def set_rect(self, mn, mx):
        mn = np.ascontiguousarray(mn, dtype=np.float32)
        mx = np.ascontiguousarray(mx, dtype=np.float32)
        if mn.size < self.dim or mx.size < self.dim:
            raise ValueError("set_rect needs dim floats for min and max")
        _check(lib().fluid_set_rect(self._h, _as_fp(mn), _as_fp(mx)))

    def add_particle(self, pos, vel=None, affine_momentum=None, mass: float = 1.0, id=None):
        """One `Particle` (3d:35-41, 104-108)."""
        rec = np.zeros(self.rec_floats, dtype=np.float32)
        d = self.dim
        rec[:d] = pos
        if vel is not None:
            rec[d:2 * d] = vel
        if affine_momentum is not None:
            rec[2 * d:2 * d + d * d] = np.asarray(affine_momentum, dtype=np.float32).reshape(-1)
        rec[-1] = mass
        self.add_particles(rec[None, :], None if id is None else np.array([id], dtype=np.int32))

    def add_particles(self, records, ids=None):
        records = np.ascontiguousarray(records, dtype=np.float32).reshape(-1, self.rec_floats)
        idp = None
        if ids is not None:
            ids = np.ascontiguousarray(ids, dtype=np.int32)
            if ids.shape[0] != records.shape[0]:
                raise ValueError("ids and records differ in length")
            idp = ids.ctypes.data_as(_ip)
        _check(lib().fluid_add_particles(self._h, _as_fp(records), idp, records.shape[0]))

    def add_particles_device(self, d_records_ptr: int, n: int, d_ids_ptr: int = 0):
        _check(lib().fluid_add_particles_device(self._h, C.c_void_p(d_records_ptr),
                                                C.c_void_p(d_ids_ptr) if d_ids_ptr else None, n))

    def add_particles_pinned(self, host_ptr: int, n: int):
        """Records in caller-owned (pinned) host memory, by address."""
        _check(lib().fluid_add_particles(self._h, C.cast(C.c_void_p(host_ptr), _fp), None, n))

    def clear_particles(self):
        _check(lib().fluid_clear_particles(self._h))

    def _mouse(self, mouse_pos):
        if mouse_pos is None:
            return None, None
        m = np.ascontiguousarray(mouse_pos, dtype=np.float32)
        return m, _as_fp(m)

    def step(self, mouse_pos=None):
        """`step(&mut self, mouse_pos: &Option<Vec2>)` (3d:110-134)."""
        keep, p = self._mouse(mouse_pos)
        _check(lib().fluid_step(self._h, p))

    def substeps(self, n: int, mouse_pos=None):
        keep, p = self._mouse(mouse_pos)
        _check(lib().fluid_substeps(self._h, int(n), p))

    def set_deterministic(self, on: bool = True):
        """Order-independent (64-bit fixed-point) node sums: bit-for-bit reproducible runs."""
        _check(lib().fluid_set_deterministic(self._h, 1 if on else 0))

    def set_resident_max(self, max_particles: int):
        """Largest particle count whose step() runs as one cooperative launch (0 = never)."""
        _check(lib().fluid_set_resident_max(self._h, int(max_particles)))

    def set_sparse(self, max_blocks: int):
        """Block-sparse node storage with a pool of `max_blocks` 8x8x4 blocks (from the next set_rect on)."""
        _check(lib().fluid_set_sparse(self._h, int(max_blocks)))

    def memory_stats(self) -> dict:
        o = (C.c_int64 * 6)()
        _check(lib().fluid_memory_stats(self._h, o))
        return dict(node_bytes=o[0], dense_node_bytes=o[1], pool_blocks=o[2], blocks_in_use=o[3], dense_nodes=o[4],
                    pool_exhausted=bool(o[5]))

    def iter_particle(self):
        """`iter_particle` (3d:383-387): yields (id, record) for every a_rect particle."""
        rec, ids = self.read_particles()
        for i in range(rec.shape[0]):
            yield int(ids[i]), rec[i]

    @property
    def dt(self) -> float:
        out = C.c_float()
        _check(lib().fluid_get_dt(self._h, C.byref(out)))
        return out.value

    @property
    def debug_elapseds(self):
        """[(label, seconds)] of the last substep, the reference's five labels (3d:112-132)."""
        sec = (C.c_double * 5)()
        _check(lib().fluid_get_phase_times(self._h, sec, None))
        return list(zip(PHASE_LABELS, list(sec)))

    def phase_times(self):
        sec = (C.c_double * 5)()
        srt = C.c_double()
        _check(lib().fluid_get_phase_times(self._h, sec, C.byref(srt)))
        d = dict(zip(PHASE_LABELS, list(sec)))
        d["sort"] = srt.value
        return d

    def profile(self, on: bool):
        _check(lib().fluid_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> dict:
        """Seconds per phase summed over every substep since profile(True), and that count."""
        sec = (C.c_double * 6)()
        n = C.c_int64()
        _check(lib().fluid_profile_read(self._h, sec, C.byref(n)))
        d = dict(zip(("sort",) + PHASE_LABELS, list(sec)))
        d["substeps"] = n.value
        return d

    # ---- headless frame: `draw` without a terminal (3d:461-500) --------------------------------
    def frame_counts(self, viewport=(64.0, 64.0), cols: int = 80, rows: int = 40) -> np.ndarray:
        vp = np.ascontiguousarray(viewport, dtype=np.float32)
        out = np.zeros((rows, cols), dtype=np.int32)
        _check(lib().fluid_render_frame(self._h, _as_fp(vp), cols, rows, out.ctypes.data_as(_ip)))
        return out

    def frame_text(self, viewport=(64.0, 64.0), cols: int = 80, rows: int = 40) -> str:
        ramp = " .-=*%$#"
        c = np.clip(self.frame_counts(viewport, cols, rows), 0, 7)
        return "\n".join("".join(ramp[v] for v in row) for row in c)

    # ---- plumbing ----------------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None):
        _check(lib().fluid_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self):
        _check(lib().fluid_synchronize(self._h))

    def particle_count(self) -> int:
        n = C.c_int64()
        _check(lib().fluid_particle_count(self._h, C.byref(n)))
        return n.value

    def particle_counts(self) -> dict:
        c = (C.c_int64 * 4)()
        _check(lib().fluid_particle_counts(self._h, c))
        return dict(active=c[0], frozen=c[1], outside=c[2], dropped=c[3])

    def slot_count(self) -> int:
        """Upper bound on the number of records read_particles returns; no device work."""
        n = C.c_int64()
        _check(lib().fluid_slot_count(self._h, C.byref(n)))
        return n.value

    def read_particles(self, sort_by_id: bool = False):
        n = self.slot_count()      # not particle_count(): that re-runs the neighbour search from scratch
        rec = np.empty((max(n, 1), self.rec_floats), dtype=np.float32)
        ids = np.empty(max(n, 1), dtype=np.int32)
        w = C.c_int64()
        _check(lib().fluid_read_particles(self._h, rec.ctypes.data_as(C.c_void_p),
                                          ids.ctypes.data_as(C.c_void_p), n, C.byref(w)))
        rec, ids = rec[: w.value], ids[: w.value]
        if sort_by_id:
            o = np.argsort(ids, kind="stable")
            rec, ids = rec[o], ids[o]
        return rec, ids

    def read_particles_into(self, host_ptr: int, capacity: int, ids_ptr: int = 0) -> int:
        w = C.c_int64()
        _check(lib().fluid_read_particles(self._h, C.c_void_p(host_ptr),
                                          C.c_void_p(ids_ptr) if ids_ptr else None, capacity,
                                          C.byref(w)))
        return w.value

    def rects(self) -> dict:
        out = [np.zeros(3, dtype=np.int32) for _ in range(6)]
        _check(lib().fluid_get_rects(self._h, *[a.ctypes.data_as(_ip) for a in out]))
        names = ("a_lo", "a_hi", "p_lo", "p_hi", "origin", "size")
        return {k: v[: self.dim].copy() for k, v in zip(names, out)}

    def read_grid(self) -> np.ndarray:
        n = C.c_int64()
        _check(lib().fluid_read_grid(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, self.dim + 1), dtype=np.float32)
        _check(lib().fluid_read_grid(self._h, _as_fp(out), n.value, C.byref(n)))
        return out

    def launch_count(self) -> int:
        n = C.c_int64()
        _check(lib().fluid_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- parity taps -------------------------------------------------------------------
    def debug_substep(self, mouse_pos=None) -> dict:
        c = self.particle_counts()
        cap = c["active"] + c["frozen"]
        d = self.dim
        ids = np.empty(max(cap, 1), dtype=np.int32)
        cell = np.empty((max(cap, 1), d), dtype=np.int32)
        key = np.empty((max(cap, 1), d), dtype=np.int32)
        den = np.empty(max(cap, 1), dtype=np.float32)
        prs = np.empty(max(cap, 1), dtype=np.float32)
        w = C.c_int64()
        keep, p = self._mouse(mouse_pos)
        _check(lib().fluid_debug_substep(self._h, p, cap, ids.ctypes.data_as(_ip),
                                         cell.ctypes.data_as(_ip), key.ctypes.data_as(_ip),
                                         _as_fp(den), _as_fp(prs), C.byref(w)))
        n = w.value
        return dict(ids=ids[:n], cell=cell[:n], key=key[:n], density=den[:n], pressure=prs[:n])

    def debug_tiles(self):
        """Active tile list {tile, first slot, N, W} of the current sort (after neighbour_table())."""
        n = C.c_int64()
        _check(lib().fluid_debug_tiles(self._h, 0, None, C.byref(n)))
        out = np.empty((max(n.value, 1), 4), dtype=np.int32)
        _check(lib().fluid_debug_tiles(self._h, n.value, out.ctypes.data_as(_ip), C.byref(n)))
        return out[: n.value]

    def debug_windows(self):
        """Per sorted slot: window * 32 + lane of the tile kernels' walk (after neighbour_table())."""
        cap = self.slot_count()    # (no count query here: that would re-run the neighbour search)
        out = np.empty(max(cap, 1), dtype=np.int32)
        w = C.c_int64()
        _check(lib().fluid_debug_windows(self._h, cap, out.ctypes.data_as(_ip), C.byref(w)))
        return out[: w.value]

    def neighbour_table(self):
        c = self.particle_counts()
        cap = c["active"] + c["frozen"]
        ids = np.empty(max(cap, 1), dtype=np.int32)
        idx = np.empty(max(cap, 1), dtype=np.int32)
        w = C.c_int64()
        _check(lib().fluid_debug_neighbour_table(self._h, cap, ids.ctypes.data_as(_ip),
                                                 idx.ctypes.data_as(_ip), C.byref(w)))
        return ids[: w.value], idx[: w.value]
