#!/usr/bin/env python
"""bench.py — particle-updates/sec of the step hot path on synthetic 3D dam-break scenes.

Contract (driver):  python bench.py --gpus N --steps K --warmup W  [--impl reference]
One JSON line on stdout (rank 0).

  step      one `step()` call = config.iterations (31) substeps of
            clear -> p2g 1 -> p2g 2 -> update -> g2p over every particle (3d:110-134)
  value     particle-updates/s with the state resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the public C ABI with HOST buffers: every step uploads the
            particle records from pinned host memory, runs step(), reads every record back
  roofline  dominant kernel: algorithmic bytes per launch / CUDA-event duration, over the
            timed region, against the measured HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the CPU oracle (a C++ restatement of the reference, "port", 1 thread because the
            reference is single-threaded) on a bounded sample of the same workload

Workload: N=1 -> BASELINE config 4 (3D dam break, 2^24 particles); N GPUs -> 2^24 particles per
GPU in z-slabs (N=8 is BASELINE config 5, 2^27 particles): weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "particle-updates/sec (3D WCSPH-reference step: MLS-MPM substeps)"
UNIT = "particle-updates/s"

# Algorithmic bytes per particle-substep (3D, A/N = 1).  The whole step uses SURVEY.md section 8(d):
# 188 B of particle streams + 92 B of node traffic = 280 B.  Per kernel the same stream table is
# re-cut along OUR kernel boundaries: the momentum scatter of p2g_1 runs inside
# the "p2g 2" kernel, so that kernel reads the particle once (64 B), the node mass (4 B) and does the
# node read-modify-write (32 B) = 100 B; "p2g 1" (mass only) reads pos+mass (16 B) and RMWs the node
# mass (8 B); clear writes the node record and the node mass (20 B); update+g2p is SURVEY's 88 B.
ALG_BYTES = {"clear": 20.0, "p2g 1": 24.0, "p2g 2": 100.0, "update": 0.0, "g2p": 88.0}
ALG_BYTES_STEP = 280.0
KERNEL_OF_PHASE = {"clear": "k_clear_tiles", "p2g 1": "k_mass_tiled", "p2g 2": "k_p2g_tiled", "g2p": "k_g2p_tiled"}
NCU_CAPTURES = ["profiles/r02_ncu_full_16M.csv", "profiles/r01_ncu_full_16M_v30.csv"]   # newest first: ncu --set full, config 4


def ncu_capture():
    for c in NCU_CAPTURES:
        if (ROOT / c).exists():
            return c
    return None


def ncu_row(kernel: str):
    """Metrics of `kernel` from the committed ncu --set full summary (profiles/summarize.py full): a dict
    {column name without unit: value in base units}, or None."""
    import csv
    cap = ncu_capture()
    if cap is None:
        return None
    rows = list(csv.reader((ROOT / cap).read_text().splitlines()))
    hdr = rows[0]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    for r in rows[1:]:
        if kernel in r[0]:
            out = {}
            for h, v in zip(hdr[1:], r[1:]):
                name, unit = h[:h.index(" [")], h[h.index("[") + 1:-1]
                try:
                    out[name] = float(v) * scale.get(unit, 1.0)
                except ValueError:
                    pass
            return out
    return None


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, or None."""
    m = ncu_row(kernel)
    if not m or "dram__bytes_read.sum" not in m:
        return None
    return m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]


def ncu_limiter(kernel: str):
    """Which unit the capture shows closest to its peak for `kernel` (derived, not a literal)."""
    m = ncu_row(kernel)
    if not m:
        return None
    cand = {"shared-memory / L1 data pipe (l1tex__data_pipe_lsu_wavefronts)":
                m.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "HBM (gpu__dram_throughput)": m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "instruction issue (smsp__issue_active)": m.get("smsp__issue_active.avg.pct_of_peak_sustained_active")}
    cand = {k: v for k, v in cand.items() if v is not None}
    if not cand:
        return None
    top = max(cand, key=cand.get)
    out = {"unit": top, "pct_of_peak": cand[top], "all_pct": cand, "source": ncu_capture()}
    wf, bc = m.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), m.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    if wf:
        out["smem_wavefronts_per_launch"] = wf
        out["smem_bank_conflict_frac"] = (bc or 0.0) / wf
    return out


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_info():
    model = "unknown"
    try:
        for line in Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return model, os.cpu_count()


def oracle_sample(scenes, seconds_target: float = 15.0, scene=None):
    """Time the CPU oracle on a scaled-down dam break of the same construction (1 thread), or on `scene`."""
    from oracle import oracle
    oracle.build()
    sc = scene or scenes.dam_break_3d(64, 64, 64)  # 262,144 particles, same column shape as 256^3
    sim = oracle.OracleSim(sc.cfg)
    sim.add_particles(sc.records())
    sim.set_rect(sc.rect_min, sc.rect_max)
    sim.substeps(2)                                # warm: first touch of the grid
    t0 = time.perf_counter()
    sim.substeps(2)
    per = (time.perf_counter() - t0) / 2
    n_sub = max(4, min(62, int(seconds_target / max(per, 1e-6))))
    t0 = time.perf_counter()
    sim.substeps(n_sub)
    dt = time.perf_counter() - t0
    phases = sim.phase_seconds()
    sim.close()
    return sc, n_sub, dt, sc.n * n_sub / dt, phases


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust binary cannot
    be built here (no rustc), so this is the oracle port: same five phases, same order, 1 thread
    (the reference has no parallel loops — SURVEY.md section 0.2).

    Workload = our arm's (BASELINE config 4, the 2^24-particle dam break; for N > 1 one GPU's share of the
    weak-scaling scene is the same 2^24 particles, and a single-threaded CPU rate does not depend on N).
    A full substep of it costs about 20 s on one core, so K steps + W warm-ups at the true size would take
    ten minutes.  One `step` of this arm is therefore a bounded sample: ONE substep over a quarter of the
    column (256 x 256 x 64 cells = 2^22 particles: the true depth and length of the column, a quarter of its
    translation-invariant z extent; `config` says so).  After the timed steps the TRUE 2^24 scene is timed for
    two substeps (`same_config_probe`): that is the same-config CPU rate (lower: the working set is larger).
    --scene runs the named small scene in full (31 substeps per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import fluidpkg
    scenes = fluidpkg.load().scenes
    from oracle import oracle
    oracle.build()
    chunk = 1 << 22

    def build(sc):
        sim = oracle.OracleSim(sc.cfg)
        for s0 in range(0, sc.n, chunk):
            sim.add_particles(sc.records(s0, min(chunk, sc.n - s0)))
        sim.set_rect(sc.rect_min, sc.rect_max)
        return sim

    full = scenes.dam_break_16m()
    if args.scene:
        sc = getattr(scenes, args.scene)()
        sub_per_step = sc.cfg["iterations"]
        sample = f"{sc.name}: all {sc.n} particles x {sub_per_step} substeps per step (the whole scene)"
    else:
        sc = scenes.dam_break_3d(256, 256, 64, "dam_break_3d_16M_quarter_z")
        sub_per_step = 1
        sample = (f"{sc.name}: {sc.n} particles (256 x 256 x 64 cells: a quarter of the 2^24-particle column along z, true "
                  f"depth and length) x 1 substep per step; the true-size rate is in same_config_probe")
    sim = build(sc)
    for _ in range(args.warmup):
        sim.substeps(sub_per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sim.substeps(sub_per_step)
    dt = time.perf_counter() - t0
    phases = sim.phase_seconds()
    sim.close()
    value = sc.n * sub_per_step * args.steps / dt
    probe = None
    if not args.scene:
        sim = build(full)
        sim.substeps(1)                            # first touch of the 1.6 GB grid
        t0 = time.perf_counter()
        sim.substeps(2)
        pdt = time.perf_counter() - t0
        sim.close()
        probe = {"workload": full.name, "particles": full.n, "substeps": 2, "seconds": pdt,
                 "value": full.n * 2 / pdt, "unit": UNIT}
    model, cores = cpu_info()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(sc.describe(), substeps_per_step=sub_per_step,
                       sample_of=(scenes.dam_break_for_gpus(args.gpus).name or full.name) if not args.scene else sc.name,
                       sample_of_particles=scenes.dam_break_for_gpus(args.gpus).n if not args.scene else sc.n,
                       sample_of_substeps_per_step=sc.cfg["iterations"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cpu": model, "host_cores": cores, "phase_seconds_last_substep": phases},
        "same_config_probe": probe,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import fluidpkg
    pkg = fluidpkg.load()
    scenes = pkg.scenes
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this engine has no CPU path (use --impl reference "
                         "for the CPU oracle)")
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    if world > 1:
        return run_slabs(args, pkg, world, rank, local_rank)

    sc = scenes.dam_break_for_gpus(args.gpus) if not args.scene else getattr(scenes, args.scene)()
    iters = sc.cfg["iterations"]
    rf = scenes.rec_floats(sc.dim)
    headline = not args.scene          # BASELINE config 4; --scene lines are the small parity configs (profiles/)

    # host records in pinned memory (also the e2e upload source)
    host = torch.empty((sc.n, rf), dtype=torch.float32, pin_memory=True)
    hnp = host.numpy()
    chunk = 1 << 21
    for s in range(0, sc.n, chunk):
        c = min(chunk, sc.n - s)
        hnp[s:s + c] = sc.records(s, c)
    back = torch.empty((sc.n, rf), dtype=torch.float32, pin_memory=True)

    # a non-default stream: the engine and the timing events share it (NULL would mean "the
    # handle's own stream" to fluid_set_stream)
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    sim = pkg.Simulation.new(sc.cfg, device=local_rank)
    sim.set_stream(stream.cuda_stream)
    if args.sparse_blocks:
        sim.set_sparse(args.sparse_blocks)      # block-sparse node storage (not the headline configuration)
    sim.set_rect(sc.rect_min, sc.rect_max)
    sim.add_particles_pinned(host.data_ptr(), sc.n)
    sim.synchronize()

    # ---- device-resident timing -------------------------------------------------------------
    # The timed region runs the way a caller runs it: no per-phase event brackets, so steady-state substeps replay
    # the engine's CUDA graph (scenes of at most 16,384 particles: one cooperative launch per step()).  The
    # per-kernel durations the roofline needs come from a REPLAY of the same region — same initial state, same
    # warm-up, same K steps — with CUDA-event brackets around every phase (plain launches: a bracket would split
    # the graph; scenes on the resident kernel report its own %globaltimer stamps of the last substep instead).
    resident = sc.n <= 16384

    def region(profiled: bool):
        sim.clear_particles()
        sim.add_particles_pinned(host.data_ptr(), sc.n)
        for _ in range(args.warmup):
            sim.step()
        torch.cuda.synchronize()
        if profiled:
            sim.profile(True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        l0 = sim.launch_count()
        torch.cuda.synchronize()
        ev[0].record(stream)
        for k in range(args.steps):
            sim.step()
            ev[k + 1].record(stream)   # an event record costs nothing on the stream; no sync inside the timed region
        torch.cuda.synchronize()
        total = ev[0].elapsed_time(ev[-1])
        by_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
        pr = None
        if profiled:
            pr = sim.profile_read()
            sim.profile(False)
        return total, by_step, sim.launch_count() - l0, pr

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, ms_by_step, launches, _ = region(False)
    clocks = sampler.stop()
    if resident:
        prof = dict(sim.phase_times(), substeps=1)
        ms_profiled = None
    else:
        ms_profiled, _, _, prof = region(True)
    # occupancy of the last timed substep (from the engine's tile list; outside the timed region)
    tiles_now = sim.debug_tiles()
    tiles_now = tiles_now[tiles_now[:, 2] > 0]
    occupancy = None
    if len(tiles_now):
        n_t, w_t = tiles_now[:, 2].astype(np.int64), tiles_now[:, 3].astype(np.int64)
        occupancy = {"active_tiles": int(len(tiles_now)), "particles_per_tile": float(n_t.mean()),
                     "window_fill": float(n_t.sum() / (32.0 * w_t.sum())),
                     "A_over_N": float(len(tiles_now) * 256 / sc.n),
                     "A_note": "A = nodes of the 8x8x4 blocks of the tiles that hold particles (SURVEY.md 8d fixes A/N = 1 "
                               "for the roofline; the rim the stencils reach adds about a quarter)"}
    counts = sim.particle_counts()
    assert counts["active"] == sc.n, counts
    value = sc.n * iters * args.steps / (ms * 1e-3)
    value_by_step = [sc.n * iters / (m * 1e-3) for m in ms_by_step]

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peak, peak_src = measured_peaks()
    nsub = max(prof["substeps"], 1)
    per_phase_ms = {k: prof[k] / nsub * 1e3 for k in ("sort", "clear", "p2g 1", "p2g 2", "update", "g2p")}
    dom = max(("clear", "p2g 1", "p2g 2", "g2p"), key=lambda k: per_phase_ms[k])
    alg, alg_step, kernel_of = ALG_BYTES, ALG_BYTES_STEP, KERNEL_OF_PHASE
    if sc.dim == 2:      # SURVEY.md 8(d), 2D, A/N = 0.25: 121 B per particle-substep; the 2D path = particle-per-thread kernels
        alg = {"clear": 3.0, "p2g 1": 36.0 + 6.0, "p2g 2": 28.0 + 5.0, "update": 0.0, "g2p": 40.0 + 3.0}
        alg_step = 121.0
        kernel_of = {"clear": "k_clear_tiles", "p2g 1": "k_mass_tiled2", "p2g 2": "k_p2g_tiled2", "g2p": "k_g2p_tiled2"}
    if resident:
        kernel_of = {k: "k_substeps_resident" for k in kernel_of}
    achieved = alg[dom] * sc.n / (per_phase_ms[dom] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "kernel": f"{kernel_of[dom]} ({dom})", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": ncu_traffic(KERNEL_OF_PHASE[dom]) if sc.n == (1 << 24) else None,
        "traffic_source": ncu_capture(), "peak_source": peak_src,
        "alg_bytes_per_particle": alg[dom], "ms_per_launch": per_phase_ms[dom],
        "limiter": ncu_limiter(KERNEL_OF_PHASE[dom]) if sc.n == (1 << 24) else
                   "launch latency: a few thousand particles per kernel (see step_latency_ms)",
        "step_frac": value * alg_step / 1e9 / peak,
        "step_frac_range": [min(value_by_step) * alg_step / 1e9 / peak, max(value_by_step) * alg_step / 1e9 / peak],
        "per_phase_ms": per_phase_ms,
        "occupancy": occupancy,
    }

    # ---- end to end through the C ABI with host buffers -----------------------------------------
    e2e_steps = max(2, min(args.steps, 4))

    def e2e_step():
        sim.clear_particles()
        sim.add_particles_pinned(host.data_ptr(), sc.n)           # H2D from pinned memory
        sim.step()
        n = sim.read_particles_into(back.data_ptr(), sc.n)        # D2H of every record
        assert n == sc.n

    e2e_step()                                                    # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = sc.n * iters * e2e_steps / e2e_s
    rec_bytes = sc.n * rf * 4

    # ---- the reference's own main loop through the API: step(mouse) then the 80x40 frame `draw` prints
    # (3d:541-560): the mouse position goes up, the frame's bin counts come back, the state stays resident
    def loop_step(k):
        sim.step(mouse_pos=(0.25 * sc.rect_max[0] + k, 0.5 * sc.rect_max[1]))
        return sim.frame_counts(viewport=(float(sc.rect_max[0]), float(sc.rect_max[1])))

    loop_step(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        frame = loop_step(k + 1)
    torch.cuda.synchronize()
    loop_s = time.perf_counter() - t0
    e2e_loop = {"value": sc.n * iters * e2e_steps / loop_s, "unit": UNIT, "h2d_bytes_per_step": 8,
                "d2h_bytes_per_step": int(frame.size * 4), "steps": e2e_steps,
                "what": "the reference's main loop: step(mouse position from the host) + the 80x40 frame's bin counts "
                        "read back (draw's binning runs on the device); particle state stays resident"}

    # ---- CPU baseline (rank 0, N=1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        bsc, n_sub, dt, cpu_value, phases = oracle_sample(scenes, scene=None if headline else sc)
        model, cores = cpu_info()
        cpu = {"value": cpu_value, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{bsc.name}: {bsc.n} particles x {n_sub} substeps in {dt:.1f} s " +
                         ("(1/64 of the 2^24-particle column, same construction)" if headline else "(the whole scene)"),
               "host_cpu": model, "host_cores": cores,
               "phase_seconds_last_substep": phases}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(sc.describe(), substeps_per_step=iters,
                       l2="inputs larger than L2 (particle state 1.1 GB, node grid 1.2 GB per GPU)" if sc.n >= (1 << 24)
                          else "working set smaller than L2 (a parity config, not the headline workload)",
                       parallelism=f"z-slabs x{args.gpus}" if args.gpus > 1 else "single GPU"),
        "value_by_step": value_by_step,
        "step_latency_ms": ms / args.steps,
        "timing": {"timed_region": "plain step() calls (CUDA-graph replay of steady-state substeps / resident kernel)",
                   "per_phase": "the kernel's %globaltimer stamps of the last substep" if resident else
                                "replay of the timed region (same state, warm-up and steps) with CUDA-event brackets per phase",
                   "ms_per_step_with_phase_brackets": None if ms_profiled is None else ms_profiled / args.steps},
        "memory": sim.memory_stats(),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rec_bytes,
                "d2h_bytes_per_step": rec_bytes, "steps": e2e_steps,
                "what": "clear + add_particles(pinned host records) + step() + read_particles(all records)"},
        "e2e_main_loop": e2e_loop,
        "gpu_launches": launches,
        "clocks": clocks,
        "ms_per_substep": ms / args.steps / iters,
    }
    if rank == 0:
        print(json.dumps(line))
    sim.close()


def run_slabs(args, pkg, world, rank, local_rank):
    """N > 1: z-slab decomposition, one rank per GPU, 2^24 particles per GPU (weak scaling; N = 8 is
    BASELINE config 5).  Halo planes and migrating particles move over NCCL (NVLink)."""
    import torch
    import torch.distributed as dist
    scenes = pkg.scenes
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # Parity first, outside every timed region: the N-rank slab run of a small sloshing scene against the one-GPU
    # run of the same scene (ids exactly once, sent == received, |dpos|, |dvel| under the stated bound).
    parity = pkg.slab.parity_check(pkg, dist, rank, world, local_rank)
    # ... and the same scene in the deterministic mode (fixed-point node sums, planes added as integers): BIT-FOR-BIT
    det = pkg.slab.parity_check(pkg, dist, rank, world, local_rank, deterministic=True)
    parity = dict(parity, ok=bool(parity["ok"] and det["ok"]), float_path_ok=parity["ok"],
                  deterministic_mode={k: det.get(k) for k in ("ok", "bitwise_equal", "migrated_out", "migrated_in", "halo")})
    sc = scenes.dam_break_for_gpus(world)
    iters = sc.cfg["iterations"]
    rf = scenes.rec_floats(3)
    fill_lo, fill_hi = float(sc.fill_lo[2]), float(sc.fill_hi[2])
    # every rank computes the same plan and the same per-rank particle counts
    tmp = pkg.Simulation.new(sc.cfg, device=local_rank)
    tmp.set_rect(sc.rect_min, sc.rect_max)
    r = tmp.rects()
    tmp.close()
    slabs = pkg.slab.plan_slabs(fill_lo, fill_hi, int(r["origin"][2]), int(r["size"][2]), world)
    spans = [(max(a, fill_lo), min(b, fill_hi)) for a, b in slabs]
    n_rank = [int(round(sc.n * (hi - lo) / (fill_hi - fill_lo))) for lo, hi in spans]
    n_rank[-1] += sc.n - sum(n_rank)
    n_local, id0 = n_rank[rank], sum(n_rank[:rank])

    sim = pkg.slab.SlabSimulation(pkg, sc.cfg, sc.rect_min, sc.rect_max, fill_lo, fill_hi, rank, world, dist,
                                  local_rank, reserve=int(n_local * 1.25) + (1 << 20))
    host = torch.empty((n_local, rf), dtype=torch.float32, pin_memory=True)
    hnp = host.numpy()
    lo = [float(sc.fill_lo[0]), float(sc.fill_lo[1]), spans[rank][0]]
    hi = [float(sc.fill_hi[0]), float(sc.fill_hi[1]), spans[rank][1]]
    chunk = 1 << 21
    for s0 in range(0, n_local, chunk):
        c = min(chunk, n_local - s0)
        hnp[s0:s0 + c] = scenes.box_records(3, lo, hi, n_local, seed=scenes.SEED + 1000 * (rank + 1), start=s0, count=c)
        # lo + u*(hi-lo) can round up to exactly hi (once in 2^24): keep every particle inside this rank's cells
        np.minimum(hnp[s0:s0 + c, 2], np.nextafter(np.float32(hi[2]), np.float32(lo[2])), out=hnp[s0:s0 + c, 2])
    ids = torch.arange(id0, id0 + n_local, dtype=torch.int32)
    back = torch.empty((n_local + (1 << 20), rf), dtype=torch.float32, pin_memory=True)

    def load():
        sim.sim.clear_particles()
        L = pkg.lib()
        import ctypes as C
        st = L.fluid_add_particles(sim.sim._h, C.cast(C.c_void_p(host.data_ptr()), C.POINTER(C.c_float)),
                                   C.cast(C.c_void_p(ids.data_ptr()), C.POINTER(C.c_int32)), n_local)
        assert st == 0, L.fluid_last_error()

    stream = sim.stream

    def region(profiled: bool):
        """warm-up + the K timed steps from the initial state; `profiled` adds CUDA-event brackets per phase."""
        load()
        sim.driver.migrated_out = sim.driver.migrated_in = 0
        for _ in range(args.warmup):
            sim.step()
        torch.cuda.synchronize()
        dist.barrier()
        l0 = sim.sim.launch_count()
        if profiled:
            sim.sim.profile(True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        torch.cuda.synchronize()
        dist.barrier()
        ev[0].record(stream)
        for k in range(args.steps):
            sim.step()
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([ev[0].elapsed_time(ev[-1])] + [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)        # the slowest rank, per step and in total
        pr = None
        if profiled:
            pr = sim.sim.profile_read()
            sim.sim.profile(False)
        return t, sim.sim.launch_count() - l0, pr

    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_t, my_launches, _ = region(False)                # the timed region: no per-phase brackets
    clocks = sampler.stop()
    ms = float(ms_t[0].item())
    value_by_step = [sc.n * iters / (float(m) * 1e-3) for m in ms_t[1:].tolist()]
    launches_t = torch.tensor([my_launches], device="cuda", dtype=torch.int64)
    dist.all_reduce(launches_t)
    migrated_timed = sim.driver.migrated_out
    ms_prof_t, _, prof = region(True)                   # its replay with CUDA-event brackets per phase (roofline)
    cnt = torch.tensor([sim.sim.particle_counts()["active"], migrated_timed], device="cuda", dtype=torch.int64)
    dist.all_reduce(cnt)
    assert int(cnt[0].item()) == sc.n, (int(cnt[0].item()), sc.n)
    value = sc.n * iters * args.steps / (ms * 1e-3)

    peak, peak_src = measured_peaks()
    nsub = max(prof["substeps"], 1)
    per_phase_ms = {k: prof[k] / nsub * 1e3 for k in ("sort", "clear", "p2g 1", "p2g 2", "update", "g2p")}
    dom = max(("clear", "p2g 1", "p2g 2", "g2p"), key=lambda k: per_phase_ms[k])
    achieved = ALG_BYTES[dom] * n_local / (per_phase_ms[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"{KERNEL_OF_PHASE[dom]} ({dom})", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(KERNEL_OF_PHASE[dom]),
                "traffic_source": f"{ncu_capture()} (single-GPU capture of the same kernel and per-GPU size)",
                "limiter": ncu_limiter(KERNEL_OF_PHASE[dom]),
                "peak_source": peak_src,
                "alg_bytes_per_particle": ALG_BYTES[dom], "ms_per_launch": per_phase_ms[dom],
                "step_frac": value * ALG_BYTES_STEP / 1e9 / (peak * world),
                "per_phase_ms": per_phase_ms,
                "note": "rank 0's kernels; phase brackets in slab runs include the halo exchange that follows the phase"}

    # end to end: every step uploads this rank's records from pinned memory and reads every record back
    e2e_steps = max(2, min(args.steps, 3))

    def e2e_step():
        load()
        sim.step()
        n = sim.sim.read_particles_into(back.data_ptr(), back.shape[0])
        return n

    e2e_step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dist.barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = sc.n * iters * e2e_steps / float(e2e_t.item())
    rec_bytes = n_local * rf * 4

    # the reference's main loop over the slabs: step(mouse) + the summed 80x40 frame, state resident
    def loop_step(k):
        sim.step(mouse=(0.25 * float(sc.rect_max[0]) + k, 0.5 * float(sc.rect_max[1])))
        return sim.frame_counts(viewport=(float(sc.rect_max[0]), float(sc.rect_max[1])))

    loop_step(0)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        frame = loop_step(k + 1)
    torch.cuda.synchronize()
    dist.barrier()
    loop_t = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(loop_t, op=dist.ReduceOp.MAX)
    e2e_loop = {"value": sc.n * iters * e2e_steps / float(loop_t.item()), "unit": UNIT, "h2d_bytes_per_step": 8,
                "d2h_bytes_per_step": int(frame.size * 4), "steps": e2e_steps,
                "what": "the reference's main loop: step(mouse position from the host) on every rank + the 80x40 frame's "
                        "bin counts summed over the ranks and read back; particle state stays resident"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(sc.describe(), substeps_per_step=iters,
                       l2="inputs larger than L2 (particle state 1.1 GB, node grid > 1 GB per GPU)",
                       parallelism=f"z-slabs x{world}", particles_per_gpu=n_rank,
                       slabs=[list(x) for x in slabs]),
        "value_by_step": value_by_step,
        "timing": {"timed_region": "plain step() calls on every rank, max over ranks",
                   "per_phase": "replay of the timed region (same state, warm-up and steps) with CUDA-event brackets per phase",
                   "ms_per_step_with_phase_brackets": float(ms_prof_t[0].item()) / args.steps},
        "roofline": roofline, "cpu_baseline": None, "parity_check": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rec_bytes, "d2h_bytes_per_step": rec_bytes,
                "steps": e2e_steps, "what": "per rank: clear + add_particles(pinned host records) + step() + "
                                            "read_particles(all records); bytes are per rank"},
        "e2e_main_loop": e2e_loop,
        "gpu_launches": int(launches_t.item()), "clocks": clocks, "ms_per_substep": ms / args.steps / iters,
        "migrated_particles": int(cnt[1].item()),
    }
    if rank == 0:
        print(json.dumps(line))
    sim.close()
    dist.barrier()
    dist.destroy_process_group()
    if not parity["ok"]:
        raise SystemExit("bench.py: the multi-GPU parity check failed: " + json.dumps(parity))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene", default="", help="scene function in scenes.py (default: dam break for --gpus)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sparse-blocks", type=int, default=0, help="N > 0: block-sparse node storage with a pool of N blocks")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything a library writes to file descriptor 1 meanwhile (NCCL's
    # version banner comes from C code) goes to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    import builtins
    real_print = builtins.print

    def capture(*a, **k):
        if k.get("file") in (None, sys.stdout) and len(a) == 1 and isinstance(a[0], str) and a[0].startswith("{"):
            lines.append(a[0])
        else:
            real_print(*a, **k)

    builtins.print = capture
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        builtins.print = real_print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for ln in lines:
        print(ln, flush=True)


if __name__ == "__main__":
    main()
