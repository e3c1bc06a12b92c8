"""In-tree build of the CUDA library (sm_100a only) — called by __graft_entry__.build()."""
from __future__ import annotations

import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = CSRC / "libfluid_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # IEEE division / sqrt and no fast-math: integer cell and key rules must match the
    # reference bit for bit (common.cuh)
    "-prec-div=true", "-prec-sqrt=true",
]


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "fluid_b200.h"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    cus, hdrs = sources()
    return any(p.stat().st_mtime > t for p in cus + hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cus, _ = sources()
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB), *[str(c) for c in cus]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
