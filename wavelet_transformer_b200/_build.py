"""In-tree build of libwavelet_sm100a.so (one shared library, sm_100a only).

    python -m wavelet_transformer_b200._build [--force]
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libwavelet_sm100a.so"

NVCC_FLAGS = [
    "-shared", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xlinker", "--no-undefined",      # a symbol missing between translation units fails the build, not the first load
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB.exists():
        return True
    newest = max(p.stat().st_mtime for p in list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh"))
                 + [PKG.parent / "include" / "wtb.h"])
    return newest > LIB.stat().st_mtime


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into lib/libwavelet_sm100a.so with nvcc."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libwavelet_sm100a.so")
    LIB.parent.mkdir(parents=True, exist_ok=True)
    cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []),
           "-o", str(LIB), *map(str, sources())]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
