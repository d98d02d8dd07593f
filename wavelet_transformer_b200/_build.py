"""In-tree build of libwavelet_sm100a.so (one shared library, sm_100a only).

    python -m wavelet_transformer_b200._build [--force] [-v]

Every .cu under csrc/ is compiled to lib/obj/<name>.o (only when it, or a header it includes, is
newer than its object), the objects are compiled in parallel and linked into the one library.
"""

from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libwavelet_sm100a.so"
OBJ = PKG / "lib" / "obj"

NVCC_COMPILE = [
    "-c", "-Xcompiler", "-fPIC", "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
]
NVCC_LINK = [
    "-shared", "-Xcompiler", "-fPIC",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xlinker", "--no-undefined",      # a symbol missing between translation units fails the build, not the first load
]
# kept for readers of the round-1 docs: the flags one nvcc invocation over all sources would take
NVCC_FLAGS = ["-shared", *NVCC_COMPILE[1:], "-Xlinker", "--no-undefined"]

_INCLUDE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps(path: Path, seen: set[Path] | None = None) -> set[Path]:
    """The file and every local header it includes, transitively."""
    seen = seen if seen is not None else set()
    path = path.resolve()
    if path in seen or not path.exists():
        return seen
    seen.add(path)
    for inc in _INCLUDE.findall(path.read_text()):
        _deps((path.parent / inc), seen)
    return seen


def _obj(src: Path) -> Path:
    return OBJ / (src.stem + ".o")


def _stale_obj(src: Path) -> bool:
    o = _obj(src)
    if not o.exists():
        return True
    return max(p.stat().st_mtime for p in _deps(src)) > o.stat().st_mtime


def _stale() -> bool:
    if not LIB.exists():
        return True
    return any(_stale_obj(s) or _obj(s).stat().st_mtime > LIB.stat().st_mtime for s in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ with nvcc for sm_100a and link lib/libwavelet_sm100a.so."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libwavelet_sm100a.so")
    OBJ.mkdir(parents=True, exist_ok=True)
    todo = [s for s in sources() if force or _stale_obj(s)]

    def compile_one(src: Path):
        cmd = [nvcc, *NVCC_COMPILE, *(["-Xptxas", "-v"] if verbose else []), "-o", str(_obj(src)), str(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4) or 1) as ex:
        results = list(ex.map(compile_one, todo))
    log = []
    for src, proc in results:
        if proc.returncode != 0:
            _obj(src).unlink(missing_ok=True)
            raise RuntimeError(f"nvcc failed on {src.name}:\n" + proc.stdout + proc.stderr)
        log.append(proc.stderr)
    known = {_obj(s) for s in sources()}
    for stray in OBJ.glob("*.o"):          # objects of sources that no longer exist
        if stray not in known:
            stray.unlink()
    cmd = [nvcc, *NVCC_LINK, "-o", str(LIB), *map(str, sorted(known))]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
