"""In-tree build of libmarlsat_b200.so (hand-written CUDA for sm_100a, C ABI in include/marl_sat_b200.h).

The library is compiled with plain ``nvcc`` (no torch extension machinery): the boundary is a C ABI and
the Python layer binds it with ctypes.  ``python marl_sat_b200/build.py`` rebuilds it (run as a script so the package, which refuses to import
without the library, is not imported first).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = CSRC / "libmarlsat_b200.so"
SOURCES = ["cabi.cu", "satenv.cu", "keys.cu", "gae.cu", "features.cu", "dimacs.cpp"]
HEADERS = ["common.cuh", "internal.h", "../../include/marl_sat_b200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build libmarlsat_b200.so (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = [CSRC / s for s in SOURCES if (CSRC / s).exists()] + [(CSRC / h).resolve() for h in HEADERS]
    return any(p.stat().st_mtime > t for p in deps if p.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``.cu`` under ``csrc/`` into one shared library for sm_100a."""
    if not force and not is_stale():
        return LIB_PATH
    srcs = [s for s in SOURCES if (CSRC / s).exists()]
    cmd = [_nvcc(), *NVCC_FLAGS] + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB_PATH), *srcs]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
