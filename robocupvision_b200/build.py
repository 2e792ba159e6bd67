"""Build librcv_b200.so (sm_100a only) in-tree with nvcc.

    python -m robocupvision_b200.build [--force] [--verbose]

The shared library is plain C ABI (include/rcv_b200.h): no torch headers, so the
build is a handful of nvcc invocations that cross-compile without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "librcv_b200.so"

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-DRCV_BUILDING=1",
    f"-I{ROOT / 'include'}",
    f"-I{CSRC}",
] + (["-DRCV_PROF=1"] if os.environ.get("RCV_PROF") else [])  # phase-timing instrumentation (debug builds)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: librcv_b200.so cannot be built")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + [ROOT / "include" / "rcv_b200.h"]):
        if f.is_file():
            h.update(f.name.encode())
            h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp_file = OBJ / "stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    OBJ.mkdir(exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: Path) -> Path:
        obj = OBJ / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(f"== {src.name}\n{r.stderr}\n")
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
