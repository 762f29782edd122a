"""Build libfr_b200.so (the C-ABI library, include/fr_b200.h) with nvcc for sm_100a, in-tree.

    python face-recognition-cpp-tensorrt_b200/build.py [--force] [--verbose]

Objects go to face-recognition-cpp-tensorrt_b200/build/, the library to face-recognition-cpp-tensorrt_b200/lib/.
nvcc cross-compiles on a host without a GPU. The CUDA runtime is linked statically so that the library
does not depend on which libcudart a host process (e.g. PyTorch in the tests) has already loaded.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# A/B builds: FR_BUILD_TAG=<name> FR_BUILD_DEFS="-DFR_AB_X=0 ..." python build.py  ->  lib_ab/<name>/libfr_b200.so (tools/ab_search.py)
TAG = os.environ.get("FR_BUILD_TAG", "")
BUILD = PKG / "build_ab" / TAG if TAG else PKG / "build"
LIBDIR = PKG / "lib_ab" / TAG if TAG else PKG / "lib"
LIB = LIBDIR / "libfr_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "-DFR_BUILDING_LIB", *os.environ.get("FR_BUILD_DEFS", "").split()]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_hash(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(ARCH + CFLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile(src: Path, force: bool, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".hash")
    want = _deps_hash(src)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [NVCC, *ARCH, *CFLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(r.stderr)
    stamp.write_text(want)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, *ARCH, "-shared", "-cudart", "static", "-o", str(LIB), *map(str, objs), "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
