"""Build oracle/_ref/libref_matmul.so from the reference's OWN sources where they lie (/root/reference/src/matmul.cpp,
common.cpp — compiled verbatim, nothing copied into this repo) plus oracle/ref_matmul_wrap.cpp and the NvInfer.h stub.

Only possible where /root/reference is mounted (the build container). oracle/_ref/ is git-ignored but travels to the GPU
box with the gpurun snapshot; the library needs a GPU + cuBLASLt to run, so it is exercised by `-m gpu` tests and
bench.py only. The rest of the reference path (retinaface.cpp, arcface.cpp) needs TensorRT + OpenCV C++ headers, which this
image lacks, and is therefore unbuildable here.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src")
OUT = HERE / "_ref"
LIB = OUT / "libref_matmul.so"


def build(force: bool = False) -> Path | None:
    if not (REF / "matmul.cpp").exists():
        return LIB if LIB.exists() else None
    srcs = [REF / "matmul.cpp", REF / "common.cpp", HERE / "ref_matmul_wrap.cpp"]
    if LIB.exists() and not force and LIB.stat().st_mtime >= max(s.stat().st_mtime for s in srcs):
        return LIB
    OUT.mkdir(exist_ok=True)
    cmd = ["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w", f"-I{HERE / 'stubs'}", f"-I{REF}", "-I/usr/local/cuda/include",
           *map(str, srcs), "-o", str(LIB), "-L/usr/local/cuda/lib64", "-lcublasLt", "-lcudart",
           "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference matmul build failed:\n" + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
