"""Build oracle/_ref/ from the reference's OWN sources where they lie under /root/reference/src (compiled verbatim, nothing
copied into this repo), plus the small extern "C" accessors beside this file and the stub headers in oracle/stubs/:

    libref_matmul.so   matmul.cpp + common.cpp + ref_matmul_wrap.cpp   (cuBLASLt: needs a GPU to RUN -> `-m gpu` tests, bench.py)
    libref_retina.so   retinaface.cpp + common.cpp + ref_retina_wrap.cpp  (constructor sizes, create_anchor_retinaface,
                       postprocessing, nms — pure host code, runs anywhere; TensorRT / OpenCV / the CUDA runtime are stubbed:
                       oracle/stubs/NvInfer.h, oracle/stubs/opencv2/, and host-memory cudaMalloc & co. in the accessor)

Only possible where /root/reference is mounted (the build container). oracle/_ref/ is git-ignored but travels to the GPU
box with the gpurun snapshot. Flags follow the reference's build (app/CMakeLists.txt: C++11, no -march, no fast-math): plain
x86-64 has no FMA, so no contraction can occur in the double/float arithmetic of postprocessing.
The network halves of retinaface.cpp / arcface.cpp (TensorRT engines) and app.cpp (Crow, SQLite) stay unbuildable here.
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/src")
OUT = HERE / "_ref"
LIB = OUT / "libref_matmul.so"
LIB_RETINA = OUT / "libref_retina.so"


def _stale(lib: Path, srcs: list[Path]) -> bool:
    deps = srcs + list((HERE / "stubs").rglob("*.h*"))
    return not lib.exists() or lib.stat().st_mtime < max(s.stat().st_mtime for s in deps)


def _run(cmd: list[str], what: str) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference {what} build failed:\n" + r.stderr)


def build(force: bool = False) -> Path | None:
    if not (REF / "matmul.cpp").exists():
        return LIB if LIB.exists() else None
    OUT.mkdir(exist_ok=True)
    base = ["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w", f"-I{HERE / 'stubs'}", f"-I{REF}", "-I/usr/local/cuda/include"]
    srcs = [REF / "matmul.cpp", REF / "common.cpp", HERE / "ref_matmul_wrap.cpp"]
    if force or _stale(LIB, srcs):
        _run([*base, *map(str, srcs), "-o", str(LIB), "-L/usr/local/cuda/lib64", "-lcublasLt", "-lcudart",
              "-Wl,-rpath,/usr/local/cuda/lib64"], "matmul")
    srcs = [REF / "retinaface.cpp", REF / "common.cpp", HERE / "ref_retina_wrap.cpp"]
    if force or _stale(LIB_RETINA, srcs):
        # no libcudart: the accessor satisfies the runtime calls of the constructor with host memory; -Bsymbolic keeps those
        # definitions private to this library even inside a process that has the real runtime loaded
        _run([*base, *map(str, srcs), "-o", str(LIB_RETINA), "-Wl,-Bsymbolic"], "retinaface")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv), LIB_RETINA)
