"""Build the oracle's plain-C parts (gcc) into oracle/_native/: libretina_post.so from oracle/retina_post.c.
-ffp-contract=off: the restatement must not fuse multiply-adds the reference's x86 build would not fuse."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_native"


def build(force: bool = False) -> Path:
    OUT.mkdir(exist_ok=True)
    src, lib = HERE / "retina_post.c", OUT / "libretina_post.so"
    if force or not lib.exists() or lib.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", str(src), "-o", str(lib), "-lm"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle C build failed:\n" + r.stderr)
    return lib


if __name__ == "__main__":
    print(build(force=True))
