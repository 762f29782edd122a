"""Throughput of the detector / embedder / end-to-end pipeline on one GPU (BASELINE.json configs[1..3]); also the target
command for ncu captures of the conv kernels. Prints one JSON object per stage. Timed with CUDA events through torch on the
public host-buffer API (H2D/D2H inside), after warm-up.

    python tools/perf_nets.py [--stages detect,embed,e2e] [--reps 10] [--gallery 1000000]
"""
from __future__ import annotations

import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))


def make_ckpts(tmp: Path, arc_mode="ir_se"):
    from tools import synth_weights as sw
    from tools import make_golden_retina as mgr
    from tools import pack_retina as pr
    from tools import pack_weights as pw

    pr.save_retina(tmp / "det.frw", sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT), False)
    pw.save_arcface(tmp / "arc.frw", sw.arcface_state_dict(arc_mode, 7), arc_mode)
    return tmp / "det.frw", tmp / "arc.frw"


def timed(fn, reps, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="detect,embed,e2e")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--gallery", type=int, default=1_000_000)
    ap.add_argument("--det-batch", type=int, default=16)
    ap.add_argument("--emb-batch", type=int, default=32)
    ap.add_argument("--e2e-batch", type=int, default=64)
    ap.add_argument("--arc-mode", default="ir_se")
    a = ap.parse_args()
    import torch

    import frb200
    from tools import make_golden_nets as mg
    from tools import make_golden_retina as mgr

    stages = a.stages.split(",")
    with tempfile.TemporaryDirectory() as td:
        det_f, arc_f = make_ckpts(Path(td), a.arc_mode)
        if "detect" in stages:
            det = frb200.Detector(det_f, (640, 640), max_batch=a.det_batch, max_faces=4)
            frames = np.ascontiguousarray(np.concatenate([mgr.det_frames(4, 640, 640, seed=13)] * (a.det_batch // 4 + 1))[: a.det_batch])
            t = timed(lambda: det.run(frames), a.reps)
            print(json.dumps({"stage": "detect", "batch": a.det_batch, "ms": t * 1e3, "frames_per_s": a.det_batch / t,
                              "hbm_floor_ms": 0.137 * a.det_batch / 16}), flush=True)
            det.close()
        if "embed" in stages:
            emb = frb200.Embedder(arc_f, max_batch=a.emb_batch)
            crops = np.ascontiguousarray(np.concatenate([mg.arcface_inputs()] * (a.emb_batch // 8 + 1))[: a.emb_batch])
            t = timed(lambda: emb.run_crops(crops), a.reps)
            gflop = 12.593 * a.emb_batch
            print(json.dumps({"stage": "embed", "mode": a.arc_mode, "batch": a.emb_batch, "ms": t * 1e3, "faces_per_s": a.emb_batch / t,
                              "tflops": gflop / t / 1e3, "tensor_floor_ms": gflop / 1381e3 * 1e3}), flush=True)
            emb.close()
        if "e2e" in stages:
            det = frb200.Detector(det_f, (640, 640), max_batch=a.e2e_batch, max_faces=4)
            emb = frb200.Embedder(arc_f, max_batch=min(256, a.e2e_batch * 4))
            gal = frb200.Gallery.synthetic(a.gallery, seed=17)
            gal.set_path(frb200.FR_PATH_TENSOR)
            pipe = frb200.Pipeline(det, emb, gal)
            frames = torch.from_numpy(np.ascontiguousarray(
                np.concatenate([mgr.det_frames(4, 640, 640, seed=13)] * (a.e2e_batch // 4 + 1))[: a.e2e_batch])).pin_memory()
            res = pipe.run(frames)
            faces = int(res["counts"].sum())
            t = timed(lambda: pipe.run(frames), a.reps)
            print(json.dumps({"stage": "e2e", "batch_frames": a.e2e_batch, "faces": faces, "gallery_rows": a.gallery, "ms": t * 1e3,
                              "faces_per_s": faces / t, "frames_per_s": a.e2e_batch / t}), flush=True)
            pipe.close()
            gal.close()
            det.close()
            emb.close()


if __name__ == "__main__":
    main()
