"""Secondary measurements for bench.py: the detect -> crop -> embed -> search pipeline of BASELINE.json configs[1..3]
(faces/sec end-to-end on 640x640 frames) on ONE GPU, with per-stage times and roofline fractions, and the CPU port of the
same stages (oracle modules, torch CPU fp32, all host threads) on a bounded sample.

Algorithmic work per unit (SURVEY §8d / BASELINE.md): embedder 12.593 GFLOP/face (tensor-bound); detector 54.9 MB/frame of
layer-wise fp16 activation traffic + 1.23 MB u8 input (HBM-bound); search rows x 1 KiB per batch (HBM-bound).
"""
from __future__ import annotations

import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "face-recognition-cpp-tensorrt_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

EMB_GFLOP_PER_FACE = 12.593
DET_MB_PER_FRAME = 54.9 + 1.23


def _event_time(torch, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def run_gpu(device: int, frames_batch=64, gallery_rows=1_000_000, reps=10, hbm_gbs=6542.1, tf_sust=1381.0) -> dict:
    import torch

    import frb200
    from oracle import synth_weights as sw
    from tools import make_golden_nets as mg
    from tools import make_golden_retina as mgr
    from tools import pack_retina as pr
    from tools import pack_weights as pw

    out: dict = {}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        pr.save_retina(td / "det.frw", sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT), False)
        pw.save_arcface(td / "arc.frw", sw.arcface_state_dict("ir_se", 7), "ir_se")
        det = frb200.Detector(td / "det.frw", (640, 640), max_batch=frames_batch, max_faces=4, device=device)
        emb = frb200.Embedder(td / "arc.frw", max_batch=256, device=device)
        gal = frb200.Gallery.synthetic(gallery_rows, seed=17, device=device)
        gal.set_path(frb200.FR_PATH_TENSOR)
        pipe = frb200.Pipeline(det, emb, gal)
        base = mgr.det_frames(4, 640, 640, seed=13)
        frames = torch.from_numpy(np.ascontiguousarray(np.concatenate([base] * (frames_batch // 4 + 1))[:frames_batch])).pin_memory()
        res = pipe.run(frames)
        faces = int(res["counts"].sum())
        l0 = frb200.launch_count()
        t = _event_time(torch, lambda: pipe.run(frames), reps)
        launches = (frb200.launch_count() - l0) // (reps + 3)
        out["e2e"] = {"metric": "faces/sec end-to-end 640x640", "value": faces / t, "unit": "faces/s", "frames_per_s": frames_batch / t,
                      "ms_per_batch": t * 1e3, "batch_frames": frames_batch, "faces_per_batch": faces, "gallery_rows": gallery_rows,
                      "h2d_bytes_per_batch": int(frames.numel()), "gpu_launches_per_batch": int(launches),
                      "config": "detect(RetinaFace mobile0.25 640x640) -> crop+bicubic 112x112 -> embed(ArcFace IR-SE-50) -> top-1 search"}
        # throughput with TWO batches in flight: a second, independent set of handles (own stream, own scratch) driven from a second
        # host thread, so that one batch's H2D copies and host-side box compaction overlap the other's kernels. Same public call
        # (fr_pipeline_run, host buffers in and out); wall clock around both threads, each call ends with its own stream sync.
        try:
            import threading

            det2 = frb200.Detector(td / "det.frw", (640, 640), max_batch=frames_batch, max_faces=4, device=device)
            emb2 = frb200.Embedder(td / "arc.frw", max_batch=256, device=device)
            gal2 = frb200.Gallery.synthetic(gallery_rows, seed=17, device=device)
            gal2.set_path(frb200.FR_PATH_TENSOR)
            pipe2 = frb200.Pipeline(det2, emb2, gal2)
            frames2 = frames.clone().pin_memory()
            res2 = pipe2.run(frames2)
            assert np.array_equal(res2["idx"], res["idx"]) and np.array_equal(res2["counts"], res["counts"])
            for _ in range(2):
                pipe.run(frames)
                pipe2.run(frames2)
            torch.cuda.synchronize()
            n_iter = reps
            start = threading.Barrier(3)

            def worker(pp, ff):
                start.wait()
                for _ in range(n_iter):
                    pp.run(ff)

            ths = [threading.Thread(target=worker, args=(pipe, frames)), threading.Thread(target=worker, args=(pipe2, frames2))]
            for th in ths:
                th.start()
            start.wait()
            t0 = time.perf_counter()
            for th in ths:
                th.join()
            torch.cuda.synchronize()
            t2 = (time.perf_counter() - t0) / (2 * n_iter)
            out["e2e_two_in_flight"] = {"value": faces / t2, "unit": "faces/s", "frames_per_s": frames_batch / t2, "ms_per_batch": t2 * 1e3,
                                        "note": "two independent handle sets on two streams / host threads, wall clock; per-batch latency is the "
                                                "single-stream ms_per_batch above"}
            pipe2.close()
            gal2.close()
            det2.close()
            emb2.close()
        except Exception as e:  # informational
            out["e2e_two_in_flight"] = {"error": f"{type(e).__name__}: {e}"}
        # stage breakdown on the same handles (host-buffer API, so H2D/D2H are inside)
        f16 = frames[:16].numpy()
        td_ = _event_time(torch, lambda: det.run(f16), reps)
        out["detect"] = {"batch": 16, "ms": td_ * 1e3, "frames_per_s": 16 / td_,
                         "roofline": {"bound": "hbm", "achieved": DET_MB_PER_FRAME * 16e-3 / td_, "peak": hbm_gbs, "unit": "GB/s",
                                      "frac": DET_MB_PER_FRAME * 16e-3 / td_ / hbm_gbs, "note": "whole detector step incl. H2D, not one kernel"}}
        crops = np.ascontiguousarray(np.concatenate([mg.arcface_inputs()] * 4))
        te = _event_time(torch, lambda: emb.run_crops(crops), reps)
        out["embed"] = {"batch": 32, "mode": "ir_se", "ms": te * 1e3, "faces_per_s": 32 / te,
                        "roofline": {"bound": "tensor", "achieved": EMB_GFLOP_PER_FACE * 32e-3 / te, "peak": tf_sust, "unit": "TFLOP/s",
                                     "frac": EMB_GFLOP_PER_FACE * 32e-3 / te / tf_sust, "note": "whole embedder step incl. H2D, not one kernel"}}
        pipe.close()
        gal.close()
        det.close()
        emb.close()
    return out


def run_cpu(det_frames=2, emb_faces=8) -> dict:
    """the reference's arithmetic on the host cores: oracle restatements of its PyTorch modules (fp32 eager) + C decode/NMS"""
    import torch

    from oracle import arcface_oracle as ao
    from oracle import retina_oracle as ro
    from oracle import synth_weights as sw
    from tools import make_golden_nets as mg
    from tools import make_golden_retina as mgr

    threads = torch.get_num_threads()
    det_sd = ro.to_torch(sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT))
    frames = mgr.det_frames(det_frames, 640, 640, seed=13)
    t0 = time.perf_counter()
    x = torch.from_numpy(np.stack([ro.preprocess(f, 640, 640) for f in frames]))
    loc, conf, _ = ro.forward(det_sd, x, False)
    for i in range(det_frames):
        ro.postprocess(loc[i].numpy(), conf[i].numpy(), None, 640, 640, 640, 640, 0.4, 0.6, 4)
    t_det = (time.perf_counter() - t0) / det_frames
    arc_sd = ao.to_torch(sw.arcface_state_dict("ir_se", 7))
    crops = mg.arcface_inputs(emb_faces)
    t0 = time.perf_counter()
    ao.forward(arc_sd, torch.from_numpy(ao.preprocess_faces(crops)), "ir_se")
    t_emb = (time.perf_counter() - t0) / emb_faces
    faces_per_frame = 4
    return {"cores": threads, "kind": "port", "detect_s_per_frame": t_det, "embed_s_per_face": t_emb,
            "faces_per_s": faces_per_frame / (t_det + faces_per_frame * t_emb),
            "sample": f"{det_frames} frames 640x640 through the detector oracle + C decode/NMS, {emb_faces} faces through the IR-SE-50 oracle "
                      f"(torch CPU fp32, {threads} threads); search excluded (see cpu_baseline.value)"}
