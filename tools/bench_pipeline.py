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


def _event_time(torch, fn, reps, warm=3, stream=None):
    """CUDA events on `stream` (the stream fn launches on; None = torch's current stream, for the host-buffer calls that synchronise
    themselves)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def _oracle_plant(frames, boxes, counts, gallery_rows, noise=0.3, seed=3):
    """the parity gate's expectation (checker, untimed): the oracle chain on the detector's boxes — OpenCV-exact crops
    (oracle/cv_resize.py) -> fp32 IR-SE-50 (oracle/arcface_oracle.py) — gives one embedding per DISTINCT face; a noisy copy of each is
    planted at a known gallery row, so every face of the batch has a known top-1 identity. Returns (rows [n_distinct], planted
    vectors [n_distinct, 512], oracle embeddings)."""
    import torch

    from oracle import arcface_oracle as ao
    from oracle import cv_resize as cr
    from oracle import search_oracle as so
    from tools import synth_weights as sw

    arc_sd = ao.to_torch(sw.arcface_state_dict("ir_se", 7))
    embs = []
    for f in range(frames.shape[0]):
        crops = cr.cropped_faces(frames[f], boxes[f, : counts[f]])
        embs.append(ao.forward(arc_sd, torch.from_numpy(ao.preprocess_faces(crops)), "ir_se").numpy())
    emb = np.concatenate(embs)
    rows = (np.arange(emb.shape[0], dtype=np.int64) * 7919 + 11) % gallery_rows
    assert len(set(rows.tolist())) == emb.shape[0]
    return rows, so.planted_queries(emb, noise=noise, seed=seed), emb


def run_gpu(device: int, frames_batch=64, gallery_rows=1_000_000, reps=10, hbm_gbs=6542.1, tf_sust=1381.0, dist=None, world=1, rank=0,
            stage_breakdown=True, distinct_frames=16) -> dict:
    """faces/sec end to end on THIS rank's GPU; with dist != None every rank runs its own replica (weights and the 1M-row gallery
    replicated, SURVEY 8e "replicas only") between common barriers and the aggregate is reported by rank 0."""
    import torch

    import frb200
    from oracle import search_oracle as so
    from tools import make_golden_nets as mg
    from tools import make_golden_retina as mgr
    from tools import pack_retina as pr
    from tools import pack_weights as pw
    from tools import synth_weights as sw

    dev = torch.device("cuda", device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
        base = mgr.det_frames(distinct_frames, 640, 640, seed=13)
        reps_f = (frames_batch + distinct_frames - 1) // distinct_frames
        frames = torch.from_numpy(np.ascontiguousarray(np.concatenate([base] * reps_f)[:frames_batch])).pin_memory()
        # ---- parity gate (untimed): identities of every face of the batch are known and checked in this run
        boxes_d, counts_d, _ = det.run(base)
        planted = torch.zeros((int(counts_d.sum()), 512), dtype=torch.float32, device=dev)
        rows_t = torch.zeros(int(counts_d.sum()), dtype=torch.int64, device=dev)
        oracle_emb = None
        if rank == 0:
            rows, vecs, oracle_emb = _oracle_plant(base, boxes_d, counts_d, gallery_rows)
            planted.copy_(torch.from_numpy(vecs))
            rows_t.copy_(torch.from_numpy(rows))
        if dist is not None:
            dist.broadcast(planted, 0)
            dist.broadcast(rows_t, 0)
        rows, vecs = rows_t.cpu().numpy(), planted.cpu().numpy()
        for r, v in zip(rows, vecs):
            gal.update(int(r), v)
        res = pipe.run(frames, want_embeddings=True)
        faces = int(res["counts"].sum())
        # expected identity of face j of frame f = the planted row of that face of base frame f % distinct_frames
        first = np.concatenate([[0], np.cumsum(counts_d)])[:-1]
        exp = np.full(res["idx"].shape, -1, np.int64)
        for f in range(frames_batch):
            bf = f % distinct_frames
            exp[f, : counts_d[bf]] = rows[first[bf] : first[bf] + counts_d[bf]]
        ok_idx = bool(np.array_equal(res["idx"], exp) and np.array_equal(res["counts"], np.tile(counts_d, reps_f)[:frames_batch]))
        gate = {"identities_exact": ok_idx, "faces_checked": faces, "distinct_faces": int(counts_d.sum())}
        if rank == 0 and oracle_emb is not None:
            got = np.concatenate([res["embeddings"][f, : counts_d[f]] for f in range(distinct_frames)])
            gate["max_abs_dembedding_vs_fp32_oracle"] = float(np.abs(got - oracle_emb).max())
            ws = np.einsum("ij,ij->i", oracle_emb.astype(np.float64), vecs.astype(np.float64))
            gs = np.concatenate([res["score"][f, : counts_d[f]] for f in range(distinct_frames)])
            gate["max_abs_dscore_vs_oracle"] = float(np.abs(gs - ws).max())
            ok_idx = ok_idx and gate["max_abs_dembedding_vs_fp32_oracle"] <= 1e-3 and gate["max_abs_dscore_vs_oracle"] <= 1e-3
        if not ok_idx:
            raise RuntimeError(f"pipeline parity gate failed: {gate}")
        # ---- timed: `reps` batches back to back through the public host-buffer calls, wall clock between barriers, max over ranks.
        # (a) fr_pipeline_run: synchronous, one batch at a time; (b) fr_pipeline_submit / fr_pipeline_collect: two batches in flight (batch
        # i + 1 crosses PCIe and is enqueued while batch i computes) - what a serving loop does. Identities are checked in both loops.
        windows_ms = {}

        def timed(step_fn, drain_fn=None, name="", windows=5):
            """median of `windows` windows of `reps` batches each (a window is ~65 ms of wall clock: one host hiccup - a page fault, a
            scheduler stall - used to decide the whole number); every window: barrier + synchronize on both sides, max over ranks"""
            for _ in range(3):
                step_fn()
            if drain_fn:
                drain_fn()
            ts, launches = [], 0
            for _ in range(windows):
                torch.cuda.synchronize()
                barrier()
                l0 = frb200.launch_count()
                t0 = time.perf_counter()
                for _ in range(reps):
                    step_fn()
                if drain_fn:
                    drain_fn()
                torch.cuda.synchronize()
                t = (time.perf_counter() - t0) / reps
                launches = (frb200.launch_count() - l0) // reps
                ts.append(max_over_ranks(t))
                barrier()
            windows_ms[name] = [round(x * 1e3, 4) for x in ts]
            return float(np.median(ts)), launches

        last = {}

        def sync_step():
            last["res"] = pipe.run(frames)

        t_sync, launches = timed(sync_step, name="synchronous")
        if not np.array_equal(last["res"]["idx"], exp):
            raise RuntimeError("pipeline parity gate failed inside the timed loop (synchronous call)")
        bad = []

        def flight_step():
            pipe.submit(frames)
            if pipe.in_flight() == 2:
                r = pipe.collect()
                if not np.array_equal(r["idx"], exp):
                    bad.append(1)

        def flight_drain():
            while pipe.in_flight():
                r = pipe.collect()
                if not np.array_equal(r["idx"], exp):
                    bad.append(1)

        t_fl, _ = timed(flight_step, flight_drain, name="in_flight")
        if bad:
            raise RuntimeError("pipeline parity gate failed inside the timed loop (two batches in flight)")
        t_all = t_fl

        def line(t, api):
            return {"metric": "faces/sec end-to-end 640x640", "value": world * faces / t, "unit": "faces/s", "n_gpus": world,
                    "faces_per_s_per_gpu": faces / t, "frames_per_s": world * frames_batch / t, "ms_per_batch": t * 1e3, "api": api}

        out["e2e"] = {**line(t_fl, "fr_pipeline_submit + fr_pipeline_collect: host frames (pinned) -> identities on the host, two batches in flight"),
                      "batch_frames": frames_batch, "faces_per_batch": faces, "gallery_rows": gallery_rows, "parallelism": "replicas (one pipeline per GPU, no collective)",
                      "h2d_bytes_per_batch": int(frames.numel()), "d2h_bytes_per_batch": faces * 12 + frames_batch * (4 * 20 + 4),
                      "gpu_launches_per_batch": int(launches), "parity_gate": gate, "timing": "median of 5 windows of `reps` batches, wall clock, barrier + synchronize on both sides, max over ranks; identities of every batch checked",
                      "ms_per_batch_windows": windows_ms,
                      "config": "detect(RetinaFace mobile0.25 640x640) -> device-side face compaction -> crop+bicubic 112x112 -> embed(ArcFace IR-SE-50) -> top-1 search (BASELINE.json configs[3])",
                      "synchronous_call": line(t_sync, "fr_pipeline_run: one batch at a time, the call returns the identities")}
        if stage_breakdown:
            # stage breakdown on the same handles, device-resident (kernels only) and through the host-buffer entry points
            f16 = frames[:16].numpy()
            td_ = _event_time(torch, lambda: det.run(f16), reps)
            out["detect_host_b16"] = {"batch": 16, "ms": td_ * 1e3, "frames_per_s": 16 / td_, "inputs": "host frames (fr_detector_run: H2D + kernels + D2H)",
                             "roofline": {"bound": "hbm", "achieved": DET_MB_PER_FRAME * 16e-3 / td_, "peak": hbm_gbs, "unit": "GB/s",
                                          "frac": DET_MB_PER_FRAME * 16e-3 / td_ / hbm_gbs, "note": "whole detector step incl. H2D, not one kernel"}}
            # the detector's kernels alone, frames resident in HBM (what the pipeline runs: its H2D rides the copy stream)
            # (an explicit stream: the library reads a NULL stream as "the handle's own stream", which torch's events would not see)
            tstream = torch.cuda.Stream(device=dev)
            st = tstream.cuda_stream
            for b in sorted({16, frames_batch}):
                fd = frames[:b].to(dev)
                bx = torch.empty((b, 4, 5), dtype=torch.int32, device=dev)
                ct = torch.empty((b,), dtype=torch.int32, device=dev)
                torch.cuda.synchronize()
                tdd = _event_time(torch, lambda: det.run_dev(fd, bx, ct, stream=st), reps, stream=tstream)
                out["detect" if b == frames_batch else f"detect_b{b}_dev"] = {"batch": b, "ms": tdd * 1e3, "frames_per_s": b / tdd, "us_per_frame": tdd * 1e6 / b,
                                            "inputs": "device-resident u8 HWC frames",
                                            "roofline": {"bound": "hbm", "achieved": DET_MB_PER_FRAME * b * 1e-3 / tdd, "peak": hbm_gbs, "unit": "GB/s",
                                                         "frac": DET_MB_PER_FRAME * b * 1e-3 / tdd / hbm_gbs,
                                                         "note": "whole detector forward + decode + NMS (all kernels), device-resident; algorithmic "
                                                                 "bytes = layer-wise fp16 activation traffic + the u8 frame"}}
            for b in (32, 256):
                crops = np.ascontiguousarray(np.concatenate([mg.arcface_inputs()] * ((b + 7) // 8))[:b])
                x = torch.from_numpy((crops[..., ::-1].transpose(0, 3, 1, 2).astype(np.float32) - 127.5) * 0.0078125).to(dev).contiguous()
                y = torch.empty((b, 512), dtype=torch.float32, device=dev)
                torch.cuda.synchronize()
                te = _event_time(torch, lambda: emb.run_dev(x, y, stream=st), reps, stream=tstream)
                out[f"embed_b{b}"] = {"batch": b, "mode": "ir_se", "ms": te * 1e3, "faces_per_s": b / te, "inputs": "device-resident f32 CHW",
                                      "roofline": {"bound": "tensor", "achieved": EMB_GFLOP_PER_FACE * b * 1e-3 / te, "peak": tf_sust, "unit": "TFLOP/s",
                                                   "frac": EMB_GFLOP_PER_FACE * b * 1e-3 / te / tf_sust, "note": "whole embedder forward (all kernels), device-resident"}}
            crops32 = np.ascontiguousarray(np.concatenate([mg.arcface_inputs()] * 4))
            te = _event_time(torch, lambda: emb.run_crops(crops32), reps)
            out["embed"] = {"batch": 32, "mode": "ir_se", "ms": te * 1e3, "faces_per_s": 32 / te,
                            "roofline": {"bound": "tensor", "achieved": EMB_GFLOP_PER_FACE * 32e-3 / te, "peak": tf_sust, "unit": "TFLOP/s",
                                         "frac": EMB_GFLOP_PER_FACE * 32e-3 / te / tf_sust, "note": "whole embedder step incl. H2D/D2H (host-buffer call), not one kernel"}}
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
    from tools import synth_weights as sw
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
