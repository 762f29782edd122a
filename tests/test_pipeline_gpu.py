"""GPU: the glue between detector and embedder (getCroppedFaces, /root/reference src/arcface.cpp:3-17) and the end-to-end
pipeline detect -> crop -> embed -> search (src/app.cpp:293-352) against the chained oracles.
  crops       vs cv2.resize(INTER_CUBIC) (the OpenCV in this image, 4.13, its own generic u8 path: IPP off): 0 differing pixels;
              vs cv2 with the closed-source IPP accelerator on (which differs from OpenCV's own path): <= 1 LSB
  embeddings  of GPU crops vs the fp32 oracle on cv2 crops: |d| <= 1e-3 (BASELINE.json north_star)
  identities  top-1 index exact for every detected face (gallery rows planted from the oracle's embeddings)
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import frb200
from oracle import arcface_oracle as ao
from oracle import retina_oracle as ro
from oracle import search_oracle as so
from tools import synth_weights as sw

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools import make_golden_retina as mgr  # noqa: E402
from tools import pack_retina as pr  # noqa: E402
from tools import pack_weights as pw  # noqa: E402

pytestmark = pytest.mark.gpu


def cv2_crops(frame, boxes, ipp=False):
    """getCroppedFaces restated with cv2 (src/arcface.cpp:5-10): Rect(Point(y1,x1), Point(y2,x2)) -> resize to 112x112, INTER_CUBIC.
    ipp=False: OpenCV's own u8 bicubic (the byte-exact target); True: with the IPP accelerator, as cv2 runs by default here."""
    import cv2

    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(bool(ipp))
    try:
        return _cv2_crops(frame, boxes)
    finally:
        cv2.ipp.setUseIPP(was)


def _cv2_crops(frame, boxes):
    import cv2

    out = []
    for b in boxes:
        r0, r1 = sorted((int(b["x1"]), int(b["x2"])))
        c0, c1 = sorted((int(b["y1"]), int(b["y2"])))
        roi = frame[r0:max(r1, r0 + 1), c0:max(c1, c0 + 1)]
        out.append(cv2.resize(roi, (112, 112), interpolation=cv2.INTER_CUBIC))
    return np.stack(out)


@pytest.fixture(scope="module")
def nets(tmp_path_factory):
    d = tmp_path_factory.mktemp("e2e")
    det_sd = sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT)
    arc_sd = sw.arcface_state_dict("ir_se", 7)
    pr.save_retina(d / "det.frw", det_sd, False)
    pw.save_arcface(d / "arc.frw", arc_sd, "ir_se")
    det = frb200.Detector(d / "det.frw", (640, 640), max_batch=8, max_faces=4)
    emb = frb200.Embedder(d / "arc.frw", max_batch=16)
    yield det, emb, arc_sd
    det.close()
    emb.close()


def test_crop_resize_matches_opencv_bicubic(nets):
    det, emb, arc_sd = nets
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    boxes = np.zeros(7, frb200.BOX_DTYPE)
    for i, (x1, y1, x2, y2) in enumerate([(10, 20, 35, 46), (0, 0, 112, 112), (100, 300, 300, 450), (470, 600, 479, 639), (5, 5, 6, 6),
                                          (200, 100, 260, 147), (0, 0, 479, 639)]):
        boxes[i] = (x1, y1, x2, y2, 0.9)
    out, crops = frb200.embed_boxes(emb, frame, boxes, want_crops=True)
    want = cv2_crops(frame, boxes)
    assert np.array_equal(crops, want), f"{int((crops != want).sum())} pixels differ from OpenCV's u8 bicubic"
    from oracle import cv_resize as cr

    assert np.array_equal(crops, cr.cropped_faces(frame, boxes))
    d = np.abs(crops.astype(int) - cv2_crops(frame, boxes, ipp=True).astype(int))
    print(f"crop vs cv2 with IPP: max |d| = {d.max()}, differing pixels = {float((d > 0).mean()):.4%}")
    assert d.max() <= 1
    assert np.array_equal(crops[1], frame[0:112, 0:112])  # 112x112 ROI: identity resize
    # high-contrast frame (overshoot saturates), random boxes including 1- and 2-pixel ROIs
    hc = (rng.integers(0, 2, (480, 640, 3)) * 255).astype(np.uint8)
    rb = np.zeros(16, frb200.BOX_DTYPE)
    for i in range(16):
        x1, y1 = int(rng.integers(0, 470)), int(rng.integers(0, 630))
        hgt, wid = (int(rng.integers(1, 4)), int(rng.integers(1, 4))) if i < 4 else (int(rng.integers(4, 479 - x1 + 2)), int(rng.integers(4, 639 - y1 + 2)))
        rb[i] = (x1, y1, min(x1 + hgt, 479), min(y1 + wid, 639), 0.5)
    _, c2 = frb200.embed_boxes(emb, hc, rb, want_crops=True)
    assert np.array_equal(c2, cv2_crops(hc, rb))
    # embeddings of the GPU crops == embeddings of the same crops fed as u8 (bitwise), and close to the oracle on cv2 crops
    again = emb.run_crops(crops)
    assert np.array_equal(again.view(np.uint32), out.view(np.uint32))
    ref = ao.forward(ao.to_torch(arc_sd), torch.from_numpy(ao.preprocess_faces(want[:3])), "ir_se").numpy()
    assert np.abs(out[:3] - ref).max() <= 1e-3


def test_end_to_end_identities(nets):
    det, emb, arc_sd = nets
    frames = mgr.det_frames(6, 640, 640, seed=13)
    # oracle chain on the GPU detector's boxes: cv2 crops -> fp32 embeddings; plant them (noisy) into a gallery
    boxes, counts, _ = det.run(frames)
    assert counts.tolist() == [4] * 6
    oracle_emb = []
    for f in range(2):
        crops = cv2_crops(frames[f], boxes[f, : counts[f]])
        oracle_emb.append(ao.forward(ao.to_torch(arc_sd), torch.from_numpy(ao.preprocess_faces(crops)), "ir_se").numpy())
    oracle_emb = np.concatenate(oracle_emb)  # 8 faces of frames 0 and 1
    n_gal = 50_000
    G = so.synth_rows(np.arange(n_gal), seed=17)
    planted = np.array([11, 4097, 20_000, 33_333, 49_999, 256, 255, 7])
    G[planted] = so.planted_queries(oracle_emb, noise=0.3, seed=3)
    gal = frb200.Gallery.from_rows(G)
    pipe = frb200.Pipeline(det, emb, gal)
    res = pipe.run(frames, want_embeddings=True)
    assert np.array_equal(res["counts"], counts) and np.array_equal(res["boxes"], boxes)
    got_emb = res["embeddings"][:2].reshape(8, 512)
    assert np.abs(got_emb - oracle_emb).max() <= 1e-3
    # identities: exact, and equal to the oracle's search on the oracle's embeddings
    want_idx, want_score = so.get_outputs(so.sims(G, oracle_emb))
    assert np.array_equal(want_idx, planted)
    assert np.array_equal(res["idx"][:2].reshape(-1), planted)
    assert np.abs(res["score"][:2].reshape(-1) - want_score).max() <= 1e-3
    # the remaining frames: the pipeline's own embeddings searched by the oracle give the same identities
    e_all = res["embeddings"].reshape(-1, 512)
    oi, ov = so.get_outputs(so.sims(G, e_all))
    assert np.array_equal(res["idx"].reshape(-1), oi)
    assert np.abs(res["score"].reshape(-1) - ov).max() <= 1e-5
    # no gallery: idx = -1
    pipe2 = frb200.Pipeline(det, emb, None)
    r2 = pipe2.run(frames[:1])
    assert np.all(r2["idx"] == -1) and r2["counts"][0] == 4
    pipe.close()
    pipe2.close()
    gal.close()


def test_two_batches_in_flight_equal_synchronous_runs(nets):
    """fr_pipeline_submit / fr_pipeline_collect (two batches in flight, speculative face count) return exactly what the synchronous
    call returns for the same frames — including a batch with MORE faces than the one before it (the guess is too small and the rest is
    embedded at collect time), a batch with fewer, and the error cases of the ring."""
    det, emb, arc_sd = nets
    frames = mgr.det_frames(6, 640, 640, seed=13)
    blank = np.full((2, 640, 640, 3), 128, np.uint8)  # no detections
    batches = [frames[:2], blank, frames[:6], frames[2:3], blank[:1], frames[1:5]]
    G = so.synth_rows(np.arange(30_000), seed=5)
    gal = frb200.Gallery.from_rows(G)
    pipe = frb200.Pipeline(det, emb, gal)
    want = [pipe.run(b, want_embeddings=True) for b in batches]
    assert [int(w["counts"].sum()) for w in want] == [8, 0, 24, 4, 0, 16]
    got = []
    pipe.submit(batches[0], want_embeddings=True)
    for b in batches[1:]:
        pipe.submit(b, want_embeddings=True)
        assert pipe.in_flight() == 2
        with pytest.raises(frb200.FrError) as e:
            pipe.submit(b)
        assert e.value.code == frb200.FR_ESTATE
        got.append(pipe.collect())
    got.append(pipe.collect())
    assert pipe.in_flight() == 0
    with pytest.raises((frb200.FrError, RuntimeError)):
        pipe.collect()
    for w, g in zip(want, got):
        for key in ("counts", "idx"):
            assert np.array_equal(w[key], g[key]), key
        assert np.array_equal(w["boxes"], g["boxes"])
        assert np.array_equal(w["score"].view(np.uint32), g["score"].view(np.uint32))
        assert np.array_equal(w["embeddings"].view(np.uint32), g["embeddings"].view(np.uint32))
    pipe.close()
    gal.close()


def test_service_batches_concurrent_requests(nets):
    """fr_service_infer from 12 threads at once: every caller gets exactly what the pipeline returns for its own frame, and the worker
    really formed batches (fewer batches than frames)."""
    import threading

    det, emb, arc_sd = nets
    frames = mgr.det_frames(6, 640, 640, seed=13)
    blank = np.full((640, 640, 3), 128, np.uint8)
    reqs = [frames[i % 6] for i in range(20)] + [blank] * 4
    G = so.synth_rows(np.arange(30_000), seed=5)
    gal = frb200.Gallery.from_rows(G)
    pipe = frb200.Pipeline(det, emb, gal)
    want = {}
    for i in list(range(6)) + [20]:
        r = pipe.run(reqs[i][None])
        want[i] = (r["boxes"][0], int(r["counts"][0]), r["idx"][0], r["score"][0])
    svc = frb200.Service(pipe, max_wait_us=2000)
    got = [None] * len(reqs)

    def call(i):
        got[i] = svc.infer(reqs[i])

    for wave in (range(0, 12), range(12, 24)):
        ts = [threading.Thread(target=call, args=(i,)) for i in wave]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    for i, g in enumerate(got):
        b, c, ix, sc = want[i % 6] if i < 20 else want[20]
        assert g["count"] == c
        assert np.array_equal(g["boxes"][:c], b[:c])
        assert np.array_equal(g["idx"], ix)
        assert np.array_equal(g["score"].view(np.uint32), sc.view(np.uint32))
    batches, served = svc.stats()
    assert served == len(reqs) and batches < served
    svc.close()
    pipe.close()
    gal.close()
