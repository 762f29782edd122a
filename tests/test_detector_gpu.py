"""GPU parity of the RetinaFace detector through the C ABI.
  network     raw loc / conf / landm vs the reference modules' golden vectors and the restated fp32 oracle: |d| <= 1e-3 on conf,
              and on loc / landm relative to their scale (they are O(1) regression outputs; BASELINE.json north_star: 1e-3)
  decode+NMS  Bbox integers bit-exact vs oracle/retina_post.c run on the GPU's OWN raw outputs; landmarks |d| <= 1e-3 px
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import frb200
from oracle import retina_oracle as ro
from tools import synth_weights as sw

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools import make_golden_retina as mgr  # noqa: E402
from tools import pack_retina as pr  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden"
TOL = 1e-3      # decoded boxes / landmarks (normalised image units)
RAW_TOL = 1e-2  # raw head outputs (O(1..5) regression values and probabilities) with fp16 activations, like the reference's fp16 engine


@pytest.fixture(scope="module", params=[False, True], ids=["trim", "full"])
def ckpt(request, tmp_path_factory):
    full = request.param
    sd = sw.retina_state_dict(full, 11, mgr.DET_CLS_SHIFT)
    f = tmp_path_factory.mktemp("det") / "retina.frw"
    pr.save_retina(f, sd, full)
    return full, sd, f


def _boxes_list(boxes, counts, i):
    return [(int(b["x1"]), int(b["y1"]), int(b["x2"]), int(b["y2"])) for b in boxes[i, : counts[i]]]


@pytest.mark.parametrize("hw,n", [((96, 128), 2), ((640, 640), 1)])
def test_raw_outputs_match_reference_golden(ckpt, hw, n):
    full, sd, f = ckpt
    gold = np.load(GOLD / f"retina_{'full' if full else 'trim'}_seed11.npz")
    sub = 1 if hw == (96, 128) else int(gold["sub"])
    det = frb200.Detector(f, hw, max_batch=4, landmarks=full)
    frames = mgr.det_frames(n, *hw)
    loc, conf, lm = det.raw(frames)
    key = f"{hw[0]}x{hw[1]}"
    assert det.anchors == ro.num_anchors(*hw) == loc.shape[1]
    e_conf = float(np.abs(conf[:, ::sub] - gold[key + ".conf"]).max())
    e_loc = float(np.abs(loc[:, ::sub] - gold[key + ".loc"]).max())
    print(f"{key} full={full}: max|d| conf {e_conf:.2e}, loc {e_loc:.2e} (loc scale {np.abs(gold[key + '.loc']).max():.2f})")
    if full:
        e_lm = float(np.abs(lm[:, ::sub] - gold[key + ".landm"]).max())
        print(f"   landm {e_lm:.2e} (scale {np.abs(gold[key + '.landm']).max():.2f})")
        assert e_lm <= RAW_TOL
    assert e_conf <= RAW_TOL and e_loc <= RAW_TOL
    # decoded boxes in normalised image units (what the 1e-3 of BASELINE.json's north_star is about): cx, cy, w, h
    a = ro.anchors(*hw)[::sub]
    def dec(l):
        return np.concatenate([a[None, :, :2] + l[..., :2] * 0.1 * a[None, :, 2:], a[None, :, 2:] * np.exp(l[..., 2:] * 0.2)], -1)
    e_box = float(np.abs(dec(loc[:, ::sub].astype(np.float64)) - dec(gold[key + ".loc"].astype(np.float64))).max())
    print(f"   decoded box max|d| {e_box:.2e} (normalised units)")
    assert e_box <= TOL
    # same tensors through the preprocessed-input hook (RetinaFace::preprocess output, src/retinaface.cpp:128-135)
    x = np.stack([ro.preprocess(fr, *hw) for fr in frames])
    loc2, conf2, _ = det.net(x)
    assert np.array_equal(loc2.view(np.uint32), loc.view(np.uint32)) and np.array_equal(conf2.view(np.uint32), conf.view(np.uint32))
    det.close()


def test_batch_and_oracle_agreement_640(ckpt):
    full, sd, f = ckpt
    det = frb200.Detector(f, (640, 640), max_batch=6, landmarks=full)
    frames = mgr.det_frames(5, 640, 640, seed=13)
    loc, conf, lm = det.raw(frames)
    x = torch.from_numpy(np.stack([ro.preprocess(fr, 640, 640) for fr in frames[:2]]))
    o_loc, o_conf, o_lm = ro.forward(ro.to_torch(sd), x, full)
    assert np.abs(conf[:2] - o_conf.numpy()).max() <= RAW_TOL
    assert np.abs(loc[:2] - o_loc.numpy()).max() <= RAW_TOL
    # batch independence: frame 3 alone gives the same bits
    l1, c1, _ = det.raw(frames[3:4])
    assert np.array_equal(l1[0].view(np.uint32), loc[3].view(np.uint32)) and np.array_equal(c1[0].view(np.uint32), conf[3].view(np.uint32))
    # findFace: boxes bit-exact vs the C restatement applied to the GPU's own raw outputs
    boxes, counts, glm = det.run(frames)
    for i in range(5):
        want, want_lm, ids = ro.postprocess(loc[i], conf[i], lm[i] if full else None, 640, 640, 640, 640, 0.4, 0.6, 4)
        assert _boxes_list(boxes, counts, i) == [w[:4] for w in want]
        assert np.array_equal(boxes["score"][i, : counts[i]], np.array([w[4] for w in want], np.float32))
        if full:
            assert np.abs(glm[i, : counts[i]] - want_lm).max() <= 1e-3
    assert counts.tolist() == [4] * 5  # the synthetic checkpoint saturates max_faces
    # ... and within 1 px of the oracle-on-oracle boxes for the two frames the fp32 oracle ran on
    for i in range(2):
        want, _, _ = ro.postprocess(o_loc[i].numpy(), o_conf[i].numpy(), None, 640, 640, 640, 640, 0.4, 0.6, 4)
        got = _boxes_list(boxes, counts, i)
        assert len(got) == len(want)
        assert max(abs(a - b) for g, w in zip(got, want) for a, b in zip(g, w[:4])) <= 1
    det.close()


def test_postprocess_hook_adversarial_cases(ckpt):
    full, sd, f = ckpt
    det = frb200.Detector(f, (640, 640), max_batch=4, max_faces=6, nms_thr=0.4, bbox_thr=0.6, landmarks=full)
    A = det.anchors
    rng = np.random.default_rng(3)
    loc = (rng.standard_normal((4, A, 4)) * 1.5).astype(np.float32)
    lm = rng.standard_normal((4, A, 10)).astype(np.float32)
    conf = np.zeros((4, A, 2), np.float32)
    p = rng.random((4, A)).astype(np.float32) ** 8          # a few hundred candidates above 0.6
    conf[..., 1], conf[..., 0] = p, 1 - p
    conf[1, :, 1] = 0.1                                       # image 1: nothing passes
    conf[1, :, 0] = 0.9
    conf[2, 100:140, 1] = 0.75                                # image 2: many equal scores (tie order = anchor order)
    conf[3, :, 1] = np.float32(0.6)                           # image 3: exactly at the threshold -> rejected (strict >)
    conf[3, 7, 1] = np.nextafter(np.float32(0.6), np.float32(1))
    loc[3, 7] = (0, 0, 30, -30)                               # huge / tiny box: clipping
    boxes, counts, glm = det.post(loc, conf, lm)
    for i in range(4):
        want, want_lm, ids = ro.postprocess(loc[i], conf[i], lm[i], 640, 640, 640, 640, 0.4, 0.6, 6)
        assert counts[i] == len(want)
        assert _boxes_list(boxes, counts, i) == [w[:4] for w in want], i
        if len(want):
            assert np.abs(glm[i, : counts[i]] - want_lm).max() <= 1e-3
    assert counts[1] == 0 and counts[3] == 1
    det.close()


@pytest.mark.parametrize("geom", [(640, 640, 640, 640), (288, 320, 480, 640), (320, 288, 640, 480), (640, 640, 720, 1280)])
def test_decode_nms_kernel_matches_compiled_reference(ckpt, geom):
    """det_decode_nms_kernel vs the reference's OWN RetinaFace::postprocessing / create_anchor_retinaface / nms
    (/root/reference/src/retinaface.cpp:154-271 compiled verbatim into oracle/_ref/libref_retina.so): Bbox integers and scores
    bit-exact on random heads, both letterbox branches, clipping."""
    from oracle import ref_retina as rr

    if not rr.available():
        pytest.skip("oracle/_ref/libref_retina.so not built")
    full, sd, f = ckpt
    nh, nw, fh, fw = geom
    for max_faces, nms_thr, bbox_thr, seed in ((4, 0.4, 0.6, 0), (32, 0.3, 0.7, 1)):
        det = frb200.Detector(f, (nh, nw), frame_hw=(fh, fw), max_batch=3, max_faces=max_faces, nms_thr=nms_thr, bbox_thr=bbox_thr, landmarks=full)
        ref = rr.RefRetinaFace(nh, nw, fh, fw, max_faces=max_faces, nms_thr=nms_thr, bbox_thr=bbox_thr)
        assert det.anchors == ref.output_size_base
        A = det.anchors
        rng = np.random.default_rng(17 + seed)
        loc = (rng.standard_normal((3, A, 4)) * (1.5 + seed)).astype(np.float32)
        p = rng.random((3, A)).astype(np.float32)
        p = np.where(rng.random((3, A)) < 0.02, 0.6 + 0.4 * p, 0.6 * p).astype(np.float32)
        conf = np.stack([1 - p, p], axis=-1).astype(np.float32)
        lm = rng.standard_normal((3, A, 10)).astype(np.float32) if full else None
        boxes, counts, _ = det.post(loc, conf, lm)
        for i in range(3):
            want = ref.postprocess(loc[i], conf[i])
            assert counts[i] == len(want) > 0
            assert _boxes_list(boxes, counts, i) == [w[:4] for w in want], (geom, i)
            assert np.array_equal(boxes["score"][i, : counts[i]].view(np.uint32), np.array([w[4] for w in want], np.float32).view(np.uint32))
        ref.close()
        det.close()


def test_detector_errors(ckpt, tmp_path):
    full, sd, f = ckpt
    with pytest.raises(frb200.FrError) as e:
        frb200.Detector(tmp_path / "nope.frw", (640, 640))
    assert e.value.code == frb200.FR_ENOENT and "Cant find engine file" in e.value.msg  # src/retinaface.cpp:53
    with pytest.raises(frb200.FrError) as e:
        frb200.Detector(f, (100, 640))
    assert e.value.code == frb200.FR_EINVAL
    if not full:
        with pytest.raises(frb200.FrError) as e:
            frb200.Detector(f, (640, 640), landmarks=True)
        assert e.value.code == frb200.FR_EFORMAT


def test_letterboxed_frames_shipped_config(ckpt):
    # the reference's shipped operating point (app/config.json:3-8): 640x480 frames into a 3x288x320 network. The GPU letterbox
    # (bilinear resize + pad 128) must reproduce RetinaFace::preprocess (cv2 on the oracle side) and the boxes must map back.
    full, sd, f = ckpt
    det = frb200.Detector(f, (288, 320), frame_hw=(480, 640), max_batch=3, landmarks=full)
    rng = np.random.default_rng(21)
    # smooth-ish frames (upsampled noise) so that the resize is non-trivial
    small = rng.integers(0, 256, (3, 60, 80, 3), dtype=np.uint8)
    frames = np.ascontiguousarray(np.repeat(np.repeat(small, 8, axis=1), 8, axis=2))
    frames = (frames.astype(np.int32) + rng.integers(-20, 21, frames.shape)).clip(0, 255).astype(np.uint8)
    loc, conf, lm = det.raw(frames)
    assert loc.shape[1] == ro.num_anchors(288, 320) == 3780
    # network on the oracle's preprocessing output (cv2 resize): same tensors up to the resize's rounding
    x = np.stack([ro.preprocess(fr, 288, 320) for fr in frames])
    loc2, conf2, _ = det.net(x)
    d = np.abs(conf - conf2).max()
    print(f"letterbox: max|d conf| GPU-resize vs cv2-resize input = {d:.2e}")
    assert d <= 2e-3 and np.abs(loc - loc2).max() <= 2e-2
    boxes, counts, _ = det.run(frames)
    for i in range(3):
        want, _, _ = ro.postprocess(loc[i], conf[i], None, 288, 320, 480, 640, 0.4, 0.6, 4)
        assert _boxes_list(boxes, counts, i) == [w[:4] for w in want]
        for b in boxes[i, : counts[i]]:
            assert 0 <= b["x1"] <= b["x2"] <= 479 and 0 <= b["y1"] <= b["y2"] <= 639
    det.close()
