"""CPU: pins the detector post-processing oracle to the REFERENCE ITSELF. oracle/_ref/libref_retina.so is
/root/reference/src/retinaface.cpp compiled verbatim (oracle/build_ref.py); these tests require oracle/retina_post.c (the plain-C
restatement every other test uses) to agree with it bit for bit on constructor sizes (a2), anchors (a5), decode / threshold /
rescale / clip (a6) and NMS + cap (a7) — random heads, both letterbox branches, clipping, threshold edges, ties."""
import numpy as np
import pytest

from oracle import ref_retina as rr
from oracle import retina_oracle as ro

pytestmark = pytest.mark.skipif(not rr.available(), reason="oracle/_ref/libref_retina.so not built (needs /root/reference at build time)")

# (net_h, net_w, frame_h, frame_w): identity, the shipped 640x480 -> 288x320 config (scale_h > scale_w), the other branch, a tall frame
GEOMS = [(640, 640, 640, 640), (288, 320, 480, 640), (320, 288, 640, 480), (96, 128, 1080, 1920), (640, 640, 720, 1280), (160, 160, 90, 300)]


def _heads(a, seed, frac=0.01, loc_scale=1.5):
    rng = np.random.default_rng(seed)
    loc = (rng.standard_normal((a, 4)) * loc_scale).astype(np.float32)
    p = rng.random(a).astype(np.float32)
    p = np.where(rng.random(a) < frac, 0.6 + 0.4 * p, 0.6 * p).astype(np.float32)
    conf = np.stack([1 - p, p], axis=1).astype(np.float32)
    return loc, conf


@pytest.mark.parametrize("geom", GEOMS)
def test_constructor_sizes_and_anchors(geom):
    nh, nw, fh, fw = geom
    r = rr.RefRetinaFace(nh, nw, fh, fw)
    assert r.output_size_base == ro.num_anchors(nh, nw)                     # m_OUTPUT_SIZE_BASE, src/retinaface.cpp:13
    sh, sw_ = r.scales
    assert (sh, sw_) == (float(np.float32(nh) / np.float32(fh)), float(np.float32(nw) / np.float32(fw)))  # :21-22
    a_ref = r.anchors()
    a_c = ro.anchors(nh, nw)
    assert a_ref.shape == a_c.shape and np.array_equal(a_ref.view(np.uint32), a_c.view(np.uint32))
    r.close()


def test_anchors_non_multiple_of_32():
    # ceil() in create_anchor_retinaface (src/retinaface.cpp:214-217) matters when the input is not a multiple of the strides
    r = rr.RefRetinaFace(640, 640, 640, 640)
    for h, w in ((300, 500), (97, 131), (33, 65)):
        a_ref = r.anchors(h, w)
        fm = [(-(-h // s), -(-w // s)) for s in (8, 16, 32)]
        assert a_ref.shape[0] == sum(2 * a * b for a, b in fm)
        import ctypes as C
        out = np.empty_like(a_ref)
        n = ro._post_lib().retina_anchors(w, h, out.ctypes.data_as(C.c_void_p), out.shape[0])
        assert n == a_ref.shape[0] and np.array_equal(a_ref.view(np.uint32), out.view(np.uint32))
    r.close()


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_postprocessing_random_heads_bit_exact(geom, seed):
    nh, nw, fh, fw = geom
    a = ro.num_anchors(nh, nw)
    for max_faces, nms_thr, bbox_thr in ((4, 0.4, 0.6), (50, 0.3, 0.6), (1000, 0.5, 0.7)):
        r = rr.RefRetinaFace(nh, nw, fh, fw, max_faces=max_faces, nms_thr=nms_thr, bbox_thr=bbox_thr)
        loc, conf = _heads(a, 100 * seed + max_faces, frac=0.02 if a > 2000 else 0.2, loc_scale=1.5 + seed)
        want = r.postprocess(loc, conf)
        got, _, _ = ro.postprocess(loc, conf, None, nh, nw, fh, fw, nms_thr, bbox_thr, max_faces)
        assert len(want) > 0
        assert [w[:4] for w in want] == [g[:4] for g in got]
        assert np.array_equal(np.array([w[4] for w in want], np.float32).view(np.uint32), np.array([g[4] for g in got], np.float32).view(np.uint32))
        r.close()


def test_postprocessing_edges():
    nh = nw = 640
    a = ro.num_anchors(nh, nw)
    r = rr.RefRetinaFace(nh, nw, 640, 640, max_faces=8)
    loc = np.zeros((a, 4), np.float32)
    conf = np.zeros((a, 2), np.float32)
    conf[:, 0] = 1
    # nothing passes
    assert r.postprocess(loc, conf) == [] and ro.postprocess(loc, conf, None, nh, nw, 640, 640, 0.4, 0.6, 8)[0] == []
    # exactly at the threshold: rejected (strict >, src/retinaface.cpp:160); one ulp above: kept
    conf[5, 1] = np.float32(0.6)
    assert r.postprocess(loc, conf) == []
    conf[5, 1] = np.nextafter(np.float32(0.6), np.float32(1))
    assert len(r.postprocess(loc, conf)) == 1
    # huge exp(): clipped to the frame; tiny box; negative offsets
    loc[5] = (0, 0, 30, -30)
    loc[9] = (-40, 55, 3, 3)
    conf[9, 1] = 0.9
    loc[12801] = (2.5, -2.5, 0.3, 0.1)
    conf[12801, 1] = 0.8
    want = r.postprocess(loc, conf)
    got = ro.postprocess(loc, conf, None, nh, nw, 640, 640, 0.4, 0.6, 8)[0]
    assert want == got and len(want) == 3
    # ties: a handful of equal scores (std::sort on < 16 elements is an insertion sort: anchor order, like the restatement)
    conf[:, 1] = 0
    for i in (100, 2000, 9000, 12900, 16000, 16700):
        conf[i, 1] = 0.75
    want = r.postprocess(loc, conf)
    got = ro.postprocess(loc, conf, None, nh, nw, 640, 640, 0.4, 0.6, 8)[0]
    assert want == got and len(want) == 6
    r.close()


def test_nms_alone_random_boxes():
    r = rr.RefRetinaFace(640, 640, 640, 640)
    rng = np.random.default_rng(5)
    for trial in range(20):
        n = int(rng.integers(1, 120))
        x1 = rng.integers(0, 600, n)
        y1 = rng.integers(0, 600, n)
        boxes = [(int(a), int(b), int(a + rng.integers(0, 80)), int(b + rng.integers(0, 80)), float(s))
                 for a, b, s in zip(x1, y1, np.sort(rng.random(n).astype(np.float32))[::-1])]
        thr = float(rng.choice([0.1, 0.3, 0.4, 0.5]))
        want = r.nms(boxes, thr)
        # the restatement's NMS, driven through retina_postprocess is covered above; here an independent numpy greedy NMS with the
        # reference's arithmetic (+1 areas, float division, >=) confirms what the compiled reference does
        keep, dead = [], np.zeros(n, bool)
        f = np.float32
        area = [f((b[2] - b[0] + 1) * (b[3] - b[1] + 1)) for b in boxes]
        for i in range(n):
            if dead[i]:
                continue
            keep.append(boxes[i])
            for j in range(i + 1, n):
                if dead[j]:
                    continue
                w = max(f(0), f(min(boxes[i][2], boxes[j][2])) - f(max(boxes[i][0], boxes[j][0])) + f(1))
                h = max(f(0), f(min(boxes[i][3], boxes[j][3])) - f(max(boxes[i][1], boxes[j][1])) + f(1))
                inter = f(w * h)
                if f(inter / f(f(area[i] + area[j]) - inter)) >= f(thr):
                    dead[j] = True
        assert [w[:4] for w in want] == [k[:4] for k in keep], trial
    r.close()
