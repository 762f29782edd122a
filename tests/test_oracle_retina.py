"""CPU: the detection oracle (oracle/retina_oracle.py + oracle/retina_post.c) against the golden vectors of the reference's
own PyTorch modules and against hand-computed cases of RetinaFace::postprocessing (src/retinaface.cpp:154-271)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import retina_oracle as ro  # noqa: E402
from tools import synth_weights as sw  # noqa: E402
from tools import make_golden_retina as mgr  # noqa: E402
from tools import pack_retina as pr  # noqa: E402
from tools import pack_weights as pw  # noqa: E402

GOLD = ROOT / "tests" / "golden"


@pytest.mark.parametrize("full", [False, True])
def test_forward_matches_reference_golden(full):
    gold = np.load(GOLD / f"retina_{'full' if full else 'trim'}_seed11.npz")
    sd = ro.to_torch(sw.retina_state_dict(full, 11, float(gold["cls_shift"])))
    for h, w, n, sub in ((96, 128, 2, 1), (640, 640, 1, int(gold["sub"]))):
        frames = mgr.det_frames(n, h, w)
        x = torch.from_numpy(np.stack([ro.preprocess(f, h, w) for f in frames]))
        loc, conf, landm = ro.forward(sd, x, full)
        key = f"{h}x{w}"
        assert loc.shape[1] == ro.num_anchors(h, w)
        assert np.abs(loc.numpy()[:, ::sub] - gold[key + ".loc"]).max() <= 1e-5
        assert np.abs(conf.numpy()[:, ::sub] - gold[key + ".conf"]).max() <= 1e-6
        if full:
            assert np.abs(landm.numpy()[:, ::sub] - gold[key + ".landm"]).max() <= 1e-5
        assert np.array_equal((conf.numpy()[..., 1] > 0.6).sum(axis=1), gold[key + ".n_pass"])


def test_preprocess_identity_and_letterbox_geometry():
    f = mgr.det_frames(1, 96, 128)[0]
    x = ro.preprocess(f, 96, 128)
    assert x.shape == (3, 96, 128)
    # planar B,G,R (no RGB swap), mean (104,117,123) subtracted (src/retinaface.cpp:129-135)
    assert np.array_equal(x[0], f[..., 0].astype(np.float32) - 104) and np.array_equal(x[2], f[..., 2].astype(np.float32) - 123)
    # shipped config: 640x480 frame into a 288x320 net (app/config.json:3-8): scale_h = 0.6 > scale_w = 0.5 -> w = 320, h = 240, y = 24
    assert ro.letterbox_params(480, 640, 288, 320) == (320, 240, 0, 24)
    g = ro.preprocess(np.full((480, 640, 3), 200, np.uint8), 288, 320)
    assert g[0, 0, 0] == 128 - 104 and g[0, 24, 0] == 200 - 104 and g[0, 263, 5] == 200 - 104 and g[0, 264, 5] == 128 - 104


def test_anchor_layout_matches_reference_formula():
    a = ro.anchors(640, 640)
    assert a.shape == (16800, 4)
    # level 0, cell (0,0): min_sizes 10, 20; cx = (0+0.5)*8/640
    assert np.allclose(a[0], [0.5 * 8 / 640, 0.5 * 8 / 640, 10 / 640, 10 / 640]) and np.allclose(a[1][2:], [20 / 640, 20 / 640])
    # level 1 starts after 80*80*2 anchors; x (cx) runs fastest
    assert np.allclose(a[12800], [0.5 * 16 / 640, 0.5 * 16 / 640, 32 / 640, 32 / 640])
    assert np.allclose(a[12802], [1.5 * 16 / 640, 0.5 * 16 / 640, 32 / 640, 32 / 640])
    b = ro.anchors(288, 320)
    assert b.shape[0] == ro.num_anchors(288, 320) == 3780
    assert np.allclose(b[0], [4 / 320, 4 / 288, 10 / 320, 10 / 288])


def test_postprocess_hand_computed_cases():
    A = ro.num_anchors(640, 640)
    loc = np.zeros((A, 4), np.float32)
    conf = np.zeros((A, 2), np.float32)
    conf[:, 0] = 1
    # anchor 12800+2*(40*10+20) = level 1 cell (row 10, col 20), size 32: cx = 20.5*16/640, cy = 10.5*16/640
    a = 12800 + 2 * (40 * 10 + 20)
    conf[a] = (0.1, 0.9)
    boxes, _, ids = ro.postprocess(loc, conf, None, 640, 640, 640, 640, 0.4, 0.6, 4)
    assert ids.tolist() == [a]
    x1, y1, x2, y2, s = boxes[0]
    # y = column from cx, x = row from cy (src/retinaface.cpp:165,171-174); truncation toward zero
    # float32 arithmetic exactly as :166-174 (loc = 0: tmp1 == anchor)
    f = np.float32
    cx, cy, sx = f(20.5 * 16.0 / 640), f(10.5 * 16.0 / 640), f(32 * 1.0 / 640)
    assert (y1, y2) == (int(f(f(cx - f(sx / f(2))) * f(640))), int(f(f(cx + f(sx / f(2))) * f(640))))
    assert (x1, x2) == (int(f(f(cy - f(sx / f(2))) * f(640))), int(f(f(cy + f(sx / f(2))) * f(640))))
    assert abs(y1 - 312) <= 1 and abs(x2 - 184) <= 1
    assert abs(s - 0.9) < 1e-7
    # score exactly at the threshold is rejected (strict >), just above is kept
    conf[a] = (0.4, 0.6)
    assert ro.postprocess(loc, conf, None, 640, 640, 640, 640, 0.4, 0.6, 4)[0] == []
    # NMS: the same-cell second anchor (size 64) overlaps the first one: IoU = 33*33/(65*65) = 0.258 < 0.4 -> both kept;
    # with threshold 0.25 the weaker one is suppressed ('>=')
    conf[a] = (0.1, 0.9)
    conf[a + 1] = (0.2, 0.8)
    assert len(ro.postprocess(loc, conf, None, 640, 640, 640, 640, 0.4, 0.6, 4)[0]) == 2
    assert ro.postprocess(loc, conf, None, 640, 640, 640, 640, 0.25, 0.6, 4)[2].tolist() == [a]
    # cap to max_faces after NMS, order by score
    far = [12800 + 2 * (40 * r + 3) for r in (1, 9, 17, 25, 33)]
    for i, idx in enumerate(far):
        conf[idx] = (0.05 * i, 1 - 0.05 * i - 0.01)
    res = ro.postprocess(loc, conf, None, 640, 640, 640, 640, 0.4, 0.6, 4)
    assert len(res[0]) == 4 and res[2].tolist()[:2] == [far[0], far[1]]
    # clipping to the frame (src/retinaface.cpp:190-193) and exp() decode
    loc[0] = (0, 0, 25, 25)
    conf[0] = (0, 1)
    b0 = ro.postprocess(loc, conf, None, 640, 640, 640, 640, 0.4, 0.6, 1)[0][0]
    assert b0[:4] == (0, 0, 639, 639)


def test_postprocess_letterboxed_frame_rescale():
    # 640x480 frame in a 288x320 net: boxes are mapped back with the smaller scale and the vertical pad of 24 rows
    A = ro.num_anchors(288, 320)
    loc = np.zeros((A, 4), np.float32)
    conf = np.zeros((A, 2), np.float32)
    conf[:, 0] = 1
    a = 2 * (40 * 18 + 20)  # level 0 (stride 8, 36x40 cells), row 18, col 20, size 10
    conf[a] = (0.1, 0.9)
    (x1, y1, x2, y2, _), = ro.postprocess(loc, conf, None, 288, 320, 480, 640, 0.4, 0.6, 4)[0]
    cx, cy, sx, sy = ro.anchors(288, 320)[a]
    ny1, nx1 = int(np.float32(cx - sx / 2) * 320), int(np.float32(cy - sy / 2) * 288)
    assert y1 == int(np.float32(ny1) / np.float32(0.5)) and x1 == int((np.float32(nx1) - np.float32(24.0)) / np.float32(0.5))


@pytest.mark.parametrize("full", [False, True])
def test_retina_packer_tensors(full, tmp_path):
    sd = sw.retina_state_dict(full, 11, -4.5)
    f = tmp_path / "det.frw"
    pr.save_retina(f, sd, full)
    kind, T = pw.read_file(f)
    assert kind == (pw.KIND_RETINA_FULL if full else pw.KIND_RETINA_TRIM)
    assert T["stem.w"].shape == (8, 27) and T["dw1.w"].shape == (9, 8) and T["pw1.w"].shape == (8, 16) and T["pw5.w"].shape == (64, 64)
    assert T["pw5.w"].dtype == np.float16 and T["ssh1.b.w"].shape == (9, 16, 16) and T["head2.w"].shape == (32, 64)
    assert bool(np.any(T["head1.w"][12:] != 0)) == full
    # folded stem equals conv + BN of the state dict at one pixel
    x = np.random.default_rng(0).standard_normal((1, 3, 8, 8)).astype(np.float32)
    import torch.nn.functional as F

    ref = ro._cbr(torch.from_numpy(x), ro.to_torch(sd), "body.stage1.0", 2, relu=False).numpy()
    w = torch.from_numpy(T["stem.w"].reshape(8, 3, 3, 3).transpose(0, 3, 1, 2).copy())
    got = F.conv2d(torch.from_numpy(x), w, torch.from_numpy(T["stem.b"].copy()), 2, 1).numpy()
    assert np.abs(ref - got).max() < 1e-4
