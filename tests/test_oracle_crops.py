"""CPU: pins oracle/cv_resize.py (the restatement of OpenCV's u8 INTER_CUBIC behind getCroppedFaces, /root/reference
src/arcface.cpp:9) against the cv2 wheel of this image with IPP off: 0 differing pixels, including 1-pixel ROIs, extreme contrast
(overshoot -> saturation) and non-square ROIs. With IPP on, cv2 differs from its OWN generic path by 1 LSB — reported, not a target."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import cv_resize as cr  # noqa: E402


@pytest.fixture()
def no_ipp():
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    yield
    cv2.ipp.setUseIPP(was)


def test_restatement_is_byte_exact_vs_cv2_generic_path(no_ipp):
    rng = np.random.default_rng(1)
    total = 0
    for t in range(120):
        h, w = (int(rng.integers(1, 8)), int(rng.integers(1, 8))) if t < 30 else (int(rng.integers(1, 400)), int(rng.integers(1, 400)))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if t % 3 == 0:
            img = (img // 128 * 255).astype(np.uint8)
        want = cv2.resize(img, (112, 112), interpolation=cv2.INTER_CUBIC)
        got = cr.resize_cubic_u8(img)
        assert np.array_equal(got, want), (t, h, w, int((got != want).sum()))
        total += want.size
    assert total > 4_000_000


def test_identity_and_roi_semantics(no_ipp):
    rng = np.random.default_rng(2)
    frame = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
    boxes = np.zeros(3, dtype=[("x1", "<i4"), ("y1", "<i4"), ("x2", "<i4"), ("y2", "<i4"), ("score", "<f4")])
    boxes[0] = (0, 0, 112, 112, 1)        # rows [0,112), columns [0,112): identity
    boxes[1] = (10, 200, 60, 230, 1)      # x = row, y = column (src/arcface.cpp:6)
    boxes[2] = (239, 319, 239, 319, 1)    # empty ROI -> 1 pixel
    c = cr.cropped_faces(frame, boxes)
    assert np.array_equal(c[0], frame[:112, :112])
    assert np.array_equal(c[1], cv2.resize(frame[10:60, 200:230], (112, 112), interpolation=cv2.INTER_CUBIC))
    assert np.all(c[2] == frame[239, 319])


def test_ipp_path_differs_from_opencv_generic_by_one_lsb():
    # documents why the oracle is "cv2 with IPP off": the accelerator is closed source and not bit-identical to OpenCV's own code
    if not cv2.ipp.useIPP():
        pytest.skip("this cv2 build has no IPP")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (200, 300, 3), dtype=np.uint8)
    d = np.abs(cv2.resize(img, (112, 112), interpolation=cv2.INTER_CUBIC).astype(int) - cr.resize_cubic_u8(img).astype(int))
    assert d.max() <= 1
