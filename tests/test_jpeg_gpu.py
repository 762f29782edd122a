"""fr_jpeg_* (the JPEG decode of the serving loop, /root/reference src/app.cpp:247-256,294-301 = cv::imdecode + cv::resize).
Oracle: cv2 (libjpeg-turbo). Two different third-party JPEG decoders agree only up to their IDCT / chroma-upsampling rounding, so the
decode is compared with a stated tolerance; the resize leg is the library's own OpenCV-exact kernel and is checked bit for bit."""
import numpy as np
import pytest

import frb200

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu


def _picture(h, w, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.stack([127 + 100 * np.sin(xx / 37.0 + 0.3 * c) * np.cos(yy / 51.0 - 0.2 * c) for c in range(3)], axis=-1)
    img += rng.normal(0, 3.0, img.shape)
    img = np.ascontiguousarray(np.clip(img, 0, 255).astype(np.uint8))
    cv2.rectangle(img, (w // 5, h // 4), (w // 2, h // 2), (30, 200, 90), -1)
    cv2.circle(img, (3 * w // 4, 2 * h // 3), min(h, w) // 6, (220, 40, 160), -1)
    return img


@pytest.fixture(scope="module")
def dec():
    d = frb200.JpegDecoder()
    yield d
    d.close()


@pytest.mark.parametrize("sampling,mean_tol,max_tol", [("444", 1.0, 8), ("420", 1.5, None)])
def test_decode_matches_cv2_imdecode(dec, sampling, mean_tol, max_tol):
    img = _picture(480, 640, seed=1)
    flag = getattr(cv2, f"IMWRITE_JPEG_SAMPLING_FACTOR_{sampling}")
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, flag])
    assert ok
    jpeg = enc.tobytes()
    assert dec.info(jpeg) == (640, 480)
    want = cv2.imdecode(enc, cv2.IMREAD_UNCHANGED)          # src/app.cpp:249,296
    got = dec.decode(jpeg)
    assert got.shape == want.shape == (480, 640, 3) and got.dtype == np.uint8
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    print(f"jpeg {sampling}: mean |d| = {d.mean():.3f}, max |d| = {d.max()}, differing = {100 * (d > 0).mean():.1f} %")
    assert d.mean() <= mean_tol                              # BGR channel order, geometry and colour conversion agree
    if max_tol is not None:
        assert d.max() <= max_tol                            # IDCT rounding only (no chroma upsampling at 4:4:4)
    assert np.abs(got.astype(np.int32) - img.astype(np.int32)).mean() < 4.0   # and both are the picture that was encoded


def test_resize_leg_is_opencv_exact(dec):
    # cv::resize(frame, frame, Size(videoFrameWidth, videoFrameHeight)) (src/app.cpp:301): same decoded pixels in, same bytes out
    img = _picture(720, 960, seed=2)
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 90])
    jpeg = enc.tobytes()
    full = dec.decode(jpeg)
    small = dec.decode(jpeg, size=(640, 480))
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        want = cv2.resize(full, (640, 480), interpolation=cv2.INTER_LINEAR)
    finally:
        cv2.ipp.setUseIPP(was)
    assert small.shape == (480, 640, 3)
    assert np.array_equal(small, want), f"{(small != want).mean():.4%} of the bytes differ"


def test_not_a_jpeg_is_the_reference_empty_image_error(dec):
    for junk in (b"", b"not a jpeg at all", bytes(2000)):
        with pytest.raises(frb200.FrError) as e:
            dec.decode(junk, size=(64, 64))
        assert e.value.code == frb200.FR_EINVAL and "Empty image" in str(e.value)
    # the decoder is still usable afterwards
    ok, enc = cv2.imencode(".jpg", _picture(64, 96, seed=3))
    assert dec.decode(enc.tobytes()).shape == (64, 96, 3)


def test_decoded_frame_feeds_the_detector(dec, tmp_path):
    # the serving loop's first two steps (src/app.cpp:294-305): decode + stretch to the configured frame size, then findFace
    from tools import make_golden_retina as mgr
    from tools import pack_retina as pr
    from tools import synth_weights as sw

    pr.save_retina(tmp_path / "det.frw", sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT), False)
    det = frb200.Detector(tmp_path / "det.frw", (288, 320), frame_hw=(480, 640), max_batch=1, max_faces=4)
    ok, enc = cv2.imencode(".jpg", _picture(600, 800, seed=4), [cv2.IMWRITE_JPEG_QUALITY, 92])
    frame = dec.decode(enc.tobytes(), size=(640, 480))
    boxes, counts, _ = det.run(frame[None])
    assert boxes.shape == (1, 4) and 0 <= int(counts[0]) <= 4
    det.close()
