"""ORACLE / TEST INFRASTRUCTURE ONLY. numpy restatement of OpenCV's own u8 INTER_CUBIC resize (opencv 4.x
modules/imgproc/src/resize.cpp: hal::resize -> resizeGeneric_<HResizeCubic<uchar,int,short>, VResizeCubic<uchar,int,short,
FixedPtCast<int,uchar,22>, VResizeCubicVec_32s8u>>) — the arithmetic behind getCroppedFaces' cv::resize(..., Size(112,112),
INTER_CUBIC) (/root/reference src/arcface.cpp:9). OpenCV is a third-party dependency of the reference (README.md:11 pins 4.5.5;
this image carries the cv2 4.13 wheel): the restatement is pinned in tests/test_oracle_crops.py against cv2.resize with the
closed-source IPP accelerator switched off (cv2.ipp.setUseIPP(False)) — 0 differing pixels; with IPP on, cv2 itself differs from
its own generic path by 1 LSB on a few per cent of noise pixels.

  tables      scale = 1 / ((double)dst / src);  f = (float)((d + 0.5) * scale - 0.5);  s = floor(f);  f -= s;
              interpolateCubic(f) in float (A = -0.75);  icoef = saturate_cast<short>(c * 2048)   (INTER_RESIZE_COEF_BITS = 11)
  horizontal  int32  H = sum_k src[clamp(s - 1 + k)] * icoef_x[k]
  vertical    the SSE-baseline SIMD kernel: float, beta * 2^-22, acc = H3*b3; acc = H2*b2 + acc; acc = H1*b1 + acc; acc = H0*b0 + acc
              (separate multiplies and adds), round half to even, saturate to u8
"""
from __future__ import annotations

import numpy as np

F = np.float32


def cubic_coeffs(x: np.ndarray) -> np.ndarray:
    A = F(-0.75)
    x = x.astype(np.float32)
    c0 = ((A * (x + F(1)) - F(5) * A) * (x + F(1)) + F(8) * A) * (x + F(1)) - F(4) * A
    c1 = ((A + F(2)) * x - (A + F(3))) * x * x + F(1)
    y = F(1) - x
    c2 = ((A + F(2)) * y - (A + F(3))) * y * y + F(1)
    c3 = F(1) - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(np.float32)


def axis_tables(src: int, dst: int):
    inv = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    ic = np.clip(np.rint(cubic_coeffs(f) * F(2048)), -32768, 32767).astype(np.int32)
    return s, ic


def resize_cubic_u8(img: np.ndarray, dst_h: int = 112, dst_w: int = 112) -> np.ndarray:
    sh, sw, _ = img.shape
    sx, ia = axis_tables(sw, dst_w)
    sy, ib = axis_tables(sh, dst_h)
    cols = np.clip(sx[:, None] - 1 + np.arange(4)[None, :], 0, sw - 1)
    H = (img.astype(np.int32)[:, cols, :] * ia[None, :, :, None]).sum(2)             # [sh, dst_w, cn]
    rows = np.clip(sy[:, None] - 1 + np.arange(4)[None, :], 0, sh - 1)
    b = ib.astype(np.float32) * F(1.0 / (2048 * 2048))
    S = H[rows].astype(np.float32)                                                    # [dst_h, 4, dst_w, cn]
    acc = (S[:, 3] * b[:, 3, None, None]).astype(np.float32)
    for k in (2, 1, 0):
        acc = ((S[:, k] * b[:, k, None, None]).astype(np.float32) + acc).astype(np.float32)
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def cropped_faces(frame: np.ndarray, boxes) -> np.ndarray:
    """getCroppedFaces (src/arcface.cpp:3-17): Rect(Point(y1, x1), Point(y2, x2)) -> 112 x 112 BGR u8, one per box"""
    out = []
    for b in boxes:
        r0, r1 = sorted((int(b["x1"]), int(b["x2"])))
        c0, c1 = sorted((int(b["y1"]), int(b["y2"])))
        out.append(resize_cubic_u8(frame[r0:max(r1, r0 + 1), c0:max(c1, c0 + 1)]))
    return np.stack(out)
