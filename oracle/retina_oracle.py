"""ORACLE (test infrastructure, never shipped or measured as the product): CPU restatement of the reference's detection
path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Restated reference code (paths relative to /root/reference):
  preprocess()   RetinaFace::preprocess        src/retinaface.cpp:106-136  (cv2 = the same OpenCV resize code)
  forward()      RetinaFace.forward (trimmed / full model), MobileNetV1, FPN, SSH, heads, softmax
                 conversion/retina/models/retinaface_trim.py:101-127, retinaface.py:99-130, models/net.py:9-124
  postprocess()  RetinaFace::postprocessing + create_anchor_retinaface + nms   src/retinaface.cpp:154-271
                 -> C restatement oracle/retina_post.c (double intermediates, both int truncations, '>=' NMS), via ctypes
  decode_landmarks()  NOT in the reference (its deployed model has no landmark head, conversion/retina/torch2trt.py:7-9):
                 centre-form decode with variance 0.1 (conversion/retina/config.py:6) per SURVEY §8c — parity unpinned.
Parity status: no golden vectors exist in the reference. forward() is pinned against the reference's own
modules by tools/make_golden_retina.py (outputs committed under tests/golden/retina_*.npz); the C post-processing is pinned
by hand-computed cases in tests/test_oracle_retina.py.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
HERE = Path(__file__).resolve().parent


# ------------------------------------------------------------------------------------------------ preprocessing
def letterbox_params(frame_h, frame_w, net_h, net_w):
    """(w, h, x, y) of the resized frame inside the net canvas, src/retinaface.cpp:21-22,111-122 (float32 scales, int truncation)"""
    scale_h = np.float32(net_h) / np.float32(frame_h)
    scale_w = np.float32(net_w) / np.float32(frame_w)
    if scale_h > scale_w:
        w, h = net_w, int(np.float32(scale_w * np.float32(frame_h)))
        x, y = 0, (net_h - h) // 2
    else:
        w, h = int(np.float32(scale_h * np.float32(frame_w))), net_h
        x, y = (net_w - w) // 2, 0
    return w, h, x, y


def preprocess(frame_bgr_u8: np.ndarray, net_h: int, net_w: int) -> np.ndarray:
    """u8 HWC BGR frame -> f32 planar CHW (B,G,R planes), letterboxed on a 128 canvas, minus (104,117,123)"""
    import cv2

    fh, fw = frame_bgr_u8.shape[:2]
    w, h, x, y = letterbox_params(fh, fw, net_h, net_w)
    re = cv2.resize(frame_bgr_u8, (w, h), interpolation=cv2.INTER_LINEAR)
    out = np.full((net_h, net_w, 3), 128, np.uint8)
    out[y:y + h, x:x + w] = re
    f = out.astype(np.float32) - np.array([104, 117, 123], np.float32)
    return np.ascontiguousarray(f.transpose(2, 0, 1))


# ------------------------------------------------------------------------------------------------ network
def to_torch(sd_np) -> dict:
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items() if v.dtype != np.int64}


def _cbr(x, sd, p, stride=1, pad=1, groups=1, relu=True, ci=0, bi=1):
    x = F.conv2d(x, sd[f"{p}.{ci}.weight"], None, stride, pad, 1, groups)
    q = f"{p}.{bi}"
    x = F.batch_norm(x, sd[q + ".running_mean"], sd[q + ".running_var"], sd[q + ".weight"], sd[q + ".bias"], False, 0.0, BN_EPS)
    return F.relu(x) if relu else x


def _dw(x, sd, p, stride):
    """conv_dw, net.py:29-38 (activations are plain ReLU: `leaky` is computed but unused, net.py:44-46)"""
    x = _cbr(x, sd, p, stride, 1, x.shape[1], True, 0, 1)
    return _cbr(x, sd, p, 1, 0, 1, True, 3, 4)


def _ssh(x, sd, p):
    a = _cbr(x, sd, p + ".conv3X3", relu=False)
    t = _cbr(x, sd, p + ".conv5X5_1")
    b = _cbr(t, sd, p + ".conv5X5_2", relu=False)
    u = _cbr(t, sd, p + ".conv7X7_2")
    c = _cbr(u, sd, p + ".conv7x7_3", relu=False)
    return F.relu(torch.cat([a, b, c], 1))


@torch.no_grad()
def forward(sd: dict, x: torch.Tensor, full: bool = False, feats: dict | None = None):
    """x: n x 3 x H x W f32 (preprocess output) -> loc [n, A, 4], conf [n, A, 2] (softmax), landm [n, A, 10] or None"""
    x = _cbr(x, sd, "body.stage1.0", 2)
    for i, s in enumerate((1, 2, 1, 2, 1), start=1):
        x = _dw(x, sd, f"body.stage1.{i}", s)
    c1 = x
    for i, s in enumerate((2, 1, 1, 1, 1, 1)):
        x = _dw(x, sd, f"body.stage2.{i}", s)
    c2 = x
    for i, s in enumerate((2, 1)):
        x = _dw(x, sd, f"body.stage3.{i}", s)
    c3 = x
    o1 = _cbr(c1, sd, "fpn.output1", 1, 0)
    o2 = _cbr(c2, sd, "fpn.output2", 1, 0)
    o3 = _cbr(c3, sd, "fpn.output3", 1, 0)
    o2 = _cbr(o2 + F.interpolate(o3, size=o2.shape[2:], mode="nearest"), sd, "fpn.merge2")
    o1 = _cbr(o1 + F.interpolate(o2, size=o1.shape[2:], mode="nearest"), sd, "fpn.merge1")
    fs = [_ssh(o1, sd, "ssh1"), _ssh(o2, sd, "ssh2"), _ssh(o3, sd, "ssh3")]
    if feats is not None:
        feats.update(c1=c1, c2=c2, c3=c3, o1=o1, o2=o2, o3=o3, f1=fs[0], f2=fs[1], f3=fs[2])

    def head(name, width):
        outs = []
        for lvl, f in enumerate(fs):
            y = F.conv2d(f, sd[f"{name}.{lvl}.conv1x1.weight"], sd[f"{name}.{lvl}.conv1x1.bias"])
            outs.append(y.permute(0, 2, 3, 1).reshape(y.shape[0], -1, width))
        return torch.cat(outs, 1)

    loc = head("BboxHead", 4)
    conf = F.softmax(head("ClassHead", 2), dim=-1)
    landm = head("LandmarkHead", 10) if full else None
    return loc, conf, landm


def num_anchors(net_h, net_w):
    """m_OUTPUT_SIZE_BASE, src/retinaface.cpp:13 (integer arithmetic, left to right)"""
    return ((net_h // 8 * net_w) // 8 + (net_h // 16 * net_w) // 16 + (net_h // 32 * net_w) // 32) * 2


# ------------------------------------------------------------------------------------------------ post-processing (C)
class _Bbox(C.Structure):
    _fields_ = [("x1", C.c_int), ("y1", C.c_int), ("x2", C.c_int), ("y2", C.c_int), ("score", C.c_float)]


_lib = None


def _post_lib():
    global _lib
    if _lib is None:
        so = HERE / "_native" / "libretina_post.so"
        if not so.exists():
            from oracle import build_native

            build_native.build()
        L = C.CDLL(str(so))
        L.retina_anchors.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.retina_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                         C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _lib = L
    return _lib


def anchors(net_h, net_w) -> np.ndarray:
    """create_anchor_retinaface, src/retinaface.cpp:210-240 -> [A, 4] f32 (cx, cy, sx, sy)"""
    n = num_anchors(net_h, net_w)
    out = np.empty((n, 4), np.float32)
    got = _post_lib().retina_anchors(net_w, net_h, out.ctypes.data_as(C.c_void_p), n)
    assert got == n, (got, n)
    return out


def postprocess(loc, conf, landm, net_h, net_w, frame_h, frame_w, nms_thr, bbox_thr, max_faces):
    """one image: loc [A,4], conf [A,2], landm [A,10] or None -> (boxes [(x1,y1,x2,y2,score)], landmarks [k,10] f32, anchor ids)"""
    loc = np.ascontiguousarray(loc, np.float32)
    conf = np.ascontiguousarray(conf, np.float32)
    a = loc.shape[0]
    lm = np.ascontiguousarray(landm, np.float32) if landm is not None else None
    cap = a
    boxes = (_Bbox * cap)()
    out_lm = np.zeros((cap, 10), np.float32)
    ids = np.zeros(cap, np.int32)
    n = _post_lib().retina_postprocess(loc.ctypes.data_as(C.c_void_p), conf.ctypes.data_as(C.c_void_p),
                                       lm.ctypes.data_as(C.c_void_p) if lm is not None else None, a, net_w, net_h, frame_w, frame_h,
                                       nms_thr, bbox_thr, max_faces, C.cast(boxes, C.c_void_p), out_lm.ctypes.data_as(C.c_void_p),
                                       ids.ctypes.data_as(C.c_void_p), cap)
    assert n >= 0
    res = [(boxes[i].x1, boxes[i].y1, boxes[i].x2, boxes[i].y2, boxes[i].score) for i in range(n)]
    return res, out_lm[:n].copy(), ids[:n].copy()
