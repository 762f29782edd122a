"""ORACLE / TEST INFRASTRUCTURE ONLY. ctypes driver for oracle/_ref/libref_retina.so = the reference's own
/root/reference/src/retinaface.cpp compiled verbatim (oracle/build_ref.py, accessor oracle/ref_retina_wrap.cpp).

It is the pin for SURVEY §8 rows a2 (constructor sizes), a5 (anchors), a6 (decode / threshold / rescale / clip), a7 (NMS): both the
restatement oracle/retina_post.c and the CUDA kernel det_decode_nms_kernel are compared with what the reference code itself returns.
Pure host code: usable in the CPU suite and on the GPU box (the .so travels in oracle/_ref/; /root/reference is not needed at run time).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libref_retina.so"


class Bbox(C.Structure):
    """struct Bbox, src/common.h:13-16"""
    _fields_ = [("x1", C.c_int), ("y1", C.c_int), ("x2", C.c_int), ("y2", C.c_int), ("score", C.c_float)]


_lib = None


def available() -> bool:
    if LIB.exists():
        return True
    try:
        from oracle import build_ref

        build_ref.build()
    except Exception:
        return False
    return LIB.exists()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB} missing (built only where /root/reference is mounted)")
        L = C.CDLL(str(LIB))
        L.ref_retina_new.restype = C.c_void_p
        L.ref_retina_new.argtypes = [C.c_char_p] + [C.c_int] * 7 + [C.c_float, C.c_float]
        L.ref_retina_free.argtypes = [C.c_void_p]
        L.ref_retina_output_size_base.argtypes = [C.c_void_p]
        L.ref_retina_scales.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_retina_anchors.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_retina_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_retina_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float]
        _lib = L
    return _lib


class RefRetinaFace:
    """the reference's RetinaFace object (its real constructor ran; the TensorRT engine behind it is a stub)"""

    def __init__(self, net_h, net_w, frame_h, frame_w, max_faces=4, nms_thr=0.4, bbox_thr=0.6, max_batch=1):
        # any readable file satisfies loadEngine's fileExists + read (src/retinaface.cpp:32-48)
        self._h = lib().ref_retina_new(str(LIB).encode(), frame_w, frame_h, 3, net_h, net_w, max_batch, max_faces, nms_thr, bbox_thr)
        if not self._h:
            raise RuntimeError("reference RetinaFace constructor threw")
        self.net_h, self.net_w = net_h, net_w

    def close(self):
        if self._h:
            lib().ref_retina_free(self._h)
            self._h = None

    __del__ = close

    @property
    def output_size_base(self) -> int:
        return lib().ref_retina_output_size_base(self._h)

    @property
    def scales(self):
        a, b = C.c_float(), C.c_float()
        lib().ref_retina_scales(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def anchors(self, net_h=None, net_w=None) -> np.ndarray:
        h, w = net_h or self.net_h, net_w or self.net_w
        n = lib().ref_retina_anchors(self._h, w, h, None, 0)
        out = np.empty((n, 4), np.float32)
        assert lib().ref_retina_anchors(self._h, w, h, out.ctypes.data_as(C.c_void_p), n) == n
        return out

    def postprocess(self, loc, conf):
        """RetinaFace::postprocessing on one image's heads -> [(x1, y1, x2, y2, score)] exactly as m_outputBbox holds them"""
        loc = np.ascontiguousarray(loc, np.float32).copy()
        conf = np.ascontiguousarray(conf, np.float32).copy()
        cap = max(loc.shape[0], 1)
        out = (Bbox * cap)()
        n = lib().ref_retina_postprocess(self._h, loc.ctypes.data_as(C.c_void_p), conf.ctypes.data_as(C.c_void_p), C.cast(out, C.c_void_p), cap)
        return [(out[i].x1, out[i].y1, out[i].x2, out[i].y2, out[i].score) for i in range(n)]

    def nms(self, boxes, thr):
        arr = (Bbox * max(len(boxes), 1))(*[Bbox(*b) for b in boxes])
        n = lib().ref_retina_nms(self._h, C.cast(arr, C.c_void_p), len(boxes), thr)
        return [(arr[i].x1, arr[i].y1, arr[i].x2, arr[i].y2, arr[i].score) for i in range(n)]
