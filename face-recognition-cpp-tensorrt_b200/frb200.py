"""ctypes binding of libfr_b200.so (include/fr_b200.h) used by tests/, bench.py and __graft_entry__.py.

This is test/bench plumbing around the C ABI, not a second implementation: every call goes straight into the CUDA
library and raises if the library is missing — there is no CPU fallback here or anywhere in the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["FR_B200_LIB"]) if os.environ.get("FR_B200_LIB") else PKG / "lib" / "libfr_b200.so"  # A/B builds (tools/ab_search.py)

FR_OK, FR_EINVAL, FR_ENODEVICE, FR_ECUDA, FR_ENOENT, FR_EFORMAT, FR_ESTATE = 0, -1, -2, -3, -4, -5, -6
FR_TOPK_MAX = 8
FR_PATH_AUTO, FR_PATH_EXACT, FR_PATH_TENSOR = 0, 1, 2
FR_SCAN_F16, FR_SCAN_F8 = 0, 1


class FrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fr_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class FrBbox(C.Structure):
    _fields_ = [("x1", C.c_int), ("y1", C.c_int), ("x2", C.c_int), ("y2", C.c_int), ("score", C.c_float)]


class FrSearchStats(C.Structure):
    _fields_ = [("scan_bytes", C.c_int64), ("flops", C.c_int64), ("launches", C.c_int), ("ctas", C.c_int)]


_lib = None


def lib() -> C.CDLL:
    """Load the library (once). Fails loudly if it was not built: run face-recognition-cpp-tensorrt_b200/build.py."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `python {PKG / 'build.py'}` (no CPU fallback exists)")
        L = C.CDLL(str(LIB_PATH))
        L.fr_last_error.restype = C.c_char_p
        L.fr_launch_count.restype = C.c_uint64
        L.fr_gallery_rows.restype = C.c_int64
        L.fr_gallery_rows.argtypes = [C.c_void_p]
        L.fr_gallery_destroy.restype = None
        L.fr_gallery_destroy.argtypes = [C.c_void_p]
        L.fr_gallery_create.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
        L.fr_gallery_create_dev.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
        L.fr_gallery_create_synthetic.argtypes = [C.c_int64, C.c_int, C.c_uint64, C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
        L.fr_gallery_set_path.argtypes = [C.c_void_p, C.c_int]
        L.fr_gallery_read_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        L.fr_gallery_sims.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.fr_gallery_sims_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.fr_gallery_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fr_gallery_topk_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_topk_merge_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.fr_gallery_last_stats.argtypes = [C.c_void_p, C.POINTER(FrSearchStats)]
        L.fr_gallery_last_flagged.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.fr_gallery_reserve.argtypes = [C.c_void_p, C.c_int64]
        L.fr_gallery_append.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.fr_gallery_remove.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.fr_gallery_clear.argtypes = [C.c_void_p]
        L.fr_gallery_capacity.argtypes = [C.c_void_p]
        L.fr_gallery_capacity.restype = C.c_int64
        L.fr_gallery_set_timing.argtypes = [C.c_void_p, C.c_int]
        L.fr_gallery_scan_time.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != FR_OK:
        raise FrError(rc, lib().fr_last_error().decode(errors="replace"))


def launch_count() -> int:
    return int(lib().fr_launch_count())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(x):
    """host numpy array -> void*, torch tensor (any device) -> data_ptr, int -> itself"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, int):
        return C.c_void_p(x)
    return C.c_void_p(x.data_ptr())


class Gallery:
    """One shard of the gallery resident on one GPU (MatMul::init, /root/reference src/matmul.cpp:9-34)."""

    def __init__(self, handle: C.c_void_p, device: int):
        self._h = handle
        self.device = device

    @classmethod
    def from_rows(cls, rows, device: int = 0, row_offset: int = 0) -> "Gallery":
        rows = _f32(rows)
        n, dim = (rows.shape if rows.ndim == 2 else (0, 512))
        h = C.c_void_p()
        check(lib().fr_gallery_create(_ptr(rows) if n else None, n, dim, device, row_offset, C.byref(h)))
        return cls(h, device)

    @classmethod
    def from_device_rows(cls, rows_t, device: int = 0, row_offset: int = 0) -> "Gallery":
        h = C.c_void_p()
        check(lib().fr_gallery_create_dev(_ptr(rows_t), rows_t.shape[0], rows_t.shape[1], device, row_offset, C.byref(h)))
        return cls(h, device)

    @classmethod
    def synthetic(cls, n: int, seed: int, device: int = 0, row_offset: int = 0, dim: int = 512) -> "Gallery":
        h = C.c_void_p()
        check(lib().fr_gallery_create_synthetic(n, dim, seed, device, row_offset, C.byref(h)))
        return cls(h, device)

    def close(self) -> None:
        if self._h:
            lib().fr_gallery_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rows(self) -> int:
        return int(lib().fr_gallery_rows(self._h))

    @property
    def capacity(self) -> int:
        return int(lib().fr_gallery_capacity(self._h))

    # gallery lifecycle without a re-upload (addEmbedding / resetEmbeddings / reload of the reference)
    def reserve(self, capacity: int) -> None:
        check(lib().fr_gallery_reserve(self._h, capacity))

    def append(self, rows) -> None:
        rows = _f32(rows)
        if rows.ndim == 1:
            rows = rows[None, :]
        check(lib().fr_gallery_append(self._h, _ptr(rows), rows.shape[0]))

    def update(self, first: int, rows) -> None:
        """replace rows [first, first + n) in place (all resident copies follow)"""
        rows = _f32(rows)
        if rows.ndim == 1:
            rows = rows[None, :]
        lib().fr_gallery_update.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        check(lib().fr_gallery_update(self._h, first, _ptr(rows), rows.shape[0]))

    def remove(self, row: int) -> int:
        """deletes local row `row`; returns the index the row that now sits in its slot had before (the old last row)"""
        moved = C.c_int64()
        check(lib().fr_gallery_remove(self._h, row, C.byref(moved)))
        return moved.value

    def clear(self) -> None:
        check(lib().fr_gallery_clear(self._h))

    def set_path(self, path: int) -> None:
        check(lib().fr_gallery_set_path(self._h, path))

    def set_scan(self, scan: int) -> None:
        lib().fr_gallery_set_scan.argtypes = [C.c_void_p, C.c_int]
        check(lib().fr_gallery_set_scan(self._h, scan))

    def debug_read(self, what: int, first: int = 0, count: int = 0) -> np.ndarray:
        """test hook (fr_gallery_debug_read): 0 = e4m3 rows [count, 512] u8, 1 = query operand image, 2 = q_margin, 3 = q_gap,
        4 = (gmax, g4max, w4max), 5 / 6 = sorted candidate lists (coarse scores / local rows) of the last k > 1 search"""
        L = lib()
        L.fr_gallery_debug_read.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        out = {0: lambda: np.empty((count, 512), np.uint8), 1: lambda: np.empty((256, 1024), np.uint8), 2: lambda: np.empty(256, np.float32),
               3: lambda: np.empty(256, np.float32), 4: lambda: np.empty(3, np.float32),
               5: lambda: np.empty((296, 256, 16), np.float32), 6: lambda: np.empty((296, 256, 16), np.int32)}[what]()
        check(L.fr_gallery_debug_read(self._h, what, first, count, _ptr(out)))
        return out

    def read_rows(self, first: int, count: int) -> np.ndarray:
        out = np.empty((count, 512), np.float32)
        check(lib().fr_gallery_read_rows(self._h, first, count, _ptr(out)))
        return out

    def sims(self, q) -> np.ndarray:
        """MatMul::calculate (src/matmul.cpp:36-77): out[i, j] = <q_i, row_j> in exact fp32."""
        q = _f32(q)
        out = np.empty((q.shape[0], max(self.rows, 0)), np.float32)
        check(lib().fr_gallery_sims(self._h, _ptr(q), q.shape[0], _ptr(out)))
        return out

    def topk(self, q, k: int = 1, scores_out=None, idx_out=None):
        """host queries -> (scores nq x k f32, idx nq x k int64), order (score desc, row asc)."""
        q = _f32(q) if isinstance(q, np.ndarray) or not hasattr(q, "data_ptr") else q
        nq = q.shape[0]
        scores = scores_out if scores_out is not None else np.empty((nq, k), np.float32)
        idx = idx_out if idx_out is not None else np.empty((nq, k), np.int64)
        check(lib().fr_gallery_topk(self._h, _ptr(q), nq, k, _ptr(scores), _ptr(idx)))
        return scores, idx

    def topk_dev(self, q_t, k, scores_t, idx_t, stream: int = 0) -> None:
        """device tensors in/out (torch), launched on `stream` (raw cudaStream_t value; 0 = the handle's own stream)."""
        check(lib().fr_gallery_topk_dev(self._h, _ptr(q_t), q_t.shape[0], k, _ptr(scores_t), _ptr(idx_t), C.c_void_p(stream)))

    def sims_dev(self, q_t, out_t, stream: int = 0) -> None:
        check(lib().fr_gallery_sims_dev(self._h, _ptr(q_t), q_t.shape[0], _ptr(out_t), C.c_void_p(stream)))

    def set_timing(self, enable) -> None:
        """False / True: pooled event pairs around every launch of the fused scan kernel (scan_time); 2: one fixed pair per scan
        copy that also works inside CUDA-graph replays (last_scan_ms)"""
        check(lib().fr_gallery_set_timing(self._h, int(enable)))

    def pool_time(self, count: int) -> float:
        """summed ms of the first `count` pooled event pairs (a captured block of `count` steps re-records them on every replay)"""
        lib().fr_gallery_pool_time.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        ms = C.c_double()
        check(lib().fr_gallery_pool_time(self._h, count, C.byref(ms)))
        return ms.value

    def last_scan_ms(self, scan: int) -> float:
        lib().fr_gallery_last_scan_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        ms = C.c_double()
        check(lib().fr_gallery_last_scan_ms(self._h, scan, C.byref(ms)))
        return ms.value

    def scan_time(self):
        """(summed ms, launches) of the fused scan kernel since timing was enabled / last read"""
        ms, n = C.c_double(), C.c_int()
        check(lib().fr_gallery_scan_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_flagged(self) -> int:
        """queries of the last tensor-path search that were recomputed by the exact scan (0 normally)"""
        n = C.c_int()
        check(lib().fr_gallery_last_flagged(self._h, C.byref(n)))
        return n.value

    def last_stats(self) -> FrSearchStats:
        st = FrSearchStats()
        check(lib().fr_gallery_last_stats(self._h, C.byref(st)))
        return st


def search_topk(gal: "Gallery", exchange, q, k: int = 1, scores_out=None, idx_out=None, stream: int = 0):
    """fr_search_topk: host queries in, host results out, through this shard's fused search and (exchange != None) the cross-GPU merge"""
    L = lib()
    L.fr_search_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    nq = q.shape[0]
    scores = scores_out if scores_out is not None else np.empty((nq, k), np.float32)
    idx = idx_out if idx_out is not None else np.empty((nq, k), np.int64)
    check(L.fr_search_topk(gal._h, exchange._h if exchange is not None else None, _ptr(q), nq, k, _ptr(scores), _ptr(idx), C.c_void_p(stream)))
    return scores, idx


class Roster:
    """fr_roster_*: row -> userId table (the reference's classNames) in step with a row-sharded gallery; see include/fr_b200.h"""

    def __init__(self, local_shard: "Gallery | None", world: int = 1, rank: int = 0):
        L = lib()
        L.fr_roster_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.fr_roster_destroy.restype = None
        L.fr_roster_destroy.argtypes = [C.c_void_p]
        L.fr_roster_rows.restype = C.c_int64
        L.fr_roster_rows.argtypes = [C.c_void_p]
        L.fr_roster_shard_rows.restype = C.c_int64
        L.fr_roster_shard_rows.argtypes = [C.c_void_p, C.c_int]
        L.fr_roster_load.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int64]
        L.fr_roster_add.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.fr_roster_remove.argtypes = [C.c_void_p, C.c_int64]
        L.fr_roster_remove_user.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
        L.fr_roster_clear.argtypes = [C.c_void_p]
        L.fr_roster_user.restype = C.c_char_p
        L.fr_roster_user.argtypes = [C.c_void_p, C.c_int64]
        h = C.c_void_p()
        check(L.fr_roster_create(local_shard._h if local_shard is not None else None, world, rank, C.byref(h)))
        self._h, self.world, self.rank, self._keep = h, world, rank, local_shard

    @property
    def rows(self) -> int:
        return int(lib().fr_roster_rows(self._h))

    def shard_rows(self, shard: int) -> int:
        return int(lib().fr_roster_shard_rows(self._h, shard))

    def load(self, user_ids, blobs) -> None:
        """rows of `SELECT * FROM FACE`: user_ids = USR_ID strings, blobs = EMBEDDING bytes objects (512 x f32 LE)"""
        n = len(user_ids)
        ids = (C.c_char_p * n)(*[u.encode() for u in user_ids])
        keep = [bytes(b) for b in blobs]
        ptrs = (C.c_void_p * n)(*[C.cast(C.c_char_p(b), C.c_void_p) for b in keep])
        sizes = (C.c_int * n)(*[len(b) for b in keep])
        check(lib().fr_roster_load(self._h, ids, ptrs, sizes, n))

    def add(self, user_id: str, embedding) -> int:
        e = _f32(embedding)
        out = C.c_int64()
        check(lib().fr_roster_add(self._h, user_id.encode(), _ptr(e), C.byref(out)))
        return out.value

    def remove(self, row_id: int) -> None:
        check(lib().fr_roster_remove(self._h, row_id))

    def remove_user(self, user_id: str) -> int:
        n = C.c_int64()
        check(lib().fr_roster_remove_user(self._h, user_id.encode(), C.byref(n)))
        return n.value

    def clear(self) -> None:
        check(lib().fr_roster_clear(self._h))

    def user(self, row_id: int):
        u = lib().fr_roster_user(self._h, int(row_id))
        return u.decode() if u is not None else None

    def close(self) -> None:
        if self._h:
            lib().fr_roster_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SearchStream:
    """fr_search_stream_*: asynchronous host-buffer search with up to three batches in flight (the serving form of search_topk)"""

    def __init__(self, gal: "Gallery", exchange=None, k: int = 1):
        L = lib()
        L.fr_search_stream_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.fr_search_stream_destroy.restype = None
        L.fr_search_stream_destroy.argtypes = [C.c_void_p]
        L.fr_search_stream_cuda_stream.restype = C.c_void_p
        L.fr_search_stream_cuda_stream.argtypes = [C.c_void_p]
        L.fr_search_stream_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.fr_search_stream_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        h = C.c_void_p()
        check(L.fr_search_stream_create(gal._h, exchange._h if exchange is not None else None, k, C.byref(h)))
        self._h, self.k, self._keep = h, k, (gal, exchange)

    @property
    def cuda_stream(self) -> int:
        return int(lib().fr_search_stream_cuda_stream(self._h) or 0)

    def submit(self, q) -> None:
        check(lib().fr_search_stream_submit(self._h, _ptr(q), q.shape[0]))

    def collect(self, scores_out, idx_out) -> int:
        n = C.c_int()
        check(lib().fr_search_stream_collect(self._h, _ptr(scores_out), _ptr(idx_out), C.byref(n)))
        return n.value

    def close(self) -> None:
        if self._h:
            lib().fr_search_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def topk_merge_dev(scores_parts_t, idx_parts_t, n_parts: int, nq: int, k: int, scores_t, idx_t, device: int, stream: int = 0) -> None:
    check(lib().fr_topk_merge_dev(_ptr(scores_parts_t), _ptr(idx_parts_t), n_parts, nq, k, _ptr(scores_t), _ptr(idx_t), device,
                                  C.c_void_p(stream)))


FR_MODE_IR, FR_MODE_IR_SE = 0, 1


def _bind_embedder(L):
    if getattr(L, "_emb_bound", False):
        return
    L.fr_embedder_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.fr_embedder_destroy.restype = None
    L.fr_embedder_destroy.argtypes = [C.c_void_p]
    L.fr_embedder_mode.argtypes = [C.c_void_p]
    L.fr_embedder_max_batch.argtypes = [C.c_void_p]
    L.fr_embedder_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.fr_embedder_run_crops.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.fr_embedder_run_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.fr_embedder_trace.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L._emb_bound = True


class Embedder:
    """ArcFaceIR50's network half (/root/reference src/arcface.cpp:21-103,131-148) on the GPU."""

    def __init__(self, weights_path, max_batch: int = 32, device: int = 0):
        _bind_embedder(lib())
        h = C.c_void_p()
        check(lib().fr_embedder_create(str(weights_path).encode(), max_batch, device, C.byref(h)))
        self._h, self.device, self.max_batch = h, device, max_batch

    def close(self) -> None:
        if self._h:
            lib().fr_embedder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def mode(self) -> int:
        return int(lib().fr_embedder_mode(self._h))

    def run(self, chw) -> np.ndarray:
        """ArcFaceIR50::doInference(float*, float*, int): n x 3 x 112 x 112 f32 -> n x 512 f32 (host)"""
        x = _f32(chw)
        out = np.empty((x.shape[0], 512), np.float32)
        check(lib().fr_embedder_run(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def run_crops(self, crops_bgr_u8) -> np.ndarray:
        x = np.ascontiguousarray(crops_bgr_u8, dtype=np.uint8)
        out = np.empty((x.shape[0], 512), np.float32)
        check(lib().fr_embedder_run_crops(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def run_dev(self, chw_t, out_t, stream: int = 0) -> None:
        check(lib().fr_embedder_run_dev(self._h, _ptr(chw_t), chw_t.shape[0], _ptr(out_t), C.c_void_p(stream)))

    def trace(self, layer: int, batch: int) -> np.ndarray:
        geo = [112, 56, 56, 56, 28, 28, 28, 28] + [14] * 14 + [7] * 3
        ch = [64, 64, 64, 64, 128, 128, 128, 128] + [256] * 14 + [512] * 3
        out = np.empty((batch, ch[layer], geo[layer], geo[layer]), np.float32)
        n = C.c_int64()
        check(lib().fr_embedder_trace(self._h, layer, _ptr(out), out.size, C.byref(n)))
        assert n.value == out.size
        return out


def _bind_detector(L):
    if getattr(L, "_det_bound", False):
        return
    L.fr_detector_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]
    L.fr_detector_destroy.restype = None
    L.fr_detector_destroy.argtypes = [C.c_void_p]
    L.fr_detector_num_anchors.argtypes = [C.c_void_p]
    L.fr_detector_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fr_detector_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fr_detector_net.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fr_detector_post.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fr_detector_run_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L._det_bound = True


class Detector:
    """RetinaFace (/root/reference src/retinaface.{h,cpp}) on the GPU. Boxes use the reference's Bbox semantics:
    x = row (vertical), y = column (horizontal)."""

    def __init__(self, weights_path, net_hw, frame_hw=None, max_batch=16, max_faces=4, nms_thr=0.4, bbox_thr=0.6, landmarks=False,
                 device=0):
        _bind_detector(lib())
        frame_hw = frame_hw or net_hw
        h = C.c_void_p()
        check(lib().fr_detector_create(str(weights_path).encode(), net_hw[0], net_hw[1], frame_hw[0], frame_hw[1], max_batch, max_faces,
                                       nms_thr, bbox_thr, int(landmarks), device, C.byref(h)))
        self._h, self.device = h, device
        self.net_hw, self.frame_hw, self.max_batch, self.max_faces, self.landmarks = tuple(net_hw), tuple(frame_hw), max_batch, max_faces, landmarks

    def close(self) -> None:
        if self._h:
            lib().fr_detector_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def anchors(self) -> int:
        return int(lib().fr_detector_num_anchors(self._h))

    def _frames(self, frames):
        f = np.ascontiguousarray(frames, dtype=np.uint8)
        assert f.ndim == 4 and f.shape[1:] == (self.frame_hw[0], self.frame_hw[1], 3), f.shape
        return f

    def run(self, frames):
        """findFace for a batch: -> (boxes structured array [n, max_faces], counts [n], landmarks [n, max_faces, 10] or None)"""
        f = self._frames(frames)
        n = f.shape[0]
        boxes = np.zeros((n, self.max_faces), dtype=[("x1", "<i4"), ("y1", "<i4"), ("x2", "<i4"), ("y2", "<i4"), ("score", "<f4")])
        counts = np.zeros(n, np.int32)
        lm = np.zeros((n, self.max_faces, 10), np.float32) if self.landmarks else None
        check(lib().fr_detector_run(self._h, _ptr(f), f.shape[2] * 3, n, _ptr(boxes), _ptr(counts), _ptr(lm)))
        return boxes, counts, lm

    def run_dev(self, frames_t, boxes_t, counts_t, stream: int = 0) -> None:
        """device-resident findFace (fr_detector_run_dev): frames_t u8 [n, H, W, 3], boxes_t [n, max_faces, 5] 32-bit words
        (FrBbox), counts_t int32 [n], all on the detector's device; asynchronous on `stream`"""
        n = frames_t.shape[0]
        check(lib().fr_detector_run_dev(self._h, _ptr(frames_t), frames_t.shape[2] * 3, n, _ptr(boxes_t), _ptr(counts_t), None,
                                        C.c_void_p(stream)))

    def _raw_out(self, n):
        a = self.anchors
        return (np.empty((n, a, 4), np.float32), np.empty((n, a, 2), np.float32), np.empty((n, a, 10), np.float32) if self.landmarks else None)

    def raw(self, frames):
        f = self._frames(frames)
        loc, conf, lm = self._raw_out(f.shape[0])
        check(lib().fr_detector_raw(self._h, _ptr(f), f.shape[2] * 3, f.shape[0], _ptr(loc), _ptr(conf), _ptr(lm)))
        return loc, conf, lm

    def net(self, chw):
        x = _f32(chw)
        loc, conf, lm = self._raw_out(x.shape[0])
        check(lib().fr_detector_net(self._h, _ptr(x), x.shape[0], _ptr(loc), _ptr(conf), _ptr(lm)))
        return loc, conf, lm

    def post(self, loc, conf, landm=None):
        loc, conf = _f32(loc), _f32(conf)
        n = loc.shape[0]
        lm_in = _f32(landm) if landm is not None else None
        boxes = np.zeros((n, self.max_faces), dtype=[("x1", "<i4"), ("y1", "<i4"), ("x2", "<i4"), ("y2", "<i4"), ("score", "<f4")])
        counts = np.zeros(n, np.int32)
        lm = np.zeros((n, self.max_faces, 10), np.float32)
        check(lib().fr_detector_post(self._h, _ptr(loc), _ptr(conf), _ptr(lm_in), n, _ptr(boxes), _ptr(counts), _ptr(lm)))
        return boxes, counts, lm


BOX_DTYPE = np.dtype([("x1", "<i4"), ("y1", "<i4"), ("x2", "<i4"), ("y2", "<i4"), ("score", "<f4")])


def embed_boxes(emb: "Embedder", frame_bgr_u8, boxes, want_crops: bool = False):
    """ArcFaceIR50::forward for one frame (/root/reference src/arcface.cpp:166-187): boxes = structured BOX_DTYPE array"""
    L = lib()
    L.fr_embedder_run_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    f = np.ascontiguousarray(frame_bgr_u8, dtype=np.uint8)
    b = np.ascontiguousarray(boxes, dtype=BOX_DTYPE)
    n = b.shape[0]
    out = np.empty((n, 512), np.float32)
    crops = np.empty((n, 112, 112, 3), np.uint8) if want_crops else None
    check(L.fr_embedder_run_boxes(emb._h, _ptr(f), f.shape[0], f.shape[1], f.shape[1] * 3, _ptr(b), n, _ptr(out), _ptr(crops)))
    return (out, crops) if want_crops else out


class Pipeline:
    """detect -> crop -> embed -> search on one GPU (/root/reference src/app.cpp:293-352 without the socket glue)"""

    def __init__(self, det: "Detector", emb: "Embedder", gal: "Gallery | None"):
        L = lib()
        L.fr_pipeline_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        L.fr_pipeline_destroy.restype = None
        L.fr_pipeline_destroy.argtypes = [C.c_void_p]
        L.fr_pipeline_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_pipeline_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.fr_pipeline_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_pipeline_in_flight.argtypes = [C.c_void_p]
        self._pending = []  # (frames kept alive, n, want_embeddings) per submitted batch
        h = C.c_void_p()
        check(L.fr_pipeline_create(det._h, emb._h, gal._h if gal is not None else None, C.byref(h)))
        self._h, self.det, self.emb, self.gal = h, det, emb, gal

    def close(self) -> None:
        if self._h:
            lib().fr_pipeline_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, frames, want_embeddings: bool = False, out=None):
        """frames: n x H x W x 3 u8 (numpy, or a pinned torch tensor) -> dict(boxes, counts, idx, score[, embeddings])"""
        if isinstance(frames, np.ndarray):
            f = np.ascontiguousarray(frames, dtype=np.uint8)
            n, w = f.shape[0], f.shape[2]
        else:
            f, n, w = frames, frames.shape[0], frames.shape[2]
        mf = self.det.max_faces
        o = out or {}
        boxes = o.get("boxes") if out else np.zeros((n, mf), BOX_DTYPE)
        counts = o.get("counts") if out else np.zeros(n, np.int32)
        idx = o.get("idx") if out else np.zeros((n, mf), np.int64)
        score = o.get("score") if out else np.zeros((n, mf), np.float32)
        emb = np.zeros((n, mf, 512), np.float32) if want_embeddings else None
        check(lib().fr_pipeline_run(self._h, _ptr(f), w * 3, n, _ptr(boxes), _ptr(counts), _ptr(idx), _ptr(score), _ptr(emb)))
        res = {"boxes": boxes, "counts": counts, "idx": idx, "score": score}
        if want_embeddings:
            res["embeddings"] = emb
        return res

    def submit(self, frames, want_embeddings: bool = False) -> None:
        """enqueue a batch without waiting for the GPU (at most two in flight); `frames` is kept alive until it is collected"""
        if isinstance(frames, np.ndarray):
            f = np.ascontiguousarray(frames, dtype=np.uint8)
            n, w = f.shape[0], f.shape[2]
        else:
            f, n, w = frames, frames.shape[0], frames.shape[2]
        check(lib().fr_pipeline_submit(self._h, _ptr(f), w * 3, n, 1 if want_embeddings else 0))
        self._pending.append((f, n, want_embeddings))

    def collect(self, out=None):
        """results of the oldest submitted batch (same dict as run)"""
        if not self._pending:
            raise RuntimeError("nothing was submitted")
        _, n, want_embeddings = self._pending[0]
        mf = self.det.max_faces
        o = out or {}
        boxes = o.get("boxes") if out else np.zeros((n, mf), BOX_DTYPE)
        counts = o.get("counts") if out else np.zeros(n, np.int32)
        idx = o.get("idx") if out else np.zeros((n, mf), np.int64)
        score = o.get("score") if out else np.zeros((n, mf), np.float32)
        emb = np.zeros((n, mf, 512), np.float32) if want_embeddings else None
        check(lib().fr_pipeline_collect(self._h, _ptr(boxes), _ptr(counts), _ptr(idx), _ptr(score), _ptr(emb)))
        self._pending.pop(0)
        res = {"boxes": boxes, "counts": counts, "idx": idx, "score": score}
        if want_embeddings:
            res["embeddings"] = emb
        return res

    def in_flight(self) -> int:
        return int(lib().fr_pipeline_in_flight(self._h))


class Service:
    """Request batcher in front of a Pipeline (fr_service_*): infer() is thread-safe and blocking, one frame per call."""

    def __init__(self, pipe: "Pipeline", max_wait_us: int = 200):
        L = lib()
        L.fr_service_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.fr_service_destroy.restype = None
        L.fr_service_destroy.argtypes = [C.c_void_p]
        L.fr_service_infer.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_service_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        h = C.c_void_p()
        check(L.fr_service_create(pipe._h, max_wait_us, C.byref(h)))
        self._h, self.pipe = h, pipe

    def close(self) -> None:
        if self._h:
            lib().fr_service_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def infer(self, frame: np.ndarray):
        """frame: H x W x 3 u8 -> dict(boxes[max_faces], count, idx[max_faces], score[max_faces]); ctypes releases the GIL while waiting"""
        f = np.ascontiguousarray(frame, dtype=np.uint8)
        mf = self.pipe.det.max_faces
        boxes = np.zeros(mf, BOX_DTYPE)
        count = C.c_int()
        idx = np.zeros(mf, np.int64)
        score = np.zeros(mf, np.float32)
        check(lib().fr_service_infer(self._h, _ptr(f), f.shape[1] * 3, _ptr(boxes), C.byref(count), _ptr(idx), _ptr(score)))
        return {"boxes": boxes, "count": count.value, "idx": idx, "score": score}

    def stats(self):
        b, f = C.c_int64(), C.c_int64()
        check(lib().fr_service_stats(self._h, C.byref(b), C.byref(f)))
        return b.value, f.value


class Exchange:
    """Fused cross-GPU exchange + merge of per-shard top-k over NVLink peer memory (csrc/exchange.cu)."""

    def __init__(self, device: int, world: int, rank: int, nq_max: int = 256, k_max: int = 8):
        L = lib()
        L.fr_exchange_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.fr_exchange_local_handle.argtypes = [C.c_void_p, C.c_void_p]
        L.fr_exchange_connect.argtypes = [C.c_void_p, C.c_void_p]
        L.fr_exchange_connect_local.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.fr_exchange_merge_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_exchange_destroy.restype = None
        L.fr_exchange_destroy.argtypes = [C.c_void_p]
        L.fr_gallery_topk_push_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_exchange_wait_merge_dev.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fr_exchange_status.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        h = C.c_void_p()
        check(L.fr_exchange_create(device, world, rank, nq_max, k_max, C.byref(h)))
        self._h, self.device, self.world, self.rank = h, device, world, rank

    def local_handle(self) -> np.ndarray:
        out = np.zeros(lib().fr_exchange_handle_bytes(), np.uint8)
        check(lib().fr_exchange_local_handle(self._h, _ptr(out)))
        return out

    def connect(self, all_handles: np.ndarray) -> None:
        a = np.ascontiguousarray(all_handles, dtype=np.uint8)
        assert a.size == self.world * lib().fr_exchange_handle_bytes()
        check(lib().fr_exchange_connect(self._h, _ptr(a)))

    def connect_local(self, group: "list[Exchange]") -> None:
        arr = (C.c_void_p * len(group))(*[g._h for g in group])
        check(lib().fr_exchange_connect_local(self._h, arr))

    def merge_dev(self, local_scores_t, local_idx_t, scores_t, idx_t, stream: int) -> None:
        """unfused: stand-alone push of results already on the device + wait/merge (local_scores_t None = empty shard)"""
        nq, k = scores_t.shape
        check(lib().fr_exchange_merge_dev(self._h, _ptr(local_scores_t), _ptr(local_idx_t), nq, k, _ptr(scores_t), _ptr(idx_t),
                                          C.c_void_p(stream)))

    def topk_push_dev(self, gal: "Gallery", q_t, k: int, local_scores_t, local_idx_t, stream: int) -> None:
        """this shard's search with the push to the peers fused into its re-rank kernel"""
        check(lib().fr_gallery_topk_push_dev(gal._h, self._h, _ptr(q_t), q_t.shape[0], k, _ptr(local_scores_t), _ptr(local_idx_t),
                                             C.c_void_p(stream)))

    def wait_merge_dev(self, scores_t, idx_t, stream: int) -> None:
        """wait for every shard's push of the oldest unmerged batch, merge into scores_t / idx_t (nq x k)"""
        nq, k = scores_t.shape
        check(lib().fr_exchange_wait_merge_dev(self._h, nq, k, _ptr(scores_t), _ptr(idx_t), C.c_void_p(stream)))

    def status(self) -> int:
        v = C.c_int()
        check(lib().fr_exchange_status(self._h, C.byref(v)))
        return v.value

    def close(self) -> None:
        if self._h:
            lib().fr_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

class JpegDecoder:
    """fr_jpeg_*: cv::imdecode (+ cv::resize) of the request handlers (/root/reference src/app.cpp:247-256,294-301) on the GPU (nvJPEG)"""

    def __init__(self, device: int = 0):
        L = lib()
        L.fr_jpeg_decoder_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.fr_jpeg_decoder_destroy.restype = None
        L.fr_jpeg_decoder_destroy.argtypes = [C.c_void_p]
        L.fr_jpeg_info.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.fr_jpeg_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int]
        h = C.c_void_p()
        check(L.fr_jpeg_decoder_create(device, C.byref(h)))
        self._h = h

    def close(self) -> None:
        if self._h:
            lib().fr_jpeg_decoder_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self, jpeg: bytes):
        """-> (width, height) of the encoded image"""
        buf = np.frombuffer(jpeg, np.uint8)
        w, h = C.c_int(), C.c_int()
        check(lib().fr_jpeg_info(self._h, _ptr(buf) if buf.size else None, buf.size, C.byref(w), C.byref(h)))
        return w.value, h.value

    def decode(self, jpeg: bytes, size=None) -> np.ndarray:
        """-> u8 BGR frame [H, W, 3]; size = (W, H) stretches it like cv::resize(frame, Size(W, H)), None keeps the decoded size"""
        buf = np.frombuffer(jpeg, np.uint8)
        w, h = size if size is not None else self.info(jpeg)
        out = np.empty((h, w, 3), np.uint8)
        check(lib().fr_jpeg_decode(self._h, _ptr(buf) if buf.size else None, buf.size, w, h, _ptr(out), w * 3))
        return out
