"""ORACLE (test infrastructure, never shipped or measured as the product): CPU restatement of the reference's
cosine-similarity search path in numpy.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Parity status: **parity unpinned by reference fixtures** — /root/reference holds no tests, golden vectors or known-answer
files for this path. The oracle is instead pinned against the reference's own code: tests/gpu
(test_search_reference_lib.py) run /root/reference/src/matmul.cpp + common.cpp, compiled verbatim into
oracle/_ref/libref_matmul.so by oracle/build_ref.py, on the same inputs and compare with sims() below.

Reference code restated here (paths relative to /root/reference):
  sims()         MatMul::init + MatMul::calculate   src/matmul.cpp:9-14,36-77 ; layout src/matmul.h:8-15
  get_outputs()  ArcFaceIR50::getOutputs            src/arcface.cpp:203-217  (std::max_element = FIRST maximum)
  topk()         k-best generalisation of get_outputs, order (score desc, row asc)  (k = 1 equals get_outputs)
synth_rows() is not reference code: it regenerates, bit for bit, the rows written by the library's synthetic gallery
generator (face-recognition-cpp-tensorrt_b200/csrc/search_kernels.cuh, synth_rows_kernel) so that 10M-row benchmarks can
be checked on the host without holding the gallery there.
"""
from __future__ import annotations

import numpy as np

DIM = 512

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def synth_rows(global_rows, seed: int) -> np.ndarray:
    """rows of the synthetic gallery with the given GLOBAL row ids -> (len, 512) float32, bit-identical to the device."""
    g = np.asarray(global_rows, dtype=np.uint64).reshape(-1)
    with np.errstate(over="ignore"):
        key = _mix64(np.uint64(seed) ^ (g * np.uint64(0xD6E8FEB86659FD93)))
        h = _mix64(key[:, None] + np.arange(DIM, dtype=np.uint64)[None, :])
    m = np.uint64(0xFFFF)
    v = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))).astype(np.int64) - 131070
    ss = (v * v).sum(axis=1)
    scale = np.where(ss > 0, 1.0 / np.sqrt(ss.astype(np.float64)), 0.0)
    return (v.astype(np.float64) * scale[:, None]).astype(np.float32)


def sims(gallery: np.ndarray, q: np.ndarray) -> np.ndarray:
    """out[i, j] = <q_i, gallery_j>, fp32 (CUDA_R_32F / CUBLAS_COMPUTE_32F, src/matmul.h:24-25), row-major n x m as
    MatMul::calculate leaves it in `outputs` (src/matmul.cpp:9-14: ldc = m with column-major C == row-major n x m)."""
    gallery = np.ascontiguousarray(gallery, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    return q @ gallery.T


def get_outputs(sim: np.ndarray):
    """per query: index of the FIRST maximum and its value (src/arcface.cpp:210-211)."""
    idx = np.argmax(sim, axis=1)  # numpy argmax returns the first occurrence, like std::max_element
    return idx.astype(np.int64), sim[np.arange(sim.shape[0]), idx]


def topk(sim: np.ndarray, k: int, row_offset: int = 0):
    """k best per query ordered by (score desc, row asc); missing entries (-inf, -1)."""
    nq, n = sim.shape
    scores = np.full((nq, k), -np.inf, np.float32)
    idx = np.full((nq, k), -1, np.int64)
    kk = min(k, n)
    if kk:
        order = np.lexsort((np.broadcast_to(np.arange(n), sim.shape), -sim.astype(np.float64)), axis=1)[:, :kk]
        scores[:, :kk] = np.take_along_axis(sim, order, axis=1)
        idx[:, :kk] = order + row_offset
    return scores, idx


def merge_topk(parts_scores, parts_idx, k: int):
    """merge per-shard results (list of nq x k) by (score desc, idx asc) — the step after the all-gather."""
    s = np.concatenate(parts_scores, axis=1)
    i = np.concatenate(parts_idx, axis=1)
    key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((key_i, -s.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(s, order, axis=1), np.take_along_axis(i, order, axis=1)


def l2_normalise(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, np.float32)
    return (x / np.linalg.norm(x.astype(np.float64), axis=1, keepdims=True)).astype(np.float32)


def planted_queries(rows: np.ndarray, noise: float, seed: int) -> np.ndarray:
    """queries = normalise(row + noise * unit gaussian direction): cos(query, row) ~ 1/sqrt(1+noise^2)."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal(rows.shape).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return l2_normalise(rows + noise * d)
