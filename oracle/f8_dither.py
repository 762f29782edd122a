"""ORACLE / TEST INFRASTRUCTURE ONLY (never imported by the product): numpy emulation of the certified e4m3 scan of
face-recognition-cpp-tensorrt_b200/csrc/search_kernels.cuh — the stochastic e4m3 rounding (f8_round_dither, dither_key / dither_bits),
the per-query margin / certificate gap (prep_queries_kernel<true>), the row bounds (make_f8_copy_kernel) and the accept / recompute
decision of append_rerank_kernel. It lets the CPU suite run the exactness argument on structured galleries (the reference's
contract is "first maximum of the exact fp32 scores", /root/reference src/arcface.cpp:203-217).

Constants are parsed from the .cuh so the emulation cannot drift from the kernels."""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

_SRC = (Path(__file__).resolve().parent.parent / "face-recognition-cpp-tensorrt_b200" / "csrc" / "search_kernels.cuh").read_text()


def _const(name: str) -> float:
    m = re.search(rf"constexpr float {name} = ([0-9.eE+-]+)f;", _SRC)
    assert m, name
    return float(m.group(1))


SCALE, LOGP, GAP_FRAC, ACC_EPS, F8_MAX = (_const(n) for n in ("kF8Scale", "kF8LogP", "kF8GapFrac", "kF8AccEps", "kF8Max"))
QUERY_SALT = np.uint64(int(re.search(r"kQueryDitherSalt = (0x[0-9A-Fa-f]+)ull", _SRC).group(1), 16))
DEFAULT_SEED = 0x5EEDF8B200
_U = np.uint64


def _mix(z):
    with np.errstate(over="ignore"):
        z = z + _U(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)
        return z ^ (z >> _U(31))


def dither_key(seed, rows):
    with np.errstate(over="ignore"):
        return _mix(_U(seed) ^ (np.asarray(rows, dtype=np.uint64) * _U(0xD6E8FEB86659FD93)))


def dither_r24(keys, dim=512):
    """[rows, dim] 24-bit uniforms: column c uses pair c // 2, low word for even c, high word for odd c"""
    pair = (np.arange(dim, dtype=np.uint64) // _U(2))[None, :]
    with np.errstate(over="ignore"):
        h = _mix(keys[:, None] + pair)
    lo = (h & _U(0xFFFFFF)).astype(np.uint32)
    hi = ((h >> _U(32)) & _U(0xFFFFFF)).astype(np.uint32)
    return np.where((np.arange(dim) % 2 == 0)[None, :], lo, hi)


def bracket(a):
    """a: f32 scaled magnitudes in [0, 448] -> (lo, step), exactly as f8_bracket"""
    a = a.astype(np.float32)
    e = ((a.view(np.uint32) >> 23) & 0xFF).astype(np.int32) - 127
    e = np.maximum(e, -6)
    step = np.ldexp(np.float32(1), e - 3).astype(np.float32)
    lo = (np.floor(a / step) * step).astype(np.float32)
    return lo, step


def round_dither(x, r24):
    """x f32 (cosine units) -> (rounded scaled value f32, step u (0 if representable), outward magnitude abar, saturated mask)"""
    x = x.astype(np.float32)
    a = np.abs(x) * np.float32(SCALE)
    sat = a > np.float32(F8_MAX)
    a = np.minimum(a, np.float32(F8_MAX))
    lo, step = bracket(a)
    frac = ((a - lo) / step).astype(np.float32)
    inexact = frac > 0
    up = r24.astype(np.float32) < frac * np.float32(16777216.0)
    r = np.where(up, lo + step, lo).astype(np.float32)
    return np.where(x < 0, -r, r).astype(np.float32), np.where(inexact, step, 0).astype(np.float32), np.where(inexact, lo + step, lo).astype(np.float32), sat


def round_nearest(x):
    """the round-to-nearest e4m3 image the scan used before (ties to even on the 3-bit mantissa) — kept to show what it gets wrong"""
    x = x.astype(np.float32)
    a = np.minimum(np.abs(x) * np.float32(SCALE), np.float32(F8_MAX))
    lo, step = bracket(a)
    k = np.floor(a / step)
    frac = a / step - k
    up = (frac > 0.5) | ((frac == 0.5) & (k % 2 == 1))
    r = np.where(up, lo + step, lo).astype(np.float32)
    return np.where(x < 0, -r, r).astype(np.float32)


def dither_gallery(G, seed=DEFAULT_SEED, first_row_id=0):
    """make_f8_copy_kernel: -> (Ghat scaled f32 [n,512], g4max, w4max)"""
    G = np.ascontiguousarray(G, np.float32)
    keys = dither_key(seed, np.arange(G.shape[0], dtype=np.uint64) + _U(first_row_id))
    gh, u, _, sat = round_dither(G, dither_r24(keys))
    assert not sat.any(), "rows must be L2-normalised for the e4m3 copy"
    g4 = float(((G.astype(np.float64) ** 4).sum(1)).max()) * 1.0001
    w4 = float((((u.astype(np.float64) / SCALE) ** 4).sum(1)).max()) * 1.0001
    return gh, g4, w4


def dither_queries(q, g4max, w4max, gmax, seed=DEFAULT_SEED, logp=LOGP):
    """prep_queries_kernel<true>: -> (Qhat scaled, margin m, gap, E) per query, cosine units"""
    q = np.ascontiguousarray(q, np.float32)
    keys = dither_key(int(_U(seed) ^ QUERY_SALT), np.arange(q.shape[0], dtype=np.uint64))
    qh, u, abar, sat = round_dither(q, dither_r24(keys))
    qbar4 = ((abar.astype(np.float64) / SCALE) ** 4).sum(1)
    uq4 = ((u.astype(np.float64) / SCALE) ** 4).sum(1)
    qbar2 = ((abar.astype(np.float64) / SCALE) ** 2).sum(1)
    V = np.sqrt(qbar4 * w4max) + np.sqrt(g4max * uq4)
    E = (np.sqrt(0.5 * logp * V) + ACC_EPS * np.sqrt(qbar2) * 1.125 * gmax) * 1.001 + 1e-7
    gap = np.where(sat.any(1), -np.inf, GAP_FRAC * E)
    return qh, (1 + GAP_FRAC) * E, gap, E


def certified_top1(G, q, seed=DEFAULT_SEED, logp=LOGP, rounding="dither"):
    """Emulates the top-1 search on the e4m3 copy. Returns dict(idx [nq] (-1 where the query is handed to the exact scan),
    flagged [nq] bool, n_cand [nq], best_in_cand [nq] bool, exact_idx [nq], margin [nq])."""
    G = np.ascontiguousarray(G, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    gmax = float(np.sqrt((G.astype(np.float64) ** 2).sum(1).max())) * 1.0000002
    gh, g4, w4 = dither_gallery(G, seed)
    qh, m, gap, E = dither_queries(q, g4, w4, gmax, seed, logp)
    if rounding == "nearest":
        gh, qh = round_nearest(G), round_nearest(q)
    coarse = (qh.astype(np.float64) @ gh.astype(np.float64).T) / (SCALE * SCALE)   # exact products, wide accumulation
    exact = (q @ G.T).astype(np.float32)
    exact_idx = exact.argmax(1)                                                   # first maximum, src/arcface.cpp:210
    ck = coarse.max(1)
    cand = coarse >= (ck - m)[:, None]
    n_margin = cand.sum(1)
    # append_rerank_kernel's exact-leader filter: L0 = best EXACT score among a few coarse leaders (the kernel takes its eight per-warp
    # leaders, which always include the overall coarse leader; here the eight best coarse rows); only rows with coarse >= L0 - E stay.
    lead = np.argsort(-np.where(cand, coarse, -np.inf), axis=1, kind="stable")[:, :8]
    rows = np.arange(q.shape[0])[:, None]
    l0 = np.where(cand[rows, lead], exact[rows, lead], -np.inf).max(1)
    cand &= coarse >= np.maximum(ck - m, l0 - E)[:, None]
    masked = np.where(cand, exact, -np.inf)
    idx = masked.argmax(1)
    L = masked.max(1)
    flagged = (L < ck - gap) | (n_margin > 4096)
    return {"idx": np.where(flagged, -1, idx), "flagged": flagged, "n_cand": cand.sum(1), "n_margin": n_margin, "best_in_cand": cand[np.arange(q.shape[0]), exact_idx],
            "exact_idx": exact_idx, "margin": m, "E": E, "coarse": coarse, "exact": exact}
