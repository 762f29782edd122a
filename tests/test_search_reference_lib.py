"""GPU: pin the search oracle (and the product) against the reference's OWN code — /root/reference/src/matmul.cpp + common.cpp
compiled verbatim into oracle/_ref/libref_matmul.so (oracle/build_ref.py) and run here on the same inputs."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

import frb200
from oracle import search_oracle as so

pytestmark = pytest.mark.gpu
LIB = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "libref_matmul.so"


def _ref():
    if not LIB.exists():
        pytest.skip("oracle/_ref/libref_matmul.so not built (needs /root/reference at build time)")
    L = C.CDLL(str(LIB))
    L.ref_matmul_new.restype = C.c_void_p
    L.ref_matmul_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    L.ref_matmul_calculate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    L.ref_matmul_free.argtypes = [C.c_void_p]
    return L


@pytest.mark.parametrize("n,nq", [(1000, 1), (1000, 4), (4096, 37), (50000, 8)])
def test_reference_matmul_vs_oracle_vs_product(n, nq):
    L = _ref()
    rng = np.random.default_rng(n + nq)
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    q = so.l2_normalise(rng.standard_normal((nq, 512)))
    q[0] = so.planted_queries(G[n // 2 : n // 2 + 1], 0.5, 1)[0]
    h = L.ref_matmul_new()
    assert h
    assert L.ref_matmul_init(h, G.ctypes.data_as(C.c_void_p), n, 512) == 0
    out = np.empty((nq, n), np.float32)
    assert L.ref_matmul_calculate(h, q.ctypes.data_as(C.c_void_p), nq, out.ctypes.data_as(C.c_void_p)) == 0
    L.ref_matmul_free(h)
    # oracle == reference
    o = so.sims(G, q)
    assert np.abs(out - o).max() <= 1e-5
    ri, rv = so.get_outputs(out)
    oi, ov = so.get_outputs(o)
    assert np.array_equal(ri, oi)
    # product == reference (dense sims and fused top-1)
    g = frb200.Gallery.from_rows(G)
    mine = g.sims(q)
    assert np.abs(mine - out).max() <= 1e-5
    g.set_path(frb200.FR_PATH_TENSOR)
    s, i = g.topk(q, 1)
    assert np.array_equal(i[:, 0], ri)
    assert np.abs(s[:, 0] - rv).max() <= 1e-5
    assert ri[0] == n // 2
    g.close()
