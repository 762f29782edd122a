"""GPU parity of the cosine-similarity search through the C ABI against oracle/search_oracle.py.
Tolerances: row indices bit-exact (top-1 identity exact, BASELINE.json north_star); scores |d| <= 1e-5 (north_star allows 1e-3)."""
import os

import numpy as np
import pytest

import frb200
from oracle import search_oracle as so

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-5


def _check_topk(g_scores, g_idx, sim, k, row_offset=0):
    o_scores, o_idx = so.topk(sim, k, row_offset)
    assert g_idx.shape == o_idx.shape
    assert np.all(np.abs(np.where(np.isinf(o_scores), 0, g_scores - o_scores)) <= SCORE_TOL)
    assert np.array_equal(np.isinf(g_scores), np.isinf(o_scores))
    bad = np.argwhere(g_idx != o_idx)
    for qi, j in bad:  # a differing index is only acceptable for an fp32-rounding-level tie between distinct rows
        gi = g_idx[qi, j] - row_offset
        assert gi >= 0 and abs(sim[qi, gi] - o_scores[qi, j]) <= 2e-6, (qi, j, g_idx[qi, j], o_idx[qi, j])
    # top-1 must be exact on these inputs (planted or well separated)
    return len(bad)


def test_synthetic_rows_bit_exact_with_oracle():
    for n, off, seed in ((1000, 0, 19), (300, 9_999_900, 19), (257, 12345678000, 7)):
        g = frb200.Gallery.synthetic(n, seed=seed, row_offset=off)
        got = g.read_rows(0, n)
        want = so.synth_rows(np.arange(off, off + n), seed)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        g.close()


@pytest.mark.parametrize("path", [frb200.FR_PATH_EXACT, frb200.FR_PATH_TENSOR, frb200.FR_PATH_AUTO])
def test_config1_one_query_1000_rows_top1_exact(path):
    # BASELINE.json configs[0] / SURVEY §8d config 1
    rng = np.random.default_rng(1234)
    G = so.l2_normalise(rng.standard_normal((1000, 512)))
    g = frb200.Gallery.from_rows(G)
    g.set_path(path)
    rq = np.random.default_rng(1235)
    for r in (0, 999, 517, 255, 256):
        q = so.l2_normalise(G[r : r + 1] + 0.5 * rq.standard_normal((1, 512)).astype(np.float32) / np.sqrt(512) * 1.0)
        s, i = g.topk(q, 1)
        oi, ov = so.get_outputs(so.sims(G, q))
        assert i[0, 0] == oi[0] == r
        assert abs(s[0, 0] - ov[0]) <= SCORE_TOL
    g.close()


@pytest.mark.parametrize("path", [frb200.FR_PATH_EXACT, frb200.FR_PATH_TENSOR])
def test_adversarial_ties(path):
    rng = np.random.default_rng(5)
    G = so.l2_normalise(rng.standard_normal((3000, 512)))
    # exact duplicate rows -> lowest index wins (std::max_element, src/arcface.cpp:210)
    G[2900] = G[40]
    G[1500] = G[40]
    G[41] = G[2999]
    q = np.stack([G[40], G[2999]])
    g = frb200.Gallery.from_rows(G)
    g.set_path(path)
    s, i = g.topk(q, 4)
    assert i[0, :3].tolist() == [40, 1500, 2900] and i[1, :2].tolist() == [41, 2999]
    assert np.all(np.abs(s[:, 0] - 1.0) < 1e-5)
    g.close()
    # all rows identical -> every score equal -> indices 0..k-1
    G2 = np.repeat(G[:1], 777, axis=0)
    g = frb200.Gallery.from_rows(G2)
    g.set_path(path)
    s, i = g.topk(G[:1], 8)
    assert i[0].tolist() == list(range(8))
    # query orthogonal to everything / zero query: all scores 0 -> first row
    s, i = g.topk(np.zeros((1, 512), np.float32), 1)
    assert i[0, 0] == 0 and s[0, 0] == 0.0
    g.close()


@pytest.mark.parametrize("n,nq", [(1, 1), (7, 3), (1000, 4), (5000, 33), (70000, 17)])
def test_dense_sims_matches_matmul_calculate_contract(n, nq):
    rng = np.random.default_rng(n + nq)
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    q = so.l2_normalise(rng.standard_normal((nq, 512)))
    g = frb200.Gallery.from_rows(G)
    out = g.sims(q)
    ref = (q.astype(np.float64) @ G.astype(np.float64).T)
    assert out.shape == (nq, n)
    assert np.abs(out - ref).max() <= 2e-6
    assert np.abs(out - so.sims(G, q)).max() <= SCORE_TOL
    g.close()


@pytest.mark.parametrize("n", [256, 1000, 4096 + 37, 33333, 150_001])
@pytest.mark.parametrize("nq,k", [(1, 1), (5, 4), (128, 1), (129, 8), (256, 1), (300, 2)])
def test_tensor_path_topk_matches_oracle(n, nq, k):
    rng = np.random.default_rng(n * 7 + nq)
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    planted = rng.integers(0, n, nq)
    q = so.planted_queries(G[planted], noise=0.75, seed=n + nq)
    g = frb200.Gallery.from_rows(G, row_offset=1000)
    g.set_path(frb200.FR_PATH_TENSOR)
    s, i = g.topk(q, k)
    sim = so.sims(G, q)
    _check_topk(s, i, sim, k, row_offset=1000)
    assert np.array_equal(i[:, 0], planted + 1000)
    st = g.last_stats()
    assert st.launches >= 2 and st.scan_bytes == n * 512 * 2 * ((nq + 255) // 256)
    g.close()


def test_tensor_and_exact_paths_agree_bitwise_on_scores():
    rng = np.random.default_rng(77)
    G = so.l2_normalise(rng.standard_normal((20000, 512)))
    q = so.l2_normalise(rng.standard_normal((64, 512)))
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_EXACT)
    s1, i1 = g.topk(q, 4)
    g.set_path(frb200.FR_PATH_TENSOR)
    s2, i2 = g.topk(q, 4)
    assert np.array_equal(i1, i2)
    assert np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    dense = g.sims(q)
    assert np.array_equal(dense[np.arange(64), i1[:, 0]].view(np.uint32), s1[:, 0].view(np.uint32))
    g.close()


def test_edge_cases_and_errors():
    G = so.l2_normalise(np.random.default_rng(1).standard_normal((5, 512)))
    g = frb200.Gallery.from_rows(G)
    s, i = g.topk(G[:2], 8)  # fewer rows than k -> (-inf, -1) padding
    assert i[0, :1].tolist() == [0] and i[1, 0] == 1 and np.all(i[:, 5:] == -1) and np.all(np.isinf(s[:, 5:]))
    g.set_path(frb200.FR_PATH_TENSOR)
    s, i = g.topk(G[:2], 8)
    assert i[0, 0] == 0 and i[1, 0] == 1 and np.all(i[:, 5:] == -1)
    with pytest.raises(frb200.FrError) as e:
        g.topk(G[:1], 9)
    assert e.value.code == frb200.FR_EINVAL
    with pytest.raises(frb200.FrError) as e:
        g.topk(np.zeros((0, 512), np.float32), 1)
    assert e.value.code == frb200.FR_EINVAL
    g.close()
    empty = frb200.Gallery.from_rows(np.zeros((0, 512), np.float32))
    with pytest.raises(frb200.FrError) as e:  # featureMatching throws on an empty database, src/arcface.cpp:195-199
        empty.topk(G[:1], 1)
    assert e.value.code == frb200.FR_ESTATE and "No faces in database" in e.value.msg
    empty.close()
    with pytest.raises(frb200.FrError) as e:
        frb200.Gallery.from_rows(np.zeros((4, 128), np.float32))
    assert e.value.code == frb200.FR_EINVAL


def test_sharded_search_merge_equals_single_gallery():
    import torch

    rng = np.random.default_rng(9)
    n, nq, k = 50_000, 200, 4
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    G[41000] = G[123]
    planted = rng.integers(0, n, nq)
    planted[0] = 123
    q = so.planted_queries(G[planted], noise=0.6, seed=4)
    q[0] = G[123]
    full = frb200.Gallery.from_rows(G)
    full.set_path(frb200.FR_PATH_TENSOR)
    fs, fi = full.topk(q, k)
    for shards in (2, 3, 8):
        bounds = np.linspace(0, n, shards + 1).astype(int)
        ps = torch.empty((shards, nq, k), dtype=torch.float32, device="cuda")
        pi = torch.empty((shards, nq, k), dtype=torch.int64, device="cuda")
        qd = torch.from_numpy(q).cuda()
        ms = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        mi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        st = torch.cuda.Stream()  # one explicit stream for the shard searches and the merge (a NULL stream means "handle's own")
        gs = []
        for r, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
            sh = frb200.Gallery.from_rows(G[a:b], row_offset=int(a))
            sh.set_path(frb200.FR_PATH_TENSOR)
            sh.topk_dev(qd, k, ps[r], pi[r], stream=st.cuda_stream)
            gs.append(sh)
        frb200.topk_merge_dev(ps, pi, shards, nq, k, ms, mi, 0, stream=st.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(mi.cpu().numpy(), fi)
        assert np.array_equal(ms.cpu().numpy().view(np.uint32), fs.view(np.uint32))
        for sh in gs:
            sh.close()
    assert fi[0, 0] == 123 and fi[0, 1] == 41000
    full.close()


def test_large_synthetic_gallery_planted_top1():
    # size-independent property at a size the host cannot hold as a matrix: planted rows must come back as top-1 with the
    # score the host computes from the regenerated row
    n, seed = 3_000_000, 19
    g = frb200.Gallery.synthetic(n, seed=seed)
    rng = np.random.default_rng(23)
    planted = np.sort(rng.integers(0, n, 256))
    planted[:3] = [0, n - 1, n - 257]
    rows = so.synth_rows(planted, seed)
    q = so.planted_queries(rows, noise=0.75, seed=29)
    s, i = g.topk(q, 2)
    assert np.array_equal(i[:, 0], planted)
    want = np.einsum("ij,ij->i", q.astype(np.float64), rows.astype(np.float64))
    assert np.abs(s[:, 0] - want).max() <= SCORE_TOL
    # runner-up is an impostor: far below, and its reported score equals the host dot with the regenerated row
    r2 = so.synth_rows(i[:, 1], seed)
    want2 = np.einsum("ij,ij->i", q.astype(np.float64), r2.astype(np.float64))
    assert np.abs(s[:, 1] - want2).max() <= SCORE_TOL and np.all(s[:, 1] < 0.4)
    g.close()


@pytest.mark.parametrize("cluster,k", [(12, 1), (40, 4), (100, 8)])
def test_near_tie_cluster_inside_fp16_margin(cluster, k):
    # many rows within the fp16 error margin of the best one: the coarse pass cannot order them; the exact re-score
    # (<= 64 rows) or, beyond that, the flagged exact scan must reproduce the fp32 order
    rng = np.random.default_rng(cluster)
    n = 30_000
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    base = G[777].copy()
    where = rng.choice(n, cluster, replace=False)
    G[where] = so.l2_normalise(base[None, :] + 2e-4 * rng.standard_normal((cluster, 512)).astype(np.float32) / np.sqrt(512))
    q = np.concatenate([base[None, :], so.l2_normalise(rng.standard_normal((3, 512)))]).astype(np.float32)
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_TENSOR)
    s, i = g.topk(q, k)
    sim = so.sims(G, q)
    _check_topk(s, i, sim, k)
    assert set(i[0].tolist()) <= set(where.tolist()) | {777}
    g.close()


@pytest.mark.parametrize("world", [1, 2, 4])
def test_fused_exchange_merge_protocol(world):
    # the peer-memory exchange (csrc/exchange_impl.cuh, unfused form) with `world` shard "ranks" sharing this GPU, one stream per rank:
    # every rank must end up with the single-gallery result; repeated calls exercise the four-slot mailbox and the device-resident counters
    import torch

    rng = np.random.default_rng(world)
    n, nq, k = 40_000, 256, 2
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    G[n - 5] = G[17]
    bounds = np.linspace(0, n, world + 1).astype(int)
    shards = [frb200.Gallery.from_rows(G[a:b], row_offset=int(a)) for a, b in zip(bounds[:-1], bounds[1:])]
    for s in shards:
        s.set_path(frb200.FR_PATH_TENSOR)
    group = [frb200.Exchange(0, world, r, nq_max=256, k_max=8) for r in range(world)]
    for x in group:
        x.connect_local(group)
    streams = [torch.cuda.Stream() for _ in range(world)]
    full = frb200.Gallery.from_rows(G)
    full.set_path(frb200.FR_PATH_TENSOR)
    for call in range(5):
        planted = rng.integers(0, n, nq)
        planted[0] = 17
        q = so.planted_queries(G[planted], 0.6, call)
        q[0] = G[17]
        qd = torch.from_numpy(q).cuda()
        outs = []
        torch.cuda.synchronize()
        for r in range(world):
            ls = torch.empty((nq, k), dtype=torch.float32, device="cuda")
            li = torch.empty((nq, k), dtype=torch.int64, device="cuda")
            os_ = torch.empty((nq, k), dtype=torch.float32, device="cuda")
            oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
            shards[r].topk_dev(qd, k, ls, li, stream=streams[r].cuda_stream)
            outs.append((os_, oi, ls, li))
        # on ONE shared GPU the spinning exchange blocks of one rank would keep the full-SM scan kernel of another from being
        # scheduled, so all scans are enqueued (and finished) first; on separate GPUs the two calls are simply back to back
        torch.cuda.synchronize()
        for r in range(world):
            os_, oi, ls, li = outs[r]
            group[r].merge_dev(ls, li, os_, oi, stream=streams[r].cuda_stream)
        torch.cuda.synchronize()
        fs, fi = full.topk(q, k)
        for os_, oi, _, _ in outs:
            assert np.array_equal(oi.cpu().numpy(), fi)
            assert np.array_equal(os_.cpu().numpy().view(np.uint32), fs.view(np.uint32))
        assert fi[0, 0] == 17 and fi[0, 1] == n - 5
    for o in shards + [full] + group:
        o.close()


@pytest.mark.parametrize("world,k,scan,lag", [(1, 1, 0, 0), (2, 1, 0, 1), (4, 1, 1, 1), (3, 4, 0, 1), (4, 2, 1, 0)])
def test_search_with_fused_push_and_lagged_merge(world, k, scan, lag):
    # fr_gallery_topk_push_dev (re-rank kernel delivers to the peers) + fr_exchange_wait_merge_dev, `world` shard ranks sharing this
    # GPU (one stream each). lag = 1: the merge of batch i is issued after the search of batch i + 1 (the throughput mode of bench.py);
    # batches alternate between different query sets so that a result delivered for the wrong batch would be caught. One shard is
    # EMPTY when world >= 3 (it still has to take part), one query per batch is forced through the exact-scan fix-up (all rows of a
    # cluster inside the margin), whose merging block must deliver it.
    import torch

    rng = np.random.default_rng(100 + world + k)
    n, nq = 50_000, 256
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    G[n - 5] = G[17]
    base = G[4000].copy()
    G[20_000:26_000] = so.l2_normalise(base[None, :] + 1e-4 * rng.standard_normal((6000, 512)).astype(np.float32) / np.sqrt(512))
    cuts = np.linspace(0, n, world + 1).astype(int)
    if world >= 3:
        cuts[2] = cuts[1]                      # rank 1 holds no rows
    shards = [frb200.Gallery.from_rows(G[a:b], row_offset=int(a)) for a, b in zip(cuts[:-1], cuts[1:])]
    for sh in shards:
        sh.set_path(frb200.FR_PATH_TENSOR)
        if sh.rows:
            sh.set_scan(scan)
    group = [frb200.Exchange(0, world, r, nq_max=256, k_max=8) for r in range(world)]
    for x in group:
        x.connect_local(group)
    streams = [torch.cuda.Stream() for _ in range(world)]
    batches, want = [], []
    for b in range(6):
        planted = rng.integers(0, n, nq)
        planted[0] = 17
        q = so.planted_queries(G[planted], 0.6, b)
        q[0] = G[17]
        q[1] = base                              # thousands of rows inside the margin: exact-scan fix-up on the shard that holds them
        batches.append(torch.from_numpy(q).cuda())
        want.append(so.topk(so.sims(G, q), k))
    loc = [(torch.empty((nq, k), dtype=torch.float32, device="cuda"), torch.empty((nq, k), dtype=torch.int64, device="cuda")) for _ in range(world)]
    out = [(torch.empty((nq, k), dtype=torch.float32, device="cuda"), torch.empty((nq, k), dtype=torch.int64, device="cuda")) for _ in range(world)]

    def check(b):
        torch.cuda.synchronize()
        ws, wi = want[b]
        for os_, oi in out:
            gi = oi.cpu().numpy()
            gs = os_.cpu().numpy()
            assert np.abs(gs - ws).max() <= SCORE_TOL
            bad = np.argwhere(gi != wi)
            for qi, j in bad:   # only fp32-level ties between distinct rows may differ
                assert abs(so.sims(G[gi[qi, j]][None, :], batches[b][qi].cpu().numpy()[None, :])[0, 0] - ws[qi, j]) <= 2e-6
            assert gi[0, 0] == 17

    merged = 0
    for b in range(6):
        for r in range(world):                   # all searches (full-SM kernels) first: the ranks share ONE GPU here
            group[r].topk_push_dev(shards[r], batches[b], k, loc[r][0], loc[r][1], stream=streams[r].cuda_stream)
        torch.cuda.synchronize()
        if b >= lag:
            for r in range(world):
                group[r].wait_merge_dev(out[r][0], out[r][1], stream=streams[r].cuda_stream)
            check(merged)
            merged += 1
    while merged < 6:                            # drain
        for r in range(world):
            group[r].wait_merge_dev(out[r][0], out[r][1], stream=streams[r].cuda_stream)
        check(merged)
        merged += 1
    assert all(x.status() == 0 for x in group)
    assert any(sh.rows and sh.last_flagged() >= 1 for sh in shards)
    # host-buffer entry point of the sharded search (what bench.py's e2e times)
    if world == 1:
        s1, i1 = frb200.search_topk(shards[0], group[0], batches[2].cpu().numpy(), k)
        assert np.abs(s1 - want[2][0]).max() <= SCORE_TOL and i1[0, 0] == 17
        s2, i2 = frb200.search_topk(shards[0], None, batches[2].cpu().numpy(), k)
        assert np.array_equal(i1, i2)
    for o in shards + group:
        o.close()


def test_exchange_reports_a_missing_peer_instead_of_trapping(monkeypatch):
    # a rank that never arrives must not kill the context. With a 200 ms timeout the waiting rank substitutes (-inf, -1)
    # for the missing shard, finishes, and fr_exchange_status names the rank; the device stays usable.
    import torch

    monkeypatch.setenv("FR_XCHG_TIMEOUT_MS", "200")
    rng = np.random.default_rng(8)
    G = so.l2_normalise(rng.standard_normal((6000, 512)))
    q = torch.from_numpy(so.planted_queries(G[rng.integers(0, 3000, 32)], 0.6, 1)).cuda()
    sh = frb200.Gallery.from_rows(G[:3000])
    sh.set_path(frb200.FR_PATH_TENSOR)
    group = [frb200.Exchange(0, 2, r, nq_max=32, k_max=1) for r in range(2)]
    for x in group:
        x.connect_local(group)
    st = torch.cuda.Stream()
    ls, li = torch.empty((32, 1), dtype=torch.float32, device="cuda"), torch.empty((32, 1), dtype=torch.int64, device="cuda")
    os_, oi = torch.empty((32, 1), dtype=torch.float32, device="cuda"), torch.empty((32, 1), dtype=torch.int64, device="cuda")
    group[0].topk_push_dev(sh, q, 1, ls, li, stream=st.cuda_stream)      # rank 1 never searches
    group[0].wait_merge_dev(os_, oi, stream=st.cuda_stream)
    torch.cuda.synchronize()
    assert group[0].status() == 2                                         # 1 + the rank that never arrived
    assert torch.equal(oi, li) and torch.equal(os_, ls)                   # the local shard's answer survives
    s, i = sh.topk(q.cpu().numpy(), 1)                                    # context still alive
    assert np.array_equal(i, li.cpu().numpy())
    for o in [sh] + group:
        o.close()


@pytest.mark.parametrize("n,nq,k", [(1000, 1, 1), (150_001, 256, 1), (33_333, 129, 4), (70_000, 64, 8), (300, 256, 1), (700, 200, 1), (20_000, 256, 1)])
def test_fp8_scan_copy_topk(n, nq, k):
    # opt-in e4m3 scan copy (FR_SCAN_F8): the coarse pass is fp8, the returned scores / order come from the exact fp32 re-score
    rng = np.random.default_rng(n + nq + k)
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    G[n - 2] = G[3]
    planted = rng.integers(0, n, nq)
    planted[0] = 3
    q = so.planted_queries(G[planted], noise=0.75, seed=n)
    q[0] = G[3]
    g = frb200.Gallery.from_rows(G, row_offset=7)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    s, i = g.topk(q, k)
    _check_topk(s, i, so.sims(G, q), k, row_offset=7)
    assert np.array_equal(i[:, 0], planted + 7)
    st = g.last_stats()
    # top-1 streams the e4m3 copy; k > 1 on an e4m3 gallery is served by the resident fp16 copy (sorted lists need its narrow margin)
    assert st.scan_bytes == n * (512 if k == 1 else 1024)
    # same answers as the provable fp16 scan, bit for bit
    g.set_scan(frb200.FR_SCAN_F16)
    s2, i2 = g.topk(q, k)
    assert np.array_equal(i, i2) and np.array_equal(s.view(np.uint32), s2.view(np.uint32))
    g.close()
    # rows that are not L2-normalised are refused
    g2 = frb200.Gallery.from_rows(2.0 * G[:100])
    with pytest.raises(frb200.FrError) as e:
        g2.set_scan(frb200.FR_SCAN_F8)
    assert e.value.code == frb200.FR_ESTATE
    g2.close()


def test_fp8_scan_unknown_queries_and_near_ties():
    # the hard case for the e4m3 coarse pass: queries WITHOUT a match (hundreds of impostors inside the fp8 margin of the best one)
    # and a cluster of near-duplicates of the match. Top-1 must still equal the exact fp32 answer, without the exact-scan fallback
    # for the unknown queries (append epilogue + wide re-score).
    rng = np.random.default_rng(808)
    n = 200_000
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    base = G[4242].copy()
    where = rng.choice(n, 40, replace=False)
    G[where] = so.l2_normalise(base[None, :] + 2e-3 * rng.standard_normal((40, 512)).astype(np.float32) / np.sqrt(512))
    q = so.l2_normalise(rng.standard_normal((256, 512))).astype(np.float32)
    q[7] = base
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    s, i = g.topk(q, 1)
    flagged = g.last_flagged()
    sim = so.sims(G, q)
    _check_topk(s, i, sim, 1)
    assert i[7, 0] in set(where.tolist()) | {4242}
    assert flagged == 0, f"{flagged} queries fell back to the exact scan"
    # k > 1 on an e4m3 gallery (served by the fp16 copy): same contract, and no exact-scan fallback either
    s4, i4 = g.topk(q[:64], 4)
    _check_topk(s4, i4, sim[:64], 4)
    assert g.last_flagged() == 0
    g.close()


def _e4m3_bytes_to_float(b):
    import torch

    return torch.from_numpy(np.ascontiguousarray(b)).view(torch.float8_e4m3fn).float().numpy()


def test_fp8_dither_image_and_margins_match_the_emulation():
    # the CPU suite proves the exactness argument on oracle/f8_dither.py's emulation; here the kernels' e4m3 images (rows AND
    # queries), the row bounds and the per-query margin / certificate gap are checked against that emulation, bit for bit where
    # they are integers/bytes
    from oracle import f8_dither as fd

    rng = np.random.default_rng(21)
    n, off = 5000, 1234567
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    G[7] = np.where(rng.random(512) < 0.5, -1, 1).astype(np.float32) / np.sqrt(512)        # binarised row
    G[8, :] = 0
    G[8, 3] = 1.0                                                                         # one-hot row (256 is representable)
    q = so.l2_normalise(rng.standard_normal((200, 512))).astype(np.float32)
    q[5] = G[7]
    g = frb200.Gallery.from_rows(G, row_offset=off)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    g.topk(q, 1)
    keys = fd.dither_key(fd.DEFAULT_SEED, np.arange(n, dtype=np.uint64) + np.uint64(off))
    want, u, _, _ = fd.round_dither(G, fd.dither_r24(keys))
    got = _e4m3_bytes_to_float(g.debug_read(0, 0, n))
    assert np.array_equal(got, want)
    gmax, g4max, w4max = g.debug_read(4)
    _, g4, w4 = fd.dither_gallery(G, first_row_id=off)
    assert abs(g4max / g4 - 1) < 1e-4 and abs(w4max / w4 - 1) < 1e-4
    qh, m, gap, E = fd.dither_queries(q, float(g4max), float(w4max), float(gmax))
    qimg = _e4m3_bytes_to_float(g.debug_read(1).reshape(-1)[: 256 * 512].reshape(256, 512))
    assert np.array_equal(qimg[:200], qh) and not qimg[200:].any()
    assert np.allclose(g.debug_read(2)[:200] / fd.SCALE ** 2, m, rtol=2e-4) and np.allclose(g.debug_read(3)[:200], gap, rtol=2e-4)
    # appended rows continue the same dither stream (global row id), removed rows move their bytes
    extra = so.l2_normalise(rng.standard_normal((300, 512)))
    g.append(extra)
    keys2 = fd.dither_key(fd.DEFAULT_SEED, np.arange(n, n + 300, dtype=np.uint64) + np.uint64(off))
    assert np.array_equal(_e4m3_bytes_to_float(g.debug_read(0, n, 300)), fd.round_dither(extra, fd.dither_r24(keys2))[0])
    g.close()


def test_fp8_tensor_core_accumulation_is_inside_eps_det():
    # the certificate budgets kF8AccEps |qbar| |ghat| for the MMA pipe's fp32 accumulation of the (exact) e4m3 products. Measure it:
    # coarse scores of a k = 8 search (sorted-list epilogue keeps them) against the exact sum of the emulated operands.
    from oracle import f8_dither as fd

    rng = np.random.default_rng(33)
    n = 40_000
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    planted = rng.integers(0, n, 256)
    q = so.planted_queries(G[planted], noise=0.75, seed=3)
    q[:64] = G[planted[:64]]                                    # cos = 1: the largest accumulators
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    os.environ["FR_F8_TOPK"] = "1"                              # k > 1 normally runs on the fp16 copy: keep it on the e4m3 copy here
    try:
        g.topk(q, 8)
    finally:
        del os.environ["FR_F8_TOPK"]
    cs, ci = g.debug_read(5), g.debug_read(6)
    gh = fd.round_dither(G, fd.dither_r24(fd.dither_key(fd.DEFAULT_SEED, np.arange(n, dtype=np.uint64))))[0].astype(np.float64)
    qh = fd.dither_queries(q, 1.0, 1.0, 1.0)[0].astype(np.float64)
    lists = 2 * min(74, (n + 255) // 256)
    worst, seen = 0.0, 0
    for l in range(lists):
        for qi in range(0, 256, 5):
            ok = ci[l, qi] >= 0
            if not ok.any():
                continue
            want = (gh[ci[l, qi][ok]] @ qh[qi]) / fd.SCALE ** 2
            scale = np.linalg.norm(qh[qi]) * np.linalg.norm(gh[ci[l, qi][ok]], axis=1) / fd.SCALE ** 2
            worst = max(worst, float(np.abs(cs[l, qi][ok] - want).max() / scale.min()))
            seen += int(ok.sum())
    assert seen > 10_000
    assert worst <= 0.5 * fd.ACC_EPS, worst                     # half the budget at most (fp32 accumulation: ~1e-6 expected)


@pytest.mark.parametrize("kind", ["counter_example", "sign", "int8", "grid", "near_dup"])
def test_fp8_structured_galleries_top1_exact(kind):
    # structured rows broke the round-to-nearest e4m3 scan (wrong identity, no fallback). Same galleries as the CPU
    # emulation tests, on the kernels: exact fp32 top-1, identical to the fp16 path, and no exact-scan fallback needed.
    rng = np.random.default_rng(abs(hash(kind)) % 997)
    n = 60_000
    if kind == "counter_example":
        sign = np.where(rng.random(512) < 0.5, -1.0, 1.0)
        A = sign / np.sqrt(512.0)
        B = sign * np.where(np.arange(512) % 2 == 0, 10.51, 12.06) / 256.0
        G = so.l2_normalise(rng.standard_normal((n, 512)))
        G[100], G[41_000] = A, B / np.linalg.norm(B)
        q = np.concatenate([A[None, :], G[41_000][None, :], so.l2_normalise(rng.standard_normal((30, 512)))]).astype(np.float32)
    elif kind == "sign":
        G = so.l2_normalise(np.where(rng.random((n, 512)) < 0.5, -1.0, 1.0))
        q = np.concatenate([G[rng.integers(0, n, 100)], so.l2_normalise(np.where(rng.random((100, 512)) < 0.5, -1.0, 1.0))]).astype(np.float32)
    elif kind == "int8":
        G = so.l2_normalise(np.clip(np.round(rng.standard_normal((n, 512)) * 24), -127, 127) + 1e-9)
        q = np.concatenate([so.planted_queries(G[rng.integers(0, n, 128)], noise=0.75, seed=1), so.l2_normalise(rng.standard_normal((128, 512)))]).astype(np.float32)
    elif kind == "grid":
        lv = np.array([8.5, 9.5, 10.5, 11.5, 12.5, 13.5]) / 256.0
        G = so.l2_normalise(rng.choice(lv, (n, 512)) * np.where(rng.random((n, 512)) < 0.5, -1.0, 1.0))
        q = np.concatenate([G[rng.integers(0, n, 64)], so.l2_normalise(rng.standard_normal((64, 512)))]).astype(np.float32)
    else:
        G = so.l2_normalise(rng.standard_normal((n, 512)))
        base = G[777].astype(np.float64)
        for j, w in enumerate(np.sort(rng.choice(np.arange(1000, n - 10), 50, replace=False))):
            c = 0.995 + 0.004 * j / 49
            v = rng.standard_normal(512)
            v -= v.dot(base) * base
            G[w] = (c * base + np.sqrt(1 - c * c) * v / np.linalg.norm(v)).astype(np.float32)
        G[n - 3] = G[777]
        q = np.concatenate([G[777][None, :], so.l2_normalise(base[None, :] + 0.02 * rng.standard_normal((63, 512)))]).astype(np.float32)
    G = np.ascontiguousarray(G, np.float32)
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    s, i = g.topk(q, 1)
    flagged = g.last_flagged()
    sim = so.sims(G, q)
    # binarised / grid rows tie EXACTLY in real arithmetic (scores are multiples of 2/512), so the winner among tied rows depends on
    # the fp32 summation order: _check_topk accepts a different row only if its score equals the oracle's to fp32 rounding; the
    # bit-for-bit comparison with the fp16 path below (same exact re-score arithmetic) is the strict check
    _check_topk(s, i, sim, 1)
    # binarised rows are where Hoeffding's bound is nearly tight (two-point errors) AND where dozens of rows tie for the best score,
    # so the best coarse score overshoots the best exact one more often: a few queries fail the certificate and are recomputed exactly
    assert flagged <= max(1, q.shape[0] // 40), (kind, flagged)
    g.set_scan(frb200.FR_SCAN_F16)
    s2, i2 = g.topk(q, 1)
    assert np.array_equal(i, i2) and np.array_equal(s.view(np.uint32), s2.view(np.uint32))
    if kind == "counter_example":
        assert i[0, 0] == 100 and i[1, 0] == 41_000
    g.close()


def test_fp8_saturating_query_goes_to_the_exact_scan():
    rng = np.random.default_rng(2)
    G = so.l2_normalise(rng.standard_normal((20_000, 512)))
    q = so.l2_normalise(rng.standard_normal((3, 512))).astype(np.float32)
    q[1] *= 50.0                                                 # components beyond 1.75: e4m3 saturates, the error model does not hold
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    s, i = g.topk(q, 1)
    oi, ov = so.get_outputs(so.sims(G, q))
    assert np.array_equal(i[:, 0], oi) and np.allclose(s[:, 0], ov, rtol=1e-5)
    assert g.last_flagged() == 1
    g.close()


def test_fp8_scan_everything_inside_margin_falls_back_exactly():
    # all rows identical up to 1e-4 noise: every row is inside the margin of the best, the append lists overflow, and the flagged
    # exact scan must still return the fp32 answer (lowest row among exact ties)
    rng = np.random.default_rng(5)
    n = 60_000
    base = so.l2_normalise(rng.standard_normal((1, 512)))
    G = so.l2_normalise(base + 1e-4 * rng.standard_normal((n, 512)).astype(np.float32) / np.sqrt(512))
    G[100] = G[50]
    q = np.concatenate([G[50][None, :], so.l2_normalise(rng.standard_normal((2, 512)))]).astype(np.float32)
    g = frb200.Gallery.from_rows(G)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    s, i = g.topk(q, 1)
    _check_topk(s, i, so.sims(G, q), 1)
    assert g.last_flagged() >= 1
    g.close()


@pytest.mark.parametrize("scan", [frb200.FR_SCAN_F16, frb200.FR_SCAN_F8])
def test_gallery_lifecycle_append_remove_clear(scan):
    # enrolment / deletion / reload without a full re-upload. After every step the resident gallery must answer
    # exactly like a gallery created from scratch with the same rows (scan copies and margin bounds maintained incrementally).
    rng = np.random.default_rng(77)
    G = so.l2_normalise(rng.standard_normal((9000, 512)))
    model = G[:5000].copy()                     # host model of the resident rows
    g = frb200.Gallery.from_rows(model)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(scan)

    def check(nq=37, k=3):
        planted = rng.integers(0, len(model), nq)
        q = so.planted_queries(model[planted], noise=0.75, seed=int(rng.integers(1 << 30)))
        assert g.rows == len(model) and g.capacity >= g.rows
        s1, i1 = g.topk(q, 1)
        sk, ik = g.topk(q, k)
        sim = so.sims(model, q)
        _check_topk(s1, i1, sim, 1)
        _check_topk(sk, ik, sim, k)
        assert np.array_equal(i1[:, 0], planted)
        assert np.array_equal(g.read_rows(0, len(model)).view(np.uint32), model.view(np.uint32))

    check()
    g.append(G[5000:5001])                      # one enrolment (grows the buffers)
    model = np.concatenate([model, G[5000:5001]])
    check()
    g.append(G[5001:8000])                      # bulk enrolment
    model = np.concatenate([model, G[5001:8000]])
    check()
    g.reserve(20_000)
    assert g.capacity >= 20_000
    check()
    for pick in (0, 4321, -1, 17):              # deletions (-1 = the last row itself): the last row moves into the slot
        row = pick if pick >= 0 else len(model) - 1
        moved = g.remove(row)
        assert moved == len(model) - 1
        model[row] = model[moved]
        model = model[:-1]
        check()
    fresh = so.l2_normalise(rng.standard_normal((300, 512)))   # re-enrolment in place: rows 1000..1299 replaced
    g.update(1000, fresh)
    model[1000:1300] = fresh
    check()
    with pytest.raises(frb200.FrError):
        g.update(len(model) - 5, fresh[:10])                   # range beyond the last row
    g.clear()
    assert g.rows == 0
    with pytest.raises(frb200.FrError) as e:
        g.topk(G[:2], 1)
    assert e.value.code == frb200.FR_ESTATE
    g.append(G[8000:9000])                      # reload after resetEmbeddings
    model = G[8000:9000].copy()
    check(k=2)
    if scan == frb200.FR_SCAN_F8:               # rows that are not L2-normalised cannot enter the e4m3 copy
        with pytest.raises(frb200.FrError):
            g.append(2.0 * G[:3])
        assert g.rows == 1000
    g.close()


def test_gallery_starts_empty_then_appends():
    rng = np.random.default_rng(3)
    G = so.l2_normalise(rng.standard_normal((3000, 512)))
    g = frb200.Gallery.from_rows(np.zeros((0, 512), np.float32))
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(frb200.FR_SCAN_F8)
    g.append(G)
    q = so.planted_queries(G[[5, 2999, 1234]], noise=0.5, seed=1)
    s, i = g.topk(q, 1)
    _check_topk(s, i, so.sims(G, q), 1)
    assert i[:, 0].tolist() == [5, 2999, 1234]
    g.close()


def test_roster_lifecycle_on_sharded_gallery(tmp_path):
    # SURVEY 8 f-2 on the device: FACE rows from an SQLite file with the reference's schema -> 3 shards (row_offset = shard << 32),
    # searched shard by shard and merged; after every enrolment / deletion / reload the identities the sharded search returns equal
    # those of a host model of the same rows, and the roster resolves the returned global ids to the right userId.
    import sqlite3

    rng = np.random.default_rng(12)
    n, world = 3000, 3
    emb = so.l2_normalise(rng.standard_normal((n, 512))).astype("<f4")
    users = [f"u{int(u):04d}" for u in rng.integers(0, n // 2, n)]
    con = sqlite3.connect(tmp_path / "face.db")
    con.execute("CREATE TABLE FACE (IMG_ID INTEGER PRIMARY KEY AUTOINCREMENT, USR_ID TEXT, IMG_PATH TEXT, EMBEDDING BLOB, UNIQUE(IMG_ID, USR_ID))")
    con.executemany("INSERT INTO FACE (USR_ID, IMG_PATH, EMBEDDING) VALUES (?, ?, ?)", [(u, "x.jpg", e.tobytes()) for u, e in zip(users, emb)])
    con.commit()
    rows = con.execute("SELECT * FROM FACE").fetchall()
    con.close()
    shards = [frb200.Gallery.from_rows(np.zeros((0, 512), np.float32), row_offset=g << 32) for g in range(world)]
    for sh in shards:
        sh.set_path(frb200.FR_PATH_EXACT)            # shards of ~1000 rows: the exact path (AUTO would choose it as well)
    rosters = [frb200.Roster(shards[g], world, g) for g in range(world)]
    model = {}                                        # global id -> (user, vector): host model of what is resident

    def reload_model():
        model.clear()
        per = (n + world - 1) // world
        for i in range(n):
            model[((i // per) << 32) | (i % per)] = (users[i], emb[i])

    def check(nq=40):
        ids = sorted(model)
        M = np.stack([model[i][1] for i in ids])
        pick = rng.integers(0, len(ids), nq)
        q = so.planted_queries(M[pick], 0.5, int(rng.integers(1 << 30)))
        parts_s, parts_i = [], []
        for sh in shards:
            if sh.rows:
                s_, i_ = sh.topk(q, 1)
            else:
                s_, i_ = np.full((nq, 1), -np.inf, np.float32), np.full((nq, 1), -1, np.int64)
            parts_s.append(s_)
            parts_i.append(i_)
        ms, mi = so.merge_topk(parts_s, parts_i, 1)
        assert [int(v) for v in mi[:, 0]] == [ids[p] for p in pick]
        for ro in rosters:
            assert [ro.user(int(v)) for v in mi[:, 0]] == [model[ids[p]][0] for p in pick]
        assert sum(sh.rows for sh in shards) == len(model) == rosters[0].rows

    for ro in rosters:
        ro.load([r[1] for r in rows], [r[3] for r in rows])
    reload_model()
    check()
    for j in range(5):                                # enrolment
        v = so.l2_normalise(rng.standard_normal((1, 512)))[0]
        got = {ro.add(f"new{j}", v) for ro in rosters}
        assert len(got) == 1
        model[got.pop()] = (f"new{j}", v)
    check()
    for _ in range(6):                                # deletion of single rows: the shard's last row moves into the slot
        gid = sorted(model)[int(rng.integers(0, len(model)))]
        shard = gid >> 32
        last = max(i for i in model if i >> 32 == shard)
        for ro in rosters:
            ro.remove(gid)
        model[gid] = model[last]
        del model[last]
    check()
    victim = users[17]                                # deletion of a user
    cnt = {ro.remove_user(victim) for ro in rosters}
    assert cnt == {sum(1 for u, _ in model.values() if u == victim)}
    ids_after = {}
    for g in range(world):                            # rebuild the model from the rosters' tables + the shards' rows
        R = shards[g].read_rows(0, shards[g].rows) if shards[g].rows else np.zeros((0, 512), np.float32)
        for l in range(rosters[0].shard_rows(g)):
            ids_after[(g << 32) | l] = (rosters[0].user((g << 32) | l), R[l])
    assert victim not in {u for u, _ in ids_after.values()}
    model.clear()
    model.update(ids_after)
    check()
    for ro in rosters:                                # /reload: reset + read everything again
        ro.load([r[1] for r in rows], [r[3] for r in rows])
    reload_model()
    check()
    for o in rosters + shards:
        o.close()


@pytest.mark.parametrize("scan", [frb200.FR_SCAN_F16, frb200.FR_SCAN_F8])
def test_search_stream_batches_in_flight(scan):
    # fr_search_stream_*: the asynchronous host-buffer search (bench.py's e2e). Results come back in submission order, equal to the
    # synchronous call, with up to three batches in flight; a fourth submit / an empty collect are refused.
    rng = np.random.default_rng(44)
    n = 80_000
    G = so.l2_normalise(rng.standard_normal((n, 512)))
    g = frb200.Gallery.from_rows(G, row_offset=5)
    g.set_path(frb200.FR_PATH_TENSOR)
    g.set_scan(scan)
    ss = frb200.SearchStream(g, None, 1)
    batches = [so.planted_queries(G[rng.integers(0, n, nq)], 0.7, b) for b, nq in enumerate((256, 100, 256, 1, 37))]
    want = [g.topk(q, 1) for q in batches]
    s_out, i_out = np.empty((256, 1), np.float32), np.empty((256, 1), np.int64)
    with pytest.raises(frb200.FrError):
        ss.collect(s_out, i_out)
    ss.submit(batches[0])
    ss.submit(batches[1])
    ss.submit(batches[2])
    with pytest.raises(frb200.FrError) as e:
        ss.submit(batches[3])
    assert e.value.code == frb200.FR_ESTATE
    for b in range(len(batches)):
        nq = ss.collect(s_out, i_out)
        assert nq == batches[b].shape[0]
        assert np.array_equal(i_out[:nq], want[b][1]) and np.array_equal(s_out[:nq].view(np.uint32), want[b][0].view(np.uint32))
        if b + 3 < len(batches):
            batches[b + 3][:] = batches[b + 3]           # the caller's buffer is free again as soon as submit returns
            ss.submit(batches[b + 3])
    ss.close()
    g.close()


def test_rows_and_queries_outside_fp16_range_take_the_exact_scan():
    # the reference's MatMul accepts arbitrary fp32 rows; the fp16 scan copy's bound only holds in fp16's normal range.
    # Huge components (inf in fp16) or a gallery of tiny rows (fp16 subnormals) must still give the exact fp32 answer.
    rng = np.random.default_rng(6)
    n = 30_000
    G = rng.standard_normal((n, 512)).astype(np.float32)
    q = rng.standard_normal((16, 512)).astype(np.float32)
    for scale in (1e5, 1e-6):
        g = frb200.Gallery.from_rows(G * np.float32(scale))
        g.set_path(frb200.FR_PATH_TENSOR)
        s, i = g.topk(q, 3)
        sim = so.sims(G * np.float32(scale), q)
        o_s, o_i = so.topk(sim, 3)
        assert np.array_equal(i, o_i) and np.allclose(s, o_s, rtol=2e-5)
        assert g.last_stats().ctas == 0                      # the exact path ran (no fused-scan CTAs)
        g.close()
    # an ordinary gallery, extraordinary queries: recomputed by the exact scan one by one
    Gn = so.l2_normalise(G)
    g = frb200.Gallery.from_rows(Gn)
    g.set_path(frb200.FR_PATH_TENSOR)
    qq = so.l2_normalise(q).astype(np.float32)
    qq[3] *= np.float32(1e-7)
    qq[5] *= np.float32(3e6)
    s, i = g.topk(qq, 1)
    o_i, o_s = so.get_outputs(so.sims(Gn, qq))
    assert np.array_equal(i[:, 0], o_i) and np.allclose(s[:, 0], o_s, rtol=2e-5)
    assert g.last_flagged() == 2
    g.close()
