"""CPU: the certified e4m3 scan (csrc/search_kernels.cuh: f8_round_dither, kF8LogP, the certificate in append_rerank_kernel), run
through its numpy emulation oracle/f8_dither.py (constants parsed from the .cuh; the GPU suite checks that the emulated e4m3 image is
bit-identical to the kernel's). The contract is the reference's: top-1 = FIRST maximum of the exact fp32 scores
(/root/reference src/arcface.cpp:203-217) — for arbitrary rows, not only isotropic Gaussians. Round-to-nearest e4m3 breaks it on
structured rows (VERDICT r1 counter-example, reproduced below); stochastic rounding + the per-query certificate must not."""
import numpy as np
import pytest

from oracle import f8_dither as fd


def unit(x):
    x = np.asarray(x, np.float64)
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def counter_example_rows():
    """row A = +-1/sqrt(512) (x256 = 11.31: rounds DOWN to 11 under round-to-nearest); row B = same signs, magnitudes alternating
    10.51/256 and 12.06/256, unit norm (x256: round UP to 11 and 12... i.e. to 11/12 from below the midpoints 10.5 / 12)."""
    rng = np.random.default_rng(42)
    sign = np.where(rng.random(512) < 0.5, -1.0, 1.0)
    A = sign / np.sqrt(512.0)
    mag = np.where(np.arange(512) % 2 == 0, 10.51, 12.06) / 256.0
    B = sign * mag
    B /= np.linalg.norm(B)
    return A.astype(np.float32), B.astype(np.float32)


def test_round_nearest_returns_the_wrong_identity_on_the_counter_example():
    A, B = counter_example_rows()
    rng = np.random.default_rng(1)
    G = unit(rng.standard_normal((2000, 512)))
    G[100], G[1500] = A, B
    q = A[None, :]
    exact = (q @ G.T)[0]
    assert exact.argmax() == 100 and 0.99 < exact[1500] < 0.999
    # the old scheme: round to nearest, keep rows within the old statistical margin 6.5 * sqrt(2) * 0.0373 * |q|_4 |g|_4max of the best coarse score
    qn, gn = fd.round_nearest(q), fd.round_nearest(G)
    coarse = (qn.astype(np.float64) @ gn.astype(np.float64).T)[0] / fd.SCALE ** 2
    old_margin = 6.5 * 2 ** 0.5 * 0.0373 * float(((q.astype(np.float64) ** 4).sum() * (G.astype(np.float64) ** 4).sum(1).max()) ** 0.25)
    assert coarse.argmax() == 1500 and coarse[100] < coarse[1500] - old_margin      # the true best was pruned, silently
    # the certified scan on the same data, many dither seeds: always the exact answer
    for seed in range(20):
        r = fd.certified_top1(G, q, seed=seed)
        assert r["best_in_cand"].all() and not r["flagged"].any() and r["idx"][0] == 100


def structured_gallery(kind, n, rng):
    if kind == "sign":           # binarised embeddings: every component +-1/sqrt(512)
        return unit(np.where(rng.random((n, 512)) < 0.5, -1.0, 1.0))
    if kind == "int8":           # int8-quantised-then-normalised rows
        return unit(np.clip(np.round(rng.standard_normal((n, 512)) * 24), -127, 127) + 1e-9)
    if kind == "sparse":         # a few large components (heavy |.|_4)
        x = rng.standard_normal((n, 512)) * (rng.random((n, 512)) < 0.05)
        x[:, 0] += 1e-3
        return unit(x)
    if kind == "student3":
        return unit(rng.standard_t(3.0, (n, 512)))
    if kind == "grid":           # components on the e4m3 midpoints: the worst case for any deterministic rounding
        lv = np.array([8.5, 9.5, 10.5, 11.5, 12.5, 13.5]) / 256.0
        return unit(rng.choice(lv, (n, 512)) * np.where(rng.random((n, 512)) < 0.5, -1.0, 1.0))
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["sign", "int8", "sparse", "student3", "grid"])
def test_structured_galleries_top1_exact(kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    n = 30_000
    G = structured_gallery(kind, n, rng)
    planted = rng.integers(0, n, 96)
    qp = unit(G[planted] + 0.75 * unit(rng.standard_normal((96, 512))))        # matched queries, cos ~ 0.8
    qu = structured_gallery(kind, 96, rng)                                        # unmatched queries of the same structure
    qe = G[planted[:32]]                                                          # exact enrolment vectors
    q = np.concatenate([qp, qu, qe])
    for seed in (0, 1, 2):
        r = fd.certified_top1(G, q, seed=seed)
        ok = r["flagged"] | (r["idx"] == r["exact_idx"])
        assert ok.all(), (kind, seed, np.nonzero(~ok)[0][:5])
        assert r["best_in_cand"].all()
        # heavy-tailed rows (large |g|_4) widen the budget E to ~0.2: unmatched queries then keep > 4096 rows and are handed to the
        # exact scan — slow, still exact. Bounded-component galleries must not need the fallback.
        if kind in ("sign", "int8", "grid"):
            assert r["flagged"].mean() <= 0.02, (kind, float(r["flagged"].mean()))
        err = r["coarse"] - r["exact"]
        # no pair exceeds the certified budget; binarised rows come closest (Hoeffding is nearly tight for two-point errors:
        # the largest of ~7 M errors is ~5 sigma = 0.6 E, as the bound predicts)
        assert float(np.abs(err / r["E"][:, None]).max()) < 0.8, kind


def test_near_duplicate_enrolments():
    # 50 near-duplicates of one identity at cos 0.995-0.999 plus an exact duplicate at a HIGHER row: first maximum must win
    rng = np.random.default_rng(9)
    n = 40_000
    G = unit(rng.standard_normal((n, 512)))
    base = G[777].astype(np.float64)
    where = np.sort(rng.choice(np.arange(1000, n), 50, replace=False))
    for j, w in enumerate(where):
        c = 0.995 + 0.004 * j / 49
        noise = unit(rng.standard_normal((1, 512)))[0]
        noise = noise - noise.dot(base) * base
        noise /= np.linalg.norm(noise)
        G[w] = (c * base + np.sqrt(1 - c * c) * noise).astype(np.float32)
    G[n - 3] = G[777]
    q = np.concatenate([G[777][None, :], unit(base[None, :] + 0.02 * rng.standard_normal((31, 512)))])
    for seed in range(8):
        r = fd.certified_top1(G, q, seed=seed)
        assert r["best_in_cand"].all()
        ok = r["flagged"] | (r["idx"] == r["exact_idx"])
        assert ok.all() and r["flagged"].sum() <= 1
        assert r["idx"][0] in (777, -1) and r["exact_idx"][0] == 777
        assert (r["n_cand"] >= 40).all()          # the whole cluster sits inside the margin and is re-scored exactly


def test_dither_is_unbiased_bounded_and_exact_on_representable_values():
    # 11.31 lies between the e4m3 values 11 and 12: the rounded value is one of them, and its mean over the dither is 11.31
    x = np.full((4000, 512), 11.31 / 256.0, np.float32)
    keys = fd.dither_key(123, np.arange(4000, dtype=np.uint64))
    r, u, abar, sat = fd.round_dither(x, fd.dither_r24(keys))
    assert set(np.unique(r).tolist()) == {11.0, 12.0} and np.all(u == 1.0) and np.all(abar == 12.0) and not sat.any()
    assert abs(float(r.mean()) - 11.31) < 3 * np.sqrt(0.31 * 0.69 / r.size) + 1e-6
    # representable values are left alone (no randomness, zero step)
    y = (np.array([0.0, 1.0, 1.125, 13.0, 448.0, 2.0 ** -9, 3 * 2.0 ** -9, -14.0], np.float32) / 256.0)[None, :].repeat(8, 0)
    r, u, _, _ = fd.round_dither(y, fd.dither_r24(fd.dither_key(5, np.arange(8, dtype=np.uint64)), 8))
    assert np.array_equal(r, y * 256.0) and np.all(u == 0)
    # steps: [8,16) -> 1, [16,32) -> 2, subnormal range -> 2^-9; the rounded value never leaves the bracket
    z = (np.array([8.3, 15.9, 16.1, 31.0, 0.001, 0.017, 255.9], np.float32) / 256.0)[None, :].repeat(64, 0)
    r, u, _, _ = fd.round_dither(z, fd.dither_r24(fd.dither_key(7, np.arange(64, dtype=np.uint64)), 7))
    assert np.array_equal(u[0], np.array([1, 1, 2, 2, 2.0 ** -9, 2.0 ** -9, 16], np.float32))
    assert np.all(np.abs(r - z * 256.0) < u)
    # matches torch's e4m3 grid: every rounded value is exactly representable
    import torch

    t = torch.from_numpy(r.copy())
    assert torch.equal(t.to(torch.float8_e4m3fn).float(), t)


def test_margin_and_candidate_counts_for_isotropic_embeddings():
    rng = np.random.default_rng(3)
    G = unit(rng.standard_normal((200_000, 512)))
    q = unit(rng.standard_normal((128, 512)))            # unmatched: the hard case (many rows inside the margin of the best impostor)
    r = fd.certified_top1(G, q)
    assert 0.035 < float(r["E"].mean()) < 0.055 and np.allclose(r["margin"], (1 + fd.GAP_FRAC) * r["E"])
    assert r["best_in_cand"].all() and (r["idx"] == r["exact_idx"]).all() and not r["flagged"].any()
    assert float(r["n_cand"].mean()) < 400 and int(r["n_cand"].max()) <= 4096
    # measured error vs the budget: sigma ~ 2.3e-3, the Hoeffding budget E is ~19 sigma of it (the bound is conservative by design)
    err = r["coarse"] - r["exact"]
    assert 1.5e-3 < float(err.std()) < 3.5e-3 and abs(float(err.mean())) < 2e-5
    assert float(np.abs(err).max()) < 0.4 * float(r["E"].min())


def test_saturating_query_is_always_recomputed():
    rng = np.random.default_rng(4)
    G = unit(rng.standard_normal((5000, 512)))
    q = unit(rng.standard_normal((4, 512)))
    q[2] *= 40.0                                           # components beyond 1.75 saturate e4m3
    r = fd.certified_top1(G, q)
    assert r["flagged"].tolist() == [False, False, True, False]


def test_exact_leader_filter_cuts_the_rerank_set_and_keeps_the_best():
    """append_rerank_kernel's second filter: rows with coarse < L0 - E are dropped, L0 = best exact score among a few coarse
    leaders. Sound under the certificate's own event (the true best A has exact_A >= L0, so dropping it needs exact_A - coarse_A > E);
    here: it never drops the true best, never changes the answer, and does shrink the set the fp16 / fp32 re-score has to gather."""
    rng = np.random.default_rng(12)
    G = unit(rng.standard_normal((120_000, 512)))
    q = unit(rng.standard_normal((96, 512)))             # unmatched queries: hundreds of rows inside the margin
    q[:8] = unit(0.8 * G[rng.integers(0, len(G), 8)] + 0.6 * q[:8])   # and a few with a match
    r = fd.certified_top1(G, q)
    assert r["best_in_cand"].all() and (r["idx"] == r["exact_idx"]).all() and not r["flagged"].any()
    assert (r["n_cand"] <= r["n_margin"]).all()
    unmatched = slice(8, None)
    assert float(r["n_cand"][unmatched].mean()) < 0.6 * float(r["n_margin"][unmatched].mean())
    # the event everything rests on, in numbers: the true best row's coarse score is within E of its exact score (by a wide factor)
    rows = np.arange(len(q))
    err_best = r["exact"][rows, r["exact_idx"]] - r["coarse"][rows, r["exact_idx"]]
    assert (err_best <= 0.5 * r["E"]).all()
