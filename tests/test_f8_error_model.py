"""CPU: the statistical margin of the e4m3 scan copy (csrc/search_kernels.cuh: kF8Delta, kF8Z, kF8Scale) against a torch
float8_e4m3fn emulation of what the kernel computes. The model sigma(q, g) = kF8Delta * sqrt(sum q_i^2 g_i^2) must describe the
rounding error (err / model has unit standard deviation), the Cauchy-Schwarz bound kF8Delta * |q|_4 * |g|_4 must dominate it, and no
error in a few million sampled pairs may come near the margin the kernel uses."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

SRC = (Path(__file__).resolve().parent.parent / "face-recognition-cpp-tensorrt_b200" / "csrc" / "search_kernels.cuh").read_text()


def _const(name):
    m = re.search(rf"constexpr float {name} = ([0-9.eE+-]+)f;", SRC)
    assert m, name
    return float(m.group(1))


K_SCALE, K_DELTA, K_Z = _const("kF8Scale"), _const("kF8Delta"), _const("kF8Z")


def f8(x):
    return (x * K_SCALE).to(torch.float8_e4m3fn).float() / K_SCALE


def unit(x):
    return x / x.norm(dim=1, keepdim=True)


@pytest.mark.parametrize("dist", ["gauss", "laplace", "student3"])
def test_error_model_and_bound(dist):
    torch.manual_seed(11)
    gen = {"gauss": lambda n: torch.randn(n, 512),
           "laplace": lambda n: torch.distributions.Laplace(0.0, 1.0).sample((n, 512)),
           "student3": lambda n: torch.distributions.StudentT(3.0).sample((n, 512))}[dist]
    g, q = unit(gen(20_000)), unit(gen(64))
    err = f8(q) @ f8(g).T - q @ g.T
    q4, g4 = (q ** 4).sum(1) ** 0.25, (g ** 4).sum(1) ** 0.25
    model = K_DELTA * torch.sqrt((q ** 2) @ (g ** 2).T)
    bound = K_DELTA * q4[:, None] * g4[None, :]
    assert abs(float((err / model).std()) - 1.0) < 0.05          # the model describes the error
    assert float((err / bound).std()) < 1.0                        # the bound dominates it
    # matched pairs (the true best of a query): equality case of the bound
    p = unit(g[:4096] + 0.75 * unit(gen(4096)))
    e2 = (f8(p) * f8(g[:4096])).sum(1) - (p * g[:4096]).sum(1)
    b2 = K_DELTA * ((p ** 4).sum(1) ** 0.25) * g4[:4096]
    assert float((e2 / b2).std()) < 1.0
    # the kernel's margin is kF8Z * sqrt(2) bound-sigmas for the DIFFERENCE of two errors; single errors must stay far inside it
    worst = max(float((err / bound).abs().max()), float((e2 / b2).abs().max()))
    assert worst < 0.7 * K_Z, worst
    assert K_Z >= 6.0 and np.isclose(K_SCALE, 256.0)


def test_true_best_survives_the_margin_for_unmatched_queries():
    # the statement the e4m3 scan relies on, emulated on the CPU for the hard case (queries with NO match, so hundreds of rows sit
    # within the margin of the best impostor): the exact fp32 top-1 is always among the rows whose coarse (e4m3) score is within
    # margin = kF8Z * sqrt(2) * kF8Delta * |q|_4 * max_rows |g|_4 of the best coarse score, and the set re-scored stays small
    torch.manual_seed(5)
    n, nq = 300_000, 256
    g, q = unit(torch.randn(n, 512)), unit(torch.randn(nq, 512))
    exact = q @ g.T
    coarse = f8(q) @ f8(g).T
    g4max = float(((g ** 4).sum(1)).max())
    margin = K_Z * 2 ** 0.5 * K_DELTA * ((q ** 4).sum(1) * g4max) ** 0.25            # per query, cosine units
    best_exact = exact.argmax(1)
    cbest = coarse.max(1).values
    survives = coarse[torch.arange(nq), best_exact] >= cbest - margin
    assert bool(survives.all())
    slack = (coarse[torch.arange(nq), best_exact] - (cbest - margin)) / margin      # 1 = at the coarse best, 0 = at the edge
    assert float(slack.min()) > 0.5, float(slack.min())                             # never closer than half the margin to being dropped
    in_margin = (coarse >= (cbest - margin)[:, None]).sum(1)
    assert int(in_margin.max()) <= 1024 and float(in_margin.float().mean()) < 200   # kAppRescoreMax bounds the re-score
