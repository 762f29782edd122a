"""Error model of the e4m3 scan copy (FR_SCAN_F8) behind the margin in csrc/search_kernels.cuh (kF8Delta, kF8Z).

    python tools/f8_error_model.py

Both operands are multiplied by kF8Scale = 256 and rounded to e4m3 (round-to-nearest, saturating); products accumulate in fp32 —
what cosine_topk_coarse<.., F8=true> computes. Model: coarse - exact is a sum of 512 independent rounding errors with
    sigma(q, g) = kF8Delta * sqrt(sum_i q_i^2 g_i^2)   <=   kF8Delta * |q|_4 * |g|_4        (Cauchy-Schwarz on the squares)
The scan keeps every row whose coarse score is within  margin = kF8Z * sqrt(2) * kF8Delta * |q|_4 * max_rows |g|_4  of the best
coarse score (two errors are involved: the true best's and the coarse best's). This script checks, for unit vectors drawn from
several distributions, that err / model has standard deviation 1 (the model is right) and err / bound stays below 1 (the bound
holds), and prints the resulting margins. Measured with torch 2.11 float8_e4m3fn:
    gauss     err/model 1.000  err/bound unrelated 0.58, matched (cos 0.8) 0.88   margin 0.029
    laplace   err/model 0.999  err/bound 0.42 / 0.90                               margin 0.053
    student-t(3)  err/model 1.01  err/bound 0.28 / 0.89                            margin 0.136
"""
import torch

K_F8_SCALE, K_F8_DELTA, K_F8_Z = 256.0, 0.0373, 6.5


def f8(x):
    return (x * K_F8_SCALE).to(torch.float8_e4m3fn).float() / K_F8_SCALE


def unit(x):
    return x / x.norm(dim=1, keepdim=True)


def main():
    torch.manual_seed(1)
    gens = {
        "gauss": lambda n: torch.randn(n, 512),
        "laplace": lambda n: torch.distributions.Laplace(0.0, 1.0).sample((n, 512)),
        "student-t(3)": lambda n: torch.distributions.StudentT(3.0).sample((n, 512)),
    }
    for name, gen in gens.items():
        g, q = unit(gen(50_000)), unit(gen(64))
        err = f8(q) @ f8(g).T - q @ g.T
        q4, g4 = (q ** 4).sum(1) ** 0.25, (g ** 4).sum(1) ** 0.25
        model = K_F8_DELTA * torch.sqrt((q ** 2) @ (g ** 2).T)
        bound = K_F8_DELTA * q4[:, None] * g4[None, :]
        p = unit(g[:4096] + 0.75 * unit(gen(4096)))
        e2 = (f8(p) * f8(g[:4096])).sum(1) - (p * g[:4096]).sum(1)
        b2 = K_F8_DELTA * ((p ** 4).sum(1) ** 0.25) * g4[:4096]
        margin = K_F8_Z * 2 ** 0.5 * K_F8_DELTA * float(q4.median()) * float(g4.max())
        print(f"{name:13s} sigma(err) {float(err.std()):.3e}  err/model std {float((err / model).std()):.3f}  err/bound std "
              f"{float((err / bound).std()):.3f} (unrelated) {float((e2 / b2).std()):.3f} (matched, cos {float((p * g[:4096]).sum(1).mean()):.2f})"
              f"  max |err/bound| {max(float((err / bound).abs().max()), float((e2 / b2).abs().max())):.2f}  margin {margin:.4f}")


if __name__ == "__main__":
    main()
