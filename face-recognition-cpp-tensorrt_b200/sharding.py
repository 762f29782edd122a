"""Host-side plan of the gallery-sharded search: contiguous row blocks per rank and the one exchange step.
Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests); the merge itself is the CUDA kernel behind fr_topk_merge_dev."""
from __future__ import annotations


def shard_bounds(n_rows: int, world: int, rank: int) -> "tuple[int, int]":
    """rank g holds global rows [g * ceil(N/G), min(N, (g+1) * ceil(N/G))): contiguous blocks keep 'lowest global row wins ties'"""
    per = (n_rows + world - 1) // world
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def all_gather_topk(dist, local_scores, local_idx, all_scores, all_idx) -> None:
    """the exchange: every rank contributes its nq x k (score f32, global idx i64) and receives all of them, rank-major.
    all_scores / all_idx: preallocated [world, nq, k] tensors on the same device as the local ones."""
    k = local_scores.shape[-1]
    dist.all_gather_into_tensor(all_scores.view(-1, k), local_scores.view(-1, k))  # concatenation form: valid for NCCL and gloo
    dist.all_gather_into_tensor(all_idx.view(-1, k), local_idx.view(-1, k))
