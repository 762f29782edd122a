"""CPU, world_size 2 (gloo): the N>1 host logic of the sharded search — shard bounds, the all-gather exchange and its rank-major
layout, merged result == single-gallery result (incl. duplicate rows across shards: lowest global row wins)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nq, k, out_dir):
    import sharding
    from oracle import search_oracle as so

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = so.synth_rows(np.arange(n), seed=5)
        rows[n - 3] = rows[10]                       # duplicate across shards
        q = so.planted_queries(rows[[10, n - 1, n // 2, 7]], 0.5, 2)[:nq]
        lo, hi = sharding.shard_bounds(n, world, rank)
        s, i = so.topk(so.sims(rows[lo:hi], q), k, row_offset=lo)
        ls, li = torch.from_numpy(s).contiguous(), torch.from_numpy(i).contiguous()
        all_s = torch.empty((world, nq, k), dtype=torch.float32)
        all_i = torch.empty((world, nq, k), dtype=torch.int64)
        sharding.all_gather_topk(dist, ls, li, all_s, all_i)
        assert torch.equal(all_s[rank], ls) and torch.equal(all_i[rank], li)      # rank-major layout
        ms, mi = so.merge_topk([all_s[g].numpy() for g in range(world)], [all_i[g].numpy() for g in range(world)], k)
        fs, fi = so.topk(so.sims(rows, q), k)
        assert np.array_equal(mi, fi) and np.allclose(ms, fs, atol=1e-6)
        assert mi[0, 0] == 10 and mi[0, 1] == n - 3
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_and_partition():
    import sharding

    for n in (0, 1, 7, 1000, 10_000_000):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            assert all(lo <= hi for lo, hi in b)
    assert sharding.shard_bounds(10_000_000, 8, 7) == (8_750_000, 10_000_000)


def test_two_rank_exchange_and_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), 999, 4, 3, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
