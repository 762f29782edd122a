"""CPU: gallery lifecycle around the search path — the roster (row -> userId table of the reference, classNames,
/root/reference src/arcface.h:38-40) fed from an SQLite FACE table with the reference's own schema (src/db.cpp:58-65) and kept in
step with a row-sharded gallery. Host logic only here (no device: local_shard = None); tests/test_search_gpu.py drives real shards.
Includes the world_size-2 gloo case: two ranks applying the same operations end in the same replicated table."""
import os
import socket
import sqlite3
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))

SCHEMA = """CREATE TABLE IF NOT EXISTS FACE (IMG_ID INTEGER PRIMARY KEY AUTOINCREMENT, USR_ID TEXT, IMG_PATH TEXT, EMBEDDING BLOB,
            UNIQUE(IMG_ID, USR_ID))"""  # src/db.cpp:58-65 (the USER table and its foreign key do not matter for the loader)


def make_db(path, n, seed=0):
    rng = np.random.default_rng(seed)
    emb = rng.standard_normal((n, 512)).astype("<f4")
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    users = [f"user{int(u):03d}" for u in rng.integers(0, max(2, n // 3), n)]
    con = sqlite3.connect(path)
    con.execute(SCHEMA)
    con.executemany("INSERT INTO FACE (USR_ID, IMG_PATH, EMBEDDING) VALUES (?, ?, ?)",
                    [(u, f"imgs/{i}.jpg", e.tobytes()) for i, (u, e) in enumerate(zip(users, emb))])
    con.commit()
    con.close()
    return users, emb


def read_faces(path):
    """Database::getEmbeddings' query (src/db.cpp:329): SELECT * FROM FACE; column 1 = USR_ID, column 3 = EMBEDDING"""
    con = sqlite3.connect(path)
    rows = con.execute("SELECT * FROM FACE").fetchall()
    con.close()
    return [r[1] for r in rows], [r[3] for r in rows]


@pytest.mark.parametrize("world", [1, 3])
def test_load_add_remove_lookup(tmp_path, world, built_lib):
    import frb200

    users, emb = make_db(tmp_path / "face.db", 50)
    ids, blobs = read_faces(tmp_path / "face.db")
    assert ids == users and np.array_equal(np.frombuffer(blobs[7], "<f4"), emb[7])
    rosters = [frb200.Roster(None, world, r) for r in range(world)]
    for ro in rosters:
        ro.load(ids, blobs)
    per = (50 + world - 1) // world
    for ro in rosters:
        assert ro.rows == 50 and [ro.shard_rows(g) for g in range(world)] == [max(0, min(50, (g + 1) * per) - min(50, g * per)) for g in range(world)]
        for i in range(50):                                  # contiguous blocks, shard 0 first: database order
            assert ro.user((i // per) << 32 | (i % per)) == users[i]
        assert ro.user(-1) is None and ro.user((world) << 32) is None and ro.user(per + 5 if world == 1 else (0 << 32) | per) is None
    # enrolment goes to the least-loaded shard (lowest on ties) at its end
    model = [[users[i] for i in range(min(50, g * per), min(50, (g + 1) * per))] for g in range(world)]
    for j in range(7):
        tgt = min(range(world), key=lambda g: (len(model[g]), g))
        got = [ro.add(f"new{j}", emb[j]) for ro in rosters]
        assert set(got) == {(tgt << 32) | len(model[tgt])}
        model[tgt].append(f"new{j}")
    # deletion: the shard's last row moves into the slot
    for shard, local in ((0, 0), (world - 1, 3), (0, -2)):
        local = local if local >= 0 else len(model[shard]) + local
        for ro in rosters:
            ro.remove((shard << 32) | local)
        model[shard][local] = model[shard][-1]
        model[shard].pop()
    victim = users[5]
    want_removed = sum(m.count(victim) for m in model)
    assert want_removed >= 1
    assert {ro.remove_user(victim) for ro in rosters} == {want_removed}
    for g in range(world):                                   # same multiset per shard; order follows the descending move-last rule
        l = len(model[g]) - 1
        while l >= 0:
            if model[g][l] == victim:
                model[g][l] = model[g][-1]
                model[g].pop()
            l -= 1
    for ro in rosters:
        for g in range(world):
            assert [ro.user((g << 32) | l) for l in range(ro.shard_rows(g))] == model[g]
    for ro in rosters:
        ro.clear()
        assert ro.rows == 0
        ro.close()


def test_malformed_blob_is_refused(tmp_path, built_lib):
    import frb200

    ro = frb200.Roster(None, 1, 0)
    with pytest.raises(frb200.FrError) as e:
        ro.load(["a", "b"], [np.zeros(512, "<f4").tobytes(), np.zeros(100, "<f4").tobytes()])
    assert e.value.code == frb200.FR_EFORMAT and "2048" in e.value.msg
    assert ro.rows == 0
    with pytest.raises(frb200.FrError):
        frb200.Roster(None, 2, 2)
    ro.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, db_path, out_dir):
    import frb200
    import torch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids, blobs = read_faces(db_path)                     # every rank reads the same database
        ro = frb200.Roster(None, world, rank)
        ro.load(ids, blobs)
        rng = np.random.default_rng(1)                       # the same operation stream on every rank (SPMD)
        for j in range(20):
            if rng.random() < 0.5:
                ro.add(f"enrol{j}", rng.standard_normal(512).astype(np.float32))
            else:
                g = int(rng.integers(0, world))
                if ro.shard_rows(g):
                    ro.remove((g << 32) | int(rng.integers(0, ro.shard_rows(g))))
        ro.remove_user(ids[3])
        table = "|".join(",".join(ro.user((g << 32) | l) for l in range(ro.shard_rows(g))) for g in range(world))
        mine = torch.tensor([hash(table) % (1 << 62), ro.rows], dtype=torch.int64)
        both = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(both, mine)
        assert all(int(b[1]) == ro.rows for b in both)
        Path(out_dir, f"table{rank}").write_text(table)
        ro.close()
    finally:
        dist.destroy_process_group()


def test_two_ranks_keep_identical_tables(tmp_path, built_lib):
    make_db(tmp_path / "face.db", 31, seed=3)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path / "face.db"), str(tmp_path)), nprocs=2, join=True)
    t0, t1 = (tmp_path / "table0").read_text(), (tmp_path / "table1").read_text()
    assert t0 == t1 and t0.count(",") > 10
