"""compute-sanitizer target: the e4m3 top-1 path (append scan, per-list walk, exact-leader filter, fp16 pre-filter, compaction, final
re-score) on a gallery small enough for racecheck, with queries that have no match (hundreds of in-margin rows per query).
compute-sanitizer --tool racecheck python tools/racecheck_f8_small.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))
import frb200  # noqa: E402

rng = np.random.default_rng(5)
n = 20_000
G = rng.standard_normal((n, 512)).astype(np.float32)
G /= np.linalg.norm(G, axis=1, keepdims=True)
q = rng.standard_normal((256, 512)).astype(np.float32)
q /= np.linalg.norm(q, axis=1, keepdims=True)
g = frb200.Gallery.from_rows(G)
g.set_path(frb200.FR_PATH_TENSOR)
s16, i16 = g.topk(q, 1)
g.set_scan(frb200.FR_SCAN_F8)
s8, i8 = g.topk(q, 1)
print("flagged", g.last_flagged(), "equal", bool(np.array_equal(i8, i16) and np.array_equal(s8.view(np.uint32), s16.view(np.uint32))))
ref = (q @ G.T).argmax(1)
print("top1 exact", bool(np.array_equal(i8[:, 0], ref)))
g.close()
