"""Launch-list target: top-1 search of 256 queries WITHOUT a match against a synthetic shard (what 7 of 8 shards see for every query
batch of the 8-GPU bench). python tools/prof_search_unknown.py [rows] [f8|f16]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))
import frb200  # noqa: E402
from oracle import search_oracle as so  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
scan = sys.argv[2] if len(sys.argv) > 2 else "f8"
g = frb200.Gallery.synthetic(rows, seed=19)
g.set_path(frb200.FR_PATH_TENSOR)
if scan == "f8":
    g.set_scan(frb200.FR_SCAN_F8)
q = so.l2_normalise(np.random.default_rng(24).standard_normal((256, 512))).astype(np.float32)
for _ in range(5):
    s, i = g.topk(q, 1)
print(scan, rows, "flagged", g.last_flagged() if hasattr(g, "last_flagged") else "?", float(s.max()))
g.close()
