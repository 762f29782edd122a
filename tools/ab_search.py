"""A/B target: sustained top-1 search time on one shard, CUDA events over `reps` back-to-back searches (graph-free, device-resident).
python tools/ab_search.py [rows] [scan:kind,...] [reps]   (FR_B200_LIB selects the library build)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))
import frb200  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
configs = (sys.argv[2] if len(sys.argv) > 2 else "f8:unknown,f8:planted,f16:unknown").split(",")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 300
g = frb200.Gallery.synthetic(rows, seed=19)
g.set_path(frb200.FR_PATH_TENSOR)
base = g.read_rows(0, 256)
noise = np.random.default_rng(24).standard_normal((256, 512)).astype(np.float32)
noise /= np.linalg.norm(noise, axis=1, keepdims=True)
s = torch.empty((256, 1), device="cuda")
i = torch.empty((256, 1), dtype=torch.int64, device="cuda")
ts_ = torch.cuda.Stream()  # not the NULL stream: the library reads NULL as "the handle's own stream"
st = ts_.cuda_stream
out = []
for cfg in configs:
    scan, kind = cfg.split(":")
    g.set_scan(frb200.FR_SCAN_F8 if scan == "f8" else frb200.FR_SCAN_F16)
    q = (0.8 * base + 0.6 * noise) if kind == "planted" else noise.copy()  # planted: cos ~0.8 to rows 0..255
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    qt = torch.from_numpy(q.astype(np.float32)).cuda()
    for _ in range(20):
        g.topk_dev(qt, 1, s, i, stream=st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts_)
        for _ in range(reps):
            g.topk_dev(qt, 1, s, i, stream=st)
        e1.record(ts_)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ok = bool((i.cpu().numpy()[:, 0] == np.arange(256)).all()) if kind == "planted" else None
    out.append(f"{cfg} " + "/".join(f"{t * 1e3:.0f}" for t in ts) + f" us flagged={g.last_flagged()}" + (f" top1_ok={ok}" if ok is not None else ""))
tag = Path(os.environ["FR_B200_LIB"]).parent.name if os.environ.get("FR_B200_LIB") else "default"
print(f"{tag:8s} rows={rows} | " + " | ".join(out))
g.close()
