"""A/B target: device-resident detector forward (fr_detector_run_dev) at several batch sizes, CUDA events on an explicit stream.
python tools/ab_detect.py [batches, default 64,16] [reps]   (FR_B200_LIB selects the library build)"""
import os
import sys
import tempfile
import zlib
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))
import frb200  # noqa: E402
from tools import make_golden_retina as mgr  # noqa: E402
from tools import pack_retina as pr  # noqa: E402
from tools import synth_weights as sw  # noqa: E402

batches = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "64,16").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = []
with tempfile.TemporaryDirectory() as td:
    pr.save_retina(Path(td) / "det.frw", sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT), False)
    det = frb200.Detector(Path(td) / "det.frw", (640, 640), max_batch=max(batches), max_faces=4)
    base = mgr.det_frames(4, 640, 640, seed=13)
    ts_ = torch.cuda.Stream()
    for b in batches:
        fd = torch.from_numpy(np.ascontiguousarray(np.concatenate([base] * (b // 4 + 1))[:b])).cuda()
        bx = torch.zeros((b, 4, 5), dtype=torch.int32, device="cuda")
        ct = torch.zeros((b,), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for _ in range(5):
            det.run_dev(fd, bx, ct, stream=ts_.cuda_stream)
        torch.cuda.synchronize()
        t = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ts_)
            for _ in range(reps):
                det.run_dev(fd, bx, ct, stream=ts_.cuda_stream)
            e1.record(ts_)
            torch.cuda.synchronize()
            t.append(e0.elapsed_time(e1) / reps)
        crc = zlib.crc32(bx.cpu().numpy().tobytes() + ct.cpu().numpy().tobytes())
        out.append(f"b{b} " + "/".join(f"{x * 1e3:.0f}" for x in t) + f" us crc={crc:08x} faces={int(ct.sum())}")
    det.close()
tag = Path(os.environ["FR_B200_LIB"]).parent.name if os.environ.get("FR_B200_LIB") else "default"
print(f"{tag:8s} | " + " | ".join(out))
