"""Target process for ncu captures of every kernel that is NOT the search scan: one detector forward (batch 16,
640x640), one embedder forward at batch 32 and 256 (IR-SE-50), one pipeline batch (compaction + crop), one exchange step.
Kernels are launched eagerly (FR_NO_GRAPHS) inside the NVTX range "prof" after an untimed warm-up outside it:

    ncu --set full --clock-control none --nvtx --nvtx-include "prof/" --csv --page raw --log-file gpurun_out/r02_nets_ncu_raw.csv \
        python tools/prof_target.py
    python tools/ncu_summary.py gpurun_out/r02_nets_ncu_raw.csv > profiles/r02_nets_ncu_summary.csv
"""
import os
import sys
import tempfile
from pathlib import Path

os.environ.setdefault("FR_NO_GRAPHS", "1")
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "face-recognition-cpp-tensorrt_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import frb200  # noqa: E402
from tools import make_golden_nets as mg  # noqa: E402
from tools import make_golden_retina as mgr  # noqa: E402
from tools import pack_retina as pr  # noqa: E402
from tools import pack_weights as pw  # noqa: E402
from tools import synth_weights as sw  # noqa: E402


def main():
    what = set((sys.argv[1] if len(sys.argv) > 1 else "detect,embed32,embed256,pipeline,exchange").split(","))
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        pr.save_retina(td / "det.frw", sw.retina_state_dict(False, 11, mgr.DET_CLS_SHIFT), False)
        pw.save_arcface(td / "arc.frw", sw.arcface_state_dict("ir_se", 7), "ir_se")
        det = frb200.Detector(td / "det.frw", (640, 640), max_batch=16, max_faces=4)
        emb = frb200.Embedder(td / "arc.frw", max_batch=256)
        frames = mgr.det_frames(4, 640, 640, seed=13)
        frames16 = np.ascontiguousarray(np.concatenate([frames] * 4))
        crops = np.ascontiguousarray(np.concatenate([mg.arcface_inputs()] * 32))
        gal = frb200.Gallery.synthetic(200_000, seed=17)
        gal.set_path(frb200.FR_PATH_TENSOR)
        pipe = frb200.Pipeline(det, emb, gal)
        x = frb200.Exchange(0, 1, 0, nq_max=256, k_max=1)
        x.connect_local([x])
        q = torch.randn((256, 512), device="cuda")
        q /= q.norm(dim=1, keepdim=True)
        ls, li = torch.empty((256, 1), device="cuda"), torch.empty((256, 1), dtype=torch.int64, device="cuda")
        os_, oi = torch.empty((256, 1), device="cuda"), torch.empty((256, 1), dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def run():
            if "detect" in what:
                det.run(frames16)
            if "embed32" in what:
                emb.run_crops(crops[:32])
            if "embed256" in what:
                emb.run_crops(crops[:256])
            if "pipeline" in what:
                pipe.run(frames16)
            if "exchange" in what:
                x.topk_push_dev(gal, q, 1, ls, li, stream=st)
                x.wait_merge_dev(os_, oi, stream=st)
            torch.cuda.synchronize()

        run()  # warm-up, outside the profiled range
        torch.cuda.nvtx.range_push("prof")
        run()
        torch.cuda.nvtx.range_pop()
        for o in (pipe, x, gal, det, emb):
            o.close()


if __name__ == "__main__":
    main()
