"""tests/golden/retina_*.npz: outputs of the reference's OWN RetinaFace modules (trimmed = deployed, full = with landmark
head; imported read-only from /root/reference/conversion/retina) on the seeded synthetic checkpoint and synthetic frames.
Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_nets.py retina
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/conversion")
GOLD = ROOT / "tests" / "golden"

from tools import synth_weights as sw  # noqa: E402

DET_SEED, DET_FRAME_SEED = 11, 11
DET_CLS_SHIFT = -4.5  # class-head bias shift: ~150 of the 16 800 anchors of a 640x640 noise frame pass the 0.6 threshold
SUB = 7               # the 640x640 golden keeps every SUB-th anchor (file size)


def det_frames(n, h, w, seed=DET_FRAME_SEED):
    """u8 BGR frames, uniform noise, hash generator (no torch RNG)"""
    return np.floor(sw.uniform(seed, f"retina.frames.{h}x{w}", (n, h, w, 3), 0.0, 256.0)).clip(0, 255).astype(np.uint8)


def make_retina():
    sys.path.insert(0, str(REF / "retina"))
    from config import cfg_mnet
    from models.retinaface import RetinaFace as RetinaFull
    from models.retinaface_trim import RetinaFace as RetinaTrim

    from oracle import retina_oracle as ro

    torch.set_grad_enabled(False)
    cfg = dict(cfg_mnet)
    cfg["pretrain"] = False
    for full, ctor in ((False, RetinaTrim), (True, RetinaFull)):
        tag = "full" if full else "trim"
        sd = sw.retina_state_dict(full, DET_SEED, DET_CLS_SHIFT)
        m = ctor(cfg, phase="test").eval()
        m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        out = {}
        for (h, w, n, sub) in ((96, 128, 2, 1), (640, 640, 1, SUB)):
            frames = det_frames(n, h, w)
            x = torch.from_numpy(np.stack([ro.preprocess(f, h, w) for f in frames]))
            ref = m(x)
            mine = ro.forward(ro.to_torch(sd), x, full)
            for a, b in zip(ref, mine):
                if b is not None:
                    assert float((a - b).abs().max()) < 1e-5, float((a - b).abs().max())
            assert ref[0].shape[1] == ro.num_anchors(h, w)
            key = f"{h}x{w}"
            out[key + ".loc"] = ref[0].numpy()[:, ::sub].astype(np.float32)
            out[key + ".conf"] = ref[1].numpy()[:, ::sub].astype(np.float32)
            if full:
                out[key + ".landm"] = ref[2].numpy()[:, ::sub].astype(np.float32)
            out[key + ".n_pass"] = (ref[1].numpy()[..., 1] > 0.6).sum(axis=1).astype(np.int32)
            print(f"retina {tag} {key}: reference module == restated oracle; anchors {ref[0].shape[1]}, >0.6: {out[key + '.n_pass']}")
        np.savez_compressed(GOLD / f"retina_{tag}_seed{DET_SEED}.npz", sub=np.int32(SUB), cls_shift=np.float32(DET_CLS_SHIFT), **out)


if __name__ == "__main__":
    make_retina()
