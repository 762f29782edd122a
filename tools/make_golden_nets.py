"""Generate tests/golden/arcface_*.npz (and retina_*.npz) HERE, in the build container, by running the reference's OWN
PyTorch modules (imported read-only from /root/reference/conversion) on the seeded synthetic checkpoints of
tools/synth_weights.py. The GPU box has no /root/reference: there the tests compare against these committed vectors and
against the restated oracles (oracle/arcface_oracle.py, oracle/retina_oracle.py), which this script also pins.

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden_nets.py [arcface] [retina]
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

ARC_SEED, ARC_INPUT_SEED, ARC_N = 7, 7, 8


def arcface_inputs(n=ARC_N, seed=ARC_INPUT_SEED):
    """u8 BGR crops, uniform [0,255], generated with the hash generator (no torch RNG)"""
    return np.floor(sw.uniform(seed, "arcface.crops", (n, 112, 112, 3), 0.0, 256.0)).clip(0, 255).astype(np.uint8)


def make_arcface():
    sys.path.insert(0, str(REF / "arcface"))
    import model_irse  # the reference's module

    from oracle import arcface_oracle as ao

    torch.set_grad_enabled(False)
    crops = arcface_inputs()
    x = torch.from_numpy(ao.preprocess_faces(crops))
    for mode, ctor in (("ir", model_irse.IR_50), ("ir_se", model_irse.IR_SE_50)):
        sd = sw.arcface_state_dict(mode, ARC_SEED)
        m = ctor([112, 112]).eval()
        missing = m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        ref = m(x).numpy()
        trace = []
        mine = ao.forward(ao.to_torch(sd), x, mode, trace).numpy()
        err = float(np.abs(ref - mine).max())
        print(f"arcface {mode}: reference module vs restated oracle max|d| = {err:.3e}; norms {np.linalg.norm(ref, axis=1)[:3]}")
        assert err < 1e-6
        # per-layer fingerprints (mean |x| and a few raw values) keep the file small but let a per-layer trace localise an error
        layer_absmean = np.array([float(t.abs().mean()) for t in trace], np.float32)
        layer_probe = np.stack([t[0, :4, 0, 0].numpy() for t in trace]).astype(np.float32)
        np.savez_compressed(GOLD / f"arcface_{mode}_seed{ARC_SEED}.npz", embeddings=ref.astype(np.float32), layer_absmean=layer_absmean,
                            layer_probe=layer_probe, n=np.int32(ARC_N), seed=np.int32(ARC_SEED), input_seed=np.int32(ARC_INPUT_SEED))


if __name__ == "__main__":
    what = sys.argv[1:] or ["arcface", "retina"]
    GOLD.mkdir(exist_ok=True)
    if "arcface" in what:
        make_arcface()
    if "retina" in what:
        from tools.make_golden_retina import make_retina  # noqa

        make_retina()
