"""CPU: the restated network oracles against the golden vectors produced by the reference's own PyTorch modules
(tools/make_golden_nets.py), and the weight packer against the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import arcface_oracle as ao  # noqa: E402
from tools import synth_weights as sw  # noqa: E402
from tools import make_golden_nets as mg  # noqa: E402
from tools import pack_weights as pw  # noqa: E402

GOLD = ROOT / "tests" / "golden"


@pytest.fixture(scope="module")
def arc_inputs():
    crops = mg.arcface_inputs()
    return crops, torch.from_numpy(ao.preprocess_faces(crops))


def test_preprocess_faces_matches_reference_arithmetic():
    crops = np.zeros((1, 112, 112, 3), np.uint8)
    crops[0, 0, 0] = (255, 0, 128)  # B, G, R
    x = ao.preprocess_faces(crops)
    assert x.shape == (1, 3, 112, 112) and x.dtype == np.float32
    # planar R, G, B; (x - 127.5) * 0.0078125 (src/arcface.cpp:118-125)
    assert x[0, 0, 0, 0] == np.float32(0.5 * 0.0078125) and x[0, 1, 0, 0] == np.float32(-127.5 * 0.0078125)
    assert x[0, 2, 0, 0] == np.float32(127.5 * 0.0078125)


@pytest.mark.parametrize("mode", ["ir", "ir_se"])
def test_arcface_oracle_matches_reference_golden(mode, arc_inputs):
    gold = np.load(GOLD / f"arcface_{mode}_seed7.npz")
    sd = sw.arcface_state_dict(mode, int(gold["seed"]))
    n = 3  # the first faces are enough to pin the restatement; keeps the CPU suite short
    trace = []
    out = ao.forward(ao.to_torch(sd), arc_inputs[1][:n], mode, trace).numpy()
    assert np.abs(out - gold["embeddings"][:n]).max() <= 1e-6
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
    probe = np.stack([t[0, :4, 0, 0].numpy() for t in trace])
    assert np.abs(probe - gold["layer_probe"]).max() <= 1e-4 * max(1.0, float(np.abs(gold["layer_probe"]).max()))


@pytest.mark.parametrize("mode", ["ir", "ir_se"])
def test_packer_folding_is_equivalent(mode, arc_inputs, tmp_path):
    from oracle.packed_forward import forward_packed

    gold = np.load(GOLD / f"arcface_{mode}_seed7.npz")
    sd = sw.arcface_state_dict(mode, 7)
    f = tmp_path / "w.frw"
    pw.save_arcface(f, sd, mode)
    kind, T = pw.read_file(f)
    assert kind == (pw.KIND_ARCFACE_IRSE if mode == "ir_se" else pw.KIND_ARCFACE_IR)
    assert T["u0.conv1.w"].dtype == np.float16 and T["u0.conv1.w"].shape == (64, 9 * 64)
    assert T["fc.w"].shape == (512, 64 * 512) and T["u3.sc.w"].shape == (128, 64)
    out = forward_packed(T, arc_inputs[1][:2], mode).numpy()
    # fp16 weights, fp32 arithmetic: well inside the 1e-3 budget
    assert np.abs(out - gold["embeddings"][:2]).max() <= 3e-4


def test_synthetic_checkpoint_is_deterministic_and_randomised():
    a = sw.arcface_state_dict("ir", 7)
    b = sw.arcface_state_dict("ir", 7)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert float(np.abs(a["body.0.res_layer.0.running_mean"]).max()) > 0.05  # BN stats are not the identity
    assert 0.5 <= float(a["body.5.res_layer.4.running_var"].min()) and float(a["body.5.res_layer.4.running_var"].max()) <= 1.5
