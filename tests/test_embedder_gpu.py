"""GPU parity of the ArcFace IR-50 / IR-SE-50 embedder through the C ABI.
Golden embeddings come from the reference's own PyTorch module (tools/make_golden_nets.py); the restated oracle
(oracle/arcface_oracle.py) provides the per-layer trace. Tolerance: |d| <= 1e-3 on unit-norm outputs (BASELINE.json north_star;
activations are fp16 like the reference's TensorRT fp16 engine, conversion/arcface/torch2trt.py:26)."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import frb200
from oracle import arcface_oracle as ao
from tools import synth_weights as sw

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tools import make_golden_nets as mg  # noqa: E402
from tools import pack_weights as pw  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden"
TOL = 1e-3


@pytest.fixture(scope="module", params=["ir", "ir_se"])
def net(request, tmp_path_factory):
    mode = request.param
    sd = sw.arcface_state_dict(mode, 7)
    f = tmp_path_factory.mktemp("w") / f"arcface_{mode}.frw"
    pw.save_arcface(f, sd, mode)
    emb = frb200.Embedder(f, max_batch=40)
    crops = mg.arcface_inputs()
    yield mode, sd, emb, crops
    emb.close()


def test_embeddings_match_reference_golden(net):
    mode, sd, emb, crops = net
    gold = np.load(GOLD / f"arcface_{mode}_seed7.npz")["embeddings"]
    assert emb.mode == (frb200.FR_MODE_IR_SE if mode == "ir_se" else frb200.FR_MODE_IR)
    out = emb.run(ao.preprocess_faces(crops))
    assert out.shape == gold.shape
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-4)
    err = float(np.abs(out - gold).max())
    cos = np.einsum("ij,ij->i", out, gold)
    print(f"{mode}: max|d| = {err:.2e}, min cosine = {cos.min():.6f}")
    assert err <= TOL
    assert cos.min() > 0.9995


def test_per_layer_trace_against_oracle(net):
    mode, sd, emb, crops = net
    x = ao.preprocess_faces(crops[:2])
    emb.run(x)
    trace = []
    ao.forward(ao.to_torch(sd), torch.from_numpy(x), mode, trace)
    worst = 0.0
    for layer in (0, 1, 2, 3, 4, 8, 21, 22, 24):
        got = emb.trace(layer, 2)
        want = trace[layer].numpy()
        assert got.shape == want.shape
        rel = float(np.abs(got - want).max() / (np.abs(want).max() + 1e-6))
        worst = max(worst, rel)
        assert rel <= 2e-2, f"layer {layer}: relative error {rel:.3e}"
    print(f"{mode}: worst per-layer relative error {worst:.2e}")


def test_u8_crop_path_equals_f32_path_and_batches(net):
    mode, sd, emb, crops = net
    ref = emb.run(ao.preprocess_faces(crops))
    via_u8 = emb.run_crops(crops)
    assert np.array_equal(ref.view(np.uint32), via_u8.view(np.uint32))
    # batch-size independence (each face is computed independently) and determinism
    one = emb.run_crops(crops[3:4])
    assert np.array_equal(one[0].view(np.uint32), ref[3].view(np.uint32))
    big = np.concatenate([crops] * 5)  # 40 faces = max_batch
    out = emb.run_crops(big)
    assert np.array_equal(out[:8].view(np.uint32), ref.view(np.uint32)) and np.array_equal(out[32:].view(np.uint32), ref.view(np.uint32))


def test_errors(net, tmp_path):
    mode, sd, emb, crops = net
    with pytest.raises(frb200.FrError) as e:
        frb200.Embedder(tmp_path / "missing.frw")
    assert e.value.code == frb200.FR_ENOENT and "Cant find engine file" in e.value.msg  # src/arcface.cpp:67
    bad = tmp_path / "bad.frw"
    bad.write_bytes(b"not a weight file at all, definitely" * 4)
    with pytest.raises(frb200.FrError) as e:
        frb200.Embedder(bad)
    assert e.value.code == frb200.FR_EFORMAT
    with pytest.raises(frb200.FrError) as e:
        emb.run_crops(np.zeros((41, 112, 112, 3), np.uint8))
    assert e.value.code == frb200.FR_EINVAL
