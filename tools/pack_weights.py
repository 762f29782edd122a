"""Weight packer: PyTorch-style state dict (numpy arrays, reference key names) -> flat "FRB2WTS1" file read by libfr_b200.

Replaces the role of the reference's TensorRT exporters (/root/reference conversion/arcface/torch2trt.py:21-24,
conversion/retina/torch2trt.py:28-64): the file path is what the C++ classes receive as `engineFile`.
`module.` prefixes are stripped like remove_prefix() there. numpy only.

ArcFace packing (consumed by csrc/embedder.cu):
  * conv weights -> [Cout][tap = ky*3+kx][Cin] fp16 (K-major rows for the tensor-core GEMM)
  * BatchNorm AFTER a conv (input_layer.1, res_layer.4, shortcut_layer.1) is folded into that conv (scale rows, fp32 bias)
  * BatchNorm BEFORE a conv (res_layer.0) cannot be folded — zero padding is applied after it (SURVEY §7 hard part 3) — it is
    kept as fp32 scale/bias and applied by the PRODUCER of the unit's input (dual write)
  * output_layer: BN2d(512) and BN1d(512) are folded into the Linear; its columns are permuted from the reference's
    NCHW flatten order (c*49 + h*7 + w) to this library's padded NHWC order ((h*8 + w)*512 + c, zero at the pad positions)
"""
from __future__ import annotations

import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

KIND_ARCFACE_IR, KIND_ARCFACE_IRSE, KIND_RETINA_TRIM, KIND_RETINA_FULL = 1, 2, 3, 4
BN_EPS = 1e-5


def strip_prefix(sd: dict) -> dict:
    return {(k[7:] if k.startswith("module.") else k): np.asarray(v) for k, v in sd.items()}


def bn_fold(sd, p):
    """(scale, shift) with y = x * scale + shift"""
    s = sd[p + ".weight"].astype(np.float64) / np.sqrt(sd[p + ".running_var"].astype(np.float64) + BN_EPS)
    b = sd[p + ".bias"].astype(np.float64) - sd[p + ".running_mean"].astype(np.float64) * s
    return s, b


def conv_rows(w: np.ndarray, out_scale=None) -> np.ndarray:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] with k = (ky*kw + kx)*Cin + c, optionally scaled per output channel"""
    w = w.astype(np.float64)
    if out_scale is not None:
        w = w * out_scale[:, None, None, None]
    return np.ascontiguousarray(w.transpose(0, 2, 3, 1).reshape(w.shape[0], -1))


def write_file(path, kind: int, tensors: "dict[str, np.ndarray]") -> None:
    names = list(tensors)
    rec = 96 + 4 + 4 + 32 + 8 + 8
    off = 24 + rec * len(names)
    off = (off + 255) // 256 * 256
    table, blobs = [], []
    for n in names:
        a = np.ascontiguousarray(tensors[n])
        assert a.dtype in (np.float32, np.float16), (n, a.dtype)
        assert a.ndim <= 4 and len(n.encode()) < 96
        dims = list(a.shape) + [1] * (4 - a.ndim)
        table.append(n.encode().ljust(96, b"\0") + struct.pack("<ii4qqq", 0 if a.dtype == np.float32 else 1, a.ndim, *dims, off, a.nbytes))
        blobs.append((off, a.tobytes()))
        off = (off + a.nbytes + 255) // 256 * 256
    with open(path, "wb") as f:
        f.write(b"FRB2WTS1" + struct.pack("<iiii", 1, kind, len(names), 0))
        for t in table:
            f.write(t)
        for o, b in blobs:
            f.seek(o)
            f.write(b)
        f.truncate(off)


# ------------------------------------------------------------------------------------------------------------ ArcFace
def pack_arcface(sd: dict, mode: str) -> "dict[str, np.ndarray]":
    from tools.synth_weights import arcface_blocks

    sd = strip_prefix(sd)
    out: "dict[str, np.ndarray]" = {}
    s, b = bn_fold(sd, "input_layer.1")
    out["stem.w"] = conv_rows(sd["input_layer.0.weight"], s).astype(np.float32)  # [64][27], CUDA-core kernel, fp32
    out["stem.b"] = b.astype(np.float32)
    out["stem.prelu"] = sd["input_layer.2.weight"].astype(np.float32)
    for i, (cin, d, stride) in enumerate(arcface_blocks()):
        p, q = f"body.{i}.", f"u{i}."
        s1, b1 = bn_fold(sd, p + "res_layer.0")
        out[q + "bn1.s"], out[q + "bn1.b"] = s1.astype(np.float32), b1.astype(np.float32)
        out[q + "conv1.w"] = conv_rows(sd[p + "res_layer.1.weight"]).astype(np.float16)
        out[q + "prelu"] = sd[p + "res_layer.2.weight"].astype(np.float32)
        s2, b2 = bn_fold(sd, p + "res_layer.4")
        out[q + "conv2.w"] = conv_rows(sd[p + "res_layer.3.weight"], s2).astype(np.float16)
        out[q + "conv2.b"] = b2.astype(np.float32)
        if cin != d:
            ss, bs = bn_fold(sd, p + "shortcut_layer.1")
            out[q + "sc.w"] = conv_rows(sd[p + "shortcut_layer.0.weight"], ss).astype(np.float16)
            out[q + "sc.b"] = bs.astype(np.float32)
        if mode == "ir_se":
            out[q + "se.fc1"] = sd[p + "res_layer.5.fc1.weight"].reshape(d // 16, d).astype(np.float32)
            out[q + "se.fc2"] = sd[p + "res_layer.5.fc2.weight"].reshape(d, d // 16).astype(np.float32)
    s0, b0 = bn_fold(sd, "output_layer.0")   # per input channel c
    s4, b4 = bn_fold(sd, "output_layer.4")   # per output feature o
    W = sd["output_layer.3.weight"].astype(np.float64).reshape(512, 512, 7, 7)  # [o][c][h][w]
    bias = s4 * (np.einsum("ochw,c->o", W, b0) + sd["output_layer.3.bias"].astype(np.float64)) + b4
    Wf = W * s0[None, :, None, None] * s4[:, None, None, None]
    Wp = np.zeros((512, 8, 8, 512), np.float64)                                  # [o][h][w][c], pad row/col 7 stay zero
    Wp[:, :7, :7, :] = Wf.transpose(0, 2, 3, 1)
    out["fc.w"] = Wp.reshape(512, 8 * 8 * 512).astype(np.float16)
    out["fc.b"] = bias.astype(np.float32)
    return out


def save_arcface(path, sd: dict, mode: str) -> None:
    write_file(path, KIND_ARCFACE_IRSE if mode == "ir_se" else KIND_ARCFACE_IR, pack_arcface(sd, mode))


if __name__ == "__main__":
    import argparse

    from tools import synth_weights as sw

    ap = argparse.ArgumentParser(description="pack a synthetic (seeded) or saved (.npz of a state dict) checkpoint")
    ap.add_argument("net", choices=["arcface_ir", "arcface_ir_se", "retina_trim", "retina_full"])
    ap.add_argument("out")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--npz", default=None, help="state dict saved with numpy.savez (keys = reference parameter names)")
    a = ap.parse_args()
    if a.net.startswith("arcface"):
        mode = a.net[len("arcface_"):]
        sd = dict(np.load(a.npz)) if a.npz else sw.arcface_state_dict(mode, a.seed if a.seed is not None else 7)
        save_arcface(a.out, sd, mode)
    else:
        from tools.pack_retina import save_retina  # noqa

        full = a.net == "retina_full"
        sd = dict(np.load(a.npz)) if a.npz else sw.retina_state_dict(full, a.seed if a.seed is not None else 11)
        save_retina(a.out, sd, full)
    print("wrote", a.out)


def read_file(path) -> "tuple[int, dict[str, np.ndarray]]":
    """inverse of write_file (tests)"""
    blob = Path(path).read_bytes()
    assert blob[:8] == b"FRB2WTS1"
    version, kind, n, _ = struct.unpack_from("<iiii", blob, 8)
    assert version == 1
    rec = 96 + 4 + 4 + 32 + 8 + 8
    out = {}
    for i in range(n):
        r = 24 + rec * i
        name = blob[r:r + 96].rstrip(b"\0").decode()
        dtype, ndim, d0, d1, d2, d3, off, nbytes = struct.unpack_from("<ii4qqq", blob, r + 96)
        dt = np.float32 if dtype == 0 else np.float16
        out[name] = np.frombuffer(blob, dt, count=nbytes // np.dtype(dt).itemsize, offset=off).reshape([d0, d1, d2, d3][:ndim])
    return kind, out
