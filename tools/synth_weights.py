"""DATA GENERATOR (tests, benches, the weight packer; not an oracle, not part of the library): deterministic synthetic checkpoints
for the two networks of the hot path.

The reference ships no weights (weight/ and *.engine are git-ignored, /root/reference/.gitignore:6-9) and there is no
network, so parity is defined on seeded synthetic checkpoints. They are generated with integer hashing + numpy only (no
torch RNG) so that the GPU box — which has no /root/reference — regenerates bit-identical tensors, while the golden
outputs under tests/golden/ were produced HERE by loading the same tensors into the reference's own PyTorch modules
(tools/make_golden_nets.py).

State-dict key names and shapes follow the reference modules:
  ArcFace  conversion/arcface/model_irse.py:139-166 (input_layer / body.N.{shortcut_layer,res_layer} / output_layer)
  Retina   conversion/retina/models/retinaface_trim.py:55-99, retinaface.py:37-46,87 and models/net.py
BatchNorm statistics, PReLU slopes and biases are randomised: the default init (mean 0, var 1) would hide
BN bugs.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _fnv1a(name: str) -> np.uint64:
    h = 0xCBF29CE484222325
    for b in name.encode():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return np.uint64(h)


def _bits(seed: int, name: str, n: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        key = _mix64(np.uint64(seed) ^ _fnv1a(name))
        return _mix64(key + np.arange(n, dtype=np.uint64))


def uniform(seed, name, shape, lo, hi) -> np.ndarray:
    n = int(np.prod(shape))
    u = (_bits(seed, name, n) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def normal(seed, name, shape, mean, std) -> np.ndarray:
    n = int(np.prod(shape))
    h = _bits(seed, name, n)
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))).astype(np.float64) - 131070.0
    z = s / 37837.226  # std of a sum of four U{0..65535}
    return (mean + std * z).astype(np.float32).reshape(shape)


def _conv(sd, seed, name, cout, cin_per_group, kh, kw, gain=1.0):
    fan_in, fan_out = cin_per_group * kh * kw, cout * kh * kw
    bound = gain * np.sqrt(6.0 / (fan_in + fan_out))  # xavier_uniform_, model_irse.py:178
    sd[name] = uniform(seed, name, (cout, cin_per_group, kh, kw), -bound, bound)


def _bn(sd, seed, prefix, c):
    sd[prefix + ".weight"] = uniform(seed, prefix + ".weight", (c,), 0.5, 1.5)
    sd[prefix + ".bias"] = normal(seed, prefix + ".bias", (c,), 0.0, 0.1)
    sd[prefix + ".running_mean"] = normal(seed, prefix + ".running_mean", (c,), 0.0, 0.1)
    sd[prefix + ".running_var"] = uniform(seed, prefix + ".running_var", (c,), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = np.zeros((), np.int64)


# ---------------------------------------------------------------------------------------------- ArcFace IR-50 / IR-SE-50
def arcface_blocks():
    """(in_channel, depth, stride) of the 24 units: get_blocks(50), model_irse.py:97-109"""
    out = []
    for cin, depth, units in ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)):
        out.append((cin, depth, 2))
        out += [(depth, depth, 1)] * (units - 1)
    return out


def arcface_state_dict(mode: str = "ir_se", seed: int = 7) -> "OrderedDict[str, np.ndarray]":
    assert mode in ("ir", "ir_se")
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    _conv(sd, seed, "input_layer.0.weight", 64, 3, 3, 3)
    _bn(sd, seed, "input_layer.1", 64)
    sd["input_layer.2.weight"] = uniform(seed, "input_layer.2.weight", (64,), 0.1, 0.4)
    for i, (cin, d, s) in enumerate(arcface_blocks()):
        p = f"body.{i}."
        if cin != d:
            _conv(sd, seed, p + "shortcut_layer.0.weight", d, cin, 1, 1)
            _bn(sd, seed, p + "shortcut_layer.1", d)
        _bn(sd, seed, p + "res_layer.0", cin)
        _conv(sd, seed, p + "res_layer.1.weight", d, cin, 3, 3)
        sd[p + "res_layer.2.weight"] = uniform(seed, p + "res_layer.2.weight", (d,), 0.1, 0.4)
        _conv(sd, seed, p + "res_layer.3.weight", d, d, 3, 3)
        _bn(sd, seed, p + "res_layer.4", d)
        if mode == "ir_se":
            _conv(sd, seed, p + "res_layer.5.fc1.weight", d // 16, d, 1, 1)
            _conv(sd, seed, p + "res_layer.5.fc2.weight", d, d // 16, 1, 1)
    _bn(sd, seed, "output_layer.0", 512)
    bound = np.sqrt(6.0 / (25088 + 512))
    sd["output_layer.3.weight"] = uniform(seed, "output_layer.3.weight", (512, 25088), -bound, bound)
    sd["output_layer.3.bias"] = normal(seed, "output_layer.3.bias", (512,), 0.0, 0.01)
    _bn(sd, seed, "output_layer.4", 512)
    return sd


# ---------------------------------------------------------------------------------------------- RetinaFace mobile0.25
def _conv_bn(sd, seed, prefix, cin, cout, k, groups=1):
    """conv_bn / conv_bn_no_relu / conv_bn1X1 / half of conv_dw: Sequential(Conv2d(bias=False), BatchNorm2d[, ReLU])"""
    _conv(sd, seed, prefix + ".0.weight", cout, cin // groups, k, k, gain=1.7)
    _bn(sd, seed, prefix + ".1", cout)


def _conv_dw(sd, seed, prefix, cin, cout):
    """conv_dw, models/net.py:29-38: dw3x3 + BN + ReLU + pw1x1 + BN + ReLU = Sequential indices 0,1,(2),3,4,(5)"""
    _conv(sd, seed, prefix + ".0.weight", cin, 1, 3, 3, gain=1.7)
    _bn(sd, seed, prefix + ".1", cin)
    _conv(sd, seed, prefix + ".3.weight", cout, cin, 1, 1, gain=1.7)
    _bn(sd, seed, prefix + ".4", cout)


def retina_state_dict(full: bool = False, seed: int = 11, cls_bias_shift: float = 0.0) -> "OrderedDict[str, np.ndarray]":
    """trimmed model (retinaface_trim.py, the deployed one) or full model with the landmark head (retinaface.py)."""
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    # MobileNetV1 (net.py:102-124), wrapped by IntermediateLayerGetter as `body`
    _conv_bn(sd, seed, "body.stage1.0", 3, 8, 3)
    for i, (a, b) in enumerate(((8, 16), (16, 32), (32, 32), (32, 64), (64, 64)), start=1):
        _conv_dw(sd, seed, f"body.stage1.{i}", a, b)
    for i, (a, b) in enumerate(((64, 128), (128, 128), (128, 128), (128, 128), (128, 128), (128, 128))):
        _conv_dw(sd, seed, f"body.stage2.{i}", a, b)
    for i, (a, b) in enumerate(((128, 256), (256, 256))):
        _conv_dw(sd, seed, f"body.stage3.{i}", a, b)
    # FPN (net.py:68-98)
    for name, cin in (("output1", 64), ("output2", 128), ("output3", 256)):
        _conv_bn(sd, seed, f"fpn.{name}", cin, 64, 1)
    _conv_bn(sd, seed, "fpn.merge1", 64, 64, 3)
    _conv_bn(sd, seed, "fpn.merge2", 64, 64, 3)
    # SSH x3 (net.py:40-66)
    for s in (1, 2, 3):
        _conv_bn(sd, seed, f"ssh{s}.conv3X3", 64, 32, 3)
        _conv_bn(sd, seed, f"ssh{s}.conv5X5_1", 64, 16, 3)
        _conv_bn(sd, seed, f"ssh{s}.conv5X5_2", 16, 16, 3)
        _conv_bn(sd, seed, f"ssh{s}.conv7X7_2", 16, 16, 3)
        _conv_bn(sd, seed, f"ssh{s}.conv7x7_3", 16, 16, 3)
    # heads (retinaface_trim.py:9-35,89-99): Conv2d 1x1 with bias, anchor_num = 2
    heads = [("ClassHead", 4), ("BboxHead", 8)] + ([("LandmarkHead", 20)] if full else [])
    for head, cout in heads:
        for lvl in range(3):
            p = f"{head}.{lvl}.conv1x1"
            _conv(sd, seed, p + ".weight", cout, 64, 1, 1, gain=1.0 if head == "ClassHead" else 0.5)
            sd[p + ".bias"] = normal(seed, p + ".bias", (cout,), 0.0, 0.05)
            if head == "ClassHead" and cls_bias_shift:
                # raise the "face" logit (odd channels: view(.., 2) -> [bg, face]) so that a controlled number of anchors
                # passes the 0.6 score threshold with random weights (SURVEY §7 hard part 9)
                sd[p + ".bias"][1::2] += np.float32(cls_bias_shift)
    return sd
