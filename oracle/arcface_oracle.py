"""ORACLE (test infrastructure, never shipped or measured as the product): fp32 CPU restatement of the reference's
embedding path in PyTorch functional ops, driven by a plain state dict (so it runs on the GPU box, where /root/reference
does not exist). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Restated reference code (paths relative to /root/reference):
  preprocess_faces()  ArcFaceIR50::preprocessFace / preprocessFaces      src/arcface.cpp:105-129
  forward()           Backbone.forward, bottleneck_IR, bottleneck_IR_SE, SEModule, output_layer, F.normalize
                      conversion/arcface/model_irse.py:22-45,48-90,139-172
Parity status: the reference has no golden vectors for this path. The restatement is pinned against the
reference's OWN module (model_irse.Backbone, imported from /root/reference by tools/make_golden_nets.py in the build
container): same synthetic checkpoint, same inputs, outputs committed under tests/golden/arcface_*.npz and compared in
tests/test_oracle_nets.py. The TensorRT fp16 engine the reference deploys is third-party arithmetic (TensorRT 8.2.2.1,
README.md:10) and cannot be run here; tolerance against the fp32 module is 1e-3 (BASELINE.json north_star).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from tools.synth_weights import arcface_blocks

BN_EPS = 1e-5  # nn.BatchNorm2d default, used by every BatchNorm in model_irse.py


def preprocess_faces(crops_bgr_u8: np.ndarray) -> np.ndarray:
    """n x 112 x 112 x 3 u8 BGR -> n x 3 x 112 x 112 f32 planar R,G,B, (x - 127.5) * 0.0078125 (src/arcface.cpp:118-125)"""
    x = crops_bgr_u8[..., ::-1].astype(np.float32)  # cvtColor BGR2RGB, convertTo CV_32F
    x = (x - np.float32(127.5)) * np.float32(0.0078125)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))  # split + push_back -> planar CHW


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def to_torch(sd_np) -> dict:
    return {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items() if v.dtype != np.int64}


@torch.no_grad()
def forward(sd: dict, x: torch.Tensor, mode: str = "ir_se", trace: list | None = None) -> torch.Tensor:
    """x: n x 3 x 112 x 112 f32 -> n x 512 unit-norm embeddings. `trace` (optional) receives the activation after the input
    layer and after every body unit (NCHW) for per-layer parity debugging."""
    x = F.conv2d(x, sd["input_layer.0.weight"], None, 1, 1)          # model_irse.py:139
    x = F.prelu(_bn(x, sd, "input_layer.1"), sd["input_layer.2.weight"])
    if trace is not None:
        trace.append(x)
    for i, (cin, d, s) in enumerate(arcface_blocks()):
        p = f"body.{i}."
        if cin == d:
            shortcut = F.max_pool2d(x, 1, s)                            # MaxPool2d(1, stride): pure subsampling, :51-52
        else:
            shortcut = _bn(F.conv2d(x, sd[p + "shortcut_layer.0.weight"], None, s, 0), sd, p + "shortcut_layer.1")
        r = _bn(x, sd, p + "res_layer.0")                               # pre-activation BN, then zero padding inside conv
        r = F.prelu(F.conv2d(r, sd[p + "res_layer.1.weight"], None, 1, 1), sd[p + "res_layer.2.weight"])
        r = _bn(F.conv2d(r, sd[p + "res_layer.3.weight"], None, s, 1), sd, p + "res_layer.4")
        if mode == "ir_se":                                             # SEModule, :22-45
            g = F.adaptive_avg_pool2d(r, 1)
            g = torch.sigmoid(F.conv2d(F.relu(F.conv2d(g, sd[p + "res_layer.5.fc1.weight"])), sd[p + "res_layer.5.fc2.weight"]))
            r = r * g
        x = r + shortcut                                                # no activation after the add, :61-65
        if trace is not None:
            trace.append(x)
    x = _bn(x, sd, "output_layer.0")                                   # Dropout is the identity in eval mode
    x = x.reshape(x.shape[0], -1)                                       # Flatten over NCHW: index = c*49 + h*7 + w
    x = F.linear(x, sd["output_layer.3.weight"], sd["output_layer.3.bias"])
    x = F.batch_norm(x, sd["output_layer.4.running_mean"], sd["output_layer.4.running_var"], sd["output_layer.4.weight"],
                     sd["output_layer.4.bias"], False, 0.0, BN_EPS)
    return F.normalize(x, p=2.0, dim=1)                                 # :171 (eps 1e-12)
