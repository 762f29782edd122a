"""ORACLE / TEST INFRASTRUCTURE: fp32 forward of the ArcFace network computed from the PACKED tensors (the file the CUDA
library loads), following the library's dataflow (folded BatchNorms, producer-side pre-activation BN, padded NHWC Linear).
It validates tools/pack_weights.py on the CPU — if this agrees with oracle/arcface_oracle.py, any GPU mismatch is a kernel
bug, not a packing bug. Never used by the product."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from tools.synth_weights import arcface_blocks


def _w(t, cout, cin, k):
    """[Cout][ (ky*k+kx)*Cin + c ] -> [Cout, Cin, k, k] f32"""
    return torch.from_numpy(np.asarray(t, np.float32).reshape(cout, k, k, cin).transpose(0, 3, 1, 2).copy())


@torch.no_grad()
def forward_packed(T: dict, x: torch.Tensor, mode: str) -> torch.Tensor:
    v = lambda n: torch.from_numpy(np.asarray(T[n], np.float32))  # noqa: E731
    y = F.conv2d(x, _w(T["stem.w"], 64, 3, 3), v("stem.b"), 1, 1)
    y = F.prelu(y, v("stem.prelu"))
    for i, (cin, d, s) in enumerate(arcface_blocks()):
        q = f"u{i}."
        xb = y * v(q + "bn1.s")[None, :, None, None] + v(q + "bn1.b")[None, :, None, None]
        t = F.prelu(F.conv2d(xb, _w(T[q + "conv1.w"], d, cin, 3), None, 1, 1), v(q + "prelu"))
        r = F.conv2d(t, _w(T[q + "conv2.w"], d, d, 3), v(q + "conv2.b"), s, 1)
        if mode == "ir_se":
            g = r.mean(dim=(2, 3))
            g = torch.sigmoid(F.linear(F.relu(F.linear(g, v(q + "se.fc1"))), v(q + "se.fc2")))
            r = r * g[:, :, None, None]
        if cin == d:
            sc = y[:, :, ::s, ::s]
        else:
            sc = F.conv2d(y[:, :, ::s, ::s], _w(T[q + "sc.w"], d, cin, 1), v(q + "sc.b"))
        y = r + sc
    B = y.shape[0]
    pad = torch.zeros(B, 8, 8, 512)
    pad[:, :7, :7, :] = y.permute(0, 2, 3, 1)
    z = F.linear(pad.reshape(B, -1), v("fc.w"), v("fc.b"))
    return F.normalize(z, p=2.0, dim=1)
