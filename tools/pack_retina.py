"""RetinaFace mobile0.25 packer: reference state dict -> tensors of the FRB2WTS1 file consumed by csrc/detector.cu.

Every BatchNorm of this network follows its conv directly (conversion/retina/models/net.py:9-38), so all of them fold:
w' = w * gamma / sqrt(var + eps) per output channel, b' = beta - mean * gamma / sqrt(var + eps).
  stem.w  [8][27] f32 (k = (ky*3+kx)*3 + c, c in B,G,R order as RetinaFace::preprocess leaves it), stem.b [8]
  dwN.w   [9][C] f32 (tap-major), dwN.b [C]                       N = 1..13 (conv_dw blocks in network order)
  pwN.w   [Cout][Cin] f16 (tensor-core GEMM rows) for Cin >= 64, f32 [Cin][Cout] for the four early layers; pwN.b [Cout]
  fpn.outputK.w [64][Cin] f16, fpn.mergeK.w [64][9*64] f16 (+ .b)
  sshL.a.w [32][576] f16, sshL.t.w [16][576] f16, sshL.{b,u,c}.w [9][16][16] f32 (tap, cin, cout)  (+ .b)
  headL.w [32][64] f16 = rows 0-7 BboxHead, 8-11 ClassHead, 12-31 LandmarkHead (zero when the model is the trimmed one), headL.b [32]
"""
from __future__ import annotations

import numpy as np

from tools.pack_weights import KIND_RETINA_FULL, KIND_RETINA_TRIM, bn_fold, conv_rows, strip_prefix, write_file

DW_BLOCKS = [("body.stage1.%d" % i) for i in range(1, 6)] + [("body.stage2.%d" % i) for i in range(6)] + [("body.stage3.%d" % i) for i in range(2)]


def _fold(sd, p, ci=0, bi=1):
    s, b = bn_fold(sd, f"{p}.{bi}")
    return sd[f"{p}.{ci}.weight"].astype(np.float64) * s[:, None, None, None], b


def pack_retina(sd: dict, full: bool) -> "dict[str, np.ndarray]":
    sd = strip_prefix(sd)
    out: "dict[str, np.ndarray]" = {}
    w, b = _fold(sd, "body.stage1.0")
    out["stem.w"] = conv_rows(w).astype(np.float32)
    out["stem.b"] = b.astype(np.float32)
    for n, p in enumerate(DW_BLOCKS, start=1):
        w, b = _fold(sd, p, 0, 1)                       # depthwise [C,1,3,3]
        out[f"dw{n}.w"] = np.ascontiguousarray(w.reshape(w.shape[0], 9).T).astype(np.float32)
        out[f"dw{n}.b"] = b.astype(np.float32)
        w, b = _fold(sd, p, 3, 4)                       # pointwise [Cout,Cin,1,1]
        w = w.reshape(w.shape[0], w.shape[1])
        out[f"pw{n}.w"] = w.astype(np.float16) if w.shape[1] >= 64 else np.ascontiguousarray(w.T).astype(np.float32)
        out[f"pw{n}.b"] = b.astype(np.float32)
    for name in ("output1", "output2", "output3", "merge1", "merge2"):
        w, b = _fold(sd, "fpn." + name)
        out[f"fpn.{name}.w"] = conv_rows(w).astype(np.float16)
        out[f"fpn.{name}.b"] = b.astype(np.float32)
    for lvl in (1, 2, 3):
        for short, name in (("a", "conv3X3"), ("t", "conv5X5_1")):
            w, b = _fold(sd, f"ssh{lvl}.{name}")
            out[f"ssh{lvl}.{short}.w"] = conv_rows(w).astype(np.float16)
            out[f"ssh{lvl}.{short}.b"] = b.astype(np.float32)
        for short, name in (("b", "conv5X5_2"), ("u", "conv7X7_2"), ("c", "conv7x7_3")):
            w, b = _fold(sd, f"ssh{lvl}.{name}")      # [16,16,3,3] -> [tap][cin][cout]
            out[f"ssh{lvl}.{short}.w"] = np.ascontiguousarray(w.transpose(2, 3, 1, 0).reshape(9, 16, 16)).astype(np.float32)
            out[f"ssh{lvl}.{short}.b"] = b.astype(np.float32)
        hw = np.zeros((32, 64), np.float64)
        hb = np.zeros(32, np.float64)
        hw[0:8] = sd[f"BboxHead.{lvl - 1}.conv1x1.weight"].reshape(8, 64)
        hb[0:8] = sd[f"BboxHead.{lvl - 1}.conv1x1.bias"]
        hw[8:12] = sd[f"ClassHead.{lvl - 1}.conv1x1.weight"].reshape(4, 64)
        hb[8:12] = sd[f"ClassHead.{lvl - 1}.conv1x1.bias"]
        if full:
            hw[12:32] = sd[f"LandmarkHead.{lvl - 1}.conv1x1.weight"].reshape(20, 64)
            hb[12:32] = sd[f"LandmarkHead.{lvl - 1}.conv1x1.bias"]
        out[f"head{lvl}.w"] = hw.astype(np.float16)
        out[f"head{lvl}.b"] = hb.astype(np.float32)
    return out


def save_retina(path, sd: dict, full: bool) -> None:
    write_file(path, KIND_RETINA_FULL if full else KIND_RETINA_TRIM, pack_retina(sd, full))
