"""ctypes driver of the plain-C restatement ``oracle/posenn_ref.c``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); pinned through the torch restatement it must agree with
(tests/test_oracle.py), which is held to the fixtures made by the reference's own code.
``build()`` compiles it with gcc into ``oracle/_build/`` (git-ignored).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "posenn_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libposenn_ref.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


class _Weights(C.Structure):
    _fp = C.POINTER(C.c_float)
    _fields_ = ([(n, C.POINTER(C.c_float)) for n in
                 ("w1", "b1", "w2", "b2", "w3", "b3", "w4", "b4", "w5", "b5")] +
                [(n, C.POINTER(C.c_float) * 2) for n in ("w6", "b6", "w7", "b7", "wp", "bp")] +
                [(n, C.POINTER(C.c_float)) for n in ("se_w1", "se_b1", "se_w2", "se_b2", "static_w")] +
                [(n, C.c_int) for n in ("cin1", "c6", "att_src", "mask_rgb", "mask_flow", "se_act",
                                        "flow_abs", "flow_norm")])


def forward(cfg: dict, img_u8: np.ndarray, flow: np.ndarray, seg: np.ndarray, weights: dict):
    """cfg: dict with the ``davo_config`` fields (att_src, mask_mode, se_act, flow_abs,
    flow_norm, in_mode, cnv6_out).  Returns (poses [B,2,6] float64, att_w [B,2,19])."""
    lib = C.CDLL(build())
    keep = []

    def fp(name):
        a = np.ascontiguousarray(weights[name], dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(C.POINTER(C.c_float))

    P = "pose_exp_net/"
    w = _Weights()
    for i in range(1, 6):
        setattr(w, "w%d" % i, fp(P + "cnv%d/weights" % i))
        setattr(w, "b%d" % i, fp(P + "cnv%d/biases" % i))
    for g, br in enumerate(("rotation", "translation")):
        for short, scope in (("6", "cnv6"), ("7", "cnv7"), ("p", "pred")):
            getattr(w, "w" + short)[g] = fp(P + "pose/%s/%s/weights" % (br, scope))
            getattr(w, "b" + short)[g] = fp(P + "pose/%s/%s/biases" % (br, scope))
    if cfg["att_src"] == 1:
        w.se_w1 = fp(P + "se_flow/bottleneck_fc/kernel")
        w.se_b1 = fp(P + "se_flow/bottleneck_fc/bias")
        w.se_w2 = fp(P + "se_flow/recover_fc/kernel")
        w.se_b2 = fp(P + "se_flow/recover_fc/bias")
    if cfg["att_src"] == 2:
        w.static_w = fp(P + "pose_exp_net/seg_channel_weight/weight")
    w.cin1 = 10 if cfg["in_mode"] == 1 else 6
    w.c6 = cfg["cnv6_out"]
    w.att_src = cfg["att_src"]
    w.mask_rgb = int(cfg["mask_mode"] != 0)
    w.mask_flow = int(cfg["mask_mode"] == 2)
    w.se_act, w.flow_abs, w.flow_norm = cfg["se_act"], cfg["flow_abs"], cfg["flow_norm"]
    B, H, W3, _ = img_u8.shape
    Wd = W3 // 3
    poses = np.zeros((B, 2, 6), np.float64)
    attw = np.zeros((B, 2, 19), np.float64)
    lib.posenn_ref_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.POINTER(_Weights), C.c_void_p, C.c_void_p]
    for b in range(B):
        im = np.ascontiguousarray(img_u8[b])
        fl = np.ascontiguousarray(flow[b], dtype=np.float32)
        sg = np.ascontiguousarray(seg[b], dtype=np.float32)
        rc = lib.posenn_ref_forward(im.ctypes.data, fl.ctypes.data, sg.ctypes.data, H, Wd, C.byref(w),
                                    poses[b].ctypes.data, attw[b].ctypes.data)
        assert rc == 0
    return poses, attw
