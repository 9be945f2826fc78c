"""ctypes front-end of the plain-C oracle (oracle/csrc/oracle_sr.c) + forward compositions.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  numpy in, numpy out,
no torch.  Compositions cite the reference lines they follow.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_sr.so")
_lib = None
_f = ctypes.POINTER(ctypes.c_float)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "csrc", "oracle_sr.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def weight_norm(g, v) -> np.ndarray:
    g, v = _c(g).reshape(-1), _c(v)
    w = np.empty_like(v)
    lib().osr_weight_norm(_p(g), _p(v), v.shape[0], int(np.prod(v.shape[1:])), _p(w))
    return w


def conv2d(x, w, b, act: int = 0) -> np.ndarray:
    x, w = _c(x), _c(w)
    n, c, h, wd = x.shape
    o, ci, k, _ = w.shape
    assert ci == c
    y = np.empty((n, o, h, wd), np.float32)
    bb = _c(b) if b is not None else None
    lib().osr_conv2d(_p(x), n, c, h, wd, _p(w), _p(bb) if bb is not None else None, o, k, act, _p(y))
    return y


def pixel_shuffle_add(x, r: int, add: float) -> np.ndarray:
    x = _c(x)
    n, c, h, w = x.shape
    y = np.empty((n, c // (r * r), h * r, w * r), np.float32)
    lib().osr_pixel_shuffle_add(_p(x), n, c, h, w, r, ctypes.c_float(add), _p(y))
    return y


def flow_warp(x, flow, padding_mode: str = "zeros") -> np.ndarray:
    x, flow = _c(x), _c(flow)
    n, c, h, w = x.shape
    assert flow.shape == (n, h, w, 2)
    y = np.empty_like(x)
    lib().osr_flow_warp(_p(x), n, c, h, w, _p(flow), 1 if padding_mode == "border" else 0, _p(y))
    return y


def resize_bilinear(x, oh: int, ow: int, align_corners: bool) -> np.ndarray:
    x = _c(x)
    n, c, h, w = x.shape
    y = np.empty((n, c, oh, ow), np.float32)
    lib().osr_resize_bilinear(_p(x), n, c, h, w, oh, ow, 1 if align_corners else 0, _p(y))
    return y


def avg_pool2(x) -> np.ndarray:
    x = _c(x)
    n, c, h, w = x.shape
    y = np.empty((n, c, h // 2, w // 2), np.float32)
    lib().osr_avg_pool2(_p(x), n, c, h, w, _p(y))
    return y


# ---------------------------------------------------------------------------------------------
# compositions
# ---------------------------------------------------------------------------------------------
SD = Dict[str, np.ndarray]


def _wn_conv(sd: SD, p: str, x, act: int = 0):
    return conv2d(x, weight_norm(sd[p + "weight_g"], sd[p + "weight_v"]), sd[p + "bias"], act)


def wdsr_block(sd: SD, p: str, x, idx=(0, 2, 3)):
    """models/basic_wdsr_b.py:96-144."""
    t = _wn_conv(sd, f"{p}body.{idx[0]}.", x, 1)
    t = _wn_conv(sd, f"{p}body.{idx[1]}.", t)
    return _wn_conv(sd, f"{p}body.{idx[2]}.", t) + x


def basic_model_forward(sd: SD, x, scale: int, image_mean: float = 0.5):
    """models/basic_wdsr_b.py:85-93."""
    nb = len({int(k.split(".")[1]) for k in sd if k.startswith("body.")})
    x0 = _c(x) - np.float32(image_mean)
    y = _wn_conv(sd, "head.", x0)
    for b in range(nb):
        y = wdsr_block(sd, f"body.{b}.", y)
    sp = "skip.0." if "skip.0.weight_v" in sd else "skip."
    t = _wn_conv(sd, "tail.", y) + _wn_conv(sd, sp, x0)
    return pixel_shuffle_add(t, scale, image_mean)


def pruned_model_forward(sd: SD, x, scale: int, image_mean: float = 0.5):
    """export_onnx.py:59-79 (no +mean at the end)."""
    ids = sorted({int(k.split(".")[1]) for k in sd if k.startswith("body.")})
    x0 = _c(x) - np.float32(image_mean)
    y = _wn_conv(sd, f"body.{ids[0]}.", x0)
    for b in ids[1:-1]:
        y = wdsr_block(sd, f"body.{b}.", y)
    t = _wn_conv(sd, f"body.{ids[-1]}.", y) + _wn_conv(sd, "skip.", x0)
    return pixel_shuffle_add(t, scale, 0.0)


def spynet_forward(sd: SD, ref, supp, prefix: str = ""):
    """models/spynet_arch.py:49-96."""
    ref, supp = _c(ref), _c(supp)
    n, _, h, w = ref.shape
    w_up = int(math.floor(math.ceil(w / 32.0) * 32.0))
    h_up = int(math.floor(math.ceil(h / 32.0) * 32.0))
    mean, std = _c(sd[prefix + "mean"]), _c(sd[prefix + "std"])
    pyr_r = [(resize_bilinear(ref, h_up, w_up, False) - mean) / std]
    pyr_s = [(resize_bilinear(supp, h_up, w_up, False) - mean) / std]
    for _ in range(5):
        pyr_r.insert(0, avg_pool2(pyr_r[0]))
        pyr_s.insert(0, avg_pool2(pyr_s[0]))
    flow = np.zeros((n, 2, pyr_r[0].shape[2] // 2, pyr_r[0].shape[3] // 2), np.float32)
    for lv in range(6):
        hh, ww = pyr_r[lv].shape[2:]
        up = resize_bilinear(flow, flow.shape[2] * 2, flow.shape[3] * 2, True) * np.float32(2.0)
        if up.shape[2] != hh:
            up = np.concatenate([up, up[:, :, -1:, :]], 2)
        if up.shape[3] != ww:
            up = np.concatenate([up, up[:, :, :, -1:]], 3)
        warped = flow_warp(pyr_s[lv], np.ascontiguousarray(up.transpose(0, 2, 3, 1)), "border")
        t = np.concatenate([pyr_r[lv], warped, up], 1)
        for j, i in enumerate((0, 2, 4, 6, 8)):
            p = f"{prefix}basic_module.{lv}.basic_module.{i}."
            t = conv2d(t, sd[p + "weight"], sd[p + "bias"], 1 if j < 4 else 0)
        flow = t + up
    out = resize_bilinear(flow, h, w, False)
    out[:, 0] *= np.float32(float(w) / float(w_up))
    out[:, 1] *= np.float32(float(h) / float(h_up))
    return out
