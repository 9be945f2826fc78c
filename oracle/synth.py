"""Deterministic synthetic weights / inputs for the parity tests.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

numpy's legacy MT19937 ``RandomState`` stream is stable across numpy versions
and machines, so a fixture only needs to record (name, shape, seed) -- the
golden generator (which has the reference) and the tests (which do not)
rebuild bit-identical tensors.  Each tensor gets its own stream keyed on its
state_dict name, so values do not depend on iteration order.

Unlike the reference's own init (zero biases, constant weight_g) these
weights have non-zero biases and varied gains: the zero-bias init hides the
"1x1s evaluated on a zero-padded trunk halo" border bug.
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Sequence

import numpy as np


def _rs(name: str, seed: int) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31 - 1))


def synth_tensor(name: str, shape: Sequence[int], seed: int = 0) -> np.ndarray:
    rs = _rs(name, seed)
    shape = tuple(int(s) for s in shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "weight_g":
        if ".body.3." in name or ".body.5." in name:      # block 3x3: reference uses res_scale gains
            return rs.uniform(0.15, 0.45, shape).astype(np.float32)
        return rs.uniform(0.6, 1.8, shape).astype(np.float32)
    if leaf == "bias":
        return rs.uniform(-0.1, 0.1, shape).astype(np.float32)
    if leaf in ("alpha1", "alpha2", "beta1", "beta2", "alpha", "beta"):
        return rs.uniform(0.0, 1.0, shape).astype(np.float32)
    if leaf == "mean":
        return np.asarray([0.485, 0.456, 0.406], np.float32).reshape(shape)
    if leaf == "std":
        return np.asarray([0.229, 0.224, 0.225], np.float32).reshape(shape)
    if leaf == "weight" and len(shape) == 4 and shape[1] == 1 and shape[2] == 1 and shape[3] == 1:
        return rs.uniform(0.2, 1.0, shape).astype(np.float32)      # BinaryConv2d mask
    if leaf in ("weight_v", "weight"):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        bound = (1.0 / fan_in) ** 0.5                                # kaiming_uniform(a=sqrt(5)) bound
        if leaf == "weight":
            bound *= 1.7                                             # plain convs: keep activations O(1) through depth
        return rs.uniform(-bound, bound, shape).astype(np.float32)
    return rs.uniform(-0.5, 0.5, shape).astype(np.float32)


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0) -> Dict[str, np.ndarray]:
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_input(shape: Sequence[int], seed: int = 1234, lo: float = 0.0, hi: float = 1.0) -> np.ndarray:
    return np.random.RandomState(seed).uniform(lo, hi, tuple(shape)).astype(np.float32)


def synth_mv_clip(shape: Sequence[int], seed: int, mv_range: float = 8.0) -> np.ndarray:
    """(b, n, 5, h, w) MotionVectorVSR input: RGB in [0,1) in channels 0:3, motion vectors in pixels (+-mv_range/2) in 3:5."""
    x = synth_input(shape, seed)
    x[:, :, 3:] = (x[:, :, 3:] - 0.5) * mv_range
    return x
