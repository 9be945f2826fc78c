"""ctypes binding of the C ABI in include/b200sr.h (the only bridge between Python and the CUDA kernels).

The library is built in-tree (``mobilesuperresolution_b200/_C/libb200sr.so``) by ``build.py``.  There is no CPU
fallback anywhere in this package: if the library is missing, or no CUDA device is visible, compute calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch  # noqa: F401  (loads libcudart.so.12 first so the library binds to the same runtime instance)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libb200sr.so")

F32, BF16, U8 = 0, 1, 2
PAD_ZEROS, PAD_BORDER = 0, 1


class WdsrDesc(Structure):
    _fields_ = [("scale", c_int32), ("num_blocks", c_int32), ("c_trunk", c_int32), ("add_mean", c_int32),
                ("image_mean", c_float), ("m1", POINTER(c_int32)), ("m2", POINTER(c_int32))]


# every symbol include/b200sr.h declares: name -> (restype, argtypes)
_FP = POINTER(c_float)
SYMBOLS = {
    "b200sr_version": (c_int, []),
    "b200sr_last_error": (c_char_p, []),
    "b200sr_device_count": (c_int, []),
    "b200sr_wdsr_create": (c_int, [POINTER(WdsrDesc), POINTER(c_void_p)]),
    "b200sr_wdsr_destroy": (None, [c_void_p]),
    "b200sr_wdsr_set_head": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b200sr_wdsr_set_block": (c_int, [c_void_p, c_int] + [c_void_p] * 6),
    "b200sr_wdsr_set_tail": (c_int, [c_void_p] + [c_void_p] * 4),
    "b200sr_wdsr_commit": (c_int, [c_void_p]),
    "b200sr_wdsr_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "b200sr_wdsr_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_size_t, c_void_p]),
    "b200sr_wdsr_forward_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200sr_wdsr_trunk_channels": (c_int, [c_void_p]),
    "b200sr_wdsr_trunk_layout": (c_int, [c_void_p, c_int]),
    "b200sr_wdsr_head": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_wdsr_block": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_wdsr_tail": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_wdsr_launches_per_forward": (c_int, [c_void_p]),
    "b200sr_wdsr_pack_block_image": (c_int, [c_int, c_int, c_int] + [c_void_p] * 6 + [c_void_p, c_size_t, POINTER(c_size_t)]),
    "b200sr_flow_warp_nchw": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "b200sr_flow_warp_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_split_create": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]),
    "b200sr_split_destroy": (None, [c_void_p]),
    "b200sr_split_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_split_set_premask": (c_int, [c_void_p, c_void_p]),
    "b200sr_nas_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "b200sr_nas_forward": (c_int, [c_void_p, POINTER(c_void_p), c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_size_t, c_void_p]),
    "b200sr_flow_warp_nhwc_into": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_flow_warp_nhwc_windows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_vsr_trunk_forward_into": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                              c_void_p]),
    "b200sr_conv_create": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_void_p)]),
    "b200sr_conv_destroy": (None, [c_void_p]),
    "b200sr_conv_forward_layout": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                           c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_conv_tcgen05_ok": (c_int, [c_void_p]),
    "b200sr_vsr_trunk_forward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200sr_vsr_conv_last_base": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                          c_void_p]),
    "b200sr_conv_set_max_ctas": (c_int, [c_void_p, c_int]),
    "b200sr_conv_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_resize_bilinear_nchw": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p]),
    "b200sr_vsr_deconv_tail": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_spynet_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "b200sr_spynet_forward": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_size_t, c_void_p]),
    "b200sr_u8_to_unit": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "b200sr_ssd_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_avg_pool2_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_spynet_level_input": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_void_p]),
    "b200sr_nhwc_plus_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_zero_async": (c_int, [c_void_p, c_size_t, c_void_p]),
    "b200sr_pad_bottom_right_async": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_nchw3_to_nhwc": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "b200sr_vsr_base_add": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "b200sr_vsr_shuffle4_base_add": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """The loaded library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m mobilesuperresolution_b200.build` "
                               "(there is no CPU / PyTorch fallback for this path)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)          # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


class B200srError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise B200srError(f"b200sr error {rc}: {lib().b200sr_last_error().decode()}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.uint8:
        return U8          # output frames of the WDSR forward only (b200sr.h B200SR_U8)
    raise TypeError(f"b200sr: unsupported tensor dtype {dt} (float32 or bfloat16)")


def precision_code(p: str) -> int:
    try:
        return {"fp32": F32, "float32": F32, "bf16": BF16, "bfloat16": BF16}[p]
    except KeyError:
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {p!r}") from None


def require_cuda_tensor(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"b200sr: {what} must be a CUDA tensor -- this package has no CPU fallback "
                           f"(got device {t.device})")


def current_stream_ptr(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
