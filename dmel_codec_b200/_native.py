"""ctypes binding of ``libdmel_b200.so`` (C ABI in ``include/dmel_b200.h``).

The library is built in-tree by ``dmel_codec_b200.build`` (nvcc, sm_100a).  If
it is missing, or there is no CUDA device, every entry point raises — there is
no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int32, c_longlong, c_uint8, c_ulonglong,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DMEL_LIB") or os.path.join(_HERE, "libdmel_b200.so")  # DMEL_LIB: A/B builds

ABI_VERSION = 9
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE = -1, -2, -3, -4

class DmelIO(ctypes.Structure):
    """``dmel_io`` of include/dmel_b200.h, field for field."""
    _fields_ = [
        ("struct_size", ctypes.c_size_t),
        ("wav_dev", c_void_p), ("wav_is_pcm16", c_int),
        ("n_rows", c_longlong), ("n_samples", c_longlong), ("row_stride", c_longlong),
        ("offsets_dev", c_void_p), ("lengths_dev", c_void_p), ("own_length", c_int), ("min_row_samples", c_longlong),
        ("row_gain_dev", c_void_p),
        ("lo_dev", c_void_p), ("scale_dev", c_void_p), ("step_dev", c_void_p), ("n_bins", c_int),
        ("codes_dev", c_void_p), ("mel_hat_dev", c_void_p), ("logmel_dev", c_void_p), ("logmel_is_bf16", c_int),
        ("mask_invalid", c_int), ("min_dev", c_void_p), ("max_dev", c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/dmel_b200.h one to one
SIGNATURES = {
    "dmel_abi_version": (c_int, []),
    "dmel_last_error": (c_char_p, []),
    "dmel_plan_create": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, POINTER(c_void_p)]),
    "dmel_plan_destroy": (None, [c_void_p]),
    "dmel_plan_describe": (c_int, [c_void_p, c_char_p, ctypes.c_size_t]),
    "dmel_plan_num_frames": (c_longlong, [c_void_p, c_longlong]),
    "dmel_logmel_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p, c_void_p]),
    "dmel_logmel_masked": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p, c_int,
                                   c_void_p, c_void_p, c_void_p]),
    "dmel_minmax_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "dmel_logmel_minmax_f32": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "dmel_encode_u8": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "dmel_encode_decode_u8": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dmel_encode_pcm16_u8": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                                     c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dmel_encode_frames_u8": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_longlong,
                                      c_longlong, c_longlong, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                      c_void_p]),
    "dmel_stream_create": (c_int, [c_void_p, c_int, c_longlong, POINTER(c_void_p)]),
    "dmel_stream_destroy": (None, [c_void_p]),
    "dmel_stream_reset": (c_int, [c_void_p]),
    "dmel_stream_pending": (c_longlong, [c_void_p, c_longlong, c_int]),
    "dmel_stream_push": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_void_p, c_void_p, c_int, c_void_p,
                                 c_longlong, POINTER(c_longlong), c_void_p]),
    "dmel_stream_flush": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_longlong, POINTER(c_longlong),
                                  c_void_p]),
    "dmel_stream_bind": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "dmel_stream_input": (c_int, [c_void_p, c_longlong, POINTER(c_void_p), POINTER(c_longlong)]),
    "dmel_stream_commit": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, POINTER(c_longlong)]),
    "dmel_encode_host_u8": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                                    c_void_p, c_void_p, c_int, c_void_p]),
    "dmel_encode_host_pcm16_u8": (c_int, [c_void_p, c_void_p, c_longlong, c_longlong, c_longlong, c_void_p,
                                          c_void_p, c_void_p, c_int, c_void_p]),
    "dmel_run": (c_int, [c_void_p, POINTER(DmelIO), c_void_p]),
    "dmel_row_peak_gain_f32": (c_int, [c_void_p, c_longlong, c_longlong, c_longlong, c_void_p, c_void_p, c_float,
                                       c_void_p, c_void_p]),
    "dmel_fsq_encode": (c_int, [c_void_p, c_longlong, c_longlong, c_int, POINTER(c_int), c_int, c_void_p, c_void_p, c_void_p,
                                c_int, c_void_p]),
    "dmel_fsq_decode": (c_int, [c_void_p, c_longlong, c_longlong, c_int, POINTER(c_int), c_int, c_void_p, c_void_p]),
    "dmel_antialias_snake_f32": (c_int, [c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "dmel_quantize_u8": (c_int, [c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_void_p, c_int,
                                 c_void_p, c_void_p]),
    "dmel_quantize_masked_u8": (c_int, [c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_void_p, c_void_p, c_int,
                                        c_void_p, c_void_p]),
    "dmel_quantizer_derive_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dmel_dequantize_f32": (c_int, [c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_int, c_void_p,
                                    c_void_p]),
    "dmel_tensor_minmax_f32": (c_int, [c_void_p, c_longlong, c_int, c_longlong, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
}

_lib = None


class DmelNativeError(RuntimeError):
    """Raised when the CUDA library is missing or a call into it fails."""


def load() -> ctypes.CDLL:
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DmelNativeError(
            f"{LIB_PATH} not found. Build it with `python -m dmel_codec_b200.build` "
            "(needs nvcc; targets sm_100a). dmel_codec_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.dmel_abi_version()
    if got != ABI_VERSION:
        raise DmelNativeError(f"libdmel_b200.so has ABI version {got}, the Python binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int) -> None:
    """Turn a negative status into the exception the reference's torch ops
    would have raised: bad shapes -> ValueError, everything else RuntimeError."""
    if rc == 0:
        return
    msg = load().dmel_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise DmelNativeError(f"dmel_b200 error {rc}: {msg}")
