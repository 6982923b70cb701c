"""ctypes binding of libpna_cuda.so (include/pna_cuda.h).  This is the stub a reference-side binding
would mirror (see INTEGRATION.md for the Rust `extern "C"` block).  There is no CPU fallback: if the
shared library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpna_cuda.so")

OK, E_INVALID_DATA, E_UNEXPECTED_EOF, E_INVALID_INPUT, E_UNSUPPORTED, E_NOSPACE, E_OOM, E_INTERNAL, E_CUDA, E_BAD_ARG = range(10)
UINT64_MAX = (1 << 64) - 1

EXPORTS = [
    "pna_cuda_init", "pna_cuda_device_count", "pna_cuda_device_id", "pna_cuda_decode_size_bound", "pna_cuda_size_hint_trusted",
    "pna_cuda_decode_plan_create_in_image", "pna_cuda_transfer_probe", "pna_cuda_destroy", "pna_cuda_strerror", "pna_cuda_last_error", "pna_cuda_host_alloc",
    "pna_cuda_host_free", "pna_cuda_stream", "pna_cuda_launch_count", "pna_cuda_crc32", "pna_cuda_crc32_image",
    "pna_cuda_decode_batch", "pna_cuda_decode_plan_create", "pna_cuda_decode_plan_create_crc", "pna_cuda_plan_crc_results", "pna_cuda_decode_plan_run", "pna_cuda_decode_plan_fetch",
    "pna_cuda_decode_plan_lengths", "pna_cuda_decode_plan_crc32_out", "pna_cuda_decode_plan_fetch_ranges",
    "pna_cuda_plan_stats", "pna_cuda_plan_counts", "pna_cuda_plan_stage_ms", "pna_cuda_stage_name", "pna_cuda_encode_stage_name", "pna_cuda_plan_destroy", "pna_cuda_encode_bound", "pna_cuda_encode_crc_count",
    "pna_cuda_encode_batch", "pna_cuda_encode_plan_create", "pna_cuda_encode_plan_run", "pna_cuda_encode_plan_lengths", "pna_cuda_encode_plan_fetch",
    "pna_cuda_encode_plan_fetch_region",
    "pna_cuda_ecb", "pna_cuda_gcm_stream_key", "pna_cuda_gcm_stream_header",
]


class Span(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("len", C.c_uint64)]


class Buf(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("cap", C.c_uint64), ("len", C.c_uint64)]


class DecodeDesc(C.Structure):
    _fields_ = [("bodies", C.POINTER(Span)), ("n_bodies", C.c_uint32), ("compression", C.c_uint8),
                ("encryption", C.c_uint8), ("cipher_mode", C.c_uint8), ("_pad", C.c_uint8), ("key", C.c_uint8 * 32),
                ("raw_size_hint", C.c_uint64)]


class EncodeDesc(C.Structure):
    _fields_ = [("plain", Span), ("compression", C.c_uint8), ("encryption", C.c_uint8), ("cipher_mode", C.c_uint8),
                ("_pad", C.c_uint8), ("level", C.c_int32), ("key", C.c_uint8 * 32), ("iv", C.c_uint8 * 16),
                ("max_chunk_size", C.c_uint32), ("stream_header", C.c_void_p)]


class PnaCudaError(RuntimeError):
    def __init__(self, code: int, what: str = ""):
        super().__init__(f"libpna_cuda: {what} (code {code})")
        self.code = code


_lib = None


def lib():
    """Load libpna_cuda.so or fail loudly -- the product path never falls back to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PnaCudaError(E_CUDA, f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32p = C.c_void_p, C.c_uint32, C.c_uint64, C.POINTER(C.c_int32)
    L.pna_cuda_init.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.pna_cuda_transfer_probe.argtypes = [vp, vp, u64, vp, u64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.pna_cuda_device_count.argtypes = [vp]
    L.pna_cuda_device_id.argtypes = [vp, C.c_int]
    L.pna_cuda_decode_size_bound.argtypes = [C.c_uint8, u64]
    L.pna_cuda_decode_size_bound.restype = u64
    L.pna_cuda_size_hint_trusted.argtypes = [C.c_uint8, u64, u64]
    L.pna_cuda_decode_plan_create_in_image.argtypes = [vp, C.POINTER(DecodeDesc), u32, vp, u64, C.POINTER(Span), C.POINTER(u32), i32p, u32, C.POINTER(vp)]
    L.pna_cuda_destroy.argtypes = [vp]
    L.pna_cuda_destroy.restype = None
    L.pna_cuda_strerror.argtypes = [C.c_int32]
    L.pna_cuda_strerror.restype = C.c_char_p
    L.pna_cuda_last_error.argtypes = [vp]
    L.pna_cuda_last_error.restype = C.c_char_p
    L.pna_cuda_host_alloc.argtypes = [vp, u64]
    L.pna_cuda_host_alloc.restype = vp
    L.pna_cuda_host_free.argtypes = [vp, vp]
    L.pna_cuda_host_free.restype = None
    L.pna_cuda_stream.argtypes = [vp]
    L.pna_cuda_stream.restype = vp
    L.pna_cuda_launch_count.argtypes = [vp]
    L.pna_cuda_launch_count.restype = u64
    L.pna_cuda_crc32.argtypes = [vp, C.POINTER(Span), u32, C.POINTER(u32)]
    L.pna_cuda_crc32_image.argtypes = [vp, vp, u64, C.POINTER(u64), C.POINTER(u64), u32, C.POINTER(u32)]
    L.pna_cuda_decode_batch.argtypes = [vp, C.POINTER(DecodeDesc), u32, C.POINTER(Buf), i32p]
    L.pna_cuda_decode_plan_create.argtypes = [vp, C.POINTER(DecodeDesc), u32, C.POINTER(vp)]
    L.pna_cuda_decode_plan_create_crc.argtypes = [vp, C.POINTER(DecodeDesc), u32, C.POINTER(Span), C.POINTER(u32), i32p, u32, C.POINTER(vp)]
    L.pna_cuda_plan_crc_results.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.pna_cuda_decode_plan_run.argtypes = [vp]
    L.pna_cuda_decode_plan_fetch.argtypes = [vp, C.POINTER(Buf), i32p]
    L.pna_cuda_decode_plan_lengths.argtypes = [vp, C.POINTER(u64), i32p]
    L.pna_cuda_plan_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.pna_cuda_encode_stage_name.restype = C.c_char_p
    L.pna_cuda_encode_stage_name.argtypes = [u32]
    L.pna_cuda_plan_counts.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.pna_cuda_plan_stage_ms.argtypes = [vp, C.POINTER(C.c_float), u32]
    L.pna_cuda_stage_name.argtypes = [u32]
    L.pna_cuda_stage_name.restype = C.c_char_p
    L.pna_cuda_plan_destroy.argtypes = [vp]
    L.pna_cuda_plan_destroy.restype = None
    L.pna_cuda_encode_bound.argtypes = [C.POINTER(EncodeDesc)]
    L.pna_cuda_encode_bound.restype = u64
    L.pna_cuda_encode_crc_count.argtypes = [C.POINTER(EncodeDesc)]
    L.pna_cuda_encode_crc_count.restype = u64
    L.pna_cuda_encode_batch.argtypes = [vp, C.POINTER(EncodeDesc), u32, C.POINTER(Buf), C.POINTER(u32), C.POINTER(u32), i32p]
    L.pna_cuda_encode_plan_create.argtypes = [vp, C.POINTER(EncodeDesc), u32, C.POINTER(vp)]
    L.pna_cuda_encode_plan_run.argtypes = [vp]
    L.pna_cuda_encode_plan_lengths.argtypes = [vp, C.POINTER(u64), i32p]
    L.pna_cuda_encode_plan_fetch.argtypes = [vp, C.POINTER(Buf), C.POINTER(u32), C.POINTER(u32), i32p]
    L.pna_cuda_encode_plan_fetch_region.argtypes = [vp, C.POINTER(Buf), vp, u64, C.POINTER(u32), C.POINTER(u32), i32p]
    L.pna_cuda_ecb.argtypes = [vp, C.c_int, C.c_int, C.c_char_p, vp, u64, vp]
    L.pna_cuda_gcm_stream_key.argtypes = [C.c_char_p, C.c_char_p, u64, C.c_char_p, C.c_char_p, u64, C.c_char_p, u64, C.c_char_p]
    L.pna_cuda_gcm_stream_key.restype = C.c_int32
    L.pna_cuda_gcm_stream_header.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, u32, C.c_char_p]
    L.pna_cuda_gcm_stream_header.restype = C.c_int32
    _lib = L
    return L
