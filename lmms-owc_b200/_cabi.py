"""ctypes binding of ``include/owc_b200.h`` (the C-ABI shared library ``libowc_b200.so``).

The library is built in-tree by ``__graft_entry__.build()`` (``nvcc -gencode
arch=compute_100a,code=sm_100a``). There is no fallback: if the library is missing or a call fails
the error propagates as :class:`OwcError`.
"""

from __future__ import annotations

import ctypes as C
import os

__all__ = ["lib", "OwcError", "OwcArch", "OwcStats", "OwcProfile", "KCATS", "check", "LIB_PATH", "PREC_BF16", "PREC_FP32",
           "MEM_HOST", "MEM_DEVICE", "EXPORTS"]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libowc_b200.so")

PREC_BF16, PREC_FP32 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1


class OwcError(RuntimeError):
    """A C-ABI call returned non-zero; the message is ``owc_last_error()``."""


class OwcArch(C.Structure):
    _fields_ = [("kind", C.c_int32), ("layers", C.c_int32), ("hidden", C.c_int32), ("heads", C.c_int32),
                ("inter", C.c_int32), ("vocab", C.c_int32), ("max_pos", C.c_int32), ("type_vocab", C.c_int32),
                ("pad_id", C.c_int32), ("ln_eps", C.c_float)]


class OwcStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("tokens_encoded", C.c_int64), ("sequences_encoded", C.c_int64)]


class OwcProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 12), ("launches", C.c_int64 * 12)]


KCATS = ("embed_ln", "gemm", "attn", "ln", "pool_l2", "sim_topk", "rescore", "misc", "gemm_qkv", "gemm_o", "gemm_ffn1",
         "gemm_ffn2")   # the last four break "gemm" down by shape on the wide path (each launch is in both)

_P = C.c_void_p
_I32, _I64 = C.c_int32, C.c_int64

ABI_VERSION = 3

# name -> (restype, argtypes); every symbol include/owc_b200.h (product ABI) and include/owc_b200_debug.h (test /
# measurement hooks) declare
EXPORTS = {
    "owc_abi_version": (_I32, []),
    "owc_last_error": (C.c_char_p, []),
    "owc_ctx_create": (_I32, [_I32, _P, _I32, C.POINTER(_P)]),
    "owc_ctx_destroy": (_I32, [_P]),
    "owc_ctx_sync": (_I32, [_P]),
    "owc_ctx_stats": (_I32, [_P, C.POINTER(OwcStats)]),
    "owc_ctx_profile_begin": (_I32, [_P]),
    "owc_ctx_profile_end": (_I32, [_P, C.POINTER(OwcProfile)]),
    "owc_encoder_load": (_I32, [_P, C.POINTER(OwcArch), _P, C.c_size_t, C.POINTER(_P)]),
    "owc_encoder_destroy": (_I32, [_P]),
    "owc_encode": (_I32, [_P, _P, _I32, _P, _I32, _I32, _P, _I32]),
    "owc_pair_dot": (_I32, [_P, _P, _I64, _P, _P, _I64, _P, _I64, _I32, _I32, _P]),
    "owc_sim_topk": (_I32, [_P, _P, _I64, _P, _I32, _I32, _I32, _I32, _I32, _P, _P]),
    "owc_score": (_I32, [_P, _P, _I32, _P, _I32, _P, _I32, _P, _I32, _I32, _P, _P, _P, _P, _I32]),
    "owc_match": (_I32, [_P, _P, _I64, _P, _I32, _I32, _P, _I32, _I32, _I32, _P, _P, _P, _I32]),
    "owc_debug_gemm": (_I32, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "owc_debug_gemm_timeline": (_I32, [_P, _I32, _I32, _I32, _I32, _P, _I32]),
    "owc_debug_gemm_loop": (_I32, [_P, _I32, _I32, _I32, _I32, _I32, C.POINTER(C.c_float)]),
    "owc_bench_sim_kernel": (_I32, [_P, _I64, _I32, _I32, _I32, _I32, C.POINTER(C.c_float)]),
    "owc_debug_fused_stamps": (_I32, [_P, _P]),
    "owc_debug_wide_counters": (_I32, [_P, _P, _I32]),
    "owc_tokenizer_last_error": (C.c_char_p, []),
    "owc_tokenizer_create": (_I32, [_P, _P, _I32, _I32, _I32, C.POINTER(_P)]),
    "owc_tokenizer_destroy": (_I32, [_P]),
    "owc_tokenize_packed": (_I32, [_P, _P, _P, _I64, _I32, _P, _I64, _P, C.POINTER(_I64)]),
    "owc_jsonl_last_error": (C.c_char_p, []),
    "owc_jsonl_count_rows": (_I32, [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "owc_jsonl_extract": (_I32, [C.c_char_p, C.c_size_t, C.c_char_p, _I32, _P, C.c_size_t, C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]),
    "owc_jsonl_extract_pairs": (_I32, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, _P, C.c_size_t,
                                       C.POINTER(C.c_size_t), _P, C.c_size_t, C.POINTER(C.c_size_t),
                                       C.POINTER(C.c_size_t)]),
    "owc_jsonl_write_columns": (_I32, [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(_I32), _P, _I32,
                                       C.c_size_t, C.c_char_p]),
    "owc_jsonl_write_columns_raw": (_I32, [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(_I32), _P,
                                           C.POINTER(C.c_char_p), _I32, C.c_size_t, C.c_char_p]),
    "owc_jsonl_extract_raw": (_I32, [C.c_char_p, C.c_size_t, C.c_char_p, _P, C.c_size_t, C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t)]),
    "owc_blob_unique": (_I32, [_P, C.c_size_t, C.c_size_t, _P, _P, C.c_size_t, C.POINTER(C.c_size_t),
                               C.POINTER(C.c_size_t)]),
    "owc_jsonl_has_members": (_I32, [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), _I32, C.POINTER(_I32)]),
    "owc_jsonl_set_columns": (_I32, [C.c_char_p, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(_I32), _P, _I32,
                                     C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OwcError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.owc_abi_version() != ABI_VERSION:
            raise OwcError("libowc_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise OwcError(lib().owc_last_error().decode("utf-8", "replace"))
