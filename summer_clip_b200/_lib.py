"""ctypes binding of libsummerclip_b200.so (the C ABI declared in include/summer_clip_b200.h).

There is deliberately NO fallback: if the shared library is missing or a symbol is absent the
import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p
from pathlib import Path

SC_F16, SC_BF16, SC_F32, SC_E4M3 = 0, 1, 2, 3
SC_E4M3_SCALE = 256.0
SC_CONF_RAW, SC_CONF_PROB = 0, 1
SC_VALUES_HARD, SC_VALUES_SOFTMAX = 0, 1

_LIB_PATH = Path(os.environ.get("SUMMER_CLIP_B200_LIB", Path(__file__).resolve().parent / "lib" / "libsummerclip_b200.so"))

# name -> (restype, argtypes); mirrors include/summer_clip_b200.h one to one
SIGNATURES = {
    "sc_version": (c_int, []),
    "sc_last_error": (c_char_p, []),
    "sc_pad_dim": (c_int64, [c_int64]),
    "sc_pad_dim_op": (c_int64, [c_int64, c_int]),
    "sc_pad_keys": (c_int64, [c_int64]),
    "sc_pad_queries": (c_int64, [c_int64]),
    "sc_pad_classes": (c_int64, [c_int64]),
    "sc_class_slice": (c_int64, [c_int64]),
    "sc_normalize_cast": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                  c_void_p, c_int, c_int64, c_int, c_void_p]),
    "sc_mean_normalize_rows": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                       c_void_p]),
    "sc_rowconf": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_float, c_int, c_void_p, c_void_p,
                           c_void_p]),
    "sc_topk_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "sc_topk_per_class": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "sc_values_prepare": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int,
                                  c_float, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p]),
    "sc_attn_splits": (c_int, [c_int64, c_int64, c_int64, c_int]),
    "sc_attn_splits_for": (c_int, [c_int64, c_int64, c_int64, c_int64, c_int]),
    "sc_attn_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                            c_float, c_int, c_void_p, c_int64, c_void_p]),
    "sc_attn_fwd_shifted": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64,
                                    c_int64, c_float, c_void_p, c_int, c_void_p, c_int64, c_void_p]),
    "sc_attn_rowmax": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "sc_attn_softmax_hard": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                     c_float, c_int, c_void_p, c_int64, c_void_p]),
    "sc_softmax_partials": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_void_p]),
    "sc_merge_softmax": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64,
                                 c_float, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "sc_pad_labels": (c_int64, [c_int64]),
    "sc_hard_labels": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                               c_int64, c_void_p]),
    "sc_hard_bank_capacity": (c_int64, [c_int64, c_int32]),
    "sc_hard_bank_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "sc_hard_bank_layout": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "sc_gather_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "sc_hard_bank_inverse": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "sc_normalize_scatter": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                     c_void_p, c_int, c_int64, c_int, c_void_p]),
    "sc_transpose_norms": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int,
                                   c_int64, c_int, c_void_p, c_void_p]),
    "sc_attn_hard_supported": (c_int, [c_int64]),
    "sc_attn_hard_splits": (c_int, [c_int64, c_int64, c_int]),
    "sc_attn_hard_splits_for": (c_int, [c_int64, c_int64, c_int64, c_int, c_int64, c_int, c_int]),
    "sc_attn_fwd_hard": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                 c_float, c_int, c_void_p, c_int64, c_void_p]),
    "sc_attn_fwd_hard_multi": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64,
                                       POINTER(c_float), c_int, c_int, c_void_p, c_int64, c_void_p]),
    "sc_merge_peer_parts": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "sc_merge_partials": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "sc_zero_shot_logits": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int, c_int64,
                                    c_int64, c_float, c_int, c_void_p, c_int64, c_void_p]),
    "sc_normalize_split": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                                   c_int, c_void_p]),
    "sc_gemm_split_nt": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                                 c_void_p, c_int64, c_void_p]),
    "sc_gemm_rows_nt": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                                c_void_p, c_int64, c_void_p]),
    "sc_rowconf_from_split": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_float,
                                      c_int, c_void_p, c_void_p, c_void_p]),
    "sc_rowconf_from_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float, c_float,
                                     c_int, c_void_p, c_void_p, c_void_p]),
    "sc_epilogue": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, POINTER(c_float),
                            c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "sc_epilogue_parts": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int64, c_void_p, c_int64, c_int64,
                                  POINTER(c_float), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}


class SummerClipError(RuntimeError):
    pass


ABI_VERSION = 3          # SC_ABI_VERSION of include/summer_clip_b200.h these signatures were written against


_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library once; raise if it (or any declared symbol) is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise SummerClipError(
            f"{_LIB_PATH} not found: build it with `python -m summer_clip_b200.build` "
            "(there is no CPU or PyTorch fallback for the CLIP-search path)")
    lib = ctypes.CDLL(str(_LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise SummerClipError(f"{_LIB_PATH} does not export {name}") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.sc_version() != ABI_VERSION:
        raise SummerClipError(f"{_LIB_PATH} implements ABI version {lib.sc_version()}, these bindings expect {ABI_VERSION}: "
                              "rebuild it with `python -m summer_clip_b200.build`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sc_last_error()
        raise SummerClipError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
