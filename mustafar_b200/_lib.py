"""ctypes binding of libmustafar_b200.so — the C-ABI boundary declared in include/mustafar_b200.h.

The library is hand-written CUDA for sm_100a.  There is NO fallback: if the shared object is
missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFB200_LIB", os.path.join(_HERE, "libmustafar_b200.so"))  # override: A/B builds only

ABI_VERSION = 5
LAYOUT_KEY = 0
LAYOUT_VALUE = 1
F_REF_SCORE_ROUNDING = 1
F_PDL = 2
F_PDL_EARLY_KV = 4

_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int


class DecodeParams(C.Structure):
    """Mirror of `mfb200_decode_params` (include/mustafar_b200.h)."""

    _fields_ = [
        ("batch", C.c_int32), ("kv_heads", C.c_int32), ("groups", C.c_int32), ("comp_len", C.c_int32),
        ("win_len", C.c_int32), ("flags", C.c_int32), ("score_div", C.c_float), ("n_split", C.c_int32),
        ("slot_kb", C.c_int32), ("workspace_kb", C.c_int32), ("plan_hint", C.c_int32), ("reserved0", C.c_int32),
        ("q", _vp), ("out", _vp),
        ("k_bmp", _vp), ("k_idx", _vp), ("k_nz", _vp), ("k_nz_off", _vp),
        ("v_bmp", _vp), ("v_idx", _vp), ("v_nz", _vp), ("v_nz_off", _vp),
        ("bmp_stride", _i64), ("idx_stride", _i64),
        ("k_win", _vp), ("v_win", _vp), ("win_stride", _i64),
        ("k_new", _vp), ("v_new", _vp),
        ("mask", _vp), ("mask_stride", _i64),
        ("workspace", _vp),
        ("peer", _vp),
        ("win_len_dev", _vp),
        ("rope_cos", _vp),
        ("rope_sin", _vp),
        ("rope_stride", C.c_int64),
    ]


MAX_PEERS = 8


class PeerOut(C.Structure):
    """Mirror of `mfb200_peer_out` (include/mustafar_b200.h): peer-to-peer output stores of the head-sharded decode."""

    _fields_ = [
        ("n_peers", C.c_int32), ("rank", C.c_int32), ("row0", C.c_int32), ("rows_total", C.c_int32),
        ("epoch", C.c_uint32), ("reserved", C.c_int32),
        ("out", _vp * MAX_PEERS), ("flags", _vp * MAX_PEERS),
    ]


# name -> (restype, argtypes); every symbol include/mustafar_b200.h declares
SIGNATURES = {
    "mfb200_abi_version": (_i32, []),
    "mfb200_last_error": (C.c_char_p, []),
    "mfb200_prune_rows": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "mfb200_prune_rows_scored": (_i32, [_vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    "mfb200_prune_token_groups": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "mfb200_compress_count": (_i32, [_vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp]),
    "mfb200_compress_scan": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp]),
    "mfb200_compress_pack": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    "mfb200_compress_append_chunk": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                            _i64, _i64, _i64, _i64, _vp, _vp]),
    "mfb200_compress_prefill": (_i32, [_vp, _vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _i32, _i32, _i64, _i32, _i32,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "mfb200_key_formulation": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32]),
    "mfb200_value_formulation": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32]),
    "mfb200_value_workspace_bytes": (C.c_size_t, [_i32, _i32]),
    "mfb200_decode_plan": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mfb200_decode_plan_check": (_i32, [_i32, _i32, _i32, _i32, _i32, _i32, _i32]),
    "mfb200_sparse_decode_attention": (_i32, [C.POINTER(DecodeParams), _vp]),
    "mfb200_decode_step": (_i32, [C.POINTER(DecodeParams), _vp, _vp, _vp, _vp, _i32, _vp]),
    "mfb200_decode_step_layers": (_i32, [C.POINTER(C.POINTER(DecodeParams)), _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "mfb200_decode_layers_static": (_i32, [C.POINTER(C.POINTER(DecodeParams)), _i32, _vp]),
    "mfb200_lengths_add": (_i32, [_vp, _i32, _i32, _vp]),
    "mfb200_decode_workspace_max": (C.c_size_t, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "mfb200_window_append": (_i32, [_vp, _vp, _i64, _vp, _vp, _i64, _i64, _vp]),
    "mfb200_peer_wait": (_i32, [_vp, _i32, C.c_uint32, _vp, _vp]),
    "mfb200_peer_alloc": (_i32, [C.c_size_t, C.POINTER(_vp)]),
    "mfb200_peer_free": (_i32, [_vp]),
    "mfb200_ipc_export": (_i32, [_vp, C.c_char_p]),
    "mfb200_ipc_open": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "mfb200_ipc_close": (_i32, [_vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C mustafar_b200/csrc`.  mustafar_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.mfb200_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libmustafar_b200 ABI {lib.mfb200_abi_version()} != expected {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc < 0:
        msg = load().mfb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")
    return rc


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
