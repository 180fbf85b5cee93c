"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mustafar_b200.h declares; argument validation works without touching a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mustafar_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from mustafar_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mustafar_b200.h")).read()
    declared = set(re.findall(r"\b(mfb200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 12
    raw = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    assert lib.mfb200_abi_version() == 1


def test_struct_layout_matches_header():
    from mustafar_b200 import _lib
    # 10 x int32/float, then 8-byte fields only
    assert _lib.DecodeParams.q.offset == 40
    assert C.sizeof(_lib.DecodeParams) == 40 + 20 * 8


def test_argument_validation_without_gpu(lib):
    from mustafar_b200 import _lib
    assert lib.mfb200_prune_rows(None, None, 4, 64, None) == -1
    assert b"null" in lib.mfb200_last_error()
    assert lib.mfb200_compress_count(C.c_void_p(8), 1, 65, 0, 0, C.c_void_p(8), C.c_void_p(8), None) == -1
    assert b"multiple of 64" in lib.mfb200_last_error()
    assert lib.mfb200_key_formulation(None, None, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), C.c_void_p(16),
                                      C.c_void_p(16), C.c_void_p(16), 256, 4, 128, None, 1, 2, 1) == -1
    assert b"N_Global" in lib.mfb200_last_error()
    ws, cb = C.c_size_t(0), C.c_size_t(0)
    n = lib.mfb200_decode_plan(1, 32, 1, 3840, 256, 148, C.byref(ws), C.byref(cb))
    assert n >= 2 and ws.value > cb.value > 0
    assert lib.mfb200_decode_plan(1, 32, 3, 3840, 256, 148, C.byref(ws), C.byref(cb)) == -1
    assert lib.mfb200_decode_plan(1, 32, 1, 3841, 256, 148, C.byref(ws), C.byref(cb)) == -1
    p = _lib.DecodeParams()
    assert lib.mfb200_sparse_decode_attention(C.byref(p), None) == -1


def test_no_cpu_fallback():
    import torch
    from mustafar_b200 import compression, pruning
    x = torch.zeros((1, 64, 128), dtype=torch.float16)
    with pytest.raises(RuntimeError):
        compression.convert_key_batched(x)
    with pytest.raises(RuntimeError):
        pruning.dh_prune_key(x.view(1, 1, 64, 128), 0.5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mustafar_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("numpy oracle", ""), fn
