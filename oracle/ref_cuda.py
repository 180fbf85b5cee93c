"""TEST INFRASTRUCTURE: the reference's own CUDA kernels (oracle/_ref/libmustafar_ref.so, built by
oracle/Makefile from /root/reference/kernel/csrc/SpMM_API.cu for sm_100a) driven through the same
glue as models/llama_mustafar_kernel.py:268-320.  Used by the -m gpu parity tests and as the
"reference CUDA kernel" comparison line of bench.py; never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libmustafar_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(REF_LIB)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(REF_LIB)
        vp, i = C.c_void_p, C.c_int
        lib.ref_key_formulation.restype = i
        lib.ref_key_formulation.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, i]
        lib.ref_value_formulation.restype = i
        lib.ref_value_formulation.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i]
        _lib = lib
    return _lib


def key_formulation(bmp, NZ, idx, NZ_Offset, B, M_Global, K_Global, Batch_Size, groups):
    """mustafar_wrapper.cu:19-133 (torch::zeros output, bmp reinterpreted as uint64)."""
    Cout = torch.zeros((Batch_Size, 8, M_Global), dtype=torch.float16, device=B.device)
    rc = load().ref_key_formulation(torch.cuda.current_stream().cuda_stream, bmp.data_ptr(), NZ.data_ptr(),
                                    idx.data_ptr(), NZ_Offset.data_ptr(), B.data_ptr(), Cout.data_ptr(), M_Global,
                                    K_Global, Batch_Size, groups)
    assert rc == 0, f"reference Key_SplitK_API returned cudaError {rc}"
    return Cout


def value_formulation(bmp, NZ, idx, NZ_Offset, B, M_Global, K_Global, Batch_Size, groups):
    """mustafar_wrapper.cu:139-263."""
    Cout = torch.zeros((Batch_Size, 8, M_Global), dtype=torch.float16, device=B.device)
    ws = torch.zeros((8,), dtype=torch.float16, device=B.device)
    rc = load().ref_value_formulation(torch.cuda.current_stream().cuda_stream, bmp.data_ptr(), NZ.data_ptr(),
                                      idx.data_ptr(), NZ_Offset.data_ptr(), B.data_ptr(), Cout.data_ptr(),
                                      ws.data_ptr(), M_Global, K_Global, Batch_Size, groups)
    assert rc == 0, f"reference Value_SplitK_API returned cudaError {rc}"
    return Cout


def pad_nz(nz_list):
    """torch.cat(k_compressed[2]) plus one spare uint4: the reference kernel reads one uint4 past a tile whose
    nonzero count is a multiple of 8 (SpMM_Kernel.cuh:71) — keep that read inside the allocation."""
    return torch.cat(list(nz_list) + [torch.zeros(64, dtype=torch.float16, device=nz_list[0].device)])


def repeat_kv(x, n_rep):
    b, h, t, d = x.shape
    if n_rep == 1:
        return x
    return x[:, :, None].expand(b, h, n_rep, t, d).reshape(b, h * n_rep, t, d)


def decode_step(query_states, k_compressed, k_local_window, v_compressed, v_local_window, compressed_length,
                groups, key_op=key_formulation, value_op=value_formulation, attention_mask=None):
    """The reference decode glue, llama_mustafar_kernel.py:268-320, with pluggable SpMV ops.

    k_local_window / v_local_window already contain the new token (`:270`, `:309`).
    """
    b, hq, _, d = query_states.shape
    tb = b * hq
    if compressed_length != 0:
        padded_query = F.pad(query_states.reshape(tb, -1, d), (0, 0, 0, 7), mode="constant", value=0).contiguous()
        att_c = key_op(k_compressed[0], pad_nz(k_compressed[2]), k_compressed[1].reshape(-1), k_compressed[3],
                       padded_query, compressed_length, d, tb, groups)
        att_c = att_c[:, 0:1, :].reshape(b, hq, 1, compressed_length)
        att_l = torch.matmul(query_states, repeat_kv(k_local_window, groups).transpose(2, 3))
        att = torch.cat([att_c, att_l], dim=-1)
    else:
        att = torch.matmul(query_states, repeat_kv(k_local_window, groups).transpose(2, 3))
    w = att / math.sqrt(d)
    if attention_mask is not None:
        w = w + attention_mask
        w = torch.max(w, torch.tensor(torch.finfo(w.dtype).min, device=w.device))
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(query_states.dtype)
    if compressed_length != 0:
        padded_score = F.pad(w[:, :, :, :compressed_length].reshape(tb, -1, compressed_length), (0, 0, 0, 7)).contiguous()
        out_c = value_op(v_compressed[0], pad_nz(v_compressed[2]), v_compressed[1].reshape(-1), v_compressed[3],
                         padded_score, d, compressed_length, tb, groups)
        out_c = out_c[:, 0:1, :].reshape(b, hq, 1, d)
        out_l = torch.matmul(w[:, :, :, compressed_length:], repeat_kv(v_local_window, groups))
        return out_c + out_l
    return torch.matmul(w, repeat_kv(v_local_window, groups))
