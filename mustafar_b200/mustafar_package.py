"""Drop-in for the reference's `mustafar_package` torch extension (kernel/kernel_wrapper/pybind.cpp:7-11).

    mustafar_key_formulation(bmp, NZ, idx, NZ_Offset, B, M_Global, K_Global, Batch_Size, num_key_value_groups)
        -> fp16 [Batch_Size, 8, M_Global]                     (mustafar_wrapper.cu:19-133)
    mustafar_value_formulation(bmp, NZ, idx, NZ_Offset, B, Reduction_Workspace, M_Global, K_Global,
                               Batch_Size, num_key_value_groups) -> fp16 [Batch_Size, 8, M_Global]   (:139-263)

Argument checking mirrors the wrapper (RuntimeError on device/dtype mismatch, contiguity required);
additionally CUDA launch errors are raised instead of being dropped (mustafar_wrapper.cu:113).  The
bitmap tensor is used in place — no `bmp.to(uint64)` copy per call (mustafar_wrapper.cu:90).
"""
from __future__ import annotations

import torch

from . import _lib

_value_ws = {}


def _check(bmp, NZ, idx, NZ_Offset, B, need_b_contig: bool):
    for t in (bmp, NZ, idx, NZ_Offset):
        if t.device != B.device:
            raise RuntimeError("All input tensors must be on the same device.")
    if B.dtype != torch.float16:
        raise RuntimeError("Tensor B must be of type float16.")
    if NZ.dtype != torch.float16:
        raise RuntimeError("Tensor NZ must be of type float16.")
    if bmp.dtype != torch.int64:
        raise RuntimeError("Tensor bmp must be of type int64.")
    if idx.dtype != torch.int32:
        raise RuntimeError("Tensor idx must be of type int.")
    if NZ_Offset.dtype != torch.int32:
        raise RuntimeError("Tensor NZ_Offset must be of type int.")
    contig = bmp.is_contiguous() and NZ.is_contiguous() and idx.is_contiguous() and NZ_Offset.is_contiguous()
    if need_b_contig:
        contig = contig and B.is_contiguous()
    if not contig:
        raise RuntimeError("bmp, NZ, idx, B, C, and Reduction_Workspace tensors must be contiguous.")
    if not (bmp.is_cuda and NZ.is_cuda and idx.is_cuda and B.is_cuda and NZ_Offset.is_cuda):
        raise RuntimeError("bmp, NZ, idx, B, C, and (not)Reduction_Workspace tensors must be on CUDA device.")


def mustafar_key_formulation(bmp, NZ, idx, NZ_Offset, B, M_Global: int, K_Global: int, Batch_Size: int,
                             num_key_value_groups: int) -> torch.Tensor:
    _check(bmp, NZ, idx, NZ_Offset, B, need_b_contig=True)
    lib = _lib.load()
    with torch.cuda.device(B.device):
        C = torch.empty((Batch_Size, 8, M_Global), dtype=torch.float16, device=B.device)
        _lib.check(lib.mfb200_key_formulation(_lib.stream_ptr(), None, bmp.data_ptr(), NZ.data_ptr(), idx.data_ptr(),
                                              NZ_Offset.data_ptr(), B.data_ptr(), C.data_ptr(), M_Global, 8, K_Global,
                                              None, 1, Batch_Size, num_key_value_groups), "mfb200_key_formulation")
    return C


def mustafar_value_formulation(bmp, NZ, idx, NZ_Offset, B, Reduction_Workspace, M_Global: int, K_Global: int,
                               Batch_Size: int, num_key_value_groups: int) -> torch.Tensor:
    # Reduction_Workspace: a 1-element dummy in the reference (llama_mustafar_kernel.py:658); kept for
    # signature compatibility and ignored.  The split-sequence merge uses a private zeroed workspace.
    _check(bmp, NZ, idx, NZ_Offset, B, need_b_contig=False)
    lib = _lib.load()
    Bc = B.contiguous()
    with torch.cuda.device(B.device):
        nbytes = lib.mfb200_value_workspace_bytes(K_Global, Batch_Size)
        key = (B.device.index, torch.cuda.current_stream().cuda_stream)
        ws = _value_ws.get(key)
        if ws is None or ws.numel() < nbytes:
            if len(_value_ws) >= 16:  # bounded: drop the workspaces of streams that are no longer in use
                _value_ws.clear()
            ws = torch.zeros((nbytes,), dtype=torch.uint8, device=B.device)
            _value_ws[key] = ws
        C = torch.empty((Batch_Size, 8, M_Global), dtype=torch.float16, device=B.device)
        _lib.check(lib.mfb200_value_formulation(_lib.stream_ptr(), None, bmp.data_ptr(), NZ.data_ptr(), idx.data_ptr(),
                                                NZ_Offset.data_ptr(), Bc.data_ptr(), C.data_ptr(), M_Global, 8,
                                                K_Global, ws.data_ptr(), 1, Batch_Size, num_key_value_groups),
                   "mfb200_value_formulation")
    return C
