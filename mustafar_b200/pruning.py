"""Per-token magnitude pruning on the GPU (reference: models/llama_mustafar_kernel.py:77-113, :117-153).

Same rule as the reference's `dh_prune_key` / `dh_prune_value`: k = max(1, int(sparsity * D)),
threshold = k-th smallest |x| of each token row, keep |x| >= threshold (every tie at the threshold
survives), dropped entries become x*0 = ±0.  One warp per token row, radix select on the 15-bit
magnitudes (csrc/prune_compress.cu).
"""
from __future__ import annotations

import torch

from . import _lib

HEAD_DIM = 128


def prune_rank(sparsity: float, dim: int = HEAD_DIM) -> int:
    """`max(1, int(target_sparsity * D))` — llama_mustafar_kernel.py:97 (named num_to_keep there)."""
    assert 0 <= sparsity < 1, "Target sparsity must be between 0 and 1"
    return max(1, int(sparsity * dim))


def _prune(x: torch.Tensor, sparsity: float, out: torch.Tensor | None = None) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("mustafar_b200.pruning: input must be a CUDA tensor (no CPU fallback)")
    if x.dtype != torch.float16:
        raise RuntimeError("mustafar_b200.pruning: input must be float16")
    if x.shape[-1] != HEAD_DIM:
        raise RuntimeError(f"mustafar_b200.pruning: last dim must be {HEAD_DIM}")
    xc = x.contiguous()
    y = torch.empty_like(xc) if out is None else out
    assert y.is_contiguous() and y.shape == xc.shape and y.dtype == xc.dtype
    rows = xc.numel() // HEAD_DIM
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.mfb200_prune_rows(xc.data_ptr(), y.data_ptr(), rows, prune_rank(sparsity), _lib.stream_ptr()),
                   "mfb200_prune_rows")
    return y


def dh_prune_key(key_states: torch.Tensor, target_sparsity: float) -> torch.Tensor:
    """[B, H, T, D] fp16 -> pruned copy, same shape (llama_mustafar_kernel.py:77-113)."""
    return _prune(key_states, target_sparsity)


def dh_prune_value(value_states: torch.Tensor, target_sparsity: float) -> torch.Tensor:
    """[B, H, T, D] fp16 -> pruned copy, same shape (llama_mustafar_kernel.py:117-153)."""
    return _prune(value_states, target_sparsity)
