"""Per-token magnitude pruning on the GPU (reference: models/llama_mustafar_kernel.py:77-113, :117-153).

Same rule as the reference's `dh_prune_key` / `dh_prune_value`: k = max(1, int(sparsity * D)),
threshold = k-th smallest |x| of each token row, keep |x| >= threshold (every tie at the threshold
survives), dropped entries become x*0 = ±0.  One warp per token row, radix select on the 15-bit
magnitudes (csrc/prune_compress.cu).

The reference's other two policies only change the mask, the compressed format and the kernels stay the same
(SURVEY.md §8(f) rank 4); they are here as well:
  * output-aware key pruning, `dh_prune_key_output_aware` (models/llama_mustafar_Kt_Opa_Vt_Mag.py:65-178);
  * channel-wise value pruning, `dh_prune_value_channelwise` (models/llama_mustafar_Kt_Mag_Vc_Mag.py:107-170).
A cache fed by them is created with sparsity 0 (its own per-token prune is then the identity) and a nonzero slab sized for
the policy's density (`nz_halves_per_token`).
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


def fold_queries(query_states: torch.Tensor, num_key_value_groups: int, group_size: int | None = None) -> torch.Tensor:
    """The per-KV-head channel weights of the output-aware policy: [B, Hq, T, 128] -> [B, Hkv, 128] fp16.

    Prefill (llama_mustafar_Kt_Opa_Vt_Mag.py:98-101): mean over the last `group_size` tokens of |q|, summed over the query
    heads of the KV head; decode (`:131-133`, T == 1): |q| summed over the query heads.  Plain torch ops in the tensors'
    dtype, exactly as the reference writes them (the arrays are tiny)."""
    b, hq, t, d = query_states.shape
    q = query_states if group_size is None else query_states[:, :, -group_size:, :]
    folded = torch.mean(torch.abs(q), dim=-2)  # [B, Hq, D]  (T == 1: the mean of one row is the row)
    return folded.view(b, hq // num_key_value_groups, num_key_value_groups, d).sum(dim=-2).contiguous()


def dh_prune_key_output_aware(key_states: torch.Tensor, query_states: torch.Tensor, target_sparsity: float,
                              num_key_value_groups: int = 1, group_size: int = 32, keep_last: bool = True) -> torch.Tensor:
    """Output-aware key pruning of a prompt (llama_mustafar_Kt_Opa_Vt_Mag.py:90-114): every token row keeps its
    int(128 * (1 - sparsity)) entries with the highest |q_folded * k|; the last `group_size` tokens stay dense (`:110`).
    key_states [B, Hkv, T, 128], query_states [B, Hq, Tq, 128] fp16 CUDA -> pruned copy of key_states.
    On a tie at the cut the reference keeps an arbitrary subset of the tied entries, this keeps all of them."""
    if not key_states.is_cuda or key_states.dtype != torch.float16 or query_states.dtype != torch.float16:
        raise RuntimeError("mustafar_b200.pruning: float16 CUDA tensors expected (no CPU fallback)")
    b, h, t, d = key_states.shape
    if d != HEAD_DIM or query_states.shape[0] != b or query_states.shape[1] != h * num_key_value_groups:
        raise RuntimeError("mustafar_b200.pruning: key [B,Hkv,T,128] and query [B,Hkv*G,Tq,128] expected")
    n_keep = int(d * (1 - target_sparsity))
    if not 1 <= n_keep <= d:
        raise ValueError("target_sparsity leaves no entry")
    w = fold_queries(query_states, num_key_value_groups, group_size)
    return prune_rows_scored(key_states, w, n_keep, keep_last=group_size if keep_last else 0)


def prune_rows_scored(x: torch.Tensor, w: torch.Tensor, n_keep: int, keep_last: int = 0) -> torch.Tensor:
    """x [B, H, T, 128] fp16 -> x * (score among the row's n_keep highest); the last `keep_last` tokens of every (b, h) stay
    as they are.  w [B, H, 128]: score = |x * w| (prefill form); w [B, H, T, 128]: w IS the score - the decode-time form
    (`:131-156`), called on the oldest window row with the accumulated score_accumulator[:, :, 0:1, :] / group_size."""
    b, h, t, d = x.shape
    xc, wc = x.contiguous(), w.contiguous()
    if wc.dtype != torch.float16 or wc.shape not in ((b, h, d), (b, h, t, d)):
        raise RuntimeError("prune_rows_scored: w must be float16 [B, H, 128] (weights) or [B, H, T, 128] (scores)")
    y = torch.empty_like(xc)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mfb200_prune_rows_scored(xc.data_ptr(), wc.data_ptr(), y.data_ptr(), b * h * t, t if wc.dim() == 3 else 0,
                                                        d - n_keep + 1, _lib.stream_ptr()), "mfb200_prune_rows_scored")
    if keep_last:
        y[:, :, -keep_last:, :] = xc[:, :, -keep_last:, :]
    return y


def dh_prune_value_channelwise(value_states: torch.Tensor, target_sparsity: float, group_size: int = 32) -> torch.Tensor:
    """Channel-wise value pruning (llama_mustafar_Kt_Mag_Vc_Mag.py:107-170): in every group of `group_size` consecutive tokens
    each channel keeps |v| >= its max(1, int(sparsity * group_size))-th smallest magnitude.  [B, H, T, 128] fp16 CUDA, T a
    multiple of group_size (the reference raises otherwise, `:131`)."""
    if not value_states.is_cuda or value_states.dtype != torch.float16 or value_states.shape[-1] != HEAD_DIM:
        raise RuntimeError("mustafar_b200.pruning: float16 CUDA [B,H,T,128] expected (no CPU fallback)")
    assert 0 <= target_sparsity <= 1, "Target sparsity must be between 0 and 1"
    b, h, t, d = value_states.shape
    if t % group_size != 0:
        raise ValueError("Token dimension must be a multiple of group_size")
    xc = value_states.contiguous()
    y = torch.empty_like(xc)
    k = max(1, int(target_sparsity * group_size))
    with torch.cuda.device(xc.device):
        _lib.check(_lib.load().mfb200_prune_token_groups(xc.data_ptr(), y.data_ptr(), b * h, t, group_size, k, _lib.stream_ptr()),
                   "mfb200_prune_token_groups")
    return y
