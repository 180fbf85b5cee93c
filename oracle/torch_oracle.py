"""Torch restatement of the oracle for shapes numpy cannot finish in seconds.  TEST INFRASTRUCTURE ONLY.

Same rule as `oracle/mustafar_oracle.py` (only tests/, smoke() and bench.py's baseline legs may import it; the
product never does).  It runs on whatever device its inputs live on — on the GPU box that makes BASELINE configs
3/4/5 (B=16 x 8 heads x 8K ... B=32 x 8 heads x 32K) checkable at their real sizes — and it is itself pinned to the
numpy oracle (which is pinned to the reference's golden vectors) at small shapes, on the CPU in the `not gpu`
suite and on the GPU in the `gpu` suite (tests/test_torch_oracle.py).

  prune_rows               models/llama_mustafar_kernel.py:97-110, :137-149 (torch.kthvalue + `>=`, exactly the
                           reference's own formulation)
  masked_dense_attention   models/llama_mustafar_Kt_Mag_Vt_Mag.py:873-874, :952-963, :974 with the reference's
                           rounding points (fp32-accumulated scores rounded to fp16, /sqrt(d) in fp16, fp32
                           softmax rounded to fp16, fp32-accumulated P.V rounded to fp16), evaluated in chunks of
                           sequences so that no fp32 copy of the whole K/V is ever materialised.
All arithmetic is plain torch (matmul in fp32 with TF32 disabled by the caller's default settings).
"""
from __future__ import annotations

import math

import torch

HEAD_DIM = 128


def prune_k(sparsity: float, dim: int = HEAD_DIM) -> int:
    return max(1, int(sparsity * dim))


def prune_rows(x: torch.Tensor, sparsity: float, chunk_rows: int = 1 << 20) -> torch.Tensor:
    """x fp16 [..., 128] -> x * (|x| >= kthvalue(|x|, k)); every tie at the threshold survives, dropped entries
    keep their sign (x * 0 = +-0)."""
    assert x.dtype == torch.float16
    d = x.shape[-1]
    k = prune_k(sparsity, d)
    flat = x.reshape(-1, d)
    out = torch.empty_like(flat)
    for r0 in range(0, flat.shape[0], chunk_rows):
        xs = flat[r0:r0 + chunk_rows]
        mag = xs.abs().float()  # kthvalue has no fp16 CPU kernel; fp16 -> fp32 is exact and order preserving
        thr = torch.kthvalue(mag, k, dim=-1, keepdim=True).values
        out[r0:r0 + chunk_rows] = xs * (mag >= thr).to(torch.float16)
    return out.reshape(x.shape)


def masked_dense_attention(q: torch.Tensor, k_full: torch.Tensor, v_full: torch.Tensor, mask: torch.Tensor | None = None,
                           seq_chunk: int = 1) -> torch.Tensor:
    """q fp16 [B,Hq,1,D]; k_full/v_full fp16 [B,Hkv,T,D] (pruned rows followed by dense rows); mask additive
    [B,1,1,T] or None.  Returns fp16 [B,Hq,1,D]."""
    b, hq, _, d = q.shape
    hkv = k_full.shape[1]
    g = hq // hkv
    out = torch.empty_like(q)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for b0 in range(0, b, seq_chunk):
            sl = slice(b0, b0 + seq_chunk)
            qf = q[sl].float().reshape(-1, hkv, g, d)                       # [b, Hkv, G, D]
            kf = k_full[sl].float()                                         # [b, Hkv, T, D]
            w = torch.matmul(qf, kf.transpose(2, 3)).to(torch.float16)      # fp32 accumulate -> fp16 (SpMM_Kernel.cuh:418)
            w = (w.float() / math.sqrt(d)).to(torch.float16)                # `/ sqrt(d)` on an fp16 tensor (:874)
            if mask is not None:
                m = mask[sl].reshape(-1, 1, 1, mask.shape[-1]).float()
                w = (w.float() + m).to(torch.float16)
                w = torch.maximum(w, torch.tensor(torch.finfo(torch.float16).min, dtype=torch.float16, device=w.device))
            p = torch.softmax(w.float(), dim=-1).to(torch.float16)          # fp32 softmax -> fp16 (:963)
            del kf
            o = torch.matmul(p.float(), v_full[sl].float()).to(torch.float16)  # (:974)
            out[sl] = o.reshape(-1, hq, 1, d)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return out


def attention_f64(q: torch.Tensor, k_full: torch.Tensor, v_full: torch.Tensor, seq_chunk: int = 1) -> torch.Tensor:
    """Rounding-free midpoint (float64) over the same pruned K/V; bounds the error of both formulations."""
    b, hq, _, d = q.shape
    hkv = k_full.shape[1]
    g = hq // hkv
    out = torch.empty(q.shape, dtype=torch.float64, device=q.device)
    for b0 in range(0, b, seq_chunk):
        sl = slice(b0, b0 + seq_chunk)
        qf = q[sl].double().reshape(-1, hkv, g, d)
        w = torch.matmul(qf, k_full[sl].double().transpose(2, 3)) / math.sqrt(d)
        p = torch.softmax(w, dim=-1)
        out[sl] = torch.matmul(p, v_full[sl].double()).reshape(-1, hq, 1, d)
    return out


def compressed_length(kv_seq_len: int, residual_length: int = 32) -> int:
    return max(0, ((kv_seq_len - residual_length) // 256) * 256)
