"""Drop-in for the reference's kernel/compression.py (Triton) — same call signatures and outputs.

    convert_key_batched(inputs)   -> (bitmaps int64 [B, 2M], accum_counts int32 [B, 2M+1], [packed fp16] * B)
    convert_value_batched(inputs) -> same                                     (compression.py:249-339, :341-432)

`inputs` is an already-pruned fp16 CUDA tensor [B, M, 128] with M % 64 == 0.  Three CUDA launches
(count, scan, pack) and ONE host read of the per-head totals (the reference needs 2B+1 `.item()` syncs,
compression.py:308, :333-334) — unavoidable for this signature because the list holds exact-size
tensors.  The sync-free path is `attention.MustafarKVCache`, which packs into a preallocated slab.

`prune_convert_*_batched(inputs, sparsity)` fuse the reference's dh_prune_* into the same launches.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib
from .pruning import HEAD_DIM, prune_rank


def compress_into(x: torch.Tensor, layout: int, prune_k: int, bitmaps: torch.Tensor, counts: torch.Tensor,
                  accum: torch.Tensor, accum_stride: int, tile_offset: int, head_total):
    """count + scan launches (no host sync).  x: [heads, tokens, 128] contiguous fp16."""
    heads, tokens, _ = x.shape
    lib = _lib.load()
    s = _lib.stream_ptr()
    _lib.check(lib.mfb200_compress_count(x.data_ptr(), heads, tokens, layout, prune_k, bitmaps.data_ptr(),
                                         counts.data_ptr(), s), "mfb200_compress_count")
    _lib.check(lib.mfb200_compress_scan(counts.data_ptr(), heads, tokens * 2, accum.data_ptr(), accum_stride,
                                        tile_offset, 0 if head_total is None else head_total.data_ptr(), s),
               "mfb200_compress_scan")


def pack_into(x: torch.Tensor, layout: int, bitmaps: torch.Tensor, accum: torch.Tensor, accum_stride: int,
              tile_offset: int, head_base: torch.Tensor, packed: torch.Tensor, head_capacity: int = 0,
              overflow: torch.Tensor | None = None):
    heads, tokens, _ = x.shape
    lib = _lib.load()
    _lib.check(lib.mfb200_compress_pack(x.data_ptr(), heads, tokens, layout, bitmaps.data_ptr(), accum.data_ptr(),
                                        accum_stride, tile_offset, head_base.data_ptr(), packed.data_ptr(),
                                        head_capacity, 0 if overflow is None else overflow.data_ptr(),
                                        _lib.stream_ptr()), "mfb200_compress_pack")


def _convert(inputs: torch.Tensor, layout: int, prune_k: int) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor]]:
    assert inputs.dim() == 3
    B, M, N = inputs.shape
    if not inputs.is_cuda:
        raise RuntimeError("mustafar_b200.compression: inputs must be a CUDA tensor (no CPU fallback)")
    assert M % 64 == 0
    if inputs.dtype != torch.float16 or N != HEAD_DIM:
        raise RuntimeError("mustafar_b200.compression: inputs must be float16 [B, M, 128]")
    x = inputs.contiguous()
    dev = x.device
    tiles = (M * N) // 64
    with torch.cuda.device(dev):
        bitmaps = torch.empty((B, tiles), dtype=torch.int64, device=dev)
        counts = torch.empty((B, tiles), dtype=torch.int32, device=dev)
        accum = torch.empty((B, tiles + 1), dtype=torch.int32, device=dev)
        totals = torch.empty((B,), dtype=torch.int32, device=dev)
        compress_into(x, layout, prune_k, bitmaps, counts, accum, tiles + 1, 0, totals)
        halves = totals.to(torch.int64) * 2
        head_base = torch.cumsum(halves, 0) - halves
        sizes = halves.cpu().tolist()  # the one host sync of this API
        packed = torch.empty((sum(sizes),), dtype=torch.float16, device=dev)
        if B > 0 and M > 0:
            pack_into(x, layout, bitmaps, accum, tiles + 1, 0, head_base, packed if packed.numel() else
                      torch.empty((8,), dtype=torch.float16, device=dev))
    return bitmaps, accum, list(torch.split(packed, sizes))


def convert_key_batched(inputs: torch.Tensor):
    return _convert(inputs, _lib.LAYOUT_KEY, 0)


def convert_value_batched(inputs: torch.Tensor):
    return _convert(inputs, _lib.LAYOUT_VALUE, 0)


def prune_convert_key_batched(inputs: torch.Tensor, sparsity: float):
    """dh_prune_key (llama_mustafar_kernel.py:77-113) fused into convert_key_batched."""
    return _convert(inputs, _lib.LAYOUT_KEY, prune_rank(sparsity))


def prune_convert_value_batched(inputs: torch.Tensor, sparsity: float):
    """dh_prune_value (llama_mustafar_kernel.py:117-153) fused into convert_value_batched."""
    return _convert(inputs, _lib.LAYOUT_VALUE, prune_rank(sparsity))


def nz_offsets(accum_counts: torch.Tensor) -> torch.Tensor:
    """uint4-unit start of every head in torch.cat(packed) — the python loop at llama_mustafar_kernel.py:329-331
    (`nz_offset[i] = nz_offset[i-1] + idx[i-1][-1] // 4`) as one device-side cumsum."""
    tot = accum_counts[:, -1].to(torch.int64) // 4
    return (torch.cumsum(tot, 0) - tot).to(torch.int32)
