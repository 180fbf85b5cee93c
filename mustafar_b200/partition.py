"""Multi-GPU partitioning of the sparse-KV decode path (new; the reference is single-GPU only —
`long_test.sh:11` pins CUDA_VISIBLE_DEVICES=0 and no NCCL/MPI call exists anywhere in it).

Every (sequence, KV head) unit owns its own bitmaps / nonzeros / window and produces the outputs of its
G query heads with no cross-unit dependency (kernel/csrc/SpMM_Kernel.cuh:174-185), so the attention
path needs NO collective:

  * batch partition (default): rank r holds sequences [lo, hi) of the global batch, all heads;
  * head partition (when batch < world): rank r holds KV heads [lo, hi) of every sequence; the
    per-rank outputs [B, Hq/W, 1, 128] are concatenated along the head axis with ONE all-gather
    (NCCL over NVLink on GPUs, gloo in the CPU tests) before the replicated o_proj.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
import torch.distributed as dist


def split_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) share of n items for `rank` (first n % world ranks get one more)."""
    assert 0 <= rank < world and n >= 0
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class Partition:
    mode: str          # "batch" or "head"
    world: int
    rank: int
    batch: Tuple[int, int]     # [lo, hi) sequences held by this rank
    kv_heads: Tuple[int, int]  # [lo, hi) KV heads held by this rank

    @property
    def local_batch(self) -> int:
        return self.batch[1] - self.batch[0]

    @property
    def local_kv_heads(self) -> int:
        return self.kv_heads[1] - self.kv_heads[0]


def make_partition(batch: int, kv_heads: int, world: int, rank: int) -> Partition:
    """Batch-partition when there are at least as many sequences as ranks, else partition KV heads."""
    if batch >= world:
        return Partition("batch", world, rank, split_range(batch, world, rank), (0, kv_heads))
    if kv_heads % world != 0:
        raise ValueError(f"head partition needs kv_heads ({kv_heads}) divisible by world ({world})")
    return Partition("head", world, rank, (0, batch), split_range(kv_heads, world, rank))


def shard_kv(part: Partition, x: torch.Tensor) -> torch.Tensor:
    """This rank's slice of a [B, Hkv, T, D] K/V tensor."""
    return x[part.batch[0]: part.batch[1], part.kv_heads[0]: part.kv_heads[1]]


def shard_q(part: Partition, q: torch.Tensor, groups: int) -> torch.Tensor:
    """This rank's slice of a [B, Hq, 1, D] query tensor (query heads follow their KV head)."""
    return q[part.batch[0]: part.batch[1], part.kv_heads[0] * groups: part.kv_heads[1] * groups]


def gather_heads(part: Partition, out_local: torch.Tensor, group=None) -> torch.Tensor:
    """head partition only: all-gather the per-rank [B, Hq/W, 1, D] outputs into [B, Hq, 1, D]."""
    if part.mode != "head" or part.world == 1:
        return out_local
    parts = [torch.empty_like(out_local) for _ in range(part.world)]
    dist.all_gather(parts, out_local.contiguous(), group=group)
    return torch.cat(parts, dim=1)
