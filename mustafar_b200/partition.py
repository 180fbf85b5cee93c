"""Multi-GPU partitioning of the sparse-KV decode path (new; the reference is single-GPU only —
`long_test.sh:11` pins CUDA_VISIBLE_DEVICES=0 and no NCCL/MPI call exists anywhere in it).

Every (sequence, KV head) unit owns its own bitmaps / nonzeros / window and produces the outputs of its
G query heads with no cross-unit dependency (kernel/csrc/SpMM_Kernel.cuh:174-185), so the attention
path needs NO collective:

  * batch partition (default): rank r holds sequences [lo, hi) of the global batch, all heads;
  * head partition (when batch < world): rank r holds KV heads [lo, hi) of every sequence; the
    per-rank outputs [B, Hq/W, 1, 128] are concatenated along the head axis before the replicated o_proj —
    either with ONE all-gather (`gather_heads`: NCCL over NVLink on GPUs, gloo in the CPU tests) or, on the GPUs of one
    node, with NO collective at all (`PeerOutput`): the fused attention launch's split-merge epilogue stores every output
    row straight into all ranks' gathered buffers (peer-to-peer stores over NVLink / NVSwitch) and raises per-unit arrival
    flags; the consumer side is one tiny wait launch.  A batch-1 layer at 8 ranks: 105.6 us per layer-step with the NCCL
    all-gather, 18.9 us with the peer stores (11.5 us for the attention launch alone; profiles/r2_bench_n8.json).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

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


class _DevMem:
    """A raw device allocation seen by torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerOutput:
    """Gathered output buffers of a head-sharded decode that the ranks fill with peer-to-peer stores.

    Every rank allocates `n_buffers` gathered outputs fp16 [B, Hq_total, 1, 128] plus arrival flags
    uint32 [n_buffers][world][B * Hkv_local] in ONE cudaMalloc'ed block, exports it as a cudaIpc handle and maps the other
    ranks' blocks (`mfb200_ipc_export/open`; the handles travel through `torch.distributed.all_gather_object`).
    `bind(cache, step)` points the cache's launches at buffer `step % n_buffers` of every rank; after the step's launch,
    `wait(step)` (one small launch) returns in stream order once every rank's rows have landed, and `gathered(step)` is
    this rank's complete [B, Hq_total, 1, 128].  Two buffers suffice: a rank cannot start step t+2 before every rank has
    finished step t+1's launch, which is stream-ordered after that rank's use of step t's buffer.
    """

    def __init__(self, part: Partition, batch: int, q_heads_total: int, groups: int, device, n_buffers: int = 2, group=None,
                 _local_peers: Optional[List["PeerOutput"]] = None):
        from . import _lib
        self._lib, self._L = _lib, _lib.load()
        self.part, self.batch, self.rows_total, self.groups, self.n_buffers = part, batch, q_heads_total, groups, n_buffers
        self.world, self.rank = part.world, part.rank
        self.units = batch * part.local_kv_heads
        self.device = torch.device(device)
        if not 2 <= self.world <= _lib.MAX_PEERS:
            raise ValueError(f"PeerOutput: 2..{_lib.MAX_PEERS} ranks, got {self.world}")
        self.out_bytes = batch * q_heads_total * 128 * 2
        self.flag_bytes = ((self.world * self.units * 4 + 255) // 256) * 256
        self.nbytes = n_buffers * (self.out_bytes + self.flag_bytes) + 256  # + the timed-out word
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.mfb200_peer_alloc(self.nbytes, C.byref(ptr)), "mfb200_peer_alloc")
        self.base = [0] * self.world
        self.base[self.rank] = ptr.value
        self._opened: List[int] = []
        self._mem = torch.as_tensor(_DevMem(ptr.value, self.nbytes), device=self.device)
        if _local_peers is None:  # one process per GPU: exchange cudaIpc handles
            handle = C.create_string_buffer(64)
            _lib.check(self._L.mfb200_ipc_export(ptr, handle), "mfb200_ipc_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, handle.raw, group=group)
            for r, h in enumerate(handles):
                if r != self.rank:
                    q = C.c_void_p()
                    with torch.cuda.device(self.device):
                        _lib.check(self._L.mfb200_ipc_open(h, C.byref(q)), "mfb200_ipc_open")
                    self.base[r] = q.value
                    self._opened.append(q.value)
        self._structs = [_lib.PeerOut() for _ in range(n_buffers)]
        self._steps_bound = 0

    @classmethod
    def local_group(cls, parts: List[Partition], batch: int, q_heads_total: int, groups: int, device, n_buffers: int = 2):
        """All ranks inside ONE process on ONE device (tests: the same stores and flags, no IPC)."""
        objs = [cls(p, batch, q_heads_total, groups, device, n_buffers, _local_peers=[]) for p in parts]
        for o in objs:
            o.base = [x.base[x.rank] for x in objs]
        return objs

    def _out_ptr(self, r: int, buf: int) -> int:
        return self.base[r] + buf * (self.out_bytes + self.flag_bytes)

    def _flag_ptr(self, r: int, buf: int) -> int:
        return self._out_ptr(r, buf) + self.out_bytes

    def bind(self, cache, step: int) -> None:
        """Point `cache`'s next launch at buffer step % n_buffers of every rank."""
        cache.set_peer_output(self.block(step))

    def block(self, step: int):
        """The `_lib.PeerOut` block of `step` (kept alive by this object): assign its address to `DecodeParams.peer`."""
        buf, st = step % self.n_buffers, self._structs[step % self.n_buffers]
        st.n_peers, st.rank, st.rows_total = self.world, self.rank, self.rows_total
        st.row0 = self.part.kv_heads[0] * self.groups
        st.epoch = step // self.n_buffers + 1
        for r in range(self.world):
            st.out[r], st.flags[r] = self._out_ptr(r, buf), self._flag_ptr(r, buf)
        return st

    def wait(self, step: int) -> None:
        """Stream-ordered: returns once every rank's rows of `step` are in this rank's gathered buffer."""
        buf = step % self.n_buffers
        with torch.cuda.device(self.device):
            self._lib.check(self._L.mfb200_peer_wait(self._flag_ptr(self.rank, buf), self.world * self.units, step // self.n_buffers + 1,
                                                     self.base[self.rank] + self.nbytes - 256,
                                                     torch.cuda.current_stream(self.device).cuda_stream), "mfb200_peer_wait")

    def gathered(self, step: int) -> torch.Tensor:
        buf = step % self.n_buffers
        off = buf * (self.out_bytes + self.flag_bytes)
        return self._mem[off: off + self.out_bytes].view(torch.float16).view(self.batch, self.rows_total, 1, 128)

    def timed_out(self) -> bool:
        """Host-syncing: did any wait give up (a peer died)?"""
        return bool(self._mem[self.nbytes - 256: self.nbytes - 252].view(torch.int32).item())

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown: the driver releases the memory with the context
            pass

    def close(self) -> None:
        """Unmap the peers' blocks and free the own one (after every rank is done with it: synchronise / barrier first)."""
        for q in self._opened:
            self._L.mfb200_ipc_close(q)
        self._opened = []
        if self.base[self.rank]:
            self._mem = None
            self._L.mfb200_peer_free(self.base[self.rank])
            self.base[self.rank] = 0
