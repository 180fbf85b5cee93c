"""CPU, world_size=2, gloo: the multi-GPU host logic (partitioning + the one all-gather of the head-sharded
case) reproduces the single-process result.  The per-rank attention is the numpy oracle here (no GPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mustafar_b200 import partition as P
from oracle import mustafar_oracle as O


def test_split_range_covers_exactly_once():
    for n in (0, 1, 7, 32, 2048):
        for w in (1, 2, 3, 8):
            spans = [P.split_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_partition_modes():
    p = P.make_partition(64, 32, 8, 3)
    assert p.mode == "batch" and p.local_batch == 8 and p.local_kv_heads == 32
    p = P.make_partition(1, 32, 8, 3)
    assert p.mode == "head" and p.local_batch == 1 and p.kv_heads == (12, 16)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        hkv, g, t, L = 4, 2, 320, 256
        k = rng.standard_normal((batch, hkv, t, 128)).astype(np.float16)
        v = rng.standard_normal((batch, hkv, t, 128)).astype(np.float16)
        q = rng.standard_normal((batch, hkv * g, 1, 128)).astype(np.float16)
        k[:, :, :L] = O.prune_rows(k[:, :, :L], 0.5)
        v[:, :, :L] = O.prune_rows(v[:, :, :L], 0.5)
        full = O.masked_dense_attention(q, k, v)
        part = P.make_partition(batch, hkv, world, rank)
        kl = P.shard_kv(part, torch.from_numpy(k)).numpy()
        vl = P.shard_kv(part, torch.from_numpy(v)).numpy()
        ql = P.shard_q(part, torch.from_numpy(q), g).numpy()
        out_l = torch.from_numpy(O.masked_dense_attention(ql, kl, vl))
        if part.mode == "head":
            out = P.gather_heads(part, out_l).numpy()
            ok = np.array_equal(out.view(np.uint16), full.view(np.uint16))
        else:
            want = full[part.batch[0]: part.batch[1]]
            ok = np.array_equal(out_l.numpy().view(np.uint16), want.view(np.uint16))
            # weak-scaling bookkeeping used by bench.py: max over ranks of a per-rank time
            tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ok = ok and float(tmax) == float(world)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run(batch):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), batch, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_head_partition_all_gather_gloo():
    _run(batch=1)


def test_batch_partition_no_collective_gloo():
    _run(batch=4)
