"""Head-sharded decode over real peer memory: one process per GPU (torchrun), cudaIpc-mapped gathered buffers, the fused
launch's P2P-store epilogue vs the NCCL all-gather.  Correctness (bit-equal to the all-gather) and timing per layer-step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mustafar_b200.attention import MustafarKVCache
from mustafar_b200.partition import PeerOutput, gather_heads, make_partition, shard_kv, shard_q


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    b, hkv, groups, T, s, layers = 1, 32, 1, 4096, 0.5, 8
    part = make_partition(b, hkv, world, rank)
    g = torch.Generator().manual_seed(0)  # same data on every rank
    caches = []
    for _ in range(layers):
        k = torch.randn(b, hkv, T, 128, generator=g).half()
        v = torch.randn(b, hkv, T, 128, generator=g).half()
        c = MustafarKVCache(b, part.local_kv_heads, groups, T + 600, s, s, device=dev)
        c.prefill(shard_kv(part, k).contiguous().to(dev), shard_kv(part, v).contiguous().to(dev))
        caches.append(c)
    po = PeerOutput(part, b, hkv * groups, groups, dev)
    steps, ok = 40, True
    qs = torch.randn(steps, b, hkv * groups, 1, 128, generator=g).half().to(dev)
    kns = torch.randn(steps, b, hkv, 1, 128, generator=g).half().to(dev)
    step_id = 0
    # correctness: peer-store result == all-gather of the local outputs
    for t in range(4):
        for c in caches:
            po.bind(c, step_id)
            o = c.decode_step(shard_q(part, qs[t], groups).contiguous(), shard_kv(part, kns[t]).contiguous(), shard_kv(part, kns[t]).contiguous())
            po.wait(step_id)
            want = gather_heads(part, o)
            ok = ok and torch.equal(po.gathered(step_id), want)
            step_id += 1
    torch.cuda.synchronize()
    ok = ok and not po.timed_out()

    def timed(fn, t0, n):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(t0, t0 + n):
            fn(t)
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() * 1e3 / (n * layers)

    def step_peer(t):
        nonlocal step_id
        for c in caches:
            po.bind(c, step_id)
            c.decode_step(shard_q(part, qs[t], groups), shard_kv(part, kns[t]), shard_kv(part, kns[t]))
            po.wait(step_id)
            step_id += 1

    def step_nccl(t):
        for c in caches:
            c.set_peer_output(None)
            gather_heads(part, c.decode_step(shard_q(part, qs[t], groups), shard_kv(part, kns[t]), shard_kv(part, kns[t])))

    def step_local(t):
        for c in caches:
            c.set_peer_output(None)
            c.decode_step(shard_q(part, qs[t], groups), shard_kv(part, kns[t]), shard_kv(part, kns[t]))

    step_peer(4); step_nccl(5); step_local(6)  # warm-up
    us_peer = timed(step_peer, 8, 10)
    us_nccl = timed(step_nccl, 18, 10)
    us_local = timed(step_local, 28, 10)
    flag = torch.tensor([1 if ok and not po.timed_out() else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"head-sharded batch-1 layer, 32 heads / {world} ranks, 4K, s=0.5: peer-store epilogue + wait {us_peer:.1f} us per layer-step, "
              f"NCCL all-gather {us_nccl:.1f} us, attention only {us_local:.1f} us; bit-equal to the all-gather on every rank: {bool(flag.item())}")
    po.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
