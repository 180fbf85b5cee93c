"""Quick GQA (tcgen05 path) correctness + timing probe: python tools/quick_gqa.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mustafar_b200.attention import MustafarKVCache
from oracle import torch_oracle as TO

def case(b, hkv, g, T, s, hint=0, steps=0):
    gen = torch.Generator(device="cuda").manual_seed(T + g)
    k = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
    v = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
    q = torch.randn(b, hkv * g, 1, 128, device="cuda", generator=gen).half()
    c = MustafarKVCache(b, hkv, g, T + 600, s, s, plan_hint=hint)
    c.prefill(k, v)
    L = c.comp_len
    k[:, :, :L] = TO.prune_rows(k[:, :, :L], s); v[:, :, :L] = TO.prune_rows(v[:, :, :L], s)
    o = c.attend(q); torch.cuda.synchronize()
    ref = TO.masked_dense_attention(q, k, v)
    d = (o.float() - ref.float()).abs()
    o2 = c.attend(q); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c.attend(q)
    e1.record(); torch.cuda.synchronize()
    print(f"B={b} Hkv={hkv} G={g} T={T} s={s} hint={hint}: max {d.max().item():.2e} mean {d.mean().item():.2e} nan={torch.isnan(o).any().item()} "
          f"repeat_equal={torch.equal(o, o2)} {e0.elapsed_time(e1) * 100:.1f} us/launch (warm)", flush=True)

if __name__ == "__main__":
    for args in [(1, 1, 4, 96 + 64, 0.5), (1, 2, 4, 600, 0.5), (1, 4, 4, 1312, 0.7), (2, 4, 8, 2112, 0.7), (1, 8, 4, 4160, 0.5),
                 (1, 8, 4, 2112, 0.7, 37), (4, 8, 4, 8192, 0.7), (16, 8, 4, 8192, 0.7), (32, 8, 4, 32768, 0.5)]:
        case(*args)
