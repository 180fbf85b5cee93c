"""Quick correctness + timing probe of the fused decode attention over the BASELINE shapes: python tools/quick_attn.py [gqa|mha|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mustafar_b200.attention import MustafarKVCache
from oracle import torch_oracle as TO

def case(b, hkv, g, T, s, hint=0, check=True):
    gen = torch.Generator(device="cuda").manual_seed(T + g)
    k = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
    v = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
    q = torch.randn(b, hkv * g, 1, 128, device="cuda", generator=gen).half()
    c = MustafarKVCache(b, hkv, g, T + 600, s, s, plan_hint=hint)
    c.prefill(k, v)
    L = c.comp_len
    if check: k[:, :, :L] = TO.prune_rows(k[:, :, :L], s); v[:, :, :L] = TO.prune_rows(v[:, :, :L], s)
    o = c.attend(q); torch.cuda.synchronize()
    ref = TO.masked_dense_attention(q, k, v) if check else o
    d = (o.float() - ref.float()).abs()
    o2 = c.attend(q); torch.cuda.synchronize()
    nbytes = c.compressed_bytes()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c.attend(q); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c.attend(q)
    e1.record(); torch.cuda.synchronize()
    print(f"B={b} Hkv={hkv} G={g} T={T} s={s} hint={hint}: max {d.max().item():.2e} mean {d.mean().item():.2e} nan={torch.isnan(o).any().item()} "
          f"repeat_equal={torch.equal(o, o2)} cold {ts[len(ts)//2]:.1f} us ({nbytes / ts[len(ts)//2] / 1e3:.0f} GB/s) warm {e0.elapsed_time(e1) * 100:.1f} us", flush=True)

GQA = [(1, 1, 4, 160, 0.5), (1, 2, 4, 600, 0.5), (2, 4, 8, 2112, 0.7), (1, 8, 4, 4160, 0.5), (1, 8, 4, 2112, 0.7, 37),
       (4, 8, 4, 8192, 0.7), (16, 8, 4, 8192, 0.7), (4, 8, 4, 32768, 0.5), (32, 8, 4, 32768, 0.5)]
MHA = [(1, 32, 1, 4096, 0.5), (8, 32, 1, 4096, 0.5), (32, 8, 1, 8192, 0.7), (4, 32, 1, 32768, 0.7), (2, 16, 2, 4096, 0.5)]
PERF = [(16, 8, 4, 8192, 0.7), (4, 8, 4, 32768, 0.5), (32, 8, 4, 32768, 0.5), (1, 32, 1, 4096, 0.5), (8, 32, 1, 4096, 0.5), (4, 32, 1, 32768, 0.7)]
if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "b1":  # batch-1 latency against the work decomposition (plan_hint: n > 0 flat plan with n CTAs, -k uniform >= k blocks/split)
        for shape in ((1, 8, 4, 4096, 0.5), (1, 8, 4, 2048, 0.7), (1, 8, 4, 16384, 0.5), (4, 8, 4, 4096, 0.5), (1, 8, 1, 4096, 0.5),
                      (1, 32, 1, 4096, 0.5), (1, 32, 1, 2048, 0.5), (1, 4, 8, 4096, 0.5)):
            for hint in (0, -2, -3, -4, -6, -8):
                case(*shape, hint, check=False)
    if which == "mid":  # mid-size launches: automatic plan vs a forced flat cut over the resident slots vs coarser uniform splits
        for shape, slots in (((2, 8, 4, 4096, 0.5), 296), ((4, 8, 4, 8192, 0.7), 296), ((8, 8, 4, 4096, 0.5), 296),
                             ((2, 32, 1, 4096, 0.5), 444), ((4, 32, 1, 4096, 0.5), 444), ((16, 32, 1, 4096, 0.5), 444)):
            for hint in (0, slots, slots // 2, -6, -8):
                case(*shape, hint, check=False)
    if which == "perf":
        for args in PERF: case(*args, check=False)
    for args in (GQA if which in ("gqa", "all") else []) + (MHA if which in ("mha", "all") else []):
        case(*args)
