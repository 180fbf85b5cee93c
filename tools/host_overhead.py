"""Host cost of issuing one layer-step (MustafarKVCache.decode_step) vs the GPU time it buys.  python -m tools.host_overhead"""
import time

import torch

from mustafar_b200.attention import MustafarKVCache


def main():
    layers, heads, ctx = 32, 32, 4096
    torch.manual_seed(0)
    caches = []
    for _ in range(layers):
        k = torch.randn(1, heads, ctx, 128, device="cuda", dtype=torch.float16)
        v = torch.randn(1, heads, ctx, 128, device="cuda", dtype=torch.float16)
        c = MustafarKVCache(1, heads, 1, ctx + 1024, 0.5, 0.5, 256)
        c.prefill(k, v)
        caches.append(c)
    x = torch.randn(layers, 3, heads, 128, device="cuda", dtype=torch.float16)
    out = torch.empty(layers, heads, 1, 128, device="cuda", dtype=torch.float16)
    vw = [(x[l, 0].view(1, heads, 1, 128), x[l, 1].view(1, heads, 1, 128), x[l, 2].view(1, heads, 1, 128), out[l:l + 1]) for l in range(layers)]

    def step():
        for c, (q, kn, vn, o) in zip(caches, vw):
            c.decode_step(q, kn, vn, out=o)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    for n in (8, 24):  # 24 steps stay inside one 256-token window (no compression event)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{n} steps: host issue {1e6*(t1-t0)/(n*layers):.2f} us/layer-step, GPU {1e3*e0.elapsed_time(e1)/(n*layers):.2f} us/layer-step, "
              f"wall incl. drain {1e6*(t2-t0)/(n*layers):.2f} us/layer-step; window {caches[0].win_len}")


if __name__ == "__main__":
    main()
