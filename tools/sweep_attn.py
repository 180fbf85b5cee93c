"""Latency sweep of the fused decode attention kernel over context length (B=1 by default)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mustafar_b200 import _lib
from mustafar_b200.attention import MustafarKVCache


def run(b, hkv, g, t, s, iters=20):
    torch.manual_seed(1)
    k = torch.randn(b, hkv, t, 128, device="cuda", dtype=torch.float16)
    v = torch.randn(b, hkv, t, 128, device="cuda", dtype=torch.float16)
    cache = MustafarKVCache(b, hkv, g, t, s, s)
    cache.prefill(k, v)
    q = torch.randn(b, hkv * g, 1, 128, device="cuda", dtype=torch.float16)
    out = torch.empty_like(q)
    lib = _lib.load()
    p = cache.make_params(q.view(b, -1, 128), out)
    sp = _lib.stream_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.mfb200_sparse_decode_attention(C.byref(p), sp)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    nbytes = cache.compressed_bytes()
    print(f"B={b} Hkv={hkv} G={g} T={t} s={s}: n_split={p.n_split} L={cache.comp_len} Lw={cache.win_len} "
          f"{nbytes/1e6:8.2f} MB  median {ts[len(ts)//2]:8.2f} us  min {ts[0]:8.2f} us  {nbytes/ts[len(ts)//2]/1e3:7.1f} GB/s")


if __name__ == "__main__":
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    hkv = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    g = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    s = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
    for t in (int(x) for x in (sys.argv[5].split(",") if len(sys.argv) > 5 else "256,576,1088,2112,4096,8192,16384".split(","))):
        run(b, hkv, g, t, s)
