"""Profiling driver: a few launches of the fused decode attention kernel on one BASELINE-shaped layer cache.
    python tools/prof_attn.py [cfg1|cfg3|cfg5s] [iters]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mustafar_b200 import _lib
from mustafar_b200.attention import MustafarKVCache

CFG = {
    "cfg1": dict(b=1, hkv=32, g=1, t=4096, s=0.5),
    "cfg3": dict(b=16, hkv=8, g=4, t=8192, s=0.7),
    "cfg4s": dict(b=4, hkv=32, g=1, t=32768, s=0.7),   # config 4 at 1/16 of the batch
    "cfg5s": dict(b=4, hkv=8, g=4, t=32768, s=0.5),    # config 5 at 1/8 of the batch
    "cfg5": dict(b=32, hkv=8, g=4, t=32768, s=0.5),    # config 5, one layer (the bench headline's kernel)
    "mid1": dict(b=8, hkv=32, g=1, t=4096, s=0.5),     # mid-size MHA launches: 256 units
    "mid2": dict(b=32, hkv=8, g=1, t=8192, s=0.7),
    "mid3": dict(b=2, hkv=32, g=1, t=4096, s=0.5),     # 64 units x 60 blocks: 8.6 blocks per CTA slot
    "mid4": dict(b=4, hkv=32, g=1, t=4096, s=0.5),     # 128 units x 60 blocks: 17.3 blocks per CTA slot
}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    c = CFG[name]
    torch.manual_seed(42)
    ncache = 4 if name == "cfg1" else 1
    caches = []
    for i in range(ncache):
        k = torch.randn(c["b"], c["hkv"], c["t"], 128, device="cuda", dtype=torch.float16)
        v = torch.randn(c["b"], c["hkv"], c["t"], 128, device="cuda", dtype=torch.float16)
        cache = MustafarKVCache(c["b"], c["hkv"], c["g"], c["t"], c["s"], c["s"])
        cache.prefill(k, v)
        caches.append(cache)
        del k, v
    q = torch.randn(c["b"], c["hkv"] * c["g"], 1, 128, device="cuda", dtype=torch.float16)
    out = torch.empty_like(q)
    lib = _lib.load()
    params = [cc.make_params(q.view(c["b"], -1, 128), out) for cc in caches]
    nbytes = caches[0].compressed_bytes()
    sp = _lib.stream_ptr()
    for _ in range(3):
        for p in params:
            lib.mfb200_sparse_decode_attention(C.byref(p), sp)
    torch.cuda.synchronize()
    # cold-L2 timing: flush with a 256 MB write between launches
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(iters):
        for p in params:
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.mfb200_sparse_decode_attention(C.byref(p), sp)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{name}: n_split={params[0].n_split} slot_kb={params[0].slot_kb} bytes={nbytes/1e6:.2f} MB  "
          f"cold median {med:.2f} us  min {ts[0]:.2f} us -> {nbytes/med/1e3:.1f} GB/s (median)")


if __name__ == "__main__":
    main()
