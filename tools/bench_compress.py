"""Throughput of the prune+compress path (prefill-side) on BASELINE shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mustafar_b200 import compression, pruning
from mustafar_b200.attention import MustafarKVCache

def t(fn, iters=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for (bk, m, s) in [(32, 3840, 0.5), (128, 7936, 0.7), (256, 32512, 0.5)]:
    x = torch.randn(bk, m, 128, device="cuda", dtype=torch.float16)
    nbytes = x.numel() * 2
    us_p = t(lambda: pruning.dh_prune_key(x.view(1, bk, m, 128), s))
    us_k = t(lambda: compression.prune_convert_key_batched(x, s), iters=5)
    us_v = t(lambda: compression.prune_convert_value_batched(x, s), iters=5)
    c = MustafarKVCache(1, bk, 1, m + 256, s, s)
    xk = x.view(1, bk, m, 128)
    us_pref = t(lambda: c.prefill(xk, xk), iters=5)
    print(f"[{bk},{m},128] s={s}: prune {us_p:.0f} us ({2*nbytes/us_p/1e3:.0f} GB/s r+w) | prune+convert_key (list API) {us_k:.0f} us | value {us_v:.0f} us | "
          f"cache.prefill K+V (slab, no sync) {us_pref:.0f} us = {2*nbytes/us_pref/1e3:.0f} GB/s of dense input")

# the other pruning policies on the largest shape
x = torch.randn(8, 32, 32512 // 8 * 8 // 32 * 32, 128, device="cuda", dtype=torch.float16)  # [8, 32, 32512, 128] = 2.1 GB
qf = torch.randn(8, 32, 32, 128, device="cuda", dtype=torch.float16)
w = pruning.fold_queries(qf, 1, 32)
nb = x.numel() * 2
us_s = t(lambda: pruning.prune_rows_scored(x, w, 64), iters=5)
us_g = t(lambda: pruning.dh_prune_value_channelwise(x, 0.5, 32), iters=5)
print(f"[256,{x.shape[2]},128] output-aware key prune (|x*w| top-64 per row) {us_s:.0f} us ({2*nb/us_s/1e3:.0f} GB/s r+w) | "
      f"channel-wise value prune (groups of 32 tokens) {us_g:.0f} us ({2*nb/us_g/1e3:.0f} GB/s r+w)")
