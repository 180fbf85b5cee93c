"""Profiling driver for the compression kernels: a few prefill compressions and decode-time chunk appends.
    python tools/prof_compress.py [heads] [tokens] [sparsity]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mustafar_b200.attention import MustafarKVCache

heads = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tokens = int(sys.argv[2]) if len(sys.argv) > 2 else 32512 + 32
s = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
torch.manual_seed(0)
k = torch.randn(1, heads, tokens, 128, device="cuda", dtype=torch.float16)
v = torch.randn(1, heads, tokens, 128, device="cuda", dtype=torch.float16)
c = MustafarKVCache(1, heads, 1, tokens + 600, s, s)
for _ in range(2):
    c.prefill(k, v)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c.prefill(k, v); e1.record(); torch.cuda.synchronize()
nbytes = 2 * heads * c.comp_len * 256
print(f"prefill K+V [{heads},{c.comp_len},128] s={s}: {e0.elapsed_time(e1) * 1e3:.0f} us = {nbytes / e0.elapsed_time(e1) / 1e6:.0f} GB/s of dense input")
# decode-time append: fill the window to 288 rows, then one compress_append_chunk launch
kn = torch.randn(1, heads, 1, 128, device="cuda", dtype=torch.float16)
while c.win_len - c.residual_length < 256:
    c.append(kn, kn)
torch.cuda.synchronize()
e0.record(); did = c.maybe_compress(); e1.record(); torch.cuda.synchronize()
print(f"compress_append_chunk ({heads} units x 256 rows, K+V): {e0.elapsed_time(e1) * 1e3:.0f} us (compressed: {did})")
