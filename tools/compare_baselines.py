"""Same-box comparison lines for DESIGN.md / BASELINE.md §5 (test/measurement infrastructure, not product):
   (a) the fused kernel, (b) the reference's own CUDA kernels (oracle/_ref, built for sm_100a) driven by the
   reference glue of models/llama_mustafar_kernel.py:268-320 — "kernels only" and "full glue",
   (c) dense FlashAttention decode (flash_attn_with_kvcache) on the unpruned fp16 KV,
   (d) masked-dense PyTorch attention on the GPU.
   python tools/compare_baselines.py cfg1|cfg3|cfg5s|cfg4s
"""
import ctypes as C
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from mustafar_b200 import _lib
from mustafar_b200.attention import MustafarKVCache
from oracle import ref_cuda
from tools.prof_attn import CFG


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    c = CFG[name]
    b, hkv, g, t, s = c["b"], c["hkv"], c["g"], c["t"], c["s"]
    hq = hkv * g
    torch.manual_seed(42)
    k = torch.randn(b, hkv, t, 128, device="cuda", dtype=torch.float16)
    v = torch.randn(b, hkv, t, 128, device="cuda", dtype=torch.float16)
    q = torch.randn(b, hq, 1, 128, device="cuda", dtype=torch.float16)
    cache = MustafarKVCache(b, hkv, g, t, s, s)
    cache.prefill(k, v)
    L = cache.comp_len
    nbytes = cache.compressed_bytes()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = torch.empty_like(q)
    p = cache.make_params(q, out)
    sp = _lib.stream_ptr()
    lib = _lib.load()
    rows = []
    t_ours = timeit(lambda: lib.mfb200_sparse_decode_attention(C.byref(p), sp), flush=flush)
    rows.append(("mustafar_b200 fused kernel (1 launch)", t_ours))

    if ref_cuda.available():
        kc, kw, vc, vw, L2, _ = cache.as_reference_tuple()
        nzk, nzv = ref_cuda.pad_nz(kc[2]), ref_cuda.pad_nz(vc[2])
        pq = F.pad(q.reshape(b * hq, 1, 128), (0, 0, 0, 7)).contiguous()
        pp = F.pad(torch.softmax(torch.randn(b * hq, 1, L, device="cuda"), -1).half(), (0, 0, 0, 7)).contiguous()
        idxk, idxv = kc[1].reshape(-1), vc[1].reshape(-1)

        def ref_kernels():
            ref_cuda.key_formulation(kc[0], nzk, idxk, kc[3], pq, L, 128, b * hq, g)
            ref_cuda.value_formulation(vc[0], nzv, idxv, vc[3], pp, 128, L, b * hq, g)

        rows.append(("reference CUDA kernels only (Key+Value SpMV, sm_100a build, NZ pre-concatenated)", timeit(ref_kernels, flush=flush)))
        rows.append(("reference CUDA kernels + reference glue (cat/pad/window matmul/softmax)",
                     timeit(lambda: ref_cuda.decode_step(q, kc, kw, vc, vw, L, g), iters=10, flush=flush)))
        del kc, vc, nzk, nzv
    try:
        from flash_attn import flash_attn_with_kvcache
        kd = k.transpose(1, 2).contiguous()  # [B, T, Hkv, D]
        vd = v.transpose(1, 2).contiguous()
        qd = q.transpose(1, 2).contiguous()  # [B, 1, Hq, D]
        rows.append(("flash_attn_with_kvcache, dense fp16 KV (unpruned)", timeit(lambda: flash_attn_with_kvcache(qd, kd, vd), flush=flush)))
        del kd, vd
    except Exception as e:  # noqa: BLE001
        rows.append((f"flash_attn unavailable: {e}", float("nan")))

    def masked_dense():
        w = torch.matmul(q, ref_cuda.repeat_kv(k, g).transpose(2, 3)) / math.sqrt(128)
        pr = torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.float16)
        return torch.matmul(pr, ref_cuda.repeat_kv(v, g))

    if b * hq * t * 128 * 2 < 8e9:
        rows.append(("masked-dense PyTorch attention on the GPU", timeit(masked_dense, iters=10, flush=flush)))
    dense_bytes = 2 * b * hkv * t * 128 * 2
    print(f"### {name}: B={b} Hkv={hkv} G={g} T={t} s={s}  compressed {nbytes/1e6:.1f} MB, dense fp16 KV {dense_bytes/1e6:.1f} MB")
    print("| implementation | us/step (median, cold L2) | vs fused | GB/s on its own bytes |")
    print("|---|---|---|---|")
    for label, us in rows:
        own = dense_bytes if ("flash" in label or "masked" in label) else nbytes
        print(f"| {label} | {us:.1f} | {us / t_ours:.2f}x | {own / us / 1e3:.0f} |")


if __name__ == "__main__":
    main()
