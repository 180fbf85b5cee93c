"""Soak: long decode runs (several compression events) for every G - CUDA-graph steps == eager steps bit-exactly, both against the
full-history masked-dense oracle.  python tools/soak.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mustafar_b200.attention import DecodeStepGraph, MustafarKVCache
from oracle import torch_oracle as TO
for (b, hkv, groups, s, T0, steps, layers) in [(1, 32, 1, 0.5, 700, 1100, 2), (4, 8, 4, 0.7, 1500, 1100, 2), (2, 4, 8, 0.5, 300, 600, 2), (3, 4, 2, 0.7, 333, 600, 1)]:
    gen = torch.Generator().manual_seed(7)
    ca, cb, ks, vs = [], [], [], []
    for _ in range(layers):
        k = torch.randn(b, hkv, T0, 128, generator=gen).half().cuda(); v = torch.randn(b, hkv, T0, 128, generator=gen).half().cuda()
        ks.append(k); vs.append(v)
        for lst in (ca, cb):
            c = MustafarKVCache(b, hkv, groups, T0 + steps + 64, s, s); c.prefill(k, v); lst.append(c)
    q = torch.zeros(layers, b, hkv * groups, 128, dtype=torch.float16, device="cuda"); kn = torch.zeros(layers, b, hkv, 128, dtype=torch.float16, device="cuda")
    vn = torch.zeros_like(kn); out = torch.zeros_like(q)
    g = DecodeStepGraph(cb, q, kn, vn, out)
    worst = 0.0
    for t in range(steps):
        q.copy_(torch.randn(q.shape, generator=gen).half()); kn.copy_(torch.randn(kn.shape, generator=gen).half()); vn.copy_(torch.randn(vn.shape, generator=gen).half())
        for l in range(layers):
            ks[l] = torch.cat([ks[l], kn[l].view(b, hkv, 1, 128)], 2); vs[l] = torch.cat([vs[l], vn[l].view(b, hkv, 1, 128)], 2)
        want = torch.stack([c.decode_step(q[l].view(b, -1, 1, 128), kn[l].view(b, hkv, 1, 128), vn[l].view(b, hkv, 1, 128)).view(b, -1, 128) for l, c in enumerate(ca)])
        got = g.step()
        assert torch.equal(got, want), t
        if t % 97 == 0 or t == steps - 1:  # against the masked-dense oracle on the full history
            for l in range(layers):
                L = ca[l].comp_len
                kp, vp = ks[l].clone(), vs[l].clone()
                kp[:, :, :L] = TO.prune_rows(kp[:, :, :L], s); vp[:, :, :L] = TO.prune_rows(vp[:, :, :L], s)
                ref = TO.masked_dense_attention(q[l].view(b, -1, 1, 128), kp, vp).view(b, -1, 128)
                worst = max(worst, (got[l].float() - ref.float()).abs().max().item())
    print(f"B={b} Hkv={hkv} G={groups} s={s}: {steps} steps, {g.captures} graph captures, comp_len {ca[0].comp_len}, regrows {ca[0].regrow_events}, "
          f"graph == eager bit-exact, max |out - masked-dense oracle| {worst:.2e}", flush=True)
    assert worst <= 2e-3
