"""Randomised differential run of the fused decode attention against the fp32 torch oracle: python tools/fuzz_attn.py [cases] [seed]
Random (batch, KV heads, G, context, sparsity, residual window, mask, forced plans, fused rotary embedding), attend + a few fused decode steps each."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mustafar_b200.attention import MustafarKVCache
from oracle import torch_oracle as TO

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    worst, worst_case = 0.0, None
    for it in range(n):
        g = rng.choice([1, 1, 2, 4, 4, 8])
        hkv = rng.choice([1, 2, 3, 4, 8, 16, 32])
        b = rng.choice([1, 1, 2, 3, 5, 8])
        T = rng.choice([rng.randint(1, 300), rng.randint(300, 3000), rng.randint(3000, 9000)])
        if b * hkv * T > 1_200_000:
            T = max(1, 1_200_000 // (b * hkv))
        s = rng.choice([0.5, 0.7, 0.3, 0.9])
        res = rng.choice([32, 32, 0, 64, 128])
        hint = rng.choice([0, 0, 0, -1, rng.randint(1, 40)])
        gen = torch.Generator(device="cuda").manual_seed(it)
        k = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
        v = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
        c = MustafarKVCache(b, hkv, g, T + 300, s, s, residual_length=res, plan_hint=hint)
        c.prefill(k, v)
        steps = rng.choice([0, 1, 3])
        for t in range(steps + 1):
            q = torch.randn(b, hkv * g, 1, 128, device="cuda", generator=gen).half()
            mask = None
            if t == 0:
                if rng.random() < 0.3:
                    mask = torch.zeros(b, 1, 1, c.kv_seq_len, device="cuda", dtype=torch.float16)
                    lo = rng.randint(0, max(0, c.kv_seq_len - 1)); hi = rng.randint(lo, c.kv_seq_len)
                    if hi - lo < c.kv_seq_len:
                        mask[rng.randrange(b), :, :, lo:hi] = torch.finfo(torch.float16).min
                o = c.attend(q, mask)
            else:
                kn = torch.randn(b, hkv, 1, 128, device="cuda", generator=gen).half(); vn = torch.randn(b, hkv, 1, 128, device="cuda", generator=gen).half()
                if rng.random() < 0.5:  # fused rotary embedding: the launch gets the unrotated rows, the oracle the rotated ones
                    shared = rng.random() < 0.3
                    ang = torch.rand(1 if shared else b, 1, 64, device="cuda", generator=gen) * 6.283
                    cos, sin = torch.cat([ang, ang], -1).cos().half(), torch.cat([ang, ang], -1).sin().half()
                    rot = lambda x: torch.cat([-x[..., 64:], x[..., :64]], -1)
                    o = c.decode_step(q, kn, vn, rope=(cos, sin))
                    q = q * cos.unsqueeze(1) + rot(q) * sin.unsqueeze(1)
                    kn = kn * cos.unsqueeze(1) + rot(kn) * sin.unsqueeze(1)
                else:
                    o = c.decode_step(q, kn, vn)
                k = torch.cat([k, kn], 2); v = torch.cat([v, vn], 2)
            L = c.comp_len if t == 0 or c.comp_len == L0 else c.comp_len
            L0 = c.comp_len
            # the oracle needs the pruned history as the cache holds it NOW (a compression event prunes 256 more rows)
            kp, vp = k.clone(), v.clone()
            Lc = c.comp_len if t == 0 else L_at_launch
            kp[:, :, :Lc] = TO.prune_rows(kp[:, :, :Lc], s); vp[:, :, :Lc] = TO.prune_rows(vp[:, :, :Lc], s)
            ref = TO.masked_dense_attention(q, kp, vp, mask)
            d = (o.float() - ref.float()).abs()
            if d.max().item() > worst:
                worst, worst_case = d.max().item(), (it, b, hkv, g, T, s, res, hint, t, c.kv_seq_len, ref.float().abs().max().item())
            assert d.max().item() <= 2e-3 and d.mean().item() <= 1e-3 and not torch.isnan(o).any(), (it, b, hkv, g, T, s, res, hint, t, d.max().item())
            L_at_launch = c.comp_len  # decode_step compresses AFTER attending: the next launch sees this length
        c.check_overflow()
        if it % 20 == 19:
            print(f"{it + 1} cases ok, worst max-abs {worst:.2e}", flush=True)
    print(f"fuzz ok: {n} cases, worst max-abs {worst:.2e} at (case, b, hkv, G, T, s, residual, hint, step, kv_len, max |ref|) = {worst_case}")

if __name__ == "__main__":
    main()
