"""GPU parity at the BASELINE.json configurations' REAL shapes (-m gpu).

The numpy oracle cannot finish these in seconds, so the checker is oracle/torch_oracle.py on the GPU (pinned to the
numpy oracle in tests/test_torch_oracle.py) and, for the compression format, the reference's OWN Triton kernels
(kernel/compression.py) cross-compiled from /root/reference into oracle/_ref/triton/*.cubin and launched here.
Tolerances: bit-exact for bitmaps / offsets / packed values; 2e-3 max-abs, 1e-3 mean-abs (north_star) for attention.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_triton
from oracle import torch_oracle as TO

pytestmark = pytest.mark.gpu
MAX_ABS, MEAN_ABS = 2e-3, 1e-3


def _randn_gpu(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda", dtype=torch.float32).to(torch.float16)


def _check(out, q, kfull, vfull, what):
    ref = TO.masked_dense_attention(q, kfull, vfull).float()
    d = (out.float() - ref).abs()
    assert not torch.isnan(out).any(), what
    assert d.max().item() <= MAX_ABS and d.mean().item() <= MEAN_ABS, (what, d.max().item(), d.mean().item())
    return d.max().item()


# (batch, kv_heads, groups, context, sparsity): configs 3, 4 (the largest batch whose dense K/V + oracle temporaries fit
# comfortably beside the cache: 8 of the 64 sequences, same per-unit shape and the same multi-segment flat CTAs), 5
CONFIGS = [
    pytest.param(16, 8, 4, 8192, 0.7, id="config3-B16-Hkv8-G4-T8192-s0.7"),
    pytest.param(8, 32, 1, 32768, 0.7, id="config4-B8of64-Hkv32-G1-T32768-s0.7"),
    pytest.param(32, 8, 4, 32768, 0.5, id="config5-B32-Hkv8-G4-T32768-s0.5"),
]


@pytest.mark.parametrize("b,hkv,groups,T,sparsity", CONFIGS)
def test_baseline_config_attend_decode_and_compression_event(b, hkv, groups, T, sparsity):
    """attend() on the prefilled cache, 3 fused decode steps, then on through a compression event (the window
    reaches residual+256 and its first 256 rows are pruned, compressed and appended in place), each checked against
    masked-dense attention over the same K/V with rows [0, compressed_length) pruned by the reference rule."""
    from mustafar_b200.attention import MustafarKVCache
    hq = hkv * groups
    k = _randn_gpu((b, hkv, T, 128), 11 * T + b)
    v = _randn_gpu((b, hkv, T, 128), 11 * T + b + 1)
    steps = 40
    extra_k = _randn_gpu((steps, b, hkv, 1, 128), 5)
    extra_v = _randn_gpu((steps, b, hkv, 1, 128), 6)
    qs = _randn_gpu((steps + 1, b, hq, 1, 128), 7)
    cache = MustafarKVCache(b, hkv, groups, max_tokens=T + 512, k_sparsity=sparsity, v_sparsity=sparsity)
    cache.prefill(k, v)
    L0 = cache.comp_len
    assert L0 == TO.compressed_length(T) and cache.win_len == T - L0
    # expected K/V: the prompt + the generated rows, rows [0, L) pruned
    kfull = torch.cat([k, extra_k.permute(1, 2, 0, 3, 4).reshape(b, hkv, steps, 128)], dim=2)
    vfull = torch.cat([v, extra_v.permute(1, 2, 0, 3, 4).reshape(b, hkv, steps, 128)], dim=2)
    del k, v
    kfull[:, :, :L0] = TO.prune_rows(kfull[:, :, :L0], sparsity)
    vfull[:, :, :L0] = TO.prune_rows(vfull[:, :, :L0], sparsity)
    worst = _check(cache.attend(qs[0]), qs[0], kfull[:, :, :T], vfull[:, :, :T], "attend")
    assert torch.equal(cache.attend(qs[0]), cache.attend(qs[0]))
    compressed_at = None
    for i in range(steps):
        before = cache.comp_len
        out = cache.decode_step(qs[i + 1], extra_k[i], extra_v[i])
        n = T + i + 1
        if i < 3 or compressed_at is not None or cache.comp_len != before:
            # the step's output is computed BEFORE its own (possible) compression: rows [0, before) are pruned
            worst = max(worst, _check(out, qs[i + 1], kfull[:, :, :n], vfull[:, :, :n], f"decode step {i}"))
        if cache.comp_len != before:
            assert cache.comp_len == before + 256 and compressed_at is None
            compressed_at = i
            kfull[:, :, before:before + 256] = TO.prune_rows(kfull[:, :, before:before + 256], sparsity)
            vfull[:, :, before:before + 256] = TO.prune_rows(vfull[:, :, before:before + 256], sparsity)
    assert compressed_at is not None and compressed_at < steps - 2, "the run must cross a compression event and go on"
    assert cache.kv_seq_len == T + steps and cache.regrow_events == 0
    assert cache.bytes_held() <= (0.72 if sparsity <= 0.5 else 0.53) * cache.dense_bytes() + 2 * cache.k_win.numel() * 2 + (64 << 20)
    cache.check_overflow()


@pytest.mark.parametrize("b,m,sparsity", [(2, 256, 0.5), (32, 3840, 0.5), (128, 7936, 0.7)])
def test_compression_equals_reference_triton_on_gpu(b, m, sparsity):
    """SURVEY.md §7 minimum slice: torch.equal on bitmaps, accum_counts and every packed value (incl. the zero
    padding) against the reference's Triton kernels at [32, 3840, 128] and [128, 7936, 128], for K and V, through
    both product paths: the list-returning compat API and the single-pass prefill into the cache slabs."""
    if not ref_triton.available(b, m):
        pytest.skip("oracle/_ref/triton cubins not built (python oracle/build_ref_triton.py where /root/reference is mounted)")
    from mustafar_b200 import compression
    from mustafar_b200.attention import MustafarKVCache
    x = TO.prune_rows(_randn_gpu((b, m, 128), m + b), sparsity)
    y = TO.prune_rows(_randn_gpu((b, m, 128), m + b + 1), sparsity)
    for which, inp, ref_fn, our_fn in (("key", x, ref_triton.convert_key_batched, compression.convert_key_batched),
                                       ("value", y, ref_triton.convert_value_batched, compression.convert_value_batched)):
        rb, ra, rp, starts, ends = ref_fn(inp)
        ob, oa, op = our_fn(inp)
        assert torch.equal(ob, rb), which
        assert torch.equal(oa, ra), which
        assert torch.equal(torch.cat(op).view(torch.int16), rp.view(torch.int16)), which
        assert [t.numel() for t in op] == (ends - starts).tolist(), which
    # the cache path: prune + compress fused, straight into the slabs (prompt of m + 32 tokens -> L = m)
    kin = _randn_gpu((b, 1, m + 32, 128), 3 * m)
    vin = _randn_gpu((b, 1, m + 32, 128), 3 * m + 1)
    cache = MustafarKVCache(b, 1, 1, max_tokens=m + 64, k_sparsity=sparsity, v_sparsity=sparsity)
    cache.prefill(kin, vin)
    assert cache.comp_len == m
    for st, inp, ref_fn in ((cache.k, kin, ref_triton.convert_key_batched), (cache.v, vin, ref_triton.convert_value_batched)):
        rb, ra, rp, starts, ends = ref_fn(TO.prune_rows(inp[:, 0, :m].contiguous(), sparsity))
        assert torch.equal(st.bmp[:, :2 * m], rb)
        assert torch.equal(st.idx[:, :2 * m + 1], ra)
        sizes = (ends - starts).tolist()
        ours = torch.cat([st.nz[u * st.head_capacity: u * st.head_capacity + sizes[u]] for u in range(b)])
        assert torch.equal(ours.view(torch.int16), rp.view(torch.int16))


def test_slabs_regrow_instead_of_overflowing():
    """Slabs are sized from the sparsity; data that keeps more (ties) must trigger a regrow, never an overrun:
    (i) a prompt that does not fit its slab is recompressed into worst-case slabs (one host read at prefill);
    (ii) during decode a chunk is appended only if even an all-ties chunk is known to fit, else the slabs are regrown
    first.  Results stay correct and the container stays bit-identical to a whole compression."""
    from mustafar_b200.attention import MustafarKVCache
    from oracle import mustafar_oracle as O
    b, hkv, g, T0, s = 1, 2, 2, 600, 0.5
    gen = torch.Generator().manual_seed(9)
    k = torch.randn(b, hkv, T0, 128, generator=gen).to(torch.float16)
    v = torch.randn(b, hkv, T0, 128, generator=gen).to(torch.float16)
    # (i) slab too small for the prompt: 768 compressed tokens x ~72 halves > 1024 x 8 + the 32K reserve
    k1 = torch.randn(b, hkv, 1000, 128, generator=gen).to(torch.float16)
    v1 = torch.randn(b, hkv, 1000, 128, generator=gen).to(torch.float16)
    q1 = torch.randn(b, hkv * g, 1, 128, generator=gen).to(torch.float16)
    c = MustafarKVCache(b, hkv, g, 1024, s, s, nz_halves_per_token=8)
    c.prefill(k1.cuda(), v1.cuda())
    assert c.regrow_events == 1 and c.k.halves_per_token == 128 and c.comp_len == 768
    c.check_overflow()
    kp, vp = k1.numpy().copy(), v1.numpy().copy()
    kp[:, :, :768] = O.prune_rows(kp[:, :, :768], s)
    vp[:, :, :768] = O.prune_rows(vp[:, :, :768], s)
    d = np.abs(c.attend(q1.cuda()).float().cpu().numpy() - O.masked_dense_attention(q1.numpy(), kp, vp).astype(np.float32))
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS
    # (ii) fits the prompt, a later decode-time chunk must regrow first
    c2 = MustafarKVCache(b, hkv, g, 1536, s, s, nz_halves_per_token=44)
    c2.prefill(k.cuda(), v.cuda())
    assert c2.regrow_events == 0
    ks, vs = [k.numpy()], [v.numpy()]
    for i in range(720):
        kn = torch.randn(b, hkv, 1, 128, generator=gen).to(torch.float16)
        vn = torch.randn(b, hkv, 1, 128, generator=gen).to(torch.float16)
        qn = torch.randn(b, hkv * g, 1, 128, generator=gen).to(torch.float16)
        ks.append(kn.numpy()); vs.append(vn.numpy())
        out = c2.decode_step(qn.cuda(), kn.cuda(), vn.cuda())
    assert c2.regrow_events == 1 and c2.comp_len == 1280 and c2.win_len == 40
    c2.check_overflow()
    kfull, vfull = np.concatenate(ks, 2), np.concatenate(vs, 2)
    Lc = c2.comp_len
    kfull[:, :, :Lc] = O.prune_rows(kfull[:, :, :Lc], s)
    vfull[:, :, :Lc] = O.prune_rows(vfull[:, :, :Lc], s)
    d = np.abs(out.float().cpu().numpy() - O.masked_dense_attention(qn.numpy(), kfull, vfull).astype(np.float32))
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS
    kc = c2.as_reference_tuple()[0]
    rb, ra, rp = O.convert_key_batched(kfull[0, :, :Lc])
    assert np.array_equal(kc[0].cpu().numpy(), rb) and np.array_equal(kc[1].cpu().numpy(), ra)
    assert all(np.array_equal(a.cpu().numpy().view(np.uint16), r.view(np.uint16)) for a, r in zip(kc[2], rp))


def test_layer_batched_step_equals_per_layer_steps():
    """mfb200_decode_step_layers (one FFI call for a whole decoder's attention path) == per-layer decode_step."""
    from mustafar_b200.attention import MustafarKVCache, decode_step_layers
    layers, b, hkv, g, T, s = 3, 2, 4, 2, 700, 0.5
    a, bb = [], []
    for l in range(layers):
        k, v = _randn_gpu((b, hkv, T, 128), 100 + l), _randn_gpu((b, hkv, T, 128), 200 + l)
        for lst in (a, bb):
            c = MustafarKVCache(b, hkv, g, T + 600, s, s)
            c.prefill(k, v)
            lst.append(c)
    out_a = torch.empty(layers, b, hkv * g, 128, device="cuda", dtype=torch.float16)
    for i in range(300):  # crosses one compression event (window 188 -> 288)
        q = _randn_gpu((layers, b, hkv * g, 128), 1000 + i)
        kn, vn = _randn_gpu((layers, b, hkv, 128), 2000 + i), _randn_gpu((layers, b, hkv, 128), 3000 + i)
        decode_step_layers(a, q, kn, vn, out_a)
        for l in range(layers):
            o = bb[l].decode_step(q[l].view(b, hkv * g, 1, 128), kn[l].view(b, hkv, 1, 128), vn[l].view(b, hkv, 1, 128))
            assert torch.equal(o.view(b, hkv * g, 128), out_a[l]), (i, l)
    assert all(x.comp_len == y.comp_len and x.win_len == y.win_len for x, y in zip(a, bb)) and a[0].comp_len == 768


@pytest.mark.gpu
@pytest.mark.parametrize("world,hkv,groups", [(2, 4, 1), (4, 8, 4)])
def test_head_sharded_peer_output_equals_all_gather(world, hkv, groups):
    """Head-sharded decode without a collective (partition.PeerOutput): every rank's launch stores its rows into all ranks'
    gathered buffers and raises arrival flags.  All ranks live in ONE process on one GPU here (same kernel code path, no
    IPC); tools/peer_check.py runs the real thing, one process per GPU over cudaIpc."""
    from mustafar_b200.attention import MustafarKVCache
    from mustafar_b200.partition import PeerOutput, make_partition, shard_kv, shard_q
    b, T, s = 1, 700, 0.5
    g = torch.Generator().manual_seed(world)
    k = torch.randn(b, hkv, T, 128, generator=g).half().cuda()
    v = torch.randn(b, hkv, T, 128, generator=g).half().cuda()
    parts = [make_partition(b, hkv, world, r) for r in range(world)]
    assert all(p.mode == "head" for p in parts)
    caches = []
    for p in parts:
        c = MustafarKVCache(b, p.local_kv_heads, groups, T + 300, s, s)
        c.prefill(shard_kv(p, k).contiguous(), shard_kv(p, v).contiguous())
        caches.append(c)
    peers = PeerOutput.local_group(parts, b, hkv * groups, groups, "cuda")
    for step in range(5):  # more steps than buffers: flags and epochs are reused
        q = torch.randn(b, hkv * groups, 1, 128, generator=g).half().cuda()
        kn = torch.randn(b, hkv, 1, 128, generator=g).half().cuda()
        vn = torch.randn(b, hkv, 1, 128, generator=g).half().cuda()
        outs = []
        for p, c, po in zip(parts, caches, peers):
            po.bind(c, step)
            outs.append(c.decode_step(shard_q(p, q, groups).contiguous(), shard_kv(p, kn).contiguous(), shard_kv(p, vn).contiguous()))
        want = torch.cat(outs, dim=1)  # what the all-gather would deliver
        for po in peers:
            po.wait(step)
            assert torch.equal(po.gathered(step), want)
        torch.cuda.synchronize()
    assert not any(po.timed_out() for po in peers)
    for po in peers:
        po.close()


@pytest.mark.gpu
@pytest.mark.parametrize("b,hkv,groups,sparsity,T0,steps,layers,captures,comp_end", [
    (1, 8, 1, 0.5, 330, 300, 3, 2, 512), (2, 2, 4, 0.7, 330, 300, 3, 2, 512),
    (16, 8, 4, 0.7, 8192 - 20, 60, 2, 2, 8192),  # BASELINE config 3's geometry: the flat plan, compression at step 52
])
def test_decode_step_graph_equals_eager_steps(b, hkv, groups, sparsity, T0, steps, layers, captures, comp_end):
    """A decode step as ONE CUDA-graph launch with the window length in device memory (attention.DecodeStepGraph) against the
    eager per-layer steps: bit-identical outputs and cache contents over 300 steps - through every window-chunk count
    (the graph is planned for the window's capacity) and a compression event (the graph is captured again)."""
    from mustafar_b200.attention import DecodeStepGraph, MustafarKVCache
    gen = torch.Generator().manual_seed(11)
    caches_a, caches_b = [], []
    for _ in range(layers):
        k = torch.randn(b, hkv, T0, 128, generator=gen).half().cuda()
        v = torch.randn(b, hkv, T0, 128, generator=gen).half().cuda()
        for lst in (caches_a, caches_b):
            c = MustafarKVCache(b, hkv, groups, T0 + steps + 64, sparsity, sparsity)
            c.prefill(k, v)
            lst.append(c)
    q = torch.zeros(layers, b, hkv * groups, 128, dtype=torch.float16, device="cuda")
    kn = torch.zeros(layers, b, hkv, 128, dtype=torch.float16, device="cuda")
    vn = torch.zeros_like(kn)
    out = torch.zeros_like(q)
    graph = DecodeStepGraph(caches_b, q, kn, vn, out)
    for t in range(steps):
        q.copy_(torch.randn(q.shape, generator=gen).half())
        kn.copy_(torch.randn(kn.shape, generator=gen).half())
        vn.copy_(torch.randn(vn.shape, generator=gen).half())
        want = torch.stack([c.decode_step(q[l].view(b, -1, 1, 128), kn[l].view(b, hkv, 1, 128), vn[l].view(b, hkv, 1, 128)).view(b, -1, 128)
                            for l, c in enumerate(caches_a)])
        got = graph.step()
        assert torch.equal(got, want), (t, (got.float() - want.float()).abs().max().item())
    assert graph.captures == captures  # one compression event inside the run (small cases: window 42 -> 288 at step 246)
    for ca, cb in zip(caches_a, caches_b):
        assert ca.comp_len == cb.comp_len == comp_end and ca.win_len == cb.win_len
        assert torch.equal(ca.k_win[:, : ca.win_len], cb.k_win[:, : cb.win_len])
        assert torch.equal(ca.k.idx, cb.k.idx) and torch.equal(ca.v.bmp, cb.v.bmp)


@pytest.mark.gpu
def test_head_sharded_peer_output_across_processes():
    """The real thing: one process per GPU, cudaIpc-mapped gathered buffers, NCCL all-gather as the reference
    (tools/peer_check.py under torchrun).  Needs two GPUs on the box; skipped otherwise."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(root, "tools", "peer_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bit-equal to the all-gather on every rank: True" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
