"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
  (1) the numpy oracle (bit-exact for prune / bitmaps / counts / packed values),
  (2) the committed golden vectors produced by the reference's own code,
  (3) the reference's own CUDA kernels compiled for sm_100a (oracle/_ref), and
  (4) the masked-dense attention restatement,
with the north-star tolerances for attention: 2e-3 max-abs, 1e-3 mean-abs (fp16).
"""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import mustafar_oracle as O
from oracle import ref_cuda

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COMPRESS = sorted(glob.glob(os.path.join(GOLDEN, "compress_*.npz")))
PRUNE = sorted(glob.glob(os.path.join(GOLDEN, "prune_*.npz")))

MAX_ABS, MEAN_ABS = 2e-3, 1e-3


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).to(torch.float16)


# --------------------------------------------------------------------------- prune
@pytest.mark.parametrize("path", PRUNE, ids=[os.path.basename(p) for p in PRUNE])
def test_prune_golden(path):
    from mustafar_b200 import pruning
    g = np.load(path)
    x = _dev(g["x"])
    y = pruning.dh_prune_key(x.view(1, *x.shape), float(g["sparsity"]))
    assert np.array_equal(_bits(y.cpu().numpy().reshape(g["y"].shape)), _bits(g["y"]))


@pytest.mark.parametrize("sparsity", [0.0, 0.3, 0.5, 0.7, 0.9, 0.99])
def test_prune_vs_oracle_large(sparsity):
    from mustafar_b200 import pruning
    x = _randn((4, 8, 1024, 128), 100 + int(sparsity * 100))
    x[0, 0, :7] = 0  # all-zero rows
    x[0, 1, :5] = 1.0  # all-equal rows: every tie survives
    x[1, 2, 3, ::2] = -0.0
    y = pruning.dh_prune_value(x.cuda(), sparsity).cpu().numpy()
    ref = O.prune_rows(x.numpy(), sparsity)
    assert np.array_equal(_bits(y), _bits(ref))


@pytest.mark.parametrize("rows", [1, 3, 5, 31, 33, 1027])
def test_prune_row_counts_and_threshold_jumps(rows):
    """A warp prunes 4 consecutive rows and starts each select from the previous row's threshold: row counts that leave partial
    warps / partial CTAs, and neighbouring rows whose thresholds are far apart (the hint misses: full search), tiny or huge."""
    from mustafar_b200 import pruning
    x = _randn((1, 1, rows, 128), rows)
    scale = torch.tensor([1.0, 1e-4, 3e4, 1.0, 6e-8, 1.0, 255.0, 0.0])[torch.arange(rows) % 8]
    x = (x.float() * scale[None, None, :, None]).half()  # includes subnormals, values near the fp16 maximum and all-zero rows
    for s in (0.5, 0.7, 0.01):
        y = pruning.dh_prune_key(x.cuda(), s).cpu().numpy()
        assert np.array_equal(_bits(y), _bits(O.prune_rows(x.numpy(), s))), (rows, s)


def test_prune_matches_torch_kthvalue_formulation():
    """The reference's literal torch expression (llama_mustafar_kernel.py:97-110), evaluated on the GPU."""
    from mustafar_b200 import pruning
    x = _randn((2, 4, 512, 128), 5).cuda()
    for s in (0.5, 0.7):
        k = max(1, int(s * 128))
        flat = x.reshape(-1, 128)
        thr, _ = torch.kthvalue(torch.abs(flat), k, dim=-1, keepdim=True)
        ref = (flat * (torch.abs(flat) >= thr)).view(x.shape)
        got = pruning.dh_prune_key(x, s)
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16))


# --------------------------------------------------------------------------- compression
@pytest.mark.parametrize("path", COMPRESS, ids=[os.path.basename(p) for p in COMPRESS])
def test_compress_golden(path):
    from mustafar_b200 import compression
    g = np.load(path)
    s = float(g["sparsity"])
    xp = _dev(g["pruned"])
    for tag, fn in (("k", compression.convert_key_batched), ("v", compression.convert_value_batched)):
        bmp, acc, packed = fn(xp)
        assert bmp.dtype == torch.int64 and acc.dtype == torch.int32
        assert np.array_equal(bmp.cpu().numpy(), g[f"{tag}_bitmaps"])
        assert np.array_equal(acc.cpu().numpy(), g[f"{tag}_accum"])
        assert [p.numel() for p in packed] == list(g[f"{tag}_packed_len"])
        flat = torch.cat(packed).cpu().numpy() if packed else np.zeros(0, np.float16)
        assert np.array_equal(_bits(flat), _bits(g[f"{tag}_packed"]))
    if s >= 0:  # fused prune + compress from the raw input
        x = _dev(g["x"])
        for tag, fn in (("k", compression.prune_convert_key_batched), ("v", compression.prune_convert_value_batched)):
            bmp, acc, packed = fn(x, s)
            assert np.array_equal(bmp.cpu().numpy(), g[f"{tag}_bitmaps"])
            assert np.array_equal(acc.cpu().numpy(), g[f"{tag}_accum"])
            flat = torch.cat(packed).cpu().numpy()
            assert np.array_equal(_bits(flat), _bits(g[f"{tag}_packed"]))


@pytest.mark.parametrize("shape,sparsity", [((32, 256, 128), 0.5), ((8, 3840, 128), 0.5), ((4, 7936, 128), 0.7),
                                            ((3, 64, 128), 0.7), ((1, 1088, 128), 0.5)])
def test_compress_vs_oracle(shape, sparsity):
    from mustafar_b200 import compression
    x = _randn(shape, shape[0] * 7 + shape[1])
    x[0, :3] = 0
    xp = O.prune_rows(x.numpy(), sparsity)
    for ofn, fn, ffn in ((O.convert_key_batched, compression.convert_key_batched, compression.prune_convert_key_batched),
                         (O.convert_value_batched, compression.convert_value_batched, compression.prune_convert_value_batched)):
        rb, ra, rp = ofn(xp)
        for got in (fn(_dev(xp)), ffn(x.cuda(), sparsity)):
            bmp, acc, packed = got
            assert np.array_equal(bmp.cpu().numpy(), rb)
            assert np.array_equal(acc.cpu().numpy(), ra)
            assert all(np.array_equal(_bits(p.cpu().numpy()), _bits(r)) for p, r in zip(packed, rp))


def test_compress_empty_and_errors():
    from mustafar_b200 import compression
    bmp, acc, packed = compression.convert_key_batched(torch.zeros((2, 64, 128), dtype=torch.float16, device="cuda"))
    assert int(bmp.abs().sum()) == 0 and int(acc.abs().sum()) == 0 and all(p.numel() == 0 for p in packed)
    with pytest.raises(AssertionError):
        compression.convert_key_batched(torch.zeros((2, 65, 128), dtype=torch.float16, device="cuda"))
    with pytest.raises(RuntimeError):
        compression.convert_key_batched(torch.zeros((2, 64, 128), dtype=torch.float16))  # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        compression.convert_value_batched(torch.zeros((2, 64, 128), dtype=torch.float32, device="cuda"))


# --------------------------------------------------------------------------- SpMV operators
def _compressed(bk, L, sparsity, seed):
    from mustafar_b200 import compression
    k = torch.from_numpy(O.prune_rows(_randn((bk, L, 128), seed).numpy(), sparsity)).cuda()
    v = torch.from_numpy(O.prune_rows(_randn((bk, L, 128), seed + 1).numpy(), sparsity)).cuda()
    kb, ki, kn = compression.convert_key_batched(k)
    vb, vi, vn = compression.convert_value_batched(v)
    return k, v, [kb, ki, kn, compression.nz_offsets(ki)], [vb, vi, vn, compression.nz_offsets(vi)]


@pytest.mark.parametrize("bk,groups,L,sparsity,full_rows", [(4, 1, 256, 0.5, False), (2, 4, 512, 0.7, False),
                                                             (3, 2, 1024, 0.5, True), (32, 1, 3840, 0.5, False)])
def test_spmv_ops(bk, groups, L, sparsity, full_rows):
    from mustafar_b200 import mustafar_package as mp
    k, v, kc, vc = _compressed(bk, L, sparsity, 11 * bk + L)
    bq = bk * groups
    q = _randn((bq, 8, 128), 3).cuda()
    p = torch.softmax(_randn((bq, 8, L), 4).float(), -1).to(torch.float16).cuda()
    if not full_rows:  # the shapes the model uses: rows 1..7 are F.pad zeros
        q[:, 1:] = 0
        p[:, 1:] = 0
    nzk, nzv = ref_cuda.pad_nz(kc[2]), ref_cuda.pad_nz(vc[2])
    got_k = mp.mustafar_key_formulation(kc[0], nzk, kc[1].reshape(-1), kc[3], q, L, 128, bq, groups)
    ws = torch.zeros(1, dtype=torch.float16, device="cuda")
    got_v = mp.mustafar_value_formulation(vc[0], nzv, vc[1].reshape(-1), vc[3], p, ws, 128, L, bq, groups)
    assert got_k.shape == (bq, 8, L) and got_v.shape == (bq, 8, 128)
    # exact math in float64 on the same pruned tensors
    kf = k.double().cpu().repeat_interleave(groups, 0)
    vf = v.double().cpu().repeat_interleave(groups, 0)
    exact_k = torch.matmul(q.double().cpu(), kf.transpose(1, 2))
    exact_v = torch.matmul(p.double().cpu(), vf)
    # fp32 accumulation + one fp16 rounding: within half an fp16 ulp (+ accumulation noise) of exact
    assert torch.allclose(got_k.double().cpu(), exact_k, rtol=2e-3, atol=2e-3)
    assert torch.allclose(got_v.double().cpu(), exact_v, rtol=2e-3, atol=1e-4)
    if ref_cuda.available() and L % 256 == 0:
        ref_k = ref_cuda.key_formulation(kc[0], nzk, kc[1].reshape(-1), kc[3], q, L, 128, bq, groups)
        ref_v = ref_cuda.value_formulation(vc[0], nzv, vc[1].reshape(-1), vc[3], p, 128, L, bq, groups)
        # same inputs, both fp32-accumulate + fp16 round: differ by accumulation order only
        dk = (got_k.float() - ref_k.float()).abs()
        assert dk.max() <= 0.0625 and (dk > 0).float().mean() < 0.05, (dk.max(), (dk > 0).float().mean())
        dv = (got_v.float() - ref_v.float()).abs()
        assert dv.max() <= 2e-3 and dv.mean() <= 1e-4, (dv.max(), dv.mean())


def test_spmv_argument_errors():
    from mustafar_b200 import mustafar_package as mp
    k, v, kc, vc = _compressed(2, 256, 0.5, 1)
    q = torch.zeros((2, 8, 128), dtype=torch.float16, device="cuda")
    nz = torch.cat(kc[2])
    with pytest.raises(RuntimeError):
        mp.mustafar_key_formulation(kc[0].int(), nz, kc[1].reshape(-1), kc[3], q, 256, 128, 2, 1)
    with pytest.raises(RuntimeError):
        mp.mustafar_key_formulation(kc[0], nz.float(), kc[1].reshape(-1), kc[3], q, 256, 128, 2, 1)
    with pytest.raises(RuntimeError):
        mp.mustafar_key_formulation(kc[0], nz, kc[1].reshape(-1), kc[3], q.cpu(), 256, 128, 2, 1)
    with pytest.raises(RuntimeError):
        mp.mustafar_key_formulation(kc[0], nz, kc[1].reshape(-1), kc[3], q, 200, 128, 2, 1)  # M % 64 != 0


# --------------------------------------------------------------------------- fused decode attention
def _attention_case(b, hkv, groups, T, sparsity, seed, residual=32):
    from mustafar_b200.attention import MustafarKVCache
    hq = hkv * groups
    k = _randn((b, hkv, T, 128), seed)
    v = _randn((b, hkv, T, 128), seed + 1)
    q = _randn((b, hq, 1, 128), seed + 2)
    cache = MustafarKVCache(b, hkv, groups, max_tokens=T + 600, k_sparsity=sparsity, v_sparsity=sparsity,
                            residual_length=residual)
    cache.prefill(k.cuda(), v.cuda())
    L = cache.comp_len
    kp, vp = k.numpy().copy(), v.numpy().copy()
    if L:
        kp[:, :, :L] = O.prune_rows(kp[:, :, :L], sparsity)
        vp[:, :, :L] = O.prune_rows(vp[:, :, :L], sparsity)
    return cache, q, kp, vp, L


def _check_attention(got, q, kp, vp, L, mask=None):
    got = got.float().cpu().numpy()
    glue = O.decode_attention_glue(q.numpy(), kp[:, :, :L], kp[:, :, L:], vp[:, :, :L], vp[:, :, L:], mask).astype(np.float32)
    dense = O.masked_dense_attention(q.numpy(), kp, vp, mask).astype(np.float32)
    for ref in (glue, dense):
        d = np.abs(got - ref)
        assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (d.max(), d.mean())


@pytest.mark.parametrize("b,hkv,groups,T,sparsity", [
    (1, 2, 1, 300, 0.5), (2, 2, 1, 1000, 0.5), (1, 4, 4, 1312, 0.7), (1, 1, 8, 576, 0.7), (2, 1, 2, 832, 0.5),
    (1, 32, 1, 4096, 0.5),  # BASELINE config 1
    (1, 2, 1, 40, 0.5), (1, 1, 1, 1, 0.5), (1, 2, 4, 287, 0.7),  # nothing compressed yet: window only
])
def test_fused_attention_vs_oracle(b, hkv, groups, T, sparsity):
    cache, q, kp, vp, L = _attention_case(b, hkv, groups, T, sparsity, seed=b * 1000 + T)
    assert L == O.compressed_length(T)
    got = cache.attend(q.cuda())
    assert got.shape == (b, hkv * groups, 1, 128)
    _check_attention(got, q, kp, vp, L)
    # second launch on the same workspace (ticket counters must have been reset)
    got2 = cache.attend(q.cuda())
    assert torch.equal(got, got2)


@pytest.mark.gpu
@pytest.mark.parametrize("groups,mode", [(1, "all"), (4, "all"), (2, "some"), (4, "some"), (8, "some")])
def test_blocks_larger_than_their_ring_slot(groups, mode):
    """Tie explosions: rows whose magnitudes are all equal keep EVERY element (`|x| >= k-th smallest`), so a 64-token block
    carries up to 16 KB of nonzeros - more than the sparsity-sized ring slot.  Such blocks are decoded straight from global
    memory (the NZ_SHARED = false paths of both the CUDA-core and the HMMA decode); `some` mixes them with normal blocks."""
    from mustafar_b200.attention import MustafarKVCache
    b, hkv, T, s = 2, 2, 712, 0.5
    gen = torch.Generator().manual_seed(groups)
    k = torch.randn(b, hkv, T, 128, generator=gen).half()
    v = torch.randn(b, hkv, T, 128, generator=gen).half()
    sign = lambda: (torch.randint(0, 2, (b, hkv, T, 128), generator=gen) * 2 - 1).half()
    tie_k, tie_v = sign() * 0.5, sign() * 0.25
    if mode == "all":
        k, v = tie_k, tie_v
    else:  # token blocks 2, 3 and 7 of every head are all-ties, the others are ordinary
        for blk in (2, 3, 7):
            k[:, :, 64 * blk: 64 * blk + 64] = tie_k[:, :, 64 * blk: 64 * blk + 64]
            v[:, :, 64 * blk: 64 * blk + 64] = tie_v[:, :, 64 * blk: 64 * blk + 64]
    q = torch.randn(b, hkv * groups, 1, 128, generator=gen).half()
    cache = MustafarKVCache(b, hkv, groups, max_tokens=T + 300, k_sparsity=s, v_sparsity=s)
    cache.prefill(k.cuda(), v.cuda())
    L = cache.comp_len
    assert L == 512 and cache.slot_kb < 16
    kp, vp = k.numpy().copy(), v.numpy().copy()
    kp[:, :, :L] = O.prune_rows(kp[:, :, :L], s)
    vp[:, :, :L] = O.prune_rows(vp[:, :, :L], s)
    if mode == "all":
        assert np.count_nonzero(kp[:, :, :L]) == kp[:, :, :L].size  # nothing was pruned
    got = cache.attend(q.cuda())
    _check_attention(got, q, kp, vp, L)
    kn, vn = torch.randn(b, hkv, 1, 128, generator=gen).half(), torch.randn(b, hkv, 1, 128, generator=gen).half()
    got = cache.decode_step(q.cuda(), kn.cuda(), vn.cuda())
    kp, vp = np.concatenate([kp, kn.numpy()], 2), np.concatenate([vp, vn.numpy()], 2)
    _check_attention(got, q, kp, vp, L)
    cache.check_overflow()


def _random_geometries(n, seed):
    """Seeded sweep over the planner's regimes: more units than CTA slots, ragged per-unit splits, the boundary
    between the flagged and the ticket merge (16 blocks per compressed CTA), tiny and full windows, every G."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        g = int(rng.choice([1, 1, 2, 4, 8]))
        hkv = int(rng.choice([1, 2, 3, 8, 32]))
        b = int(rng.choice([1, 1, 2, 5, 16]))
        while b * hkv * g > 512:
            b = max(1, b // 2)
        blocks = int(rng.choice([1, 2, 4, 17, 33, 70]))
        while b * hkv * blocks * 64 > 160000:  # keep the numpy oracle fast
            blocks = max(1, blocks // 2)
        T = 32 + blocks * 64 + int(rng.integers(0, 256))  # L = 256 * floor(...) <= blocks * 64, window 32..287+
        out.append((b, hkv, g, T, float(rng.choice([0.5, 0.7, 0.9]))))
    return out


@pytest.mark.parametrize("b,hkv,groups,T,sparsity", _random_geometries(14, seed=2024) + [
    (16, 32, 1, 352, 0.5),    # 512 units: more units than resident CTA slots, one compressed CTA per unit
    (1, 1, 1, 2600, 0.7),     # a single unit cut into one-block splits (flagged merge, 40+ partials)
    (2, 32, 1, 8736, 0.7),    # uniform plan with 23 blocks per CTA: ticket merge (just below the flat threshold)
])
def test_fused_attention_random_geometries(b, hkv, groups, T, sparsity):
    cache, q, kp, vp, L = _attention_case(b, hkv, groups, T, sparsity, seed=7 * T + b)
    _check_attention(cache.attend(q.cuda()), q, kp, vp, L)
    # three fused decode steps (append + attention) on top, checked against the grown dense cache
    g = torch.Generator().manual_seed(T)
    for _ in range(3):
        kn = torch.randn(b, hkv, 1, 128, generator=g).to(torch.float16)
        vn = torch.randn(b, hkv, 1, 128, generator=g).to(torch.float16)
        qn = torch.randn(b, hkv * groups, 1, 128, generator=g).to(torch.float16)
        out = cache.decode_step(qn.cuda(), kn.cuda(), vn.cuda())
        kp, vp = np.concatenate([kp, kn.numpy()], axis=2), np.concatenate([vp, vn.numpy()], axis=2)
        if cache.comp_len != L:  # the step crossed the compression schedule: those 256 rows are pruned from now on
            kp[:, :, L:cache.comp_len] = O.prune_rows(kp[:, :, L:cache.comp_len], sparsity)
            vp[:, :, L:cache.comp_len] = O.prune_rows(vp[:, :, L:cache.comp_len], sparsity)
            L = cache.comp_len
    # the last output was computed before that step's (possible) compression
    dense = O.masked_dense_attention(qn.numpy(), kp, vp).astype(np.float32)
    d = np.abs(out.float().cpu().numpy() - dense)
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (d.max(), d.mean())
    cache.check_overflow()


def test_fused_attention_mask():
    b, hkv, groups, T = 2, 2, 2, 700
    cache, q, kp, vp, L = _attention_case(b, hkv, groups, T, 0.5, seed=77)
    mask = np.zeros((b, 1, 1, T), dtype=np.float16)
    mask[0, :, :, :100] = np.finfo(np.float16).min  # left padding of sequence 0
    mask[1, :, :, 500:650] = np.finfo(np.float16).min
    got = cache.attend(q.cuda(), torch.from_numpy(mask).cuda())
    _check_attention(got, q, kp, vp, L, mask)


@pytest.mark.parametrize("groups,sparsity", [(1, 0.5), (4, 0.7)])
def test_fused_vs_reference_cuda_pipeline(groups, sparsity):
    """Same compressed cache through (a) the reference's CUDA kernels + its torch glue and (b) the fused kernel."""
    if not ref_cuda.available():
        pytest.skip("oracle/_ref not built")
    b, hkv, T = 2, 4, 1312
    cache, q, kp, vp, L = _attention_case(b, hkv, groups, T, sparsity, seed=9)
    kc, kw, vc, vw, L2, _ = cache.as_reference_tuple()
    assert L2 == L and L % 256 == 0
    ref = ref_cuda.decode_step(q.cuda(), kc, kw, vc, vw, L, groups)
    got = cache.attend(q.cuda())
    d = (got.float() - ref.float()).abs()
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (d.max(), d.mean())
    # and the reference glue driven by OUR drop-in SpMV ops gives the same
    from mustafar_b200 import mustafar_package as mp
    ws = torch.zeros(1, dtype=torch.float16, device="cuda")
    ours = ref_cuda.decode_step(q.cuda(), kc, kw, vc, vw, L, groups, key_op=mp.mustafar_key_formulation,
                                value_op=lambda bmp, nz, idx, off, B, m, kk, bs, g: mp.mustafar_value_formulation(
                                    bmp, nz, idx, off, B, ws, m, kk, bs, g))
    d2 = (ours.float() - ref.float()).abs()
    assert d2.max() <= MAX_ABS and d2.mean() <= MEAN_ABS, (d2.max(), d2.mean())


def test_cache_matches_whole_compression_and_reference_container():
    """prefill + 2 appends of 256 tokens == compressing everything at once (SURVEY App. A append rule)."""
    from mustafar_b200 import compression
    b, hkv, groups, T0 = 1, 3, 1, 300
    cache, q, kp, vp, L = _attention_case(b, hkv, groups, T0, 0.5, seed=5)
    g = torch.Generator().manual_seed(123)
    ks, vs = [], []
    steps = 2 * 256 + 20
    for _ in range(steps):
        kn = torch.randn(b, hkv, 1, 128, generator=g).to(torch.float16)
        vn = torch.randn(b, hkv, 1, 128, generator=g).to(torch.float16)
        qn = torch.randn(b, hkv * groups, 1, 128, generator=g).to(torch.float16)
        ks.append(kn)
        vs.append(vn)
        out = cache.decode_step(qn.cuda(), kn.cuda(), vn.cuda())
    T = T0 + steps
    assert cache.kv_seq_len == T and cache.comp_len == 256 + 512 and cache.win_len == T - 768
    kfull = np.concatenate([kp] + [x.numpy() for x in ks], axis=2)
    vfull = np.concatenate([vp] + [x.numpy() for x in vs], axis=2)
    Lc = cache.comp_len
    kfull[:, :, L:Lc] = O.prune_rows(kfull[:, :, L:Lc], 0.5)
    vfull[:, :, L:Lc] = O.prune_rows(vfull[:, :, L:Lc], 0.5)
    kc, kw, vc, vw, L2, t2 = cache.as_reference_tuple()
    rb, ra, rp = O.convert_key_batched(kfull[0, :, :Lc])
    assert np.array_equal(kc[0].cpu().numpy(), rb) and np.array_equal(kc[1].cpu().numpy(), ra)
    assert all(np.array_equal(_bits(a.cpu().numpy()), _bits(r)) for a, r in zip(kc[2], rp))
    assert np.array_equal(kc[3].cpu().numpy(), O.nz_offsets(ra))
    rb, ra, rp = O.convert_value_batched(vfull[0, :, :Lc])
    assert np.array_equal(vc[0].cpu().numpy(), rb) and np.array_equal(vc[1].cpu().numpy(), ra)
    assert all(np.array_equal(_bits(a.cpu().numpy()), _bits(r)) for a, r in zip(vc[2], rp))
    assert np.array_equal(_bits(kw.cpu().numpy()), _bits(kfull[:, :, Lc:]))
    # the last step's output was computed BEFORE that step's (non-)compression: window 276 incl. new token
    dense = O.masked_dense_attention(qn.numpy(), kfull, vfull).astype(np.float32)
    d = np.abs(out.float().cpu().numpy() - dense)
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS


def _check_prompt_streams(cache, kp, vp, L, ks, vs):
    """cache streams == oracle compression of the pruned prompt rows [0, L) (bit-exact)."""
    kc, _, vc, _, L2, _ = cache.as_reference_tuple()
    assert L2 == L
    b, hkv = kp.shape[:2]
    kpr = O.prune_rows(kp[:, :, :L], ks).reshape(b * hkv, L, 128)
    vpr = O.prune_rows(vp[:, :, :L], vs).reshape(b * hkv, L, 128)
    for got, (rb, ra, rp) in ((kc, O.convert_key_batched(kpr)), (vc, O.convert_value_batched(vpr))):
        assert np.array_equal(got[0].cpu().numpy(), rb)
        assert np.array_equal(got[1].cpu().numpy(), ra)
        assert all(np.array_equal(_bits(a.cpu().numpy()), _bits(r)) for a, r in zip(got[2], rp))
        assert np.array_equal(got[3].cpu().numpy(), O.nz_offsets(ra))


@pytest.mark.parametrize("b,hkv,T,ks,vs,layout", [
    (2, 3, 256 * 3 + 40, 0.5, 0.7, "bhtd"),     # a few blocks, different K / V sparsity
    (1, 2, 64 * 70 + 32, 0.7, 0.5, "bthd"),     # look-back across more than 32 predecessors; token-major source
    (3, 2, 1024 + 32, 0.0, 0.9, "sliced"),      # no pruning for K; source = a slice of a longer buffer
])
def test_prefill_single_pass_bit_exact(b, hkv, T, ks, vs, layout):
    """mfb200_compress_prefill (one launch, strided input, look-back offsets) vs the oracle's prune + convert."""
    from mustafar_b200.attention import MustafarKVCache
    kp, vp = _randn((b, hkv, T, 128), 21).numpy(), _randn((b, hkv, T, 128), 22).numpy()
    kp[0, 0, 5] = 0.0          # an all-zero row (every tile bit 0 there)
    kp[0, 1, 7, :64] = 0.25    # ties at the threshold: all kept
    kt, vt = torch.from_numpy(kp).cuda(), torch.from_numpy(vp).cuda()
    if layout == "bthd":       # [B, T, H, D] storage viewed as [B, H, T, D]
        kt = kt.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
        vt = vt.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
        assert not kt.is_contiguous()
    elif layout == "sliced":
        kt = torch.cat([kt, torch.zeros_like(kt)], dim=2)[:, :, :T]
        vt = torch.cat([vt, torch.zeros_like(vt)], dim=2)[:, :, :T]
        assert not kt.is_contiguous()
    cache = MustafarKVCache(b, hkv, 1, T + 300, ks, vs)
    cache.prefill(kt, vt)
    L = O.compressed_length(T)
    assert cache.comp_len == L and cache.win_len == T - L
    _check_prompt_streams(cache, kp, vp, L, ks, vs)
    cache.check_overflow()


def test_compression_with_thresholds_that_jump_between_rows():
    """The select inside the compression kernels starts from the previous row's threshold (a 1024-ulp window) and falls back to
    the full search when the answer is elsewhere: rows whose magnitudes differ by orders of magnitude, subnormal rows, rows near
    the fp16 maximum and all-zero rows, through the prefill launch, the decode-time chunk append and the two-pass list API."""
    from mustafar_b200 import compression
    from mustafar_b200.attention import MustafarKVCache
    b, hkv, T, s = 2, 2, 256 * 2 + 40, 0.5
    scale = torch.tensor([1.0, 1e-4, 3e4, 1.0, 6e-8, 1.0, 255.0, 0.0, 1.0, 1.0, 0.01])[torch.arange(T + 256) % 11]
    mk = lambda seed: (_randn((b, hkv, T + 256, 128), seed).float() * scale[None, None, :, None]).half()
    k, v = mk(61), mk(62)
    cache = MustafarKVCache(b, hkv, 1, T + 600, s, s)
    cache.prefill(k[:, :, :T].cuda(), v[:, :, :T].cuda())
    L = O.compressed_length(T)
    _check_prompt_streams(cache, k[:, :, :T].numpy(), v[:, :, :T].numpy(), L, s, s)
    for t in range(T, T + 256):  # the window reaches 288 rows: one chunk append
        cache.append(k[:, :, t:t + 1].cuda(), v[:, :, t:t + 1].cuda())
        cache.maybe_compress()
    assert cache.comp_len == L + 256
    _check_prompt_streams(cache, k.numpy(), v.numpy(), L + 256, s, s)
    cache.check_overflow()
    flat = k[:, :, :512].reshape(b * hkv, 512, 128)
    bmp, acc, packed = compression.prune_convert_key_batched(flat.cuda(), s)
    rb, ra, rp = O.convert_key_batched(O.prune_rows(flat.numpy(), s))
    assert np.array_equal(bmp.cpu().numpy(), rb) and np.array_equal(acc.cpu().numpy(), ra)
    assert all(np.array_equal(_bits(a.cpu().numpy()), _bits(r)) for a, r in zip(packed, rp))


def test_prefill_long_chain_and_append_offset():
    """One unit with 1024 blocks (the look-back chain spans the whole grid), then a second compress call that
    appends at a non-zero tile offset == compressing everything at once."""
    from mustafar_b200.attention import MustafarKVCache
    b, hkv, T = 1, 1, 64 * 1024 + 32
    kp, vp = _randn((b, hkv, T, 128), 31).numpy(), _randn((b, hkv, T, 128), 32).numpy()
    kt, vt = torch.from_numpy(kp).cuda(), torch.from_numpy(vp).cuda()
    cache = MustafarKVCache(b, hkv, 1, T + 300, 0.5, 0.5)
    cache.prefill(kt, vt)
    L = O.compressed_length(T)
    _check_prompt_streams(cache, kp, vp, L, 0.5, 0.5)
    # two-stage: first 256 * 3 tokens, then the rest at tile offset 1536
    c2 = MustafarKVCache(b, hkv, 1, T + 300, 0.5, 0.5)
    L1 = 768
    c2._compress_prompt(kt, vt, L1)
    c2.comp_len = L1
    c2._compress_prompt(kt[:, :, L1:], vt[:, :, L1:], L - L1)
    c2.comp_len = L
    for a, r in ((c2.k, cache.k), (c2.v, cache.v)):
        assert torch.equal(a.bmp[:, : 2 * L], r.bmp[:, : 2 * L]) and torch.equal(a.idx[:, : 2 * L + 1], r.idx[:, : 2 * L + 1])
        n = int(r.idx[0, 2 * L].item()) * 2
        assert torch.equal(a.nz[:n].view(torch.int16), r.nz[:n].view(torch.int16))


def test_launch_refuses_a_workspace_that_is_too_small():
    import ctypes as C
    from mustafar_b200 import _lib
    cache, q, kp, vp, L = _attention_case(1, 4, 1, 1500, 0.5, seed=90)
    qd = q.cuda()
    ref = cache.attend(qd).clone()
    p = cache.make_params(qd.view(1, -1, 128), torch.empty_like(qd))
    assert p.workspace_kb > 0
    small = _lib.DecodeParams()
    C.memmove(C.byref(small), C.byref(p), C.sizeof(p))
    small.workspace_kb = 1
    rc = _lib.load().mfb200_sparse_decode_attention(C.byref(small), _lib.stream_ptr())
    assert rc < 0 and b"workspace" in _lib.load().mfb200_last_error()
    assert torch.equal(cache.attend(qd), ref)  # the refused launch touched nothing


def test_slab_overflow_is_detected_not_corrupting():
    """C-ABI level guard (callers that manage their own slabs): a tile that does not fit its slab is not written and the
    device flag is raised.  (MustafarKVCache never gets there: tests/test_gpu_configs.py::test_slabs_regrow_...)"""
    from mustafar_b200 import _lib, compression
    x = _randn((2, 256, 128), 1).cuda()
    bitmaps = torch.empty((2, 512), dtype=torch.int64, device="cuda")
    counts = torch.empty((2, 512), dtype=torch.int32, device="cuda")
    accum = torch.empty((2, 513), dtype=torch.int32, device="cuda")
    compression.compress_into(x, _lib.LAYOUT_VALUE, 64, bitmaps, counts, accum, 513, 0, None)
    cap = 4096  # halves per head: far less than the ~18K the chunk needs
    packed = torch.full((2 * cap + 64,), 7.0, dtype=torch.float16, device="cuda")
    overflow = torch.zeros((1,), dtype=torch.int32, device="cuda")
    base = torch.tensor([0, cap], dtype=torch.int64, device="cuda")
    compression.pack_into(x, _lib.LAYOUT_VALUE, bitmaps, accum, 513, 0, base, packed, cap, overflow)
    assert int(overflow.item()) == 1
    assert torch.all(packed[2 * cap:] == 7.0)  # nothing was written past the second head's slab


@pytest.mark.parametrize("b,hkv,groups,T,sparsity", [(4, 8, 4, 8192, 0.7), (2, 8, 8, 4160, 0.5), (1, 8, 4, 8192, 0.5),
                                                    (3, 16, 2, 4096, 0.7)])
def test_repeated_launches_are_stable_and_correct(b, hkv, groups, T, sparsity):
    """Many back-to-back launches on the same cache (ring slots are recycled, PDL overlaps launches): every
    launch must return bit-identical output, and that output must match the masked-dense oracle.  Regression
    test for a ring-slot race that showed up as sporadic NaNs in one query head of the G=4 build."""
    cache, q, kp, vp, L = _attention_case(b, hkv, groups, T, sparsity, seed=31 * b + T)
    qd = q.cuda()
    outs = [cache.attend(qd).clone() for _ in range(10)]
    torch.cuda.synchronize()
    assert not any(torch.isnan(o).any().item() for o in outs)
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    dense = O.masked_dense_attention(q.numpy(), kp, vp).astype(np.float32)
    d = np.abs(outs[-1].float().cpu().numpy() - dense)
    assert d.max() <= MAX_ABS and d.mean() <= MEAN_ABS, (d.max(), d.mean())


def test_decode_launches_replay_in_a_cuda_graph():
    """The fused launch is graph-capturable (single kernel, PDL attribute, self-resetting tickets): capture three
    launches on different caches, replay twice, compare with the eager launches."""
    import ctypes as C
    from mustafar_b200 import _lib
    lib = _lib.load()
    cases = [_attention_case(2, 4, g, 1500, 0.5, seed=70 + g) for g in (1, 4)] + [_attention_case(1, 8, 1, 4200, 0.7, seed=77)]
    params, outs, eager = [], [], []
    for cache, q, kp, vp, L in cases:
        qd = q.cuda()
        eager.append(cache.attend(qd).clone())
        o = torch.zeros_like(qd)
        p = cache.make_params(qd.view(cache.batch, -1, 128), o)
        p.flags |= _lib.F_PDL
        params.append((p, qd))
        outs.append(o)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for p, _ in params:  # warm-up on the capture stream (first-call attribute setup must not happen under capture)
            _lib.check(lib.mfb200_sparse_decode_attention(C.byref(p), side.cuda_stream), "warm-up")
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for p, _ in params:
            _lib.check(lib.mfb200_sparse_decode_attention(C.byref(p), torch.cuda.current_stream().cuda_stream), "capture")
    for _ in range(2):
        for o in outs:
            o.zero_()
        g.replay()
        torch.cuda.synchronize()
        for o, e in zip(outs, eager):
            assert torch.equal(o.view_as(e), e)


def test_flat_partition_mode():
    """Mid-size launches use the flat work partition (all blocks of all units divided evenly over the resident
    CTA slots; CTAs cross unit boundaries and process several segments).  Forced here on small shapes through
    mfb200_decode_params::plan_hint, plus one shape that selects it naturally; checked against the masked-dense oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = """
import sys, numpy as np, torch
sys.path.insert(0, %r)
from mustafar_b200.attention import MustafarKVCache
from oracle import mustafar_oracle as O
for (b, hkv, g, T, s) in SHAPES:
    gen = torch.Generator().manual_seed(T + g)
    k = torch.randn(b, hkv, T, 128, generator=gen).half(); v = torch.randn(b, hkv, T, 128, generator=gen).half()
    q = torch.randn(b, hkv * g, 1, 128, generator=gen).half()
    c = MustafarKVCache(b, hkv, g, T + 64, s, s, plan_hint=HINT); c.prefill(k.cuda(), v.cuda())
    outs = [c.attend(q.cuda()).clone() for _ in range(3)]
    torch.cuda.synchronize()
    assert all(torch.equal(outs[0], o) for o in outs[1:])
    L = c.comp_len
    kp = k.numpy().copy(); vp = v.numpy().copy()
    kp[:, :, :L] = O.prune_rows(kp[:, :, :L], s); vp[:, :, :L] = O.prune_rows(vp[:, :, :L], s)
    d = np.abs(outs[-1].float().cpu().numpy() - O.masked_dense_attention(q.numpy(), kp, vp).astype(np.float32))
    assert d.max() <= 2e-3 and d.mean() <= 1e-3, (b, hkv, g, T, s, d.max(), d.mean())
print("FLAT-OK")
""" % root
    cases = [("5", "[(1, 8, 1, 2112, 0.5), (1, 4, 4, 1312, 0.7)]"), ("12", "[(1, 8, 1, 2112, 0.5), (2, 4, 2, 1088, 0.5)]"),
             ("37", "[(1, 8, 4, 2112, 0.7), (1, 8, 8, 1088, 0.5)]"), (None, "[(8, 16, 1, 4160, 0.7)]")]
    for forced, shapes in cases:
        r = subprocess.run([sys.executable, "-c", "HINT = %s\nSHAPES = " % (forced or "0") + shapes + code], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "FLAT-OK" in r.stdout, (forced, r.stdout[-1500:] + r.stderr[-1500:])


def test_no_uninitialised_shared_memory_reads():
    """Runs attention cases through the debug build whose CTAs start by filling their dynamic shared memory
    with fp16 NaNs (make poison): any read of uninitialised / not-yet-published shared memory turns the
    output into NaN.  Runs in a subprocess because the library path is fixed at first load."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "mustafar_b200", "libmustafar_b200_poison.so")
    if not os.path.exists(lib):
        pytest.skip("poison build not present (make -C mustafar_b200/csrc poison)")
    code = """
import sys, numpy as np, torch
sys.path.insert(0, %r)
from mustafar_b200.attention import MustafarKVCache
from oracle import mustafar_oracle as O
for (b, hkv, g, T, s) in [(1, 8, 4, 4160, 0.5), (2, 4, 8, 2112, 0.7), (1, 32, 1, 4096, 0.5), (2, 8, 2, 2368, 0.7)]:
    gen = torch.Generator().manual_seed(T)
    k = torch.randn(b, hkv, T, 128, generator=gen).half(); v = torch.randn(b, hkv, T, 128, generator=gen).half()
    q = torch.randn(b, hkv * g, 1, 128, generator=gen).half()
    c = MustafarKVCache(b, hkv, g, T + 64, s, s); c.prefill(k.cuda(), v.cuda())
    outs = [c.attend(q.cuda()).clone() for _ in range(3)]
    kn = torch.randn(b, hkv, 1, 128, generator=gen).half(); vn = torch.randn(b, hkv, 1, 128, generator=gen).half()
    outs.append(c.decode_step(q.cuda(), kn.cuda(), vn.cuda()))
    torch.cuda.synchronize()
    assert not any(torch.isnan(o).any().item() for o in outs), (b, hkv, g, T, s)
    L = c.comp_len
    kp = np.concatenate([k.numpy(), kn.numpy()], 2); vp = np.concatenate([v.numpy(), vn.numpy()], 2)
    kp[:, :, :L] = O.prune_rows(kp[:, :, :L], s); vp[:, :, :L] = O.prune_rows(vp[:, :, :L], s)
    d = np.abs(outs[-1].float().cpu().numpy() - O.masked_dense_attention(q.numpy(), kp, vp).astype(np.float32))
    assert d.max() <= 2e-3 and d.mean() <= 1e-3, (d.max(), d.mean())
print("POISON-OK")
""" % root
    env = dict(os.environ, MFB200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "POISON-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_tcgen05_gqa_variant_still_correct():
    """The G >= 4 contraction on tcgen05 / TMEM (csrc/gqa_tc.cuh, `make tc`) lost the A/B against the register-fragment HMMA
    path and is not shipped, but it is kept buildable and correct: same cases, same tolerance, through its own library."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "mustafar_b200", "libmustafar_b200_tc.so")
    if not os.path.exists(lib):
        pytest.skip("tcgen05 A/B build not present (make -C mustafar_b200/csrc tc)")
    code = """
import sys, torch
sys.path.insert(0, %r)
from mustafar_b200.attention import MustafarKVCache
from oracle import torch_oracle as TO
for (b, hkv, g, T, s) in [(1, 8, 4, 4160, 0.5), (2, 4, 8, 2112, 0.7), (4, 8, 4, 8192, 0.7)]:
    gen = torch.Generator(device="cuda").manual_seed(T)
    k = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half(); v = torch.randn(b, hkv, T, 128, device="cuda", generator=gen).half()
    q = torch.randn(b, hkv * g, 1, 128, device="cuda", generator=gen).half()
    c = MustafarKVCache(b, hkv, g, T + 64, s, s); c.prefill(k, v)
    L = c.comp_len
    k[:, :, :L] = TO.prune_rows(k[:, :, :L], s); v[:, :, :L] = TO.prune_rows(v[:, :, :L], s)
    o = c.attend(q); torch.cuda.synchronize()
    d = (o.float() - TO.masked_dense_attention(q, k, v).float()).abs()
    assert d.max().item() <= 2e-3 and d.mean().item() <= 1e-3, (b, hkv, g, T, s, d.max().item())
print("TC-OK")
""" % root
    env = dict(os.environ, MFB200_LIB=lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "TC-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# --------------------------------------------------------------------------- size-independent properties
def test_properties_full_size_config3_shape():
    """B=16 x 8 KV heads, G=4, T=8192, s=0.7 is too big for the numpy oracle; check structural properties:
    (i) V-linearity: attention(V) with V scaled by 2 doubles the output; (ii) a one-hot mask that only
    leaves token j visible returns exactly the stored (pruned) V row j; (iii) determinism."""
    from mustafar_b200.attention import MustafarKVCache
    b, hkv, groups, T = 4, 8, 4, 8192  # batch reduced 16 -> 4 to keep the test in seconds; same per-unit shape
    k = _randn((b, hkv, T, 128), 1).cuda()
    v = _randn((b, hkv, T, 128), 2).cuda()
    q = _randn((b, hkv * groups, 1, 128), 3).cuda()
    c1 = MustafarKVCache(b, hkv, groups, T, 0.7, 0.7)
    c1.prefill(k, v)
    c2 = MustafarKVCache(b, hkv, groups, T, 0.7, 0.7)
    c2.prefill(k, v * 2)
    o1, o2 = c1.attend(q), c2.attend(q)
    assert torch.equal(o1, c1.attend(q))
    assert torch.allclose(o2.float(), 2 * o1.float(), atol=2e-3, rtol=2e-3)
    for j in (5, 4097, T - 3):
        mask = torch.full((b, 1, 1, T), torch.finfo(torch.float16).min, dtype=torch.float16, device="cuda")
        mask[..., j] = 0
        o = c1.attend(q, mask)
        row = v[:, :, j]
        if j < c1.comp_len:
            row = torch.from_numpy(O.prune_rows(row.cpu().numpy(), 0.7)).cuda()
        want = row[:, :, None, :].expand(b, hkv, groups, 128).reshape(b, hkv * groups, 1, 128)
        assert torch.allclose(o.float(), want.float(), atol=1e-3, rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("b,hkv,groups,T,shared_row", [(2, 4, 1, 790, False), (3, 2, 4, 1310, False), (1, 8, 2, 545, True),
                                                        (2, 1, 8, 400, True)])
def test_fused_rotary_embedding_is_bit_identical_to_rotating_first(b, hkv, groups, T, shared_row):
    """mfb200_decode_params::rope_cos (the apply_rotary_pos_emb of llama_mustafar_kernel.py:238-253 inside the attention
    launch): decode steps fed UNROTATED q / k with (cos, sin) must produce the same bits - outputs and appended window rows -
    as steps fed the fp16 tensors transformers' apply_rotary_pos_emb computes, across a compression event."""
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    from mustafar_b200.attention import MustafarKVCache
    torch.manual_seed(T)
    k0 = torch.randn(b, hkv, T, 128, device="cuda", dtype=torch.float16)
    v0 = torch.randn(b, hkv, T, 128, device="cuda", dtype=torch.float16)
    caches = [MustafarKVCache(b, hkv, groups, T + 64, 0.5, 0.5) for _ in range(2)]
    for c in caches:
        c.prefill(k0, v0)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, device="cuda").float() / 128))
    steps = 40
    for t in range(steps):
        pos = torch.full((1 if shared_row else b, 1), T + t, device="cuda") + (0 if shared_row else torch.arange(b, device="cuda")[:, None] * 3)
        freqs = pos.float()[:, :, None] * inv_freq[None, None, :]
        emb = torch.cat((freqs, freqs), dim=-1)
        cos, sin = emb.cos().half(), emb.sin().half()  # [B or 1, 1, 128]
        q = torch.randn(b, hkv * groups, 1, 128, device="cuda", dtype=torch.float16)
        k = torch.randn(b, hkv, 1, 128, device="cuda", dtype=torch.float16)
        v = torch.randn(b, hkv, 1, 128, device="cuda", dtype=torch.float16)
        q_rot, k_rot = apply_rotary_pos_emb(q, k, cos, sin)
        want = caches[0].decode_step(q_rot, k_rot, v)
        got = caches[1].decode_step(q, k, v, rope=(cos, sin))
        assert torch.equal(got, want), (t, (got.float() - want.float()).abs().max().item())
        # a step without rope right after one with it must not inherit the fields
        assert not caches[1]._p.rope_cos and not caches[1]._p.rope_sin
    assert caches[0].comp_len == caches[1].comp_len and caches[0].win_len == caches[1].win_len
    n = caches[0].win_len
    assert torch.equal(caches[0].k_win[:, :n], caches[1].k_win[:, :n]) and torch.equal(caches[0].v_win[:, :n], caches[1].v_win[:, :n])
    # the compressed streams were built from those window rows (the steps after the compression event attend over them)
    tiles = caches[0].comp_len * 2
    assert torch.equal(caches[0].k.bmp[:, :tiles], caches[1].k.bmp[:, :tiles])
    assert torch.equal(caches[0].k.idx[:, :tiles + 1], caches[1].k.idx[:, :tiles + 1])
    with pytest.raises(ValueError):
        caches[1].decode_step(q, k, v, rope=(cos.float(), sin.float()))


# --------------------------------------------------------------------------- f4: the other pruning policies
POLICY_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["policy_opa_mha_s50", "policy_opa_gqa4_s70"])
def test_output_aware_key_pruning_golden_and_oracle(name):
    """mfb200_prune_rows_scored against the vectors the reference's own dh_prune_key produced
    (llama_mustafar_Kt_Opa_Vt_Mag.py:65-178, oracle/make_golden_policies.py) and against the oracle on a larger case."""
    from mustafar_b200 import pruning
    check_scored_rows = O.check_scored_rows
    g = np.load(os.path.join(POLICY_GOLDEN, name + ".npz"))
    s, gs, groups = float(g["sparsity"]), int(g["group_size"]), int(g["groups"])
    n_keep = int(128 * (1 - s))
    key, q = torch.from_numpy(g["key"]).cuda(), torch.from_numpy(g["q"]).cuda()
    w = pruning.fold_queries(q, groups, gs)
    assert np.allclose(w.float().cpu().numpy(), g["w"].astype(np.float32), rtol=2e-3, atol=0)
    # the prune itself on the reference's own weights: bit-exact off the tied rows, the documented superset on them
    got = pruning.prune_rows_scored(key, torch.from_numpy(g["w"]).cuda(), n_keep, keep_last=gs).cpu().numpy()
    tied = g["tied_rows"].copy()
    tied[:, :, -gs:] = False
    check_scored_rows(got, g["pruned"], g["key"], np.abs(g["key"] * g["w"][:, :, None, :]), n_keep, tied)
    assert np.array_equal(_bits(got), _bits(O.prune_rows_scored(g["key"], g["w"], n_keep, keep_last=gs)))
    for t in range(g["dec_window"].shape[0]):  # decode form: explicit score rows
        oldest = g["dec_window"][t][:, :, :1, :]
        sc = g["dec_acc_before"][t][:, :, 0:1, :] / np.float16(gs)
        got = pruning.prune_rows_scored(torch.from_numpy(oldest).cuda(), torch.from_numpy(sc).cuda(), n_keep).cpu().numpy()
        check_scored_rows(got, g["dec_pruned"][t], oldest, sc, n_keep, g["dec_tied"][t])
    # the public entry point end to end, larger than the golden case, against the oracle fed the same folded weights
    b, hkv, t = 2, 4, 1000
    key = _randn((b, hkv, t, 128), 77).cuda()
    q = _randn((b, hkv * groups, t, 128), 78).cuda()
    got = pruning.dh_prune_key_output_aware(key, q, s, groups, gs)
    w = pruning.fold_queries(q, groups, gs)
    want = O.prune_rows_scored(key.cpu().numpy(), w.cpu().numpy(), n_keep, keep_last=gs)
    assert np.array_equal(_bits(got.cpu().numpy()), _bits(want))
    kept = np.count_nonzero(got[:, :, :-gs].cpu().numpy()) / got[:, :, :-gs].numel()
    assert abs(kept - n_keep / 128) < 2e-3


@pytest.mark.gpu
def test_channelwise_value_pruning_golden_and_oracle():
    """mfb200_prune_token_groups against the reference's own dh_prune_value (llama_mustafar_Kt_Mag_Vc_Mag.py:107-170)."""
    from mustafar_b200 import pruning
    for path in sorted(glob.glob(os.path.join(POLICY_GOLDEN, "policy_vc_*.npz"))):
        g = np.load(path)
        got = pruning.dh_prune_value_channelwise(torch.from_numpy(g["x"]).cuda(), float(g["sparsity"]), int(g["group_size"]))
        assert np.array_equal(_bits(got.cpu().numpy()), _bits(g["y"])), path
    x = _randn((3, 5, 32 * 37, 128), 5)
    x[:, :, 100:140] = torch.round(x[:, :, 100:140] * 2) / 2  # ties, zeros and negative zeros
    x[:, :, 200:232] = 0
    for s, gs in ((0.5, 32), (0.7, 32), (0.9, 8), (0.25, 4), (1.0, 2)):
        got = pruning.dh_prune_value_channelwise(x.cuda(), s, gs)
        assert np.array_equal(_bits(got.cpu().numpy()), _bits(O.prune_token_groups(x.numpy(), s, gs))), (s, gs)
    with pytest.raises(ValueError):
        pruning.dh_prune_value_channelwise(x[:, :, :33].cuda(), 0.5, 32)
    with pytest.raises(RuntimeError):
        pruning.dh_prune_value_channelwise(x, 0.5, 32)  # CPU tensor: no fallback


@pytest.mark.gpu
def test_policy_pruned_cache_feeds_the_same_format_and_kernels():
    """Kt_Opa_Vt_Mag / Kt_Mag_Vc_Mag end to end: keys pruned output-aware, values channel-wise, then the SAME compressed format
    and attention launch (cache created with sparsity 0 = its own per-token prune is the identity)."""
    from mustafar_b200 import pruning
    from mustafar_b200.attention import MustafarKVCache
    b, hkv, groups, T, s = 2, 2, 4, 800, 0.5
    k, v = _randn((b, hkv, T, 128), 1), _randn((b, hkv, T, 128), 2)
    qp = _randn((b, hkv * groups, T, 128), 3)
    q = _randn((b, hkv * groups, 1, 128), 4)
    L = O.compressed_length(T)
    kp = pruning.dh_prune_key_output_aware(k[:, :, :L].cuda(), qp[:, :, :L].cuda(), s, groups, 32, keep_last=False)
    vp = pruning.dh_prune_value_channelwise(v[:, :, :L].cuda(), s, 32)
    k2, v2 = k.clone().cuda(), v.clone().cuda()
    k2[:, :, :L], v2[:, :, :L] = kp, vp
    cache = MustafarKVCache(b, hkv, groups, max_tokens=T + 300, k_sparsity=0.0, v_sparsity=0.0, nz_halves_per_token=88)
    cache.prefill(k2, v2)
    cache.check_overflow()
    assert cache.comp_len == L
    # the cache holds exactly the policy-pruned tensors
    kc, _, vc, _, _, _ = cache.as_reference_tuple()
    kd, vd = (x[:, :, :L].reshape(b * hkv, L, 128).cpu().numpy() for x in (k2, v2))
    for got, (rb, ra, rp) in ((kc, O.convert_key_batched(kd)), (vc, O.convert_value_batched(vd))):
        assert np.array_equal(got[0].cpu().numpy(), rb) and np.array_equal(got[1].cpu().numpy(), ra)
        assert all(np.array_equal(_bits(a.cpu().numpy()), _bits(r)) for a, r in zip(got[2], rp))
    got = cache.attend(q.cuda())
    _check_attention(got, q, k2.cpu().numpy(), v2.cpu().numpy(), L)
