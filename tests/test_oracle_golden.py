"""CPU tests: the numpy oracle must reproduce, bit for bit, the golden vectors that
oracle/make_golden.py produced by running the reference's own code (Triton compression under
TRITON_INTERPRET=1; dh_prune_key executed from the reference source)."""
import glob
import os

import numpy as np
import pytest

from oracle import mustafar_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COMPRESS = sorted(glob.glob(os.path.join(GOLDEN, "compress_*.npz")))
PRUNE = sorted(glob.glob(os.path.join(GOLDEN, "prune_*.npz")))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def test_golden_present():
    assert len(COMPRESS) >= 10 and len(PRUNE) >= 7


@pytest.mark.parametrize("path", PRUNE, ids=[os.path.basename(p) for p in PRUNE])
def test_prune_matches_reference(path):
    g = np.load(path)
    y = O.prune_rows(g["x"], float(g["sparsity"]))
    # bit-exact including the sign of zeros produced by x*False
    assert np.array_equal(_bits(y), _bits(g["y"]))


@pytest.mark.parametrize("path", COMPRESS, ids=[os.path.basename(p) for p in COMPRESS])
def test_compress_matches_reference(path):
    g = np.load(path)
    s = float(g["sparsity"])
    xp = g["pruned"] if s < 0 else O.prune_rows(g["x"], s)
    assert np.array_equal(_bits(xp), _bits(g["pruned"]))
    for tag, fn, dec in (("k", O.convert_key_batched, O.decode_key), ("v", O.convert_value_batched, O.decode_value)):
        bmp, acc, packed = fn(xp)
        assert bmp.dtype == np.int64 and acc.dtype == np.int32
        assert np.array_equal(bmp, g[f"{tag}_bitmaps"])
        assert np.array_equal(acc, g[f"{tag}_accum"])
        assert [p.size for p in packed] == list(g[f"{tag}_packed_len"])
        flat = np.concatenate(packed)
        assert np.array_equal(_bits(flat), _bits(g[f"{tag}_packed"]))
        # round trip through the kernel-side addressing rules
        back = dec(bmp, acc, flat, O.nz_offsets(acc), xp.shape[1])
        assert np.array_equal(back == 0, xp == 0)
        assert np.array_equal(_bits(back)[xp != 0], _bits(xp)[xp != 0])


def test_prune_keeps_ties_and_counts():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 128)).astype(np.float16)
    for s, kept in ((0.5, 65), (0.7, 40)):
        y = O.prune_rows(x, s)
        nz = (y != 0).sum(-1)
        assert nz.min() >= kept - 1  # an exact input zero can be among the kept
        assert (nz >= kept).mean() > 0.95
    x[:] = 1.0
    assert (O.prune_rows(x, 0.7) == 1.0).all()  # all ties survive
    x[:] = 0.0
    assert (O.prune_rows(x, 0.5) == 0).all()


def test_append_equals_whole():
    """compressing [0,L) then appending a 256-token chunk equals compressing [0,L+256)."""
    rng = np.random.default_rng(1)
    x = O.prune_rows(rng.standard_normal((2, 512, 128)).astype(np.float16), 0.5)
    for fn in (O.convert_key_batched, O.convert_value_batched):
        b0, a0, p0 = fn(x[:, :256])
        b1, a1, p1 = fn(x[:, 256:])
        bw, aw, pw = fn(x)
        assert np.array_equal(np.concatenate([b0, b1], 1), bw)
        acc = np.concatenate([a0[:, :-1], a1 + a0[:, -1:]], 1)
        assert np.array_equal(acc, aw)
        for h in range(2):
            assert np.array_equal(np.concatenate([p0[h], p1[h]]), pw[h])


def test_glue_vs_masked_dense_close():
    rng = np.random.default_rng(2)
    b, hq, hkv, t, l = 1, 4, 2, 320, 256
    k = rng.standard_normal((b, hkv, t, 128)).astype(np.float16)
    v = rng.standard_normal((b, hkv, t, 128)).astype(np.float16)
    q = rng.standard_normal((b, hq, 1, 128)).astype(np.float16)
    k[:, :, :l] = O.prune_rows(k[:, :, :l], 0.5)
    v[:, :, :l] = O.prune_rows(v[:, :, :l], 0.5)
    a = O.decode_attention_glue(q, k[:, :, :l], k[:, :, l:], v[:, :, :l], v[:, :, l:]).astype(np.float32)
    m = O.masked_dense_attention(q, k, v).astype(np.float32)
    e = O.attention_exact_f64(q, k, v)
    assert np.abs(a - m).max() < 2e-3
    assert np.abs(a - e).max() < 2e-3 and np.abs(m - e).max() < 2e-3


def test_compressed_length():
    assert O.compressed_length(4096) == 3840
    assert O.compressed_length(8192) == 7936
    assert O.compressed_length(32768) == 32512
    assert O.compressed_length(300) == 256
    assert O.compressed_length(287) == 0
    assert O.compressed_length(5) == 0


# ---- f4: the other pruning policies (golden vectors from the reference's own functions, oracle/make_golden_policies.py) ----
POLICY_OPA = sorted(glob.glob(os.path.join(GOLDEN, "policy_opa_*.npz")))
POLICY_VC = sorted(glob.glob(os.path.join(GOLDEN, "policy_vc_*.npz")))


check_scored_rows = O.check_scored_rows


def test_policy_golden_present():
    assert len(POLICY_OPA) >= 2 and len(POLICY_VC) >= 5


@pytest.mark.parametrize("path", POLICY_OPA, ids=[os.path.basename(p) for p in POLICY_OPA])
def test_output_aware_key_pruning_matches_reference(path):
    g = np.load(path)
    s, gs, groups = float(g["sparsity"]), int(g["group_size"]), int(g["groups"])
    n_keep = int(128 * (1 - s))
    w = O.fold_queries(g["q"], groups, gs)
    # the fold is two fp16 reductions: same values up to the last bit of the fp32 accumulation order
    assert np.allclose(w.astype(np.float32), g["w"].astype(np.float32), rtol=2e-3, atol=0)
    got = O.prune_rows_scored(g["key"], g["w"], n_keep, keep_last=gs)
    score = np.abs(g["key"] * g["w"][:, :, None, :])
    tied = g["tied_rows"].copy()
    tied[:, :, -gs:] = False  # those rows stay dense whatever their scores
    check_scored_rows(got, g["pruned"], g["key"], score, n_keep, tied)
    assert np.array_equal(_bits(got[:, :, -gs:]), _bits(g["key"][:, :, -gs:]))
    # decode form: the oldest window row, scored by its accumulated score / group_size (`:131-145`)
    for t in range(g["dec_window"].shape[0]):
        oldest = g["dec_window"][t][:, :, :1, :]
        sc = g["dec_acc_before"][t][:, :, 0:1, :] / np.float16(gs)
        assert sc.dtype == np.float16
        got = O.prune_rows_scored(oldest, sc, n_keep)
        check_scored_rows(got, g["dec_pruned"][t], oldest, sc, n_keep, g["dec_tied"][t])


@pytest.mark.parametrize("path", POLICY_VC, ids=[os.path.basename(p) for p in POLICY_VC])
def test_channelwise_value_pruning_matches_reference(path):
    g = np.load(path)
    y = O.prune_token_groups(g["x"], float(g["sparsity"]), int(g["group_size"]))
    assert np.array_equal(_bits(y), _bits(g["y"]))
    with pytest.raises(ValueError):
        O.prune_token_groups(g["x"][:, :, :-1], float(g["sparsity"]), int(g["group_size"]))
