"""CPU oracle for the Mustafar sparse-KV decode path.  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy restatement of the reference algorithm.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``mustafar_b200``) never imports it and has no CPU fallback.

Pinning status: the reference ships no tests or golden vectors.
The oracle is pinned instead against OUTPUTS OF THE REFERENCE ITSELF, generated
in the build container by ``oracle/make_golden.py`` (reference Triton
``kernel/compression.py`` under ``TRITON_INTERPRET=1`` and the reference
``dh_prune_key`` source executed verbatim) and committed under ``tests/golden``.
On the GPU box the attention half is additionally checked against the reference
CUDA kernels compiled from ``/root/reference`` into ``oracle/_ref``.

All ``file:line`` citations are relative to ``/root/reference``.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

TILE = 64  # elements per bitmap tile (kernel/compression.py:35, :96)
HEAD_DIM = 128


# ----------------------------------------------------------------------------
# a1. per-token magnitude threshold pruning
# ----------------------------------------------------------------------------
def prune_k(sparsity: float, dim: int = HEAD_DIM) -> int:
    """`num_to_keep = max(1, int(target_sparsity * D))` — models/llama_mustafar_kernel.py:97, :137.

    Despite its name in the reference this is the RANK of the threshold among the
    ascending magnitudes, not a number of kept elements.
    """
    assert 0 <= sparsity < 1
    return max(1, int(sparsity * dim))


def prune_rows(x: np.ndarray, sparsity: float) -> np.ndarray:
    """Threshold prune along the last axis.

    Follows models/llama_mustafar_kernel.py:97-110 (dh_prune_key) and :137-149
    (dh_prune_value): thr = k-th smallest |x| of the row; keep |x| >= thr (all
    ties at the threshold survive); dropped entries become ``x * False`` = ±0.
    """
    assert x.dtype == np.float16
    d = x.shape[-1]
    k = prune_k(sparsity, d)
    flat = x.reshape(-1, d)
    mag = np.abs(flat)
    thr = np.partition(mag, k - 1, axis=-1)[:, k - 1 : k]
    keep = mag >= thr
    # x * mask in fp16: kept -> x, dropped -> +0 or -0 depending on sign of x.
    out = flat * keep.astype(np.float16)
    return out.reshape(x.shape)


# ----------------------------------------------------------------------------
# f4. the other pruning policies (they only change the mask; format and kernels are the same)
# ----------------------------------------------------------------------------
def fold_queries(q: np.ndarray, groups: int, group_size=None) -> np.ndarray:
    """[B, Hq, T, D] fp16 -> [B, Hkv, D] fp16 channel weights of the output-aware policy.

    models/llama_mustafar_Kt_Opa_Vt_Mag.py:98-100 (prefill: mean of |q| over the last group_size tokens, then the sum over
    the KV head's query heads) and :131-132 (decode, T == 1: |q| summed over the query heads).  torch reduces fp16 tensors
    with an fp32 accumulator and rounds once per reduction; so does this."""
    assert q.dtype == np.float16
    b, hq, t, d = q.shape
    a = np.abs(q if group_size is None else q[:, :, -group_size:, :])
    folded = a.astype(np.float32).mean(axis=-2).astype(np.float16)
    return folded.reshape(b, hq // groups, groups, d).astype(np.float32).sum(axis=-2).astype(np.float16)


def prune_rows_scored(x: np.ndarray, w: np.ndarray, n_keep: int, keep_last: int = 0) -> np.ndarray:
    """Output-aware key pruning, threshold form.  x [B, H, T, D]; w [B, H, D] (score = |x*w|, an fp16 product:
    llama_mustafar_Kt_Opa_Vt_Mag.py:104) or [B, H, T, D] (w is the score: the accumulated decode-time score, :131-145).

    The reference sorts the scores descending and scatters True at the first n_keep indices (:106-109, :139-150); that is
    `score >= n_keep-th highest score` whenever that score differs from the next one.  On a tie at the cut the reference's
    choice among the tied entries is implementation-defined (torch.sort is not stable); this restatement - and the CUDA
    path - keeps all of them.  Dropped entries become x * False = +-0; the last keep_last tokens stay dense (:110)."""
    assert x.dtype == np.float16 and w.dtype == np.float16
    b, h, t, d = x.shape
    score = np.abs(x * w[:, :, None, :]) if w.ndim == 3 else np.abs(w)
    assert score.dtype == np.float16
    thr = np.partition(score, d - n_keep, axis=-1)[..., d - n_keep : d - n_keep + 1]  # the n_keep-th highest
    out = x * (score >= thr).astype(np.float16)
    if keep_last:
        out[:, :, -keep_last:, :] = x[:, :, -keep_last:, :]
    return out


def _bits16(a):
    return np.ascontiguousarray(a).view(np.uint16)


def check_scored_rows(got, want, x, score, n_keep, tied):
    """Rows without a tie at the cut: bit-exact.  Tied rows: the reference keeps an arbitrary n_keep of the candidates; the
    threshold form keeps every entry whose score reaches the cut - a superset that differs only in tied entries."""
    got, want, x, score = (a.reshape(-1, a.shape[-1]) for a in (got, want, x, score))
    tied = tied.reshape(-1)
    assert np.array_equal(_bits16(got[~tied]), _bits16(want[~tied]))
    for r in np.nonzero(tied)[0]:
        thr = np.sort(score[r].astype(np.float32))[::-1][n_keep - 1]
        keep = score[r].astype(np.float32) >= thr
        assert np.array_equal(_bits16(got[r]), _bits16(x[r] * keep.astype(np.float16)))
        ref_keep = (_bits16(want[r]) & 0x7fff) != 0
        assert not np.any(ref_keep & ~keep) and keep.sum() > n_keep  # the reference's survivors are among ours


def prune_token_groups(x: np.ndarray, sparsity: float, group_size: int = 32) -> np.ndarray:
    """Channel-wise value pruning, models/llama_mustafar_Kt_Mag_Vc_Mag.py:107-170: in every group of group_size consecutive
    tokens each channel keeps |v| >= its k-th smallest magnitude, k = max(1, int(sparsity * group_size)) (:142-153)."""
    assert x.dtype == np.float16 and 0 <= sparsity <= 1
    b, h, t, d = x.shape
    if t % group_size != 0:
        raise ValueError("Token dimension must be a multiple of group_size")
    k = max(1, int(sparsity * group_size))
    g = x.reshape(b, h, t // group_size, group_size, d)
    mag = np.abs(g)
    thr = np.partition(mag, k - 1, axis=3)[:, :, :, k - 1 : k, :]
    return (g * (mag >= thr).astype(np.float16)).reshape(x.shape)


# ----------------------------------------------------------------------------
# a2-a6. bitmap + packed-nonzero format
# ----------------------------------------------------------------------------
def _tiles_key(x: np.ndarray) -> np.ndarray:
    """[Bk, M, D] -> [Bk, tiles, 64] in the K tile order.

    kernel/compression.py:32-36 on the transposed input (`:255`): tile id
    = token_block * D + channel; element e = K[64*token_block + e, channel].
    """
    bk, m, d = x.shape
    assert m % TILE == 0
    t = x.reshape(bk, m // TILE, TILE, d)  # [b, tb, e, c]
    t = np.transpose(t, (0, 1, 3, 2))  # [b, tb, c, e]
    return np.ascontiguousarray(t).reshape(bk, (m // TILE) * d, TILE)


def _tiles_value(x: np.ndarray) -> np.ndarray:
    """[Bk, M, D] -> [Bk, tiles, 64] in the V tile order.

    kernel/compression.py:87-97: tile id = token_block*(D/64*64) + col_tile*64 + r;
    element e = V[64*token_block + r, 64*col_tile + e].
    """
    bk, m, d = x.shape
    assert m % TILE == 0 and d % TILE == 0
    t = x.reshape(bk, m // TILE, TILE, d // TILE, TILE)  # [b, tb, r, h, e]
    t = np.transpose(t, (0, 1, 3, 2, 4))  # [b, tb, h, r, e]
    return np.ascontiguousarray(t).reshape(bk, (m // TILE) * d, TILE)


def _untile_key(tiles: np.ndarray, m: int, d: int) -> np.ndarray:
    bk = tiles.shape[0]
    t = tiles.reshape(bk, m // TILE, d, TILE)
    return np.ascontiguousarray(np.transpose(t, (0, 1, 3, 2))).reshape(bk, m, d)


def _untile_value(tiles: np.ndarray, m: int, d: int) -> np.ndarray:
    bk = tiles.shape[0]
    t = tiles.reshape(bk, m // TILE, d // TILE, TILE, TILE)
    return np.ascontiguousarray(np.transpose(t, (0, 1, 3, 2, 4))).reshape(bk, m, d)


_SHIFTS = (np.uint64(1) << np.arange(63, -1, -1, dtype=np.uint64))  # MSB = element 0


def _compress_tiles(tiles: np.ndarray) -> Tuple[np.ndarray, np.ndarray, List[np.ndarray]]:
    """Common tail of convert_{key,value}_batched (kernel/compression.py:282-335, :375-428).

    bitmap bit (63-e) = tiles[..., e] != 0.0 (`:42-44`); count = ((popc+7)&~7)>>1 in
    units of two halves (`:48`); accum_counts = [0, cumsum(count)] int32 (`:294-298`);
    nonzeros of tile t at halves [2*accum[t], 2*accum[t]+popc), zero padded (`:309`, `:168-174`).
    """
    bk, nt, _ = tiles.shape
    nzmask = tiles != np.float16(0.0)  # -0.0 != 0.0 is False; NaN != 0.0 is True
    bitmaps = (nzmask.astype(np.uint64) * _SHIFTS).sum(axis=-1, dtype=np.uint64)
    popc = nzmask.sum(axis=-1).astype(np.int64)
    counts = ((popc + 7) & ~np.int64(7)) >> 1
    accum = np.zeros((bk, nt + 1), dtype=np.int32)
    accum[:, 1:] = np.cumsum(counts, axis=1).astype(np.int32)
    packed: List[np.ndarray] = []
    for b in range(bk):
        buf = np.zeros(2 * int(accum[b, -1]), dtype=np.float16)
        # destination of every nonzero = 2*accum[tile] + rank within the tile
        rank = np.cumsum(nzmask[b], axis=-1) - 1
        dst = (2 * accum[b, :-1].astype(np.int64))[:, None] + rank
        buf[dst[nzmask[b]]] = tiles[b][nzmask[b]]
        packed.append(buf)
    return bitmaps.view(np.int64), accum, packed


def convert_key_batched(x: np.ndarray):
    """Restatement of kernel/compression.py:249-339 for an already-pruned K [Bk, M, 128]."""
    assert x.ndim == 3 and x.dtype == np.float16 and x.shape[1] % TILE == 0
    return _compress_tiles(_tiles_key(x))


def convert_value_batched(x: np.ndarray):
    """Restatement of kernel/compression.py:341-432 for an already-pruned V [Bk, M, 128]."""
    assert x.ndim == 3 and x.dtype == np.float16 and x.shape[1] % TILE == 0
    return _compress_tiles(_tiles_value(x))


def nz_offsets(accum: np.ndarray) -> np.ndarray:
    """`nz_offset[i] = nz_offset[i-1] + idx[i-1][-1] // 4` — models/llama_mustafar_kernel.py:329-331.

    Start of head i inside the concatenated NZ buffer in uint4 (16 B) units.
    """
    tot = accum[:, -1].astype(np.int64) // 4
    out = np.zeros(accum.shape[0], dtype=np.int32)
    out[1:] = np.cumsum(tot[:-1]).astype(np.int32)
    return out


def _decode_tiles(bitmaps: np.ndarray, accum: np.ndarray, nz_flat: np.ndarray,
                  nz_off: np.ndarray) -> np.ndarray:
    """Decode the format the way the reference kernels address it.

    kernel/csrc/SpMM_Kernel.cuh:174-185 (per-head bases: NZ + NZ_offset[h] in uint4,
    idx + h*(tiles+1), bmp + h*tiles), :55-77 (tile t's values start at uint4 index
    idx[t]/4 of the head), :138-149 (k-th value goes to the position of the k-th set
    bit counted from the MSB, `__clzll`).
    """
    bk, nt = bitmaps.shape
    ub = bitmaps.view(np.uint64)
    bits = ((ub[..., None] >> np.arange(63, -1, -1, dtype=np.uint64)) & np.uint64(1)).astype(bool)
    tiles = np.zeros((bk, nt, TILE), dtype=np.float16)
    for b in range(bk):
        base = int(nz_off[b]) * 8  # uint4 -> halves
        rank = np.cumsum(bits[b], axis=-1) - 1
        src = base + (2 * accum[b, :-1].astype(np.int64))[:, None] + rank
        tiles[b][bits[b]] = nz_flat[src[bits[b]]]
    return tiles


def decode_key(bitmaps, accum, nz_flat, nz_off, m: int, d: int = HEAD_DIM) -> np.ndarray:
    return _untile_key(_decode_tiles(bitmaps, accum, nz_flat, nz_off), m, d)


def decode_value(bitmaps, accum, nz_flat, nz_off, m: int, d: int = HEAD_DIM) -> np.ndarray:
    return _untile_value(_decode_tiles(bitmaps, accum, nz_flat, nz_off), m, d)


# ----------------------------------------------------------------------------
# a8-a13. the two batched SpMV operators (mustafar_package)
# ----------------------------------------------------------------------------
def key_formulation(bitmaps, nz_flat, accum, nz_off, b_pad: np.ndarray, m_global: int,
                    k_global: int, batch_size: int, groups: int) -> np.ndarray:
    """C[bq, n, t] = sum_c K_h[t, c] * B[bq, n, c],  h = bq // groups.

    kernel/kernel_wrapper/mustafar_wrapper.cu:19-133 → Key_Kernel
    (kernel/csrc/SpMM_Kernel.cuh:156-419): fp16 inputs, fp32 accumulate,
    `__float2half_rn` on store (`:418`); output [Bq, 8, M] (`mustafar_wrapper.cu:81`).
    """
    k_dense = decode_key(bitmaps.reshape(-1, m_global * k_global // TILE),
                         accum.reshape(-1, m_global * k_global // TILE + 1),
                         nz_flat, nz_off, m_global, k_global).astype(np.float32)
    out = np.empty((batch_size, 8, m_global), dtype=np.float16)
    bf = b_pad.astype(np.float32)
    for bq in range(batch_size):
        out[bq] = (bf[bq] @ k_dense[bq // groups].T).astype(np.float16)
    return out


def value_formulation(bitmaps, nz_flat, accum, nz_off, b_pad: np.ndarray, m_global: int,
                      k_global: int, batch_size: int, groups: int) -> np.ndarray:
    """C[bq, n, c] = sum_t V_h[t, c] * B[bq, n, t],  h = bq // groups.

    kernel/kernel_wrapper/mustafar_wrapper.cu:139-263 → Value_Kernel
    (kernel/csrc/SpMM_Kernel.cuh:421-676); M_Global = 128 channels, K_Global = L tokens.
    """
    v_dense = decode_value(bitmaps.reshape(-1, m_global * k_global // TILE),
                           accum.reshape(-1, m_global * k_global // TILE + 1),
                           nz_flat, nz_off, k_global, m_global).astype(np.float32)
    out = np.empty((batch_size, 8, m_global), dtype=np.float16)
    bf = b_pad.astype(np.float32)
    for bq in range(batch_size):
        out[bq] = (bf[bq] @ v_dense[bq // groups]).astype(np.float16)
    return out


# ----------------------------------------------------------------------------
# a15 / a16. decode attention (glue around the SpMV ops, and masked-dense)
# ----------------------------------------------------------------------------
def _softmax_f32(w: np.ndarray) -> np.ndarray:
    w = w.astype(np.float32)
    w = w - w.max(axis=-1, keepdims=True)
    e = np.exp(w)
    return e / e.sum(axis=-1, keepdims=True)


def repeat_kv(x: np.ndarray, groups: int) -> np.ndarray:
    """[B, Hkv, T, D] -> [B, Hkv*groups, T, D] (HF `repeat_kv`, used at llama_mustafar_kernel.py:278)."""
    return np.repeat(x, groups, axis=1)


def decode_attention_glue(q: np.ndarray, k_comp_dense: np.ndarray, k_win: np.ndarray,
                          v_comp_dense: np.ndarray, v_win: np.ndarray,
                          mask: np.ndarray | None = None) -> np.ndarray:
    """The reference decode step with its fp16 rounding points (llama_mustafar_kernel.py:268-320).

    q [B,Hq,1,D]; k_comp_dense/v_comp_dense [B,Hkv,L,D] = the pruned rows that live in the
    compressed cache; k_win/v_win [B,Hkv,Lw,D] dense window (new token already appended, `:270`,
    `:309`).  scores: fp32 accumulate → fp16 (`SpMM_Kernel.cuh:418` / fp16 matmul), `/ sqrt(D)` in
    fp16 (`:284`), optional additive mask clamped at finfo.min (`:293-301`), softmax in fp32 →
    fp16 (`:304`), P·V fp32 accumulate → fp16 for each part, parts added in fp16 (`:317`).
    """
    b, hq, _, d = q.shape
    hkv = k_win.shape[1]
    g = hq // hkv
    qf = q.astype(np.float32)
    kc = repeat_kv(k_comp_dense, g).astype(np.float32)
    kw = repeat_kv(k_win, g).astype(np.float32)
    att_c = np.matmul(qf, np.swapaxes(kc, 2, 3)).astype(np.float16)
    att_l = np.matmul(qf, np.swapaxes(kw, 2, 3)).astype(np.float16)
    att = np.concatenate([att_c, att_l], axis=-1)
    w = (att.astype(np.float32) / np.float32(math.sqrt(d))).astype(np.float16)
    if mask is not None:
        w = (w.astype(np.float32) + mask.astype(np.float32)).astype(np.float16)
        w = np.maximum(w, np.float16(np.finfo(np.float16).min))
    p = _softmax_f32(w).astype(np.float16)
    lc = k_comp_dense.shape[2]
    vc = repeat_kv(v_comp_dense, g).astype(np.float32)
    vw = repeat_kv(v_win, g).astype(np.float32)
    out_c = np.matmul(p[..., :lc].astype(np.float32), vc).astype(np.float16)
    out_l = np.matmul(p[..., lc:].astype(np.float32), vw).astype(np.float16)
    return (out_c.astype(np.float32) + out_l.astype(np.float32)).astype(np.float16)


def masked_dense_attention(q: np.ndarray, k_full: np.ndarray, v_full: np.ndarray,
                           mask: np.ndarray | None = None) -> np.ndarray:
    """The masked-dense decode step: models/llama_mustafar_Kt_Mag_Vt_Mag.py:873-874, :952-963, :974.

    k_full/v_full [B,Hkv,T,D] hold the pruned-in-place rows followed by the dense rows.
    """
    b, hq, _, d = q.shape
    g = hq // k_full.shape[1]
    kf = repeat_kv(k_full, g).astype(np.float32)
    vf = repeat_kv(v_full, g).astype(np.float32)
    w = np.matmul(q.astype(np.float32), np.swapaxes(kf, 2, 3)).astype(np.float16)
    w = (w.astype(np.float32) / np.float32(math.sqrt(d))).astype(np.float16)
    if mask is not None:
        w = (w.astype(np.float32) + mask.astype(np.float32)).astype(np.float16)
        w = np.maximum(w, np.float16(np.finfo(np.float16).min))
    p = _softmax_f32(w).astype(np.float16)
    return np.matmul(p.astype(np.float32), vf).astype(np.float16)


def attention_exact_f64(q, k_full, v_full, mask=None) -> np.ndarray:
    """Rounding-free (float64) attention over the same pruned K/V: the midpoint both the
    reference and the fused kernel approximate.  Used to bound errors, not as parity target."""
    b, hq, _, d = q.shape
    g = hq // k_full.shape[1]
    kf = repeat_kv(k_full, g).astype(np.float64)
    vf = repeat_kv(v_full, g).astype(np.float64)
    w = np.matmul(q.astype(np.float64), np.swapaxes(kf, 2, 3)) / math.sqrt(d)
    if mask is not None:
        w = w + mask.astype(np.float64)
    w = w - w.max(axis=-1, keepdims=True)
    e = np.exp(w)
    p = e / e.sum(axis=-1, keepdims=True)
    return np.matmul(p, vf)


def compressed_length(kv_seq_len: int, residual_length: int = 32) -> int:
    """`((kv_seq_len - residual_length) // 256) * 256` — llama_mustafar_kernel.py:416.

    The reference yields a negative value for kv_seq_len < residual_length (floor division);
    that quirk is NOT replicated: clamp at 0.
    """
    return max(0, ((kv_seq_len - residual_length) // 256) * 256)


def algorithmic_bytes(accum_k: np.ndarray, accum_v: np.ndarray, l: int, lw: int, bk: int,
                      bq: int, d: int = HEAD_DIM) -> int:
    """bitmaps + padded NZ as stored (K and V) + dense window K,V + q,out; idx excluded."""
    tiles = l * d // TILE
    bmp = 2 * bk * tiles * 8
    nz = 2 * 2 * (int(accum_k[:, -1].astype(np.int64).sum()) + int(accum_v[:, -1].astype(np.int64).sum()))
    return bmp + nz + bk * lw * d * 2 * 2 + bq * d * 2 * 2
