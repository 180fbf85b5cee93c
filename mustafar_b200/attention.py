"""Fused sparse-KV decode attention and the slab KV cache that feeds it.

`MustafarKVCache` holds what the reference keeps in its python tuple cache
(models/llama_mustafar_kernel.py:445: k_compressed=[bitmaps, accum_counts, [NZ per head], nz_offset],
k_local_window, v_compressed, v_local_window, compressed_length, kv_seq_len) as PREALLOCATED device
slabs with the same per-head format, so that
  * the decode step is one CUDA launch (csrc/decode_attn.cu) instead of the ~15 launches and two
    whole-cache `torch.cat` copies of llama_mustafar_kernel.py:268-320,
  * the every-256-token compression (llama_mustafar_kernel.py:324-398) appends in place with no host
    sync, no `torch.cat` of bitmaps and no `torch.cuda.empty_cache()`.
The slabs can be handed to the reference-compatible ops too (`as_reference_tuple`).

Schedule (same as the reference): compressed_length = ((T - residual_length)//256)*256 at prefill
(`:416`); during decode the new token is appended to the dense window BEFORE attention (`:270`,
`:309`) and when window == residual_length + 256 its first 256 rows are pruned, compressed and
dropped (`:324`, `:392-398`).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from .pruning import HEAD_DIM, prune_rank

COMPRESS_CHUNK = 256  # llama_mustafar_kernel.py:324


def compressed_length(kv_seq_len: int, residual_length: int) -> int:
    """llama_mustafar_kernel.py:416, without its negative result for prompts shorter than the residual."""
    return max(0, ((kv_seq_len - residual_length) // COMPRESS_CHUNK) * COMPRESS_CHUNK)


WORST_CHUNK_HALVES = COMPRESS_CHUNK * HEAD_DIM  # a 256-token chunk in which every element survives (all ties)


def default_halves_per_token(sparsity: float) -> int:
    """Slab budget per token and stream, in halves: the D-k+1 survivors of the reference rule
    (llama_mustafar_kernel.py:97-110) + the expected 8-half tile padding (two tiles per token, 3.5 halves each)
    + 4 % head-room for ties, rounded up to 4.  76 at sparsity 0.5 and 52 at 0.7 (dense = 128): with the 16 B of
    bitmaps and 8 B of offsets per token the cache holds 0.69x / 0.50x of the dense fp16 bytes."""
    kept = HEAD_DIM - prune_rank(sparsity) + 1
    return min(HEAD_DIM, 4 * math.ceil((kept + 7.2) * 1.04 / 4))


class _Stream:
    """bitmaps / accum_counts / packed nonzeros of K or of V for all (sequence, kv-head) units.

    The nonzero slab of a unit holds `cap_tokens * halves_per_token` halves plus one worst-case chunk of reserve
    (see MustafarKVCache._ensure_room: a chunk is only ever appended when the host KNOWS it fits)."""

    def __init__(self, units: int, cap_tokens: int, halves_per_token: int, device):
        self.units, self.cap_tokens, self.device = units, cap_tokens, device
        self.cap_tiles = cap_tokens * 2
        self.bmp = torch.zeros((units, self.cap_tiles), dtype=torch.int64, device=device)
        self.idx = torch.zeros((units, self.cap_tiles + 1), dtype=torch.int32, device=device)
        self.nz = None
        self._allocate(halves_per_token, None)

    def _allocate(self, halves_per_token: int, keep_halves):
        """(Re)allocates the nonzero slab; `keep_halves` = leading halves of every unit to carry over."""
        self.halves_per_token = halves_per_token
        cap = self.cap_tokens * halves_per_token
        if halves_per_token < HEAD_DIM:
            cap += WORST_CHUNK_HALVES
        old, old_cap = self.nz, getattr(self, "head_capacity", 0)
        self.head_capacity = cap  # halves, multiple of 8 (cap_tokens % 64 == 0)
        self.nz = torch.empty((self.units * cap,), dtype=torch.float16, device=self.device)
        if old is not None and keep_halves:
            self.nz.view(self.units, cap)[:, :keep_halves].copy_(old.view(self.units, old_cap)[:, :keep_halves])
        self.head_base = torch.arange(self.units, dtype=torch.int64, device=self.device) * cap  # halves
        self.nz_off = self.head_base // 8  # uint4 units, int64

    def bytes_held(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.bmp, self.idx, self.nz))


class MustafarKVCache:
    """Preallocated bitmap+packed-nonzero KV cache of one layer with a dense fp16 residual window.

    Memory: the nonzero slabs are sized from the sparsity (`default_halves_per_token`), not for the worst case, so
    the cache holds ~0.69x (s=0.5) / ~0.50x (s=0.7) of the dense fp16 bytes (`bytes_held()` vs `dense_bytes()`).
    A slab can still never be overrun: a chunk is appended only when the host knows that even an all-ties chunk
    fits (the per-unit fill level is read back asynchronously after every compression, i.e. it is 256 decode steps
    old and long complete when it is needed); otherwise the slabs are regrown first (`regrow_events` counts them).
    The prompt is the one place where the fill level cannot be known in advance: `prefill` reads the overflow
    flag back (one host sync; the reference's prefill has 2*B*Hkv+1 of them, compression.py:308, :333-334) and
    repeats the compression into worst-case slabs if it was raised."""

    def __init__(self, batch: int, kv_heads: int, groups: int, max_tokens: int, k_sparsity: float, v_sparsity: float,
                 residual_length: int = 32, device="cuda", nz_halves_per_token: Optional[int] = None,
                 ref_score_rounding: bool = True, pdl: bool = True, plan_hint: int = 0):
        if not 0 <= residual_length <= COMPRESS_CHUNK:
            raise ValueError(f"residual_length={residual_length} outside [0, {COMPRESS_CHUNK}] "
                             "(the window is compressed in 256-token chunks, llama_mustafar_kernel.py:324)")
        self.batch, self.kv_heads, self.groups = batch, kv_heads, groups
        self.units = batch * kv_heads
        self.k_sparsity, self.v_sparsity = k_sparsity, v_sparsity
        self.residual_length = residual_length
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.ref_score_rounding = ref_score_rounding
        # programmatic dependent launch: overlap a launch's prologue/first fetches with its predecessor's tail
        self.pdl = pdl
        self.plan_hint = plan_hint  # tests / tuning only (mfb200_decode_params::plan_hint)
        self._streams_dirty = True  # the last launch on this cache rewrote idx/bitmaps/nonzeros
        cap = ((max_tokens + COMPRESS_CHUNK - 1) // COMPRESS_CHUNK) * COMPRESS_CHUNK
        self.cap_tokens = cap
        self.win_cap = residual_length + COMPRESS_CHUNK + 1 if residual_length + COMPRESS_CHUNK < max_tokens else max_tokens + 1
        self.win_cap = max(self.win_cap, residual_length + COMPRESS_CHUNK + 1)
        self.win_cap = (self.win_cap + 7) // 8 * 8

        def per_token(s):
            hpt = default_halves_per_token(s) if nz_halves_per_token is None else int(nz_halves_per_token)
            if hpt % 4 or not 8 <= hpt <= HEAD_DIM:
                raise ValueError("nz_halves_per_token must be a multiple of 4 in [8, 128]")
            return hpt

        with torch.cuda.device(self.device):
            self.k = _Stream(self.units, cap, per_token(k_sparsity), self.device)
            self.v = _Stream(self.units, cap, per_token(v_sparsity), self.device)
            self.k_win = torch.zeros((self.units, self.win_cap, HEAD_DIM), dtype=torch.float16, device=self.device)
            self.v_win = torch.zeros_like(self.k_win)
            self.overflow = torch.zeros((1,), dtype=torch.int32, device=self.device)
            self._fill_host = torch.zeros((2, self.units), dtype=torch.int32).pin_memory()
            self._fill_event = torch.cuda.Event()
        self._fill_pending = False
        self._fill_known = [0, 0]  # max over units of the halves in use (K, V), as of the last read-back
        self.regrow_events = 0
        self.comp_len = 0
        self.win_len = 0
        self._prefill_status = None
        self._ws = None
        self._ws_bytes = 0
        self._plan_cache = {}
        self._p = None
        self._p_ref = None
        self._p_stale = True
        self._step = None
        self._sm_count = 0
        self._attn = _lib.load().mfb200_sparse_decode_attention
        # staging capacity for one 64-token block of nonzeros, in KB: survivors + the expected tile padding (7 halves per
        # token) + 2 halves per token for ties; the rare block that is larger takes the kernel's slower global-load path
        def slot_kb(s):
            kept = HEAD_DIM - prune_rank(s) + 1
            return min(16, max(1, math.ceil(64 * (kept + 7 + 2) * 2 / 1024)))
        self.slot_kb = max(slot_kb(k_sparsity), slot_kb(v_sparsity))

    # ------------------------------------------------------------------ properties
    @property
    def kv_seq_len(self) -> int:
        return self.comp_len + self.win_len

    def bytes_held(self) -> int:
        """Device bytes this layer cache occupies (bitmaps, offsets, nonzero slabs, dense windows)."""
        return self.k.bytes_held() + self.v.bytes_held() + 2 * self.k_win.numel() * 2

    def dense_bytes(self) -> int:
        """What a dense fp16 KV cache of the same capacity would occupy."""
        return 2 * self.units * self.cap_tokens * HEAD_DIM * 2

    # ------------------------------------------------------------------ slab capacity management
    def _worst_case(self) -> bool:
        return self.k.halves_per_token >= HEAD_DIM and self.v.halves_per_token >= HEAD_DIM

    def _resolve_fill(self):
        if self._fill_pending:
            self._fill_event.synchronize()  # issued >= 256 decode steps ago: complete, this does not wait
            self._fill_known = [2 * int(self._fill_host[0].max()), 2 * int(self._fill_host[1].max())]
            self._fill_pending = False

    def _request_fill(self):
        """Asynchronous read-back of every unit's fill level (idx[u, tiles] in 2-half units) after a compression."""
        if self._worst_case():
            return
        tiles = self.comp_len * 2
        self._fill_host[0].copy_(self.k.idx[:, tiles], non_blocking=True)
        self._fill_host[1].copy_(self.v.idx[:, tiles], non_blocking=True)
        self._fill_event.record()
        self._fill_pending = True

    def _regrow(self, keep: bool):
        """Moves both streams to worst-case slabs (every element of every token kept)."""
        self.regrow_events += 1
        for st, used in ((self.k, self._fill_known[0]), (self.v, self._fill_known[1])):
            if st.halves_per_token < HEAD_DIM:
                st._allocate(HEAD_DIM, used if keep else None)
        if self._p is not None:  # the long-lived parameter block holds slab pointers: refresh it IN PLACE (layer-batched
            self._fill_static(self._p)  # steps keep pointers to the block itself)
        self._p_stale = True
        self._streams_dirty = True

    def _ensure_room(self):
        """Called before a 256-token chunk is appended: regrow unless even an all-ties chunk is known to fit."""
        if self._worst_case():
            return
        self._resolve_fill()
        if (self._fill_known[0] + WORST_CHUNK_HALVES > self.k.head_capacity
                or self._fill_known[1] + WORST_CHUNK_HALVES > self.v.head_capacity):
            self._regrow(keep=True)

    # ------------------------------------------------------------------ compression
    def _compress_prompt(self, key_states: torch.Tensor, value_states: torch.Tensor, L: int, unit0: int = 0):
        """Prune + compress tokens [0, L) of the prompt, K and V, with ONE single-pass launch
        (`mfb200_compress_prefill`); the inputs are read in place through their strides.  `unit0`: first unit the
        given sequences map to (partial-batch prefill)."""
        self._streams_dirty = True

        def strided(x):
            ok = (x.dtype == torch.float16 and x.stride(3) == 1 and x.data_ptr() % 8 == 0
                  and all(st % 4 == 0 for st in x.stride()[:3]))
            x = x if ok else x.contiguous()
            return x, (C.c_int64 * 3)(x.stride(0), x.stride(1), x.stride(2))

        k, ks = strided(key_states)
        v, vs = strided(value_states)
        nblk = L // 64
        units = k.shape[0] * self.kv_heads
        if self._prefill_status is None or self._prefill_status.numel() < 2 * units * (nblk + 1):
            self._prefill_status = torch.empty(2 * units * (nblk + 1), dtype=torch.int64, device=self.device)
        sk, sv = self.k, self.v
        _lib.check(_lib.load().mfb200_compress_prefill(
            k.data_ptr(), v.data_ptr(), ks, vs, k.shape[0], self.kv_heads, L, prune_rank(self.k_sparsity),
            prune_rank(self.v_sparsity), sk.bmp[unit0:].data_ptr(), sk.idx[unit0:].data_ptr(), sk.nz.data_ptr(),
            sk.head_base[unit0:].data_ptr(), sv.bmp[unit0:].data_ptr(), sv.idx[unit0:].data_ptr(), sv.nz.data_ptr(),
            sv.head_base[unit0:].data_ptr(), sk.cap_tiles, sk.cap_tiles + 1,
            self.comp_len * 2, min(sk.head_capacity, sv.head_capacity), self.overflow.data_ptr(), self._prefill_status.data_ptr(),
            _lib.stream_ptr()), "mfb200_compress_prefill")

    def prefill(self, key_states: torch.Tensor, value_states: torch.Tensor, batch_start: int = 0):
        """key/value_states: fp16 [B, Hkv, T, 128] (post-RoPE).  llama_mustafar_kernel.py:416-442.

        `batch_start`: the given tensors hold sequences [batch_start, batch_start + B) of the cache's batch (lets a
        large batch be prefilled in slices without materialising all of its dense K/V at once); every slice must
        have the same length T."""
        b, h, t, d = key_states.shape
        assert (h, d) == (self.kv_heads, HEAD_DIM) and t <= self.cap_tokens and batch_start + b <= self.batch
        L = compressed_length(t, self.residual_length)
        u0, u1 = batch_start * h, (batch_start + b) * h
        self.comp_len = 0  # the prompt is compressed from tile 0
        with torch.cuda.device(self.device):
            if L > 0:
                self._compress_prompt(key_states, value_states, L, u0)
                if not self._worst_case():
                    # the prompt's fill level cannot be known in advance: one host read, then (rarely) a redo
                    tiles = L * 2
                    fill = torch.stack([self.k.idx[u0:u1, tiles].max(), self.v.idx[u0:u1, tiles].max(), self.overflow[0]]).cpu()
                    self._fill_known = [max(self._fill_known[0], 2 * int(fill[0])), max(self._fill_known[1], 2 * int(fill[1]))]
                    if int(fill[2]) != 0 or (self._fill_known[0] > self.k.head_capacity or self._fill_known[1] > self.v.head_capacity):
                        if batch_start != 0:
                            raise RuntimeError("MustafarKVCache.prefill: slab overflow in a later batch slice; "
                                               "construct the cache with nz_halves_per_token=128 for this data")
                        self.overflow.zero_()
                        self._fill_known = [0, 0]
                        self._regrow(keep=False)
                        self._compress_prompt(key_states, value_states, L, u0)
            lw = t - L
            assert lw <= self.win_cap
            self.k_win[u0:u1, :lw].copy_(key_states[:, :, L:].reshape(u1 - u0, lw, d))
            self.v_win[u0:u1, :lw].copy_(value_states[:, :, L:].reshape(u1 - u0, lw, d))
        self.comp_len = L
        self.win_len = lw
        self._p_stale = True

    def append(self, key_states: torch.Tensor, value_states: torch.Tensor):
        """Append the new token's k/v rows [B, Hkv, 1, 128] to the dense window (`:270`, `:309`)."""
        assert self.win_len < self.win_cap
        k = key_states.reshape(self.units, HEAD_DIM)
        v = value_states.reshape(self.units, HEAD_DIM)
        if not k.is_contiguous():
            k = k.contiguous()
        if not v.is_contiguous():
            v = v.contiguous()
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.mfb200_window_append(self.k_win.data_ptr(), self.v_win.data_ptr(), self.win_cap * HEAD_DIM,
                                                k.data_ptr(), v.data_ptr(), self.units, self.win_len, _lib.stream_ptr()),
                       "mfb200_window_append")
        self.win_len += 1
        self._p_stale = True

    def maybe_compress(self) -> bool:
        """`if (kv_seq_len - residual_length - compressed_length) % 256 == 0` — llama_mustafar_kernel.py:324-398."""
        if self.win_len - self.residual_length != COMPRESS_CHUNK:
            return False
        assert self.comp_len + COMPRESS_CHUNK <= self.cap_tokens, "cache capacity exceeded"
        rest = self.win_len - COMPRESS_CHUNK
        with torch.cuda.device(self.device):
            self._ensure_room()
            self._streams_dirty = True
            k, v = self.k, self.v
            _lib.check(_lib.load().mfb200_compress_append_chunk(
                self.k_win.data_ptr(), self.v_win.data_ptr(), self.win_cap * HEAD_DIM, self.units, self.win_len,
                prune_rank(self.k_sparsity), prune_rank(self.v_sparsity),
                k.bmp.data_ptr(), k.idx.data_ptr(), k.nz.data_ptr(), k.head_base.data_ptr(),
                v.bmp.data_ptr(), v.idx.data_ptr(), v.nz.data_ptr(), v.head_base.data_ptr(),
                k.cap_tiles, k.cap_tiles + 1, self.comp_len * 2, min(k.head_capacity, v.head_capacity), self.overflow.data_ptr(),
                torch.cuda.current_stream(self.device).cuda_stream), "mfb200_compress_append_chunk")
            self.comp_len += COMPRESS_CHUNK
            self.win_len = rest
            self._request_fill()
        self._p_stale = True
        return True

    def check_overflow(self):
        """Host-syncing check of the device-side overflow flag.  The capacity management above keeps it clear; it is
        the last line of defence for callers that drive the C ABI themselves."""
        if int(self.overflow.item()) != 0:
            raise RuntimeError("MustafarKVCache: packed-nonzero slab overflow; raise nz_halves_per_token")

    # ------------------------------------------------------------------ attention
    def _plan(self):
        key = (self.comp_len, self.win_len)
        hit = self._plan_cache.get(key)
        if hit is not None:
            return hit
        lib = _lib.load()
        ws, cb = C.c_size_t(0), C.c_size_t(0)
        n_split = _lib.check(lib.mfb200_decode_plan(self.batch, self.kv_heads, self.groups, self.comp_len,
                                                    self.win_len, self._sm_count, self.plan_hint, C.byref(ws), C.byref(cb)),
                             "mfb200_decode_plan")
        if len(self._plan_cache) > 4096:
            self._plan_cache.clear()
        self._plan_cache[key] = (n_split, ws.value)
        return n_split, ws.value

    def _params(self) -> _lib.DecodeParams:
        """The long-lived C parameter block of this cache (created on first use, with its workspace)."""
        if self._p is None:
            lib = _lib.load()
            self._sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
            self._p = self._static_params()
            if self._ws is None:
                nbytes = lib.mfb200_decode_workspace_max(self.batch, self.kv_heads, self.groups, self.cap_tokens,
                                                         self.win_cap, self._sm_count)
                if self.plan_hint:  # forced plans: up to one partial per block and window chunk
                    nbytes = max(nbytes, 8192 + self.units * (self.cap_tokens // 64 + self.win_cap // 64 + 2) * self.groups * 132 * 8)
                with torch.cuda.device(self.device):
                    self._ws = torch.zeros((nbytes,), dtype=torch.uint8, device=self.device)
                self._ws_bytes = nbytes
            self._p.workspace = self._ws.data_ptr()
            self._p.workspace_kb = self._ws_bytes // 1024  # the library refuses launches that would need more
            self._p_ref = C.byref(self._p)
            self._step = lib.mfb200_decode_step
            self._p_stale = True
        return self._p

    def _static_params(self) -> _lib.DecodeParams:
        """The fields of the C struct that never change for this cache (pointers of the slabs, strides)."""
        return self._fill_static(_lib.DecodeParams())

    def _fill_static(self, p: _lib.DecodeParams) -> _lib.DecodeParams:
        p.batch, p.kv_heads, p.groups = self.batch, self.kv_heads, self.groups
        p.score_div = math.sqrt(HEAD_DIM)
        p.slot_kb = self.slot_kb
        p.plan_hint = self.plan_hint
        p.k_bmp, p.k_idx, p.k_nz, p.k_nz_off = self.k.bmp.data_ptr(), self.k.idx.data_ptr(), self.k.nz.data_ptr(), self.k.nz_off.data_ptr()
        p.v_bmp, p.v_idx, p.v_nz, p.v_nz_off = self.v.bmp.data_ptr(), self.v.idx.data_ptr(), self.v.nz.data_ptr(), self.v.nz_off.data_ptr()
        p.bmp_stride, p.idx_stride = self.k.cap_tiles, self.k.cap_tiles + 1
        p.k_win, p.v_win, p.win_stride = self.k_win.data_ptr(), self.v_win.data_ptr(), self.win_cap * HEAD_DIM
        return p

    def _rope_fields(self, p: _lib.DecodeParams, rope) -> None:
        """rope = (cos, sin): fp16 rows of 128 entries, one per sequence ([B, 1, 128], what transformers' rotary_emb returns
        for position_ids [B, 1]) or one for all; None = q / k_new are already rotated."""
        if rope is None:
            p.rope_cos, p.rope_sin, p.rope_stride = None, None, 0
            return
        cos, sin = rope
        for t in (cos, sin):
            if not (t.is_cuda and t.dtype == torch.float16 and t.is_contiguous() and t.shape[-1] == HEAD_DIM
                    and t.numel() in (HEAD_DIM, self.batch * HEAD_DIM)):
                raise ValueError("rope: contiguous float16 CUDA cos/sin of shape [B, 1, 128] or [1, 1, 128] expected")
        if cos.numel() != sin.numel():
            raise ValueError("rope: cos and sin differ in shape")
        p.rope_cos, p.rope_sin = cos.data_ptr(), sin.data_ptr()
        p.rope_stride = HEAD_DIM if cos.numel() == self.batch * HEAD_DIM else 0

    def make_params(self, q: torch.Tensor, out: torch.Tensor, mask: Optional[torch.Tensor] = None,
                    k_new: Optional[torch.Tensor] = None, v_new: Optional[torch.Tensor] = None, rope=None) -> _lib.DecodeParams:
        """Fills the (cached) C parameter block for one launch at the cache's current lengths.  When k_new/v_new
        are given, self.win_len must already count the new token.  rope: see `_rope_fields` (fused rotary embedding)."""
        p = self._params()
        n_split, ws_bytes = self._plan()
        assert ws_bytes <= self._ws_bytes
        p.comp_len, p.win_len, p.n_split = self.comp_len, self.win_len, n_split
        self._p_stale = True  # the next fast-path step re-syncs lengths and clears the mask fields
        p.flags = self._flags()
        p.q, p.out = q.data_ptr(), out.data_ptr()
        if k_new is not None:
            p.k_new, p.v_new = k_new.data_ptr(), v_new.data_ptr()
        else:
            p.k_new, p.v_new = None, None
        if mask is not None:
            p.mask, p.mask_stride = mask.data_ptr(), mask.stride(0)
        else:
            p.mask, p.mask_stride = None, 0
        self._rope_fields(p, rope)
        return p

    def static_step_params(self, q: torch.Tensor, out: torch.Tensor, k_new: Optional[torch.Tensor], v_new: Optional[torch.Tensor],
                           win_len_dev: int, rope=None) -> _lib.DecodeParams:
        """An OWN parameter block (the cache's long-lived block stays host-stepped) for a STATIC decode step: the launch is
        planned for the most rows the window ever holds and reads the live window length from the int32 at the DEVICE address
        `win_len_dev`, so the same launch is valid at every step until the next compression - what a CUDA graph needs
        (include/mustafar_b200.h: mfb200_decode_params::win_len_dev).  Unmasked decode only."""
        b = _lib.DecodeParams()
        C.memmove(C.byref(b), C.byref(self.make_params(q, out, None, k_new, v_new, rope)), C.sizeof(b))
        self._rope_fields(self._p, None)  # the long-lived block is shared with the host-stepped fast path
        b.win_len = self.residual_length + COMPRESS_CHUNK
        b.n_split = 0  # planned by the library for that capacity
        b.win_len_dev = win_len_dev
        # inside a captured step nothing rewrites this cache's compressed streams: the early KV prefetch is safe
        b.flags = (_lib.F_REF_SCORE_ROUNDING if self.ref_score_rounding else 0) | ((_lib.F_PDL | _lib.F_PDL_EARLY_KV) if self.pdl else 0)
        return b

    def set_peer_output(self, peer) -> None:
        """Head-sharded decode: `peer` = a `_lib.PeerOut` block (partition.PeerOutput keeps it alive and up to date) whose
        buffers every launch of this cache also stores its output rows into; None switches it off again."""
        self._peer = peer
        self._params().peer = C.addressof(peer) if peer is not None else None

    def _flags(self) -> int:
        flags = _lib.F_REF_SCORE_ROUNDING if self.ref_score_rounding else 0
        if self.pdl:
            # early KV fetch only if this cache's compressed streams were not just rewritten
            flags |= _lib.F_PDL if self._streams_dirty else (_lib.F_PDL | _lib.F_PDL_EARLY_KV)
        self._streams_dirty = False
        return flags

    def _check_q(self, query_states):
        if not query_states.is_cuda or query_states.dtype != torch.float16:
            raise RuntimeError("sparse_decode_attention: query must be a float16 CUDA tensor (no CPU fallback)")
        b, hq, ql, d = query_states.shape
        assert ql == 1 and d == HEAD_DIM and b == self.batch and hq == self.kv_heads * self.groups
        return query_states if query_states.is_contiguous() else query_states.contiguous()

    def _mask2d(self, attention_mask):
        if attention_mask is None:
            return None
        if attention_mask.size() != (self.batch, 1, 1, self.kv_seq_len):
            raise ValueError(f"Attention mask should be of size {(self.batch, 1, 1, self.kv_seq_len)}, but is {attention_mask.size()}")
        return attention_mask.reshape(self.batch, self.kv_seq_len).to(torch.float16).contiguous()

    def _launch(self, p):
        with torch.cuda.device(self.device):
            rc = self._attn(C.byref(p), torch.cuda.current_stream(self.device).cuda_stream)
        if rc < 0:
            _lib.check(rc, "mfb200_sparse_decode_attention")

    def attend(self, query_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """query_states fp16 [B, Hq, 1, 128] -> attention output fp16 [B, Hq, 1, 128] over the cache as it is.

        attention_mask: HF additive mask [B, 1, 1, kv_seq_len] (llama_mustafar_kernel.py:293-301) or None.
        """
        q = self._check_q(query_states)
        if out is None:
            out = torch.empty_like(q)
        self._launch(self.make_params(q, out, self._mask2d(attention_mask)))
        return out

    def _sync_step_params(self):
        """Brings the long-lived parameter block up to date for a fast-path step (unmasked, fused append)."""
        p = self._p or self._params()
        if self._p_stale:  # lengths changed behind the block's back (prefill / compression / masked launch)
            p.comp_len, p.win_len, p.mask, p.mask_stride = self.comp_len, self.win_len, None, 0
            self._p_stale = False
        if self._streams_dirty or not (p.flags & _lib.F_PDL_EARLY_KV):
            p.flags = self._flags()
        return p

    def decode_step(self, query_states, key_states, value_states, attention_mask=None, out=None, rope=None):
        """One reference decode step of the attention block (llama_mustafar_kernel.py:256-398) in ONE launch:
        the new token's K/V rows [B, Hkv, 1, 128] are appended to the window by the attention kernel itself,
        which attends over compressed + window (incl. the new token); then the periodic compression.
        The unmasked case is a single FFI call (mfb200_decode_step) on the cache's long-lived parameter block.
        rope = (cos, sin) fuses the rotary embedding of `:238-253` into the launch: query_states / key_states are then the
        UNROTATED projections (see `_rope_fields`); the window receives the rotated key row."""
        q, k, v = query_states, key_states, value_states
        if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        if q.dtype != torch.float16 or not q.is_cuda or k.dtype != torch.float16 or v.dtype != torch.float16:
            raise RuntimeError("sparse_decode_attention: q/k/v must be float16 CUDA tensors (no CPU fallback)")
        if (q.numel() != self.units * self.groups * HEAD_DIM or k.numel() != self.units * HEAD_DIM
                or v.numel() != k.numel() or self.win_len >= self.win_cap):
            raise ValueError("decode_step: q [B,Hq,1,128], k/v [B,Hkv,1,128] expected and window capacity not exceeded")
        if out is None:
            out = torch.empty_like(q)
        if attention_mask is None and rope is None:
            self._sync_step_params()
            with torch.cuda.device(self.device):
                rc = self._step(self._p_ref, q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), self._sm_count,
                                torch.cuda.current_stream(self.device).cuda_stream)
            if rc < 0:
                self._p_stale = True  # the block was restored by the library; re-sync it before the next step anyway
                _lib.check(rc, "mfb200_decode_step")
            self.win_len += 1
        else:
            self.win_len += 1
            try:
                self._launch(self.make_params(q, out, self._mask2d(attention_mask), k, v, rope))
            except Exception:
                self.win_len -= 1
                raise
            finally:
                if self._p is not None:
                    self._rope_fields(self._p, None)
        self.maybe_compress()
        return out

    # ------------------------------------------------------------------ interop with the reference-shaped ops
    def as_reference_tuple(self):
        """(k_compressed, k_local_window, v_compressed, v_local_window, compressed_length, kv_seq_len) with the
        reference's exact container shapes (llama_mustafar_kernel.py:332, :337, :445) — copies, for interop/tests."""
        L = self.comp_len

        def one(st: _Stream):
            if L == 0:
                return None
            tiles = L * 2
            bmp = st.bmp[:, :tiles].contiguous()
            idx = st.idx[:, : tiles + 1].contiguous()
            tot = (idx[:, -1].to(torch.int64) * 2).cpu().tolist()
            nz = [st.nz[u * st.head_capacity: u * st.head_capacity + tot[u]].clone() for u in range(self.units)]
            t4 = idx[:, -1].to(torch.int64) // 4
            off = (torch.cumsum(t4, 0) - t4).to(torch.int32)
            return [bmp, idx, nz, off]

        kw = self.k_win[:, : self.win_len].reshape(self.batch, self.kv_heads, self.win_len, HEAD_DIM).clone()
        vw = self.v_win[:, : self.win_len].reshape(self.batch, self.kv_heads, self.win_len, HEAD_DIM).clone()
        return one(self.k), kw, one(self.v), vw, L, self.kv_seq_len

    def compressed_bytes(self):
        """(algorithmic bytes read by one attend(), idx excluded) — SURVEY.md §8(d).  Host-syncing."""
        tiles = self.comp_len * 2
        bmp = 2 * self.units * tiles * 8
        nz = 0
        if tiles:
            nz = 4 * int(self.k.idx[:, tiles].to(torch.int64).sum().item() + self.v.idx[:, tiles].to(torch.int64).sum().item())
        win = self.units * self.win_len * HEAD_DIM * 2 * 2
        qo = self.batch * self.kv_heads * self.groups * HEAD_DIM * 2 * 2
        return bmp + nz + win + qo


def mustafar_sparse_decode_attention(query_states: torch.Tensor, cache: MustafarKVCache,
                                     attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Functional form of `MustafarKVCache.attend` (the fused replacement of llama_mustafar_kernel.py:268-320)."""
    return cache.attend(query_states, attention_mask)


def decode_step_layers(caches, q: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """One decode step of a whole decoder's attention path as ONE FFI call (`mfb200_decode_step_layers`).

    caches: the per-layer `MustafarKVCache`s (same geometry, same device, unmasked);
    q [layers, B, Hq, 128], k_new / v_new [layers, B, Hkv, 128], out [layers, B, Hq, 128]: fp16, contiguous.
    Equivalent to `caches[l].decode_step(q[l], k_new[l], v_new[l], out=out[l])` for every layer."""
    n = len(caches)
    c0 = caches[0]
    if not (q.is_cuda and q.dtype == k_new.dtype == v_new.dtype == out.dtype == torch.float16):
        raise RuntimeError("decode_step_layers: float16 CUDA tensors expected (no CPU fallback)")
    if not (q.is_contiguous() and k_new.is_contiguous() and v_new.is_contiguous() and out.is_contiguous()):
        raise ValueError("decode_step_layers: contiguous [layers, ...] tensors expected")
    if (q.shape[0] != n or k_new.shape[0] != n or v_new.shape[0] != n or out.shape[0] != n
            or q[0].numel() != c0.units * c0.groups * HEAD_DIM or k_new[0].numel() != c0.units * HEAD_DIM
            or v_new[0].numel() != k_new[0].numel() or out[0].numel() != q[0].numel()):
        raise ValueError("decode_step_layers: q/out [layers,B,Hq,128], k_new/v_new [layers,B,Hkv,128] expected")
    arr = getattr(c0, "_layer_arr", None)
    if arr is None or arr[1] != tuple(id(c._p) for c in caches) or any(c._p is None for c in caches):
        for c in caches:
            c._params()
        ptrs = (C.POINTER(_lib.DecodeParams) * n)(*[C.pointer(c._p) for c in caches])
        arr = c0._layer_arr = (ptrs, tuple(id(c._p) for c in caches))
    for c in caches:
        if c.win_len >= c.win_cap:
            raise ValueError("decode_step_layers: window capacity exceeded")
        c._sync_step_params()
    with torch.cuda.device(c0.device):
        rc = _lib.load().mfb200_decode_step_layers(arr[0], n, q.data_ptr(), k_new.data_ptr(), v_new.data_ptr(), out.data_ptr(),
                                                   q[0].numel(), k_new[0].numel(), out[0].numel(),
                                                   torch.cuda.current_stream(c0.device).cuda_stream)
    if rc < 0:
        for c in caches:  # layers before the failing one advanced: re-read the lengths from the blocks
            c.win_len = c._p.win_len
            c._p_stale = True
        _lib.check(rc, "mfb200_decode_step_layers")
    for c in caches:
        c.win_len += 1
        c.maybe_compress()
    return out


class DecodeStepGraph:
    """A decoder's attention path of one decode step as ONE CUDA-graph launch (`mfb200_decode_layers_static`).

    The window length of every layer lives in device memory (`mfb200_decode_params::win_len_dev`), so the launches of a step
    are identical from step to step: `lengths += 1` followed by one fused (append + attention) launch per layer, captured
    once and replayed.  The launch is planned for the window's CAPACITY; window chunks that are still empty exit at once.
    Every 256 tokens the reference schedule compresses the window (llama_mustafar_kernel.py:324): that step runs the
    compression launches eagerly, rewrites the device lengths and captures the graph again (the compressed length, and with
    it the work decomposition, has changed).

    q [layers, B, Hq, 128], k_new / v_new [layers, B, Hkv, 128], out [layers, B, Hq, 128] are STATIC fp16 buffers: the caller
    writes the step's inputs into them before `step()` and reads `out` after it (stream order).  Unmasked decode only.
    """

    def __init__(self, caches, q: torch.Tensor, k_new: torch.Tensor, v_new: torch.Tensor, out: torch.Tensor):
        self.caches, self.q, self.k_new, self.v_new, self.out = list(caches), q, k_new, v_new, out
        c0 = self.caches[0]
        n = len(self.caches)
        if not (q.is_cuda and q.dtype == k_new.dtype == v_new.dtype == out.dtype == torch.float16
                and q.is_contiguous() and k_new.is_contiguous() and v_new.is_contiguous() and out.is_contiguous()):
            raise RuntimeError("DecodeStepGraph: contiguous float16 CUDA buffers expected (no CPU fallback)")
        if (q.shape[0] != n or q[0].numel() != c0.units * c0.groups * HEAD_DIM or k_new[0].numel() != c0.units * HEAD_DIM
                or v_new.shape != k_new.shape or out.shape != q.shape):
            raise ValueError("DecodeStepGraph: q/out [layers,B,Hq,128], k_new/v_new [layers,B,Hkv,128] expected")
        self.device = c0.device
        with torch.cuda.device(self.device):
            self.lengths = torch.zeros((n,), dtype=torch.int32, device=self.device)
        self._lib = _lib.load()
        self._blocks = [_lib.DecodeParams() for _ in range(n)]  # own parameter blocks: the caches' blocks stay host-stepped
        self._arr = (C.POINTER(_lib.DecodeParams) * n)(*[C.pointer(b) for b in self._blocks])
        self._graph = None
        self._captured_comp = None
        self.captures = 0

    def _capture(self):
        n = len(self.caches)
        for l, (c, b) in enumerate(zip(self.caches, self._blocks)):
            C.memmove(C.byref(b), C.byref(c.static_step_params(self.q[l], self.out[l], self.k_new[l], self.v_new[l],
                                                               self.lengths[l:].data_ptr())), C.sizeof(b))
        self.lengths.copy_(torch.tensor([c.win_len for c in self.caches], dtype=torch.int32), non_blocking=False)
        sp = torch.cuda.current_stream(self.device).cuda_stream
        # one eager launch of every kernel the graph contains before capturing (lazy module loading, kernel attributes):
        # lengths += 0, and the layers without the fused append (side-effect free: `out` is overwritten by the real step)
        _lib.check(self._lib.mfb200_lengths_add(self.lengths.data_ptr(), n, 0, sp), "mfb200_lengths_add")
        for b in self._blocks:
            b.k_new, b.v_new = None, None
        _lib.check(self._lib.mfb200_decode_layers_static(self._arr, n, sp), "mfb200_decode_layers_static")
        for l, b in enumerate(self._blocks):
            b.k_new, b.v_new = self.k_new[l].data_ptr(), self.v_new[l].data_ptr()
        torch.cuda.current_stream(self.device).synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self._lib.mfb200_lengths_add(self.lengths.data_ptr(), n, 1, cs), "mfb200_lengths_add")
            _lib.check(self._lib.mfb200_decode_layers_static(self._arr, n, cs), "mfb200_decode_layers_static")
        self._graph = g
        self._captured_comp = tuple(c.comp_len for c in self.caches)
        self.captures += 1

    def step(self) -> torch.Tensor:
        """One decode step for all layers: replays the graph; returns `out` (valid in stream order)."""
        for c in self.caches:
            if c.win_len + 1 > c.residual_length + COMPRESS_CHUNK:
                raise ValueError("DecodeStepGraph: window capacity exceeded")
        if self._graph is None or self._captured_comp != tuple(c.comp_len for c in self.caches):
            with torch.cuda.device(self.device):
                self._capture()
        self._graph.replay()
        compressed = False
        for c in self.caches:
            c.win_len += 1
            c._p_stale = True
            compressed = c.maybe_compress() or compressed
        if compressed:
            self._graph = None  # lengths and comp_len changed: the next step captures again
        return self.out
