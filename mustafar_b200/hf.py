"""Model-level caller for transformers >= 5: a `Cache` + a registered attention function.

The reference ships 970-line clones of the HF Llama/Mistral model files whose attention block carries the
sparse cache as a tuple in `past_key_value` (models/llama_mustafar_kernel.py:199-457, tuple layout `:445`);
those clones no longer import under transformers 5.  Here the same per-layer schedule is driven
from the stock model classes through the two extension points transformers 5 offers:

* `MustafarCache` — a `transformers.Cache` whose layers own a `MustafarKVCache` each (compressed streams + dense
  residual window), created from the model config;
* `mustafar_attention_forward` — registered as attention implementation ``"mustafar"``:
  - prefill (`q_len > 1`, `:400-445`): dense causal attention through the library (the reference calls
    `flash_attn_func`, `:509`; here torch SDPA), then prune + compress tokens `[0, L)` and keep the rest as window;
  - decode (`q_len == 1`, `:256-398`): ONE fused launch (append + compressed/window attention + split merge),
    then the every-256-token compression.

Usage::

    import mustafar_b200.hf as mhf            # registers "mustafar"
    model = AutoModelForCausalLM.from_pretrained(..., torch_dtype=torch.float16, attn_implementation="mustafar")
    cache = mhf.MustafarCache(model.config, k_sparsity=0.5, v_sparsity=0.5, max_tokens=8192)
    model.generate(**inputs, past_key_values=cache)

    # or, with the host out of the decode loop: the WHOLE decode step (every layer's projections, RoPE, the fused append +
    # sparse attention launch, MLP, lm_head, argmax) captured once as a CUDA graph and replayed per token
    tokens = mhf.GraphedDecoder(model, cache).generate(input_ids, max_new_tokens=1024)

There is no CPU fallback: the cache and the decode path need the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math
import sys
import threading
from typing import Optional

import torch
from transformers import AttentionInterface, AttentionMaskInterface
from transformers.cache_utils import Cache, CacheLayerMixin
from transformers.masking_utils import sdpa_mask

from . import _lib
from .attention import HEAD_DIM, MustafarKVCache

ATTN_NAME = "mustafar"

# The attention-interface function receives the attention MODULE (with its `layer_idx`) but not `past_key_values`.
# `Cache.update` runs in the same module forward immediately before it, so the cache that was updated last on this
# thread is the one the attention call belongs to.
_ACTIVE = threading.local()


class MustafarLayer(CacheLayerMixin):
    """One decoder layer's sparse KV cache.  `update` only records the step's new K/V rows; the registered attention
    function consumes them (the fused kernel appends the new row itself, so nothing is copied here)."""

    is_sliding = False
    is_compileable = False

    def __init__(self, groups: int, k_sparsity: float, v_sparsity: float, residual_length: int, max_tokens: int):
        super().__init__()
        self.groups, self.k_sparsity, self.v_sparsity = groups, k_sparsity, v_sparsity
        self.residual_length, self.max_tokens = residual_length, max_tokens
        self.kv: Optional[MustafarKVCache] = None
        self.seen_tokens = 0

    def lazy_initialization(self, key_states: torch.Tensor, value_states: torch.Tensor) -> None:
        b, hkv, _, d = key_states.shape
        if d != HEAD_DIM or key_states.dtype != torch.float16 or not key_states.is_cuda:
            raise RuntimeError("MustafarCache: fp16 CUDA key/value states with head_dim 128 expected (no CPU fallback)")
        self.dtype, self.device = key_states.dtype, key_states.device
        self.kv = MustafarKVCache(b, hkv, self.groups, self.max_tokens, self.k_sparsity, self.v_sparsity,
                                  self.residual_length, device=self.device)
        self.is_initialized = True

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, *args, **kwargs):
        if not self.is_initialized:
            self.lazy_initialization(key_states, value_states)
        n = key_states.shape[-2]
        if n > 1 and self.seen_tokens > 0:
            raise NotImplementedError("MustafarCache: multi-token continuation (chunked prefill) is not supported; "
                                      "the reference supports a single prefill followed by 1-token decode steps")
        self.seen_tokens += n
        return key_states, value_states

    def get_seq_length(self) -> int:
        return self.seen_tokens

    def get_mask_sizes(self, query_length: int) -> tuple[int, int]:
        return self.seen_tokens + query_length, 0

    def get_max_cache_shape(self) -> int:
        return -1

    def reset(self) -> None:
        self.seen_tokens = 0
        if self.kv is not None:
            self.kv.comp_len = 0
            self.kv.win_len = 0
            self.kv._p_stale = True

    def offload(self):
        raise NotImplementedError("MustafarCache layers live on the GPU")

    def prefetch(self):
        pass

    def reorder_cache(self, beam_idx):
        raise NotImplementedError("beam search is not supported by the compressed cache")


class MustafarCache(Cache):
    """`past_key_values` for attention implementation "mustafar" (reference tuple: llama_mustafar_kernel.py:445)."""

    def __init__(self, config, k_sparsity: float = 0.5, v_sparsity: float = 0.5, residual_length: int = 32,
                 max_tokens: int = 4096):
        text = config.get_text_config() if hasattr(config, "get_text_config") else config
        groups = text.num_attention_heads // text.num_key_value_heads
        head_dim = getattr(text, "head_dim", None) or text.hidden_size // text.num_attention_heads
        if head_dim != HEAD_DIM or groups not in (1, 2, 4, 8):
            raise ValueError(f"MustafarCache: head_dim {head_dim} / {groups} query heads per KV head not supported "
                             "(head_dim 128 and 1, 2, 4 or 8 heads per KV head)")
        super().__init__(layers=[MustafarLayer(groups, k_sparsity, v_sparsity, residual_length, max_tokens)
                                 for _ in range(text.num_hidden_layers)])
        self._static: Optional["_StaticStep"] = None  # set by GraphedDecoder while it captures a decode step

    def update(self, key_states, value_states, layer_idx, *args, **kwargs):
        _ACTIVE.cache = self  # consumed by mustafar_attention_forward through module.layer_idx
        return super().update(key_states, value_states, layer_idx, *args, **kwargs)

    def bytes_held(self) -> int:
        """Device bytes of all layer caches (the reference reports torch.cuda.max_memory_allocated, mem_spd_test.py:95)."""
        return sum(l.kv.bytes_held() for l in self.layers if l.kv is not None)


def mustafar_attention_forward(module, query, key, value, attention_mask, scaling=None, dropout=0.0, **kwargs):
    """transformers attention-interface function; see the module docstring.  Returns ([B, q_len, Hq, 128], None)."""
    cache: Optional[MustafarCache] = getattr(_ACTIVE, "cache", None)
    idx = getattr(module, "layer_idx", None)
    if cache is None or idx is None or idx >= len(cache.layers) or not cache.layers[idx].is_initialized:
        raise RuntimeError('attn_implementation="mustafar" needs past_key_values=MustafarCache(config, ...)')
    layer: MustafarLayer = cache.layers[idx]
    if dropout:
        raise RuntimeError("mustafar attention: dropout is not supported (inference path)")
    b, hq, q_len, d = query.shape
    if scaling is not None and abs(scaling * math.sqrt(HEAD_DIM) - 1.0) > 1e-6:
        raise ValueError("mustafar attention: scores are scaled by 1/sqrt(128) (llama_mustafar_kernel.py:284)")
    if q_len > 1:
        # prefill: the whole prompt attends densely (the reference's flash_attn_func call), then it is compressed
        out = torch.nn.functional.scaled_dot_product_attention(
            query, key, value, attn_mask=attention_mask, is_causal=attention_mask is None, scale=scaling,
            enable_gqa=hq != key.shape[1])
        layer.kv.prefill(key, value)
        return out.transpose(1, 2).contiguous(), None
    if cache._static is not None:  # being captured into a CUDA graph: window lengths live on the device
        if attention_mask is not None:
            raise RuntimeError("mustafar attention: the graphed decode step is unmasked (no padded sequences)")
        return cache._static.launch(idx, layer.kv, query, key, value).transpose(1, 2), None
    mask = None
    if attention_mask is not None:  # [B, 1, 1, kv_len]: bool (True = attend) or additive
        if attention_mask.dtype == torch.bool:
            # one additive mask per decode step, shared by all layers (the model hands every layer the same tensor)
            memo = getattr(_ACTIVE, "mask_memo", None)
            if memo is None or memo[0] is not attention_mask:
                add = torch.zeros(attention_mask.shape, dtype=torch.float16, device=query.device)
                add.masked_fill_(~attention_mask, float("-inf"))
                memo = _ACTIVE.mask_memo = (attention_mask, add)
            mask = memo[1]
        else:
            mask = attention_mask
    out = layer.kv.decode_step(query, key, value, mask)  # [B, Hq, 1, 128]
    return out.transpose(1, 2), None


AttentionInterface.register(ATTN_NAME, mustafar_attention_forward)


def mustafar_mask(*args, **kwargs):
    """Mask function of attention implementation "mustafar": None when nothing is masked, else a boolean [B,1,q,kv] mask.
    A 1-token decode step without a padding mask attends to everything; `sdpa_mask` knows that too, but refuses to skip the
    mask while a CUDA stream is capturing (transformers.masking_utils._ignore_causal_mask_sdpa: `not is_tracing()`)."""
    if kwargs.get("q_length") == 1 and kwargs.get("attention_mask") is None:
        return None
    return sdpa_mask(*args, **kwargs)


AttentionMaskInterface.register(ATTN_NAME, mustafar_mask)


class _StaticStep:
    """Per-capture state of a graphed decode step: the window length of every layer in DEVICE memory, bumped by one tiny
    launch in front of layer 0, and read by every layer's fused launch (`MustafarKVCache.static_step_params`)."""

    def __init__(self, layers, device):
        self.n = len(layers)
        self.device = device
        self.lengths = torch.zeros((self.n,), dtype=torch.int32, device=device)
        self.lengths.copy_(torch.tensor([l.kv.win_len for l in layers], dtype=torch.int32))
        self.dry = False  # dry run before the capture: no append, lengths unchanged (nothing in the cache is modified)
        self.rope = None  # (cos, sin) of the step when the model's rotary embedding was deferred to the attention launch
        self._lib = _lib.load()

    def launch(self, idx: int, kv: MustafarKVCache, q, k, v):
        if not (q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
            q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        if q.dtype != torch.float16 or k.dtype != torch.float16 or v.dtype != torch.float16 or not q.is_cuda:
            raise RuntimeError("mustafar attention: float16 CUDA q/k/v expected (no CPU fallback)")
        out = torch.empty_like(q)
        if idx == 0:
            with torch.cuda.device(self.device):
                _lib.check(self._lib.mfb200_lengths_add(self.lengths.data_ptr(), self.n, 0 if self.dry else 1,
                                                        torch.cuda.current_stream(self.device).cuda_stream), "mfb200_lengths_add")
        p = kv.static_step_params(q, out, None if self.dry else k, None if self.dry else v, self.lengths[idx:].data_ptr(), self.rope)
        kv._launch(p)
        return out


class _DeferredRope:
    """While a decode step is captured, the model file's `apply_rotary_pos_emb` hands q / k through unrotated and leaves
    (cos, sin) with the step: the fused attention launch rotates q at staging and the new K row at the append
    (mfb200_decode_params::rope_cos) - ten elementwise launches per layer less.  Model files without that function, or
    calls that are not a 1-token fp16 head_dim-128 rotation, keep the stock path."""

    def __init__(self, model, step: Optional[_StaticStep], enabled: bool):
        self.step = step
        self.mod = sys.modules.get(type(model).__module__) if (enabled and step is not None) else None
        self.orig = getattr(self.mod, "apply_rotary_pos_emb", None) if self.mod is not None else None

    def __enter__(self):
        if self.orig is None:
            return self
        orig, step, owner = self.orig, self.step, threading.get_ident()

        def deferred(q, k, cos, sin, *args, **kwargs):
            if threading.get_ident() != owner:  # the function is a module global: other threads keep the stock behaviour
                return orig(q, k, cos, sin, *args, **kwargs)
            if (q.shape[-2] == 1 and k.shape[-2] == 1 and q.shape[-1] == HEAD_DIM and cos.shape[-1] == HEAD_DIM
                    and q.dtype == cos.dtype == sin.dtype == torch.float16 and cos.is_contiguous() and sin.is_contiguous()
                    and cos.numel() in (HEAD_DIM, q.shape[0] * HEAD_DIM)):
                step.rope = (cos, sin)
                return q, k
            step.rope = None
            return orig(q, k, cos, sin, *args, **kwargs)

        self.mod.apply_rotary_pos_emb = deferred
        return self

    def __exit__(self, *exc):
        if self.orig is not None:
            self.mod.apply_rotary_pos_emb = self.orig
        return False


class GraphedDecoder:
    """Greedy decoding of a stock transformers causal LM over a `MustafarCache` with the host out of the decode loop.

    The reference's decode step is ~15 launches of glue around two kernels per layer, driven from Python
    (llama_mustafar_kernel.py:238-320, :453); stock `generate()` spends ~0.4 ms of host time per layer at batch 1.  Here the
    step - embedding, every layer's norms / projections / RoPE / fused (append + sparse attention) launch / MLP, final norm,
    lm_head, argmax and the feedback of the new token - is captured ONCE into a CUDA graph and replayed per token: every
    layer's window length lives in device memory, so the launches are identical from step to step.  The rotary embedding
    of q and of the new K row is done inside the attention launch (`fuse_rope`, see `_DeferredRope`).  Every 256 tokens the
    reference schedule compresses the window (`:324`); that step runs the compression launches eagerly and the graph is
    captured again (the compressed length, and with it the work decomposition, changed).

    Prefill runs eagerly through the same model.  Unmasked sequences only (all prompts of one length, no padding)."""

    def __init__(self, model, cache: MustafarCache, max_new_tokens: int = 4096, fuse_rope: bool = True):
        if getattr(model.config, "_attn_implementation", None) != ATTN_NAME:
            raise ValueError(f'GraphedDecoder: the model must use attn_implementation="{ATTN_NAME}"')
        self.model, self.cache = model, cache
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("GraphedDecoder: CUDA model expected (no CPU fallback)")
        self.capacity = max_new_tokens
        self.fuse_rope = fuse_rope
        self._graph = None
        self._stale = True  # the captured launches no longer match the cache (or nothing is captured yet)
        self.captures = 0
        self.rope_fused = False
        self._filled = 0
        self.ids = self.pos = self.slot = self.tokens = self.logits = None

    def _buffers(self, batch: int):
        dev = self.device
        self.ids = torch.zeros((batch, 1), dtype=torch.long, device=dev)      # the token every sequence feeds next
        self.pos = torch.zeros((batch, 1), dtype=torch.long, device=dev)      # its position
        self.slot = torch.zeros((batch, 1), dtype=torch.long, device=dev)     # where the step's new token goes in `tokens`
        self.tokens = torch.zeros((batch, self.capacity), dtype=torch.long, device=dev)

    def _body(self):
        with _DeferredRope(self.model, self.cache._static, self.fuse_rope):
            out = self.model(input_ids=self.ids, position_ids=self.pos, past_key_values=self.cache, use_cache=True)
        logits = out.logits[:, -1]
        nxt = logits.argmax(-1, keepdim=True)
        self.tokens.scatter_(1, self.slot, nxt)
        self.ids.copy_(nxt)
        self.pos.add_(1)
        self.slot.add_(1)
        return logits

    @torch.no_grad()
    def _capture(self):
        layers = self.cache.layers
        seen = [l.seen_tokens for l in layers]
        st = self.cache._static = _StaticStep(layers, self.device)
        try:
            # dry run of every kernel of the step on a side stream (library handles, lazy module loading) with no effect on
            # the cache; the feedback buffers it overwrites are restored
            if self.captures == 0:  # re-captures after a compression launch the same kernels: nothing left to warm up
                keep = (self.ids.clone(), self.pos.clone(), self.slot.clone())
                st.dry = True
                side = torch.cuda.Stream(self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._body()
                torch.cuda.current_stream(self.device).wait_stream(side)
                st.dry = False
                self.ids.copy_(keep[0]); self.pos.copy_(keep[1]); self.slot.copy_(keep[2])
            torch.cuda.current_stream(self.device).synchronize()
            # the previous capture stays alive until the new one exists: both live in one memory pool, whose blocks the new
            # capture reuses as soon as the old graph is dropped
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self._graph.pool() if self._graph is not None else None):
                self.logits = self._body()
            self._graph, self._stale = g, False
            self.rope_fused = st.rope is not None  # the rotary embedding runs inside the attention launches of this graph
            self._static_keepalive = st  # the graph's launches read st.lengths
            self.captures += 1
        finally:
            self.cache._static = None
            for l, n in zip(layers, seen):  # the dry run and the capture went through Cache.update without decoding anything
                l.seen_tokens = n

    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor) -> torch.Tensor:
        """Runs the prompt [B, T] eagerly (dense attention + prune/compress of tokens [0, L)); returns the first generated
        token of every sequence [B] and arms the step buffers."""
        if self.cache.get_seq_length() != 0:
            raise ValueError("GraphedDecoder.prefill: the cache already holds tokens")
        b, t = input_ids.shape
        self._buffers(b)
        out = self.model(input_ids=input_ids, past_key_values=self.cache, use_cache=True, logits_to_keep=1)
        first = out.logits[:, -1].argmax(-1, keepdim=True)
        self.tokens[:, :1] = first
        self.ids.copy_(first)
        self.pos.fill_(t)
        self.slot.fill_(1)
        self._filled = 1  # host mirror of `slot`: tokens recorded so far
        self._stale = True
        return first[:, 0]

    @torch.no_grad()
    def step(self, token_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One decode step for every sequence: feeds `token_ids` [B, 1] (default: the tokens the previous step chose), replays
        the graph and returns the step's logits [B, vocab] (a static buffer, valid in stream order until the next step)."""
        layers = self.cache.layers
        if self.ids is None:
            raise ValueError("GraphedDecoder.step: call prefill() first")
        if self._filled >= self.capacity:
            raise ValueError(f"GraphedDecoder.step: the token buffer holds max_new_tokens = {self.capacity} tokens")
        for l in layers:
            if l.kv.win_len + 1 > l.kv.win_cap:
                raise ValueError("GraphedDecoder.step: window capacity exceeded")
        if token_ids is not None:
            self.ids.copy_(token_ids.view_as(self.ids))
        if self._stale:
            self._capture()
        self._graph.replay()
        self._filled += 1
        compressed = False
        for l in layers:  # host mirrors of what the graph did on the device, then the reference's compression schedule
            l.seen_tokens += 1
            l.kv.win_len += 1
            l.kv._p_stale = True
            compressed = l.kv.maybe_compress() or compressed
        if compressed:
            self._stale = True  # lengths and compressed length changed: capture again at the next step
        return self.logits

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, max_new_tokens: int) -> torch.Tensor:
        """Greedy continuation: [B, T] -> [B, T + max_new_tokens] (no eos handling: the benchmark harness disables it too)."""
        if max_new_tokens > self.capacity:
            raise ValueError(f"GraphedDecoder.generate: max_new_tokens {max_new_tokens} > capacity {self.capacity}")
        self.prefill(input_ids)
        for _ in range(max_new_tokens - 1):
            self.step()
        return torch.cat([input_ids, self.tokens[:, :max_new_tokens]], dim=1)
