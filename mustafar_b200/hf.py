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

There is no CPU fallback: the cache and the decode path need the CUDA library.
"""
from __future__ import annotations

import math
import threading
from typing import Optional

import torch
from transformers import AttentionInterface, AttentionMaskInterface
from transformers.cache_utils import Cache, CacheLayerMixin
from transformers.masking_utils import sdpa_mask

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
AttentionMaskInterface.register(ATTN_NAME, sdpa_mask)  # None when nothing is masked, else a boolean [B,1,q,kv] mask
