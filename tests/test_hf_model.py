"""Model-level parity: stock transformers-5 Llama driven through `mustafar_b200.hf` (MustafarCache + the "mustafar"
attention function) against the reference's masked-dense formulation (llama_mustafar_Kt_Mag_Vt_Mag.py:863-974: dense
attention over a KV cache whose rows are pruned in place on the compression schedule), same weights, teacher forced."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mustafar_oracle as O  # noqa: E402  (test infrastructure)


def test_hf_adapter_imports_and_validates_on_cpu():
    from transformers import AttentionInterface, LlamaConfig
    import mustafar_b200.hf as mhf
    assert mhf.ATTN_NAME in AttentionInterface._global_mapping
    good = LlamaConfig(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128, num_hidden_layers=3)
    cache = mhf.MustafarCache(good, 0.5, 0.7, max_tokens=1024)
    assert len(cache.layers) == 3 and cache.get_seq_length() == 0
    with pytest.raises(ValueError):
        mhf.MustafarCache(LlamaConfig(hidden_size=256, num_attention_heads=4, num_key_value_heads=4, num_hidden_layers=1))  # head_dim 64
    with pytest.raises(RuntimeError):  # CPU tensors: no fallback
        cache.update(torch.zeros(1, 1, 4, 128, dtype=torch.float16), torch.zeros(1, 1, 4, 128, dtype=torch.float16), 0)


def _masked_dense_cache(sparsity, residual=32):
    from transformers.cache_utils import Cache, DynamicLayer

    class MaskedDenseLayer(DynamicLayer):
        """Dense K/V whose rows are pruned in place exactly when the sparse cache would compress them."""

        def __init__(self):
            super().__init__()
            self.comp_len = 0

        def _prune(self, lo, hi):
            for name in ("keys", "values"):
                x = getattr(self, name).clone()  # out of place: the tensors handed to attention stay as they were
                rows = x[:, :, lo:hi].cpu().numpy()
                x[:, :, lo:hi] = torch.from_numpy(O.prune_rows(rows, sparsity)).to(x.device)
                setattr(self, name, x)

        def update(self, key_states, value_states, *args, **kwargs):
            first = self.get_seq_length() == 0
            k, v = super().update(key_states, value_states, *args, **kwargs)
            T = k.shape[-2]
            if first:  # prefill: the prompt attends densely, then [0, L) is pruned (llama_mustafar_kernel.py:416-442)
                L = O.compressed_length(T, residual)
                if L:
                    self._prune(0, L)
                self.comp_len = L
            elif T - self.comp_len - residual == 256:  # `:324`: compress the 256 oldest window rows after attending
                self._prune(self.comp_len, self.comp_len + 256)
                self.comp_len += 256
            return k, v

    return Cache(layer_class_to_replicate=MaskedDenseLayer)


@pytest.mark.gpu
@pytest.mark.parametrize("family,heads,hkv,sparsity", [("llama", 2, 2, 0.5), ("llama", 2, 1, 0.7),
                                                       ("mistral", 4, 1, 0.5)])  # Mistral-7B's 4 query heads per KV head
def test_llama_decode_matches_masked_dense(family, heads, hkv, sparsity):
    """Stock Llama and Mistral classes (the reference clones both: llama_mustafar_kernel.py,
    mistral_mustafar_Kt_Mag_Vt_Mag.py:440-669) through the same cache / attention function."""
    import mustafar_b200.hf as mhf
    if family == "llama":
        from transformers import LlamaConfig as Config, LlamaForCausalLM as Model
        extra = {}
    else:
        from transformers import MistralConfig as Config, MistralForCausalLM as Model
        extra = {"sliding_window": None}  # Mistral-7B-Instruct-v0.2 attends over the full context
    cfg = Config(vocab_size=512, hidden_size=128 * heads, intermediate_size=512, num_hidden_layers=2, num_attention_heads=heads,
                 num_key_value_heads=hkv, head_dim=128, max_position_embeddings=2048, **extra)
    torch.manual_seed(0)
    model = Model(cfg).half().cuda().eval()
    batch, T0, steps = 2, 300, 270  # window 44 -> 288 at T = 544: one compression event inside the run
    ids = torch.randint(0, cfg.vocab_size, (batch, T0 + steps), generator=torch.Generator().manual_seed(1)).cuda()

    def run(impl, cache):
        model.config._attn_implementation = impl
        outs = []
        with torch.no_grad():
            outs.append(model(input_ids=ids[:, :T0], past_key_values=cache, use_cache=True).logits[:, -1].float())
            for t in range(T0, T0 + steps):
                outs.append(model(input_ids=ids[:, t:t + 1], past_key_values=cache, use_cache=True).logits[:, -1].float())
        return torch.stack(outs)

    ref = run("eager", _masked_dense_cache(sparsity))
    cache = mhf.MustafarCache(cfg, sparsity, sparsity, max_tokens=T0 + steps + 8)
    got = run(mhf.ATTN_NAME, cache)
    kv = cache.layers[0].kv
    assert kv.comp_len == 512 and kv.kv_seq_len == T0 + steps and cache.get_seq_length() == T0 + steps
    kv.check_overflow()
    d = (got - ref).abs()
    scale = ref.abs().max().item()
    # fp16 model, two layers: attention outputs agree to ~2.5e-4 (test_gpu_parity), logits to ~1e-3 of their range
    assert d.max().item() <= 2e-2 * scale and d.mean().item() <= 2e-3 * scale, (d.max().item(), d.mean().item(), scale)
    assert (got.argmax(-1) == ref.argmax(-1)).float().mean().item() > 0.97


@pytest.mark.gpu
@pytest.mark.parametrize("family,heads,hkv,batch,fuse_rope", [("llama", 2, 2, 2, True), ("llama", 4, 1, 3, True),
                                                              ("mistral", 4, 1, 2, True), ("llama", 2, 1, 1, False)])
def test_graphed_decoder_equals_eager_steps(family, heads, hkv, batch, fuse_rope):
    """`GraphedDecoder` (whole decode step = one CUDA-graph replay, device-side window lengths) against the same model stepped
    eagerly through `MustafarCache`, teacher forced, across a compression event (window 44 -> 288 -> 32) and the re-capture."""
    import mustafar_b200.hf as mhf
    if family == "llama":
        from transformers import LlamaConfig as Config, LlamaForCausalLM as Model
        extra = {}
    else:
        from transformers import MistralConfig as Config, MistralForCausalLM as Model
        extra = {"sliding_window": None}
    cfg = Config(vocab_size=512, hidden_size=128 * heads, intermediate_size=512, num_hidden_layers=3, num_attention_heads=heads,
                 num_key_value_heads=hkv, head_dim=128, max_position_embeddings=2048, **extra)
    torch.manual_seed(0)
    model = Model(cfg).half().cuda().eval()
    model.config._attn_implementation = mhf.ATTN_NAME
    T0, steps = 300, 270
    ids = torch.randint(0, cfg.vocab_size, (batch, T0 + steps), generator=torch.Generator().manual_seed(2)).cuda()

    eager_cache = mhf.MustafarCache(cfg, 0.5, 0.5, max_tokens=T0 + steps + 8)
    ref = []
    with torch.no_grad():
        model(input_ids=ids[:, :T0], past_key_values=eager_cache, use_cache=True)
        for t in range(T0, T0 + steps):
            ref.append(model(input_ids=ids[:, t:t + 1], past_key_values=eager_cache, use_cache=True).logits[:, -1].float())
    ref = torch.stack(ref)

    cache = mhf.MustafarCache(cfg, 0.5, 0.5, max_tokens=T0 + steps + 8)
    dec = mhf.GraphedDecoder(model, cache, max_new_tokens=steps + 1, fuse_rope=fuse_rope)
    dec.prefill(ids[:, :T0])
    got = torch.stack([dec.step(ids[:, t:t + 1]).float().clone() for t in range(T0, T0 + steps)])
    kv, kv_e = cache.layers[0].kv, eager_cache.layers[0].kv
    assert dec.captures == 2  # once after the prefill, once after the compression at T = 544
    assert dec.rope_fused == fuse_rope  # the model file's apply_rotary_pos_emb was deferred to the attention launch (or not)
    import sys
    assert "deferred" not in sys.modules[type(model).__module__].apply_rotary_pos_emb.__name__  # and restored afterwards
    assert (kv.comp_len, kv.win_len) == (kv_e.comp_len, kv_e.win_len) == (512, T0 + steps - 512)
    assert cache.get_seq_length() == T0 + steps
    # the caches must hold the same bytes: same appended rows, same compression
    for l, le in zip(cache.layers, eager_cache.layers):
        assert torch.equal(l.kv.k_win[:, :kv.win_len], le.kv.k_win[:, :kv.win_len])
        assert torch.equal(l.kv.v.bmp[:, :1024], le.kv.v.bmp[:, :1024])
    d = (got - ref).abs()
    scale = ref.abs().max().item()
    # same kernels, another split plan (planned for the window's capacity): fp32 merge-order noise only
    assert d.max().item() <= 5e-3 * scale, (d.max().item(), scale)
    # the chosen tokens were recorded on the device: slot t holds argmax of step t's logits
    assert torch.equal(dec.tokens[:, 1:steps + 1], got.argmax(-1).T)


@pytest.mark.gpu
def test_graphed_decoder_generate_matches_hf_generate():
    import mustafar_b200.hf as mhf
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=512, hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, head_dim=128, max_position_embeddings=2048)
    torch.manual_seed(3)
    model = LlamaForCausalLM(cfg).half().cuda().eval()
    model.config._attn_implementation = mhf.ATTN_NAME
    ids = torch.randint(1, cfg.vocab_size, (2, 290), generator=torch.Generator().manual_seed(4)).cuda()
    new = 40
    with torch.no_grad():
        want = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=new, do_sample=False,
                              past_key_values=mhf.MustafarCache(cfg, 0.5, 0.5, max_tokens=400), eos_token_id=None, pad_token_id=0)
    got = mhf.GraphedDecoder(model, mhf.MustafarCache(cfg, 0.5, 0.5, max_tokens=400), max_new_tokens=new).generate(ids, new)
    assert got.shape == want.shape and torch.equal(got[:, :291], want[:, :291])  # the prefill's token is bit-identical
    # a random-init model's logits are nearly flat, so one flipped near-tie changes the continuation: require a long common prefix
    same = (got == want).all(0).float()
    first_diff = int(same.argmin().item()) if same.min().item() == 0 else got.shape[1]
    assert first_diff >= 290 + 8, first_diff
    with pytest.raises(ValueError):
        mhf.GraphedDecoder(model, mhf.MustafarCache(cfg, 0.5, 0.5, max_tokens=400), max_new_tokens=4).generate(ids, 5)
    small = mhf.GraphedDecoder(model, mhf.MustafarCache(cfg, 0.5, 0.5, max_tokens=400), max_new_tokens=3)
    with pytest.raises(ValueError):
        small.step()  # no prompt yet
    small.generate(ids, 3)
    with pytest.raises(ValueError):
        small.step()  # the token buffer is full: refused on the host, not a device-side index error


def test_deferred_rope_patch_is_scoped_to_the_capture_and_the_thread():
    """`_DeferredRope` (CPU logic only): inside the context the model file's apply_rotary_pos_emb hands 1-token fp16 q / k through
    and leaves (cos, sin) with the step; anything else - several tokens, another dtype, another thread - gets the stock function;
    the module global is restored on exit."""
    import threading
    import types
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.models.llama import modeling_llama
    import mustafar_b200.hf as mhf
    cfg = LlamaConfig(vocab_size=64, hidden_size=128, intermediate_size=64, num_hidden_layers=1, num_attention_heads=1,
                      num_key_value_heads=1, head_dim=128)
    model = LlamaForCausalLM(cfg)
    stock = modeling_llama.apply_rotary_pos_emb
    step = types.SimpleNamespace(rope=None)
    q, k = torch.randn(2, 1, 1, 128).half(), torch.randn(2, 1, 1, 128).half()
    ang = torch.rand(2, 1, 64)
    cos, sin = torch.cat([ang, ang], -1).cos().half(), torch.cat([ang, ang], -1).sin().half()
    want_q, want_k = stock(q.float(), k.float(), cos.float(), sin.float())
    with mhf._DeferredRope(model, step, enabled=True):
        assert modeling_llama.apply_rotary_pos_emb is not stock
        got_q, got_k = modeling_llama.apply_rotary_pos_emb(q, k, cos, sin)
        assert got_q is q and got_k is k and step.rope[0] is cos and step.rope[1] is sin
        # two tokens: not a decode step -> stock path, and the stale (cos, sin) are dropped
        q2, k2 = torch.randn(2, 1, 2, 128).half(), torch.randn(2, 1, 2, 128).half()
        cos2, sin2 = cos.expand(2, 2, 128).contiguous(), sin.expand(2, 2, 128).contiguous()
        r_q, _ = modeling_llama.apply_rotary_pos_emb(q2, k2, cos2, sin2)
        assert r_q is not q2 and step.rope is None
        # fp32 tensors: stock path
        f_q, _ = modeling_llama.apply_rotary_pos_emb(q.float(), k.float(), cos.float(), sin.float())
        assert torch.equal(f_q, want_q) and step.rope is None
        # another thread keeps the stock behaviour
        box = {}
        t = threading.Thread(target=lambda: box.update(r=modeling_llama.apply_rotary_pos_emb(q.float(), k.float(), cos.float(), sin.float())))
        t.start(); t.join()
        assert torch.equal(box["r"][0], want_q) and torch.equal(box["r"][1], want_k)
    assert modeling_llama.apply_rotary_pos_emb is stock
    with mhf._DeferredRope(model, step, enabled=False):  # fuse_rope=False, or no static step: nothing is patched
        assert modeling_llama.apply_rotary_pos_emb is stock
    with mhf._DeferredRope(model, None, enabled=True):
        assert modeling_llama.apply_rotary_pos_emb is stock
