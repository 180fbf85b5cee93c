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
