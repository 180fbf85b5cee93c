"""Model-level decode benchmark (BASELINE.json configs[1]): random-init Llama-2-7B geometry, 4096-token prompt + 1024
generated tokens, K/V sparsity 0.5/0.5, on one B200 — the reference's own harness method (mem_spd_test.py:81-96:
wall-clock around `model.generate`, then `torch.cuda.max_memory_allocated`), on the STOCK transformers model class.

    python tools/model_bench.py [--batch 1] [--prompt 4096] [--new 1024] [--arms mustafar,sdpa,flash,ref_cuda] [--json out.json]

Arms (same weights, same prompt ids, greedy decoding, eos disabled):
  mustafar  attn_implementation="mustafar" + MustafarCache (this repo: fused append + sparse attention launch per layer)
  mustafar_graph  the same model and cache through `mustafar_b200.hf.GraphedDecoder`: the whole decode step is ONE CUDA-graph replay
  sdpa_graph      dense StaticCache + torch SDPA with the decode step captured the same way (the strongest dense baseline here:
                  attention always spans the cache's full capacity, prompt + new + 64 tokens, behind a mask)
  sdpa      dense DynamicCache + torch SDPA
  flash     dense DynamicCache + flash_attention_2 (if transformers accepts the installed flash_attn)
  ref_cuda  the reference's decode glue (llama_mustafar_kernel.py:268-320) on the reference's own CUDA kernels built for
            sm_100a (oracle/_ref), over the same compressed cache container  [bench-only baseline; needs oracle/_ref]
Per arm: prefill seconds, decode tok/s (generate(new) minus generate(1), per sequence and aggregate), peak allocated GB,
KV bytes held.  HF's eager decode step is dominated by Python/launch overhead at batch 1 (about 100 launches per layer),
so the attention kernel's share is small there; the batch sweep shows where it starts to matter.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def build_model(layers, hidden, heads, kv_heads, inter, vocab, max_pos):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                      num_attention_heads=heads, num_key_value_heads=kv_heads, head_dim=128, max_position_embeddings=max_pos,
                      rms_norm_eps=1e-5, tie_word_embeddings=False)
    torch.manual_seed(0)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float16)
    try:
        with torch.device("cuda"):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    return model.eval()


def register_ref_cuda():
    """Attention function "mustafar_ref": same cache container, decode through the reference's kernels + glue."""
    from transformers import AttentionInterface, AttentionMaskInterface
    from transformers.masking_utils import sdpa_mask
    import mustafar_b200.hf as mhf
    from oracle import ref_cuda

    def fwd(module, query, key, value, attention_mask, scaling=None, dropout=0.0, **kw):
        cache = mhf._ACTIVE.cache
        layer = cache.layers[module.layer_idx]
        if query.shape[2] > 1:
            return mhf.mustafar_attention_forward(module, query, key, value, attention_mask, scaling, dropout, **kw)
        kv = layer.kv
        kv.append(key, value)  # the reference's torch.cat of the new row (:270, :309)
        memo = getattr(layer, "_ref_memo", None)
        if memo is None or memo[0] != kv.comp_len:  # the reference's per-head containers, rebuilt at compression events only
            kc, _, vc, _, _, _ = kv.as_reference_tuple()
            memo = layer._ref_memo = (kv.comp_len, kc, vc)
        kw_ = kv.k_win[:, : kv.win_len].reshape(kv.batch, kv.kv_heads, kv.win_len, 128)
        vw = kv.v_win[:, : kv.win_len].reshape(kv.batch, kv.kv_heads, kv.win_len, 128)
        out = ref_cuda.decode_step(query, memo[1], kw_, memo[2], vw, kv.comp_len, kv.groups)  # incl. its per-step torch.cat of NZ
        kv.maybe_compress()
        return out.transpose(1, 2), None

    AttentionInterface.register("mustafar_ref", fwd)
    AttentionMaskInterface.register("mustafar_ref", sdpa_mask)


class DenseGraphDecoder:
    """Greedy decode of the stock model over a dense StaticCache with the decode step captured in a CUDA graph."""

    def __init__(self, model, cache, capacity):
        self.model, self.cache, self.capacity = model, cache, capacity
        self.graph = None

    @torch.no_grad()
    def _body(self):
        out = self.model(input_ids=self.ids, position_ids=self.pos, past_key_values=self.cache, use_cache=True)
        nxt = out.logits[:, -1].argmax(-1, keepdim=True)
        self.tokens.scatter_(1, self.slot, nxt)
        self.ids.copy_(nxt)
        self.pos.add_(1)
        self.slot.add_(1)

    @torch.no_grad()
    def generate(self, ids, new):
        b, t = ids.shape
        dev = ids.device
        out = self.model(input_ids=ids, past_key_values=self.cache, use_cache=True, logits_to_keep=1)
        first = out.logits[:, -1].argmax(-1, keepdim=True)
        self.ids, self.pos = first.clone(), torch.full((b, 1), t, dtype=torch.long, device=dev)
        self.slot = torch.ones((b, 1), dtype=torch.long, device=dev)
        self.tokens = torch.zeros((b, self.capacity), dtype=torch.long, device=dev)
        self.tokens[:, :1] = first
        if new > 1:
            # warm-up step on a side stream (it really decodes one token), then capture and replay the rest
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if new > 2:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._body()
                # the capture ran `StaticLayer.update` on the host once without executing: nothing device-side moved
                for _ in range(new - 2):
                    g.replay()
                self.graph = g
        return torch.cat([ids, self.tokens[:, :new]], 1)


def run_arm(model, arm, ids, new, k_sparsity, v_sparsity):
    import mustafar_b200.hf as mhf
    from transformers import DynamicCache
    b, t0 = ids.shape
    impl = {"mustafar": "mustafar", "mustafar_graph": "mustafar", "sdpa": "sdpa", "sdpa_graph": "sdpa", "flash": "flash_attention_2",
            "ref_cuda": "mustafar_ref"}[arm]
    model.config._attn_implementation = impl

    def make_cache():
        if arm in ("mustafar", "mustafar_graph", "ref_cuda"):
            return mhf.MustafarCache(model.config, k_sparsity, v_sparsity, max_tokens=t0 + new + 64)
        if arm == "sdpa_graph":
            from transformers import StaticCache
            return StaticCache(config=model.config, max_cache_len=t0 + new + 64)
        return DynamicCache(config=model.config)

    def gen(n):
        cache = make_cache()
        torch.cuda.synchronize()
        t = time.time()
        with torch.no_grad():
            if arm == "mustafar_graph":
                dec = mhf.GraphedDecoder(model, cache, max_new_tokens=max(n, 2))
                out = dec.generate(ids, n)
                cache._captures = dec.captures
            elif arm == "sdpa_graph":
                out = DenseGraphDecoder(model, cache, max(n, 2)).generate(ids, n)
            else:
                out = model.generate(input_ids=ids, attention_mask=torch.ones_like(ids), max_new_tokens=n, do_sample=False,
                                     past_key_values=cache, eos_token_id=None, pad_token_id=0)
        torch.cuda.synchronize()
        return time.time() - t, out, cache

    gen(2)  # warm-up (lazy allocations, autotuning)
    mhf._ACTIVE.cache = None  # the adapter's thread-local keeps the last cache alive: drop it before measuring memory
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    t1, _, _ = gen(1)
    tn, out, cache = gen(new)
    peak = torch.cuda.max_memory_allocated()
    held = cache.bytes_held() if hasattr(cache, "bytes_held") else sum(
        l.keys.numel() * 2 + l.values.numel() * 2 for l in cache.layers if getattr(l, "keys", None) is not None)
    captures = getattr(cache, "_captures", None)
    dec = (tn - t1) / max(new - 1, 1)
    res = {"arm": arm, "prefill_s": round(t1, 4), "total_s": round(tn, 3), "decode_ms_per_token": round(dec * 1e3, 3),
           "decode_tok_s_per_seq": round(1.0 / dec, 2), "decode_tok_s_aggregate": round(b / dec, 2),
           "peak_allocated_GB": round(peak / 2**30, 3), "allocated_before_GB": round(base / 2**30, 3), "kv_bytes_held_GB": round(held / 2**30, 3)}
    if captures is not None:
        res["graph_captures"] = captures  # capture time is inside decode_ms_per_token
    del cache
    mhf._ACTIVE.cache = None
    return res, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt", type=int, default=4096)
    ap.add_argument("--new", type=int, default=1024)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=32)
    ap.add_argument("--sparsity", type=float, default=0.5)
    ap.add_argument("--arms", default="mustafar_graph,mustafar,sdpa_graph,flash,ref_cuda")
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    import mustafar_b200.hf  # noqa: F401  registers "mustafar"
    arms = [x for x in a.arms.split(",") if x]
    if "ref_cuda" in arms:
        try:
            register_ref_cuda()
        except Exception as e:  # oracle/_ref not built
            print(f"ref_cuda arm unavailable: {e}", file=sys.stderr)
            arms.remove("ref_cuda")
    model = build_model(a.layers, 4096, 32, a.kv_heads, 11008 if a.kv_heads == 32 else 14336, 32000, a.prompt + a.new + 64)
    ids = torch.randint(1, 32000, (a.batch, a.prompt), generator=torch.Generator().manual_seed(1)).cuda()
    out = {"workload": f"configs[1]: random-init Llama geometry ({a.layers} layers, 32 q heads, {a.kv_heads} KV heads, hidden 4096), batch {a.batch}, "
                       f"{a.prompt} prompt + {a.new} generated, K/V sparsity {a.sparsity}/{a.sparsity}, HF generate (greedy), 1 x B200",
           "method": "wall clock around model.generate (mem_spd_test.py:81-96); decode = generate(new) - generate(1)", "arms": []}
    ref_tokens = None
    for arm in arms:
        try:
            res, toks = run_arm(model, arm, ids, a.new, a.sparsity, a.sparsity)
        except Exception as e:
            res, toks = {"arm": arm, "unavailable": f"{type(e).__name__}: {e}"[:200]}, None
        if toks is not None and arm in ("mustafar", "mustafar_graph", "ref_cuda"):
            if ref_tokens is None:
                ref_tokens = toks
            else:  # same cache contents, two attention implementations; a random-init model's near-flat logits flip on 1e-4
                gen_a, gen_b = toks[:, a.prompt:], ref_tokens[:, a.prompt:]
                res["generated_tokens_equal_until"] = int((gen_a != gen_b).any(0).float().argmax().item()) if (gen_a != gen_b).any() else int(gen_a.shape[1])
        out["arms"].append(res)
        print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()
    if a.json:
        with open(a.json, "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
