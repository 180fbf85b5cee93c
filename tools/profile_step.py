"""Where one graphed decode step of the config-2 model spends its time: kernel-name table of a few CUDA-graph replays
(torch.profiler / CUPTI).   python tools/profile_step.py [--batch 1] [--prompt 4096] [--layers 32] [--kv-heads 32]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools.model_bench import build_model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--prompt", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    a = ap.parse_args()
    import mustafar_b200.hf as mhf
    model = build_model(a.layers, 4096, 32, a.kv_heads, 11008 if a.kv_heads == 32 else 14336, 32000, a.prompt + 512)
    model.config._attn_implementation = "mustafar"
    ids = torch.randint(1, 32000, (a.batch, a.prompt), generator=torch.Generator().manual_seed(1)).cuda()
    cache = mhf.MustafarCache(model.config, 0.5, 0.5, max_tokens=a.prompt + 512)
    dec = mhf.GraphedDecoder(model, cache, max_new_tokens=256)
    dec.prefill(ids)
    for _ in range(4):
        dec.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        dec.step()
    e1.record()
    torch.cuda.synchronize()
    print(f"replayed step: {e0.elapsed_time(e1) / a.steps:.3f} ms (CUDA events, {a.steps} steps)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            dec.step()
        torch.cuda.synchronize()
    tot = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            t = tot[ev.name[:110]]
            t[0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            t[1] += 1
    allt = sum(v[0] for v in tot.values())
    print(f"kernel time per step: {allt / a.steps / 1e3:.3f} ms")
    for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{us / a.steps:9.1f} us/step  {n / a.steps:6.1f} launches/step  {us / n:8.2f} us each  {name}")


if __name__ == "__main__":
    main()
