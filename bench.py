#!/usr/bin/env python
"""bench.py — sparse-KV decode attention (the Mustafar hot path) on B200.

Headline workload = BASELINE.json configs[4], the largest configuration that fits one B200 with ALL its layers
resident: Mistral-7B KV geometry (32 layers x 8 KV heads x 128, 4 query heads per KV head), batch 32, 32K-token
context, K/V sparsity 0.5/0.5, "streaming compress+attend per step".  One step = one decode step of the attention
path of all 32 layers for all 32 sequences: the new token's K/V rows are appended to the dense residual window by the
fused attention launch itself, attention runs over [compressed | window], and every 256 tokens the window's first
256 rows are pruned + compressed in place (models/llama_mustafar_kernel.py:256-398).  ~86 GB of compressed KV are
read per step (>> the 126 MB L2: every byte comes from HBM).  The layer loop is ONE FFI call per step
(mfb200_decode_step_layers).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]

N > 1 (torchrun, one process per GPU): STRONG scaling — the 32 sequences are batch-partitioned over the ranks
(mustafar_b200.partition.make_partition), no collective on the attention path; value = 32 tokens per step / the
max-over-ranks device time.  Extra keys (not the headline): BASELINE configs[3]'s per-layer attention with its
2048 (sequence, head) units partitioned over the N ranks, configs[0]/[1]'s batch-1 layer (in a PDL chain and
isolated; head-partitioned + NCCL all-gather when N > 1), configs[2]'s GQA layer, and on one GPU the same-box
comparison lines: dense FlashAttention decode and the reference's own CUDA kernels (oracle/_ref, sm_100a build).
--impl reference: the reference's masked-dense PyTorch attention
(models/llama_mustafar_Kt_Mag_Vt_Mag.py:873-874, :963, :974) on the host cores, rank 0 only, on a bounded sample
of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# headline workload (BASELINE.json configs[4])
LAYERS, BATCH, KV_HEADS, GROUPS, CTX, SPARSITY, RESIDUAL = 32, 32, 8, 4, 32768, 0.5, 32
METRIC = "sparse-KV decode attention: end-to-end decode tok/s over all layers (us/step and HBM GB/s in roofline)"
WORKLOAD = ("configs[4]: Mistral-7B KV geometry (32 layers x 8 KV heads x 128, G=4), batch 32 x 32K context, K/V sparsity "
            "0.5/0.5, streaming compress+attend per step, attention path of all 32 layers, bitmap+packed-nonzero cache, "
            "residual window 32..288")


def config_dict(layers):
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "global_batch": BATCH, "layers": layers, "context": CTX, "kv_heads": KV_HEADS,
            "q_heads_per_kv_head": GROUPS, "sparsity": SPARSITY,
            "l2": "inputs larger than L2 (about 2.7 GB of compressed KV per layer, 86 GB per step), no flush needed",
            "partition": "32 sequences batch-partitioned over the ranks (strong scaling), no collective on the attention path"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only (skip the other BASELINE configs and comparison lines)")
    ap.add_argument("--layers", type=int, default=LAYERS, help=argparse.SUPPRESS)
    return ap.parse_args()


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions (B200_PROFILING.md).  NVML in a thread;
    `nvidia-smi -lms` as a fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.sm, self.mask, self.max_mhz, self._stop, self.active = [], 0, None, False, False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.index)
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            return pynvml, pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            if self.active:  # only while a timed region runs
                try:
                    self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                    self.mask |= int(get_reasons(self.handle))
                except Exception:
                    pass
            time.sleep(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=1.0)
            sm = sorted(self.sm)
            reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms period, inside the timed regions only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------------------------ CPU reference arm
CPU_SAMPLE_LAYERS, CPU_SAMPLE_SEQS = 1, 2  # a sample-step = this slice of the full step (1/512 of its work)


def cpu_masked_dense(steps, warmup, distinct=2):
    """The reference's masked-dense decode attention (llama_mustafar_Kt_Mag_Vt_Mag.py:873-874, :963, :974, with its
    repeat_kv materialisation) on the host cores, on a bounded sample of the headline workload: CPU_SAMPLE_LAYERS
    layer(s) x CPU_SAMPLE_SEQS of the 32 sequences per sample-step, `distinct` different caches cycled (each 2 x 134 MB,
    larger than the host caches).  Returns (seconds per sample-step, cores)."""
    import math
    import torch
    from oracle import torch_oracle as TO
    L = ((CTX - RESIDUAL) // 256) * 256
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    caches = []
    g = torch.Generator().manual_seed(42)
    for i in range(distinct):
        k = torch.randn(CPU_SAMPLE_SEQS, KV_HEADS, CTX, 128, generator=g).to(torch.float16)
        v = torch.randn(CPU_SAMPLE_SEQS, KV_HEADS, CTX, 128, generator=g).to(torch.float16)
        k[:, :, :L] = TO.prune_rows(k[:, :, :L], SPARSITY)
        v[:, :, :L] = TO.prune_rows(v[:, :, :L], SPARSITY)
        caches.append((k, v))
    q = torch.randn(CPU_SAMPLE_SEQS, KV_HEADS * GROUPS, 1, 128, generator=g).to(torch.float16)

    def repeat_kv(x):
        b, h, t, d = x.shape
        return x[:, :, None].expand(b, h, GROUPS, t, d).reshape(b, h * GROUPS, t, d)

    def one(k, v):
        w = torch.matmul(q, repeat_kv(k).transpose(2, 3)) / math.sqrt(128)
        p = torch.softmax(w, dim=-1, dtype=torch.float32).to(torch.float16)
        return torch.matmul(p, repeat_kv(v))

    n = [0]

    def step():
        for _ in range(CPU_SAMPLE_LAYERS):
            one(*caches[n[0] % distinct])
            n[0] += 1

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps, cores


def cpu_line_fields(dt, cores, layers, steps):
    scale = (layers / CPU_SAMPLE_LAYERS) * (BATCH / CPU_SAMPLE_SEQS)
    full_step_s = dt * scale
    sample = (f"{steps} sample-steps, each {CPU_SAMPLE_LAYERS} layer x {CPU_SAMPLE_SEQS} of {BATCH} sequences of masked-dense fp16 "
              f"attention (q[{CPU_SAMPLE_SEQS},32,1,128] x K/V[{CPU_SAMPLE_SEQS},8,{CTX},128], repeat_kv as in the reference), torch CPU, "
              f"{cores} threads, {dt * 1e3:.1f} ms each; full step = x{scale:.0f} (all {layers} layers, all {BATCH} sequences)")
    return BATCH / full_step_s, full_step_s, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, max(args.warmup, 3)
    dt, cores = cpu_masked_dense(K, W)
    val, full_step_s, sample = cpu_line_fields(dt, cores, args.layers, K)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "tok/s", "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": full_step_s * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": config_dict(args.layers),
            "cpu_baseline": {"value": val, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from mustafar_b200 import _lib
    from mustafar_b200.attention import HEAD_DIM, MustafarKVCache, decode_step_layers
    from mustafar_b200.partition import gather_heads, make_partition

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    attn = lib.mfb200_sparse_decode_attention
    layers = args.layers
    K, W = args.steps, max(args.warmup, 3)
    part = make_partition(BATCH, KV_HEADS, world, rank)
    assert part.mode == "batch"
    bl = part.local_batch  # sequences held by this rank
    hq = KV_HEADS * GROUPS
    sampler = ClockSampler(local)
    sampler.start()

    def timed(fn, n, sample_clocks=True):
        """n calls of fn(i) between barrier + synchronize on both sides, CUDA events on the launch stream, max over ranks."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler.active = sample_clocks
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        sampler.active = False
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- build the layer caches at a 32K context (prompt compression through the CUDA path) ------------------------
    total_steps = 3 * (K + W) + 16
    torch.manual_seed(42 + rank)
    caches = []
    t_build = time.perf_counter()
    for l in range(layers):
        k = torch.randn(bl, KV_HEADS, CTX, HEAD_DIM, device=dev, dtype=torch.float16)
        v = torch.randn(bl, KV_HEADS, CTX, HEAD_DIM, device=dev, dtype=torch.float16)
        c = MustafarKVCache(bl, KV_HEADS, GROUPS, CTX + total_steps + 8, SPARSITY, SPARSITY, RESIDUAL, device=dev)
        c.prefill(k, v)
        # the prompt's last token plays the role of "the newest token": drop one row so that the first timed step appends
        # to a 255-row window and attends over L = 32512 + Lw = 256 = 32768 tokens
        c.win_len -= 1
        caches.append(c)
    del k, v
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    held = sum(c.bytes_held() for c in caches)
    dense_equiv = sum(c.dense_bytes() + 2 * c.k_win.numel() * 2 for c in caches)

    NSETS = 4
    q_dev = torch.randn(NSETS, layers, bl, hq, HEAD_DIM, device=dev, dtype=torch.float16)
    kv_dev = torch.randn(NSETS, 2, layers, bl, KV_HEADS, HEAD_DIM, device=dev, dtype=torch.float16)
    out_dev = torch.empty(layers, bl, hq, HEAD_DIM, device=dev, dtype=torch.float16)
    q_host = torch.randn(NSETS, layers, bl, hq, HEAD_DIM, dtype=torch.float16).pin_memory()
    kv_host = torch.randn(NSETS, 2, layers, bl, KV_HEADS, HEAD_DIM, dtype=torch.float16).pin_memory()
    out_host = torch.empty(layers, bl, hq, HEAD_DIM, dtype=torch.float16).pin_memory()
    q_stage, kv_stage = torch.empty_like(q_dev[0]), torch.empty_like(kv_dev[0])
    launches = [0]

    def step_device(q, kv):
        before = caches[0].comp_len
        decode_step_layers(caches, q, kv[0], kv[1], out_dev)
        launches[0] += layers  # fused append + attention: sparse_decode_attn_kernel, one per layer
        if caches[0].comp_len != before:
            launches[0] += layers  # compress_append_chunk_kernel: prune + compress 256 window rows of K and V, per layer

    # ---- (1) device-resident throughput ----------------------------------------------------------------------------
    for i in range(W):
        step_device(q_dev[i % NSETS], kv_dev[i % NSETS])
    launches[0] = 0
    ms_dev = timed(lambda i: step_device(q_dev[i % NSETS], kv_dev[i % NSETS]), K)
    gpu_launches = launches[0]

    # ---- (2) the dominant kernel alone: one attend per layer cache at the current state ------------------------------
    algo_bytes = sum(c.compressed_bytes() for c in caches) / layers
    sp = _lib.stream_ptr()

    def layer_params(cs, qs, outs, pdl=True):
        ps = []
        for c, q, o in zip(cs, qs, outs):
            pp = _lib.DecodeParams()
            C.memmove(C.byref(pp), C.byref(c.make_params(q, o)), C.sizeof(pp))
            pp.flags = (pp.flags & ~(_lib.F_PDL | _lib.F_PDL_EARLY_KV)) | ((_lib.F_PDL | _lib.F_PDL_EARLY_KV) if pdl else 0)
            ps.append(pp)
        return ps

    def launch_all(ps):
        for p in ps:
            attn(C.byref(p), sp)

    params = layer_params(caches, [q_dev[0, l] for l in range(layers)], [out_dev[l] for l in range(layers)])
    launch_all(params)
    ms_k = timed(lambda i: launch_all(params), K)
    us_per_launch = ms_k * 1e3 / (K * layers)

    # ---- (3) end to end through the public API with HOST buffers --------------------------------------------------------
    def step_e2e(i):
        q_stage.copy_(q_host[i % NSETS], non_blocking=True)
        kv_stage.copy_(kv_host[i % NSETS], non_blocking=True)
        step_device(q_stage, kv_stage)
        out_host.copy_(out_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller needs this step's result before the next token

    for i in range(W):
        step_e2e(i)
    ms_e2e = timed(step_e2e, K)
    h2d = (q_stage.numel() + kv_stage.numel()) * 2
    d2h = out_dev.numel() * 2
    ctx_end = caches[0].kv_seq_len

    extras = {}
    if not args.no_extras:
        del caches, params, q_dev, kv_dev, q_host, kv_host, q_stage, kv_stage, out_dev, out_host
        torch.cuda.empty_cache()
        extras = run_extras(args, world, rank, dev, timed, layer_params, launch_all, K)
    clocks = sampler.stop()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    achieved = algo_bytes / (us_per_launch * 1e-6) / 1e9
    traffic = None
    try:  # DRAM bytes of the same kernel/shape from the committed ncu --set full capture (per launch)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["sparse_decode_attn_kernel<4>@cfg5"]["dram_bytes_per_launch"]
        if world > 1:
            traffic = traffic / world
    except Exception:
        pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        dt, cores = cpu_masked_dense(4, 1)
        val, _, sample = cpu_line_fields(dt, cores, layers, 4)
        cpu = {"value": val, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": BATCH * 1e3 * K / ms_dev, "unit": "tok/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic", "config": config_dict(layers),
        "us_per_layer_step": ms_dev * 1e3 / (K * layers),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "sparse_decode_attn_kernel<4, *> (GQA variant: register-fragment HMMA with a block-diagonal operand)",
                     "us_per_launch": us_per_launch, "algorithmic_bytes_per_launch": algo_bytes,
                     "frac_of_8TBps_spec": achieved / 8000.0,
                     "how": f"{layers} x K back-to-back launches (one per layer cache of this rank, {bl} sequences each), CUDA events "
                            "on the launch stream; algorithmic bytes = bitmaps + padded nonzeros as stored + window + q/out (idx excluded)"},
        "e2e": {"value": BATCH * 1e3 * K / ms_e2e, "unit": "tok/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K},
        "gpu_launches": gpu_launches, "clocks": clocks, "cpu_baseline": cpu,
        "kv_memory": {"held_bytes_per_rank": held, "dense_fp16_equivalent_bytes": dense_equiv, "ratio": held / dense_equiv,
                      "note": "bitmaps + offsets + sparsity-sized nonzero slabs + dense windows of all layers vs a dense fp16 cache of the same capacity"},
        "build_s": t_build, "context_end": ctx_end, "local_batch": bl,
    }
    line.update(extras)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, world, rank, dev, timed, layer_params, launch_all, K):
    """The other BASELINE configurations and the same-box comparison lines (extra keys, never the headline)."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from mustafar_b200 import _lib
    from mustafar_b200.attention import HEAD_DIM, MustafarKVCache
    from mustafar_b200.partition import gather_heads, make_partition

    out = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs") or 6650.0
    attn = _lib.load().mfb200_sparse_decode_attention
    sp = _lib.stream_ptr()

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001 - an extra must never cost the headline
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    def build(b, hkv, g, T, s, n, seed, chunk=None):
        cs = []
        for i in range(n):
            c = MustafarKVCache(b, hkv, g, T + 8, s, s, RESIDUAL, device=dev)
            step = chunk or b
            for b0 in range(0, b, step):
                gen = torch.Generator(device=dev).manual_seed(seed + 100 * i + b0)
                k = torch.randn(min(step, b - b0), hkv, T, HEAD_DIM, device=dev, dtype=torch.float16, generator=gen)
                v = torch.randn(min(step, b - b0), hkv, T, HEAD_DIM, device=dev, dtype=torch.float16, generator=gen)
                c.prefill(k, v, batch_start=b0)
            cs.append(c)
        return cs

    def attend_us(cs, g, pdl, iters):
        b, hkv = cs[0].batch, cs[0].kv_heads
        qs = [torch.randn(b, hkv * g, HEAD_DIM, device=dev, dtype=torch.float16) for _ in cs]
        os_ = [torch.empty_like(q) for q in qs]
        ps = layer_params(cs, qs, os_, pdl=pdl)
        launch_all(ps)
        ms = timed(lambda i: launch_all(ps), iters, sample_clocks=False)
        return ms * 1e3 / (iters * len(cs)), sum(c.compressed_bytes() for c in cs) / len(cs)

    # ---- configs[3]: Llama-2-7B heads, batch 64 x 32K, s = 0.7: ONE layer, its 2048 units partitioned over the ranks ----
    def cfg4():
        p4 = make_partition(64, 32, world, rank)
        cs = build(p4.local_batch, p4.local_kv_heads, 1, 32768, 0.7, 1, seed=4000 + rank, chunk=4)
        us, nbytes = attend_us(cs, 1, True, max(4, min(K, 10)))
        tot = torch.tensor([nbytes], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot)
        agg = float(tot.item()) / (us * 1e-6) / 1e9
        return {"workload": "configs[3]: batch 64 x 32 heads x 32K, s=0.7, per-layer attention, 2048 units / N ranks (strong)",
                "units_per_rank": p4.local_batch * p4.local_kv_heads, "us_per_layer": us, "aggregate_GBps": agg,
                "frac_of_measured_hbm_per_gpu": agg / world / peak, "held_bytes_per_rank": cs[0].bytes_held(),
                "dense_fp16_equivalent_bytes_per_rank": cs[0].dense_bytes()}
    guarded("cfg4_layer", cfg4)

    # ---- configs[0]/[1]: batch 1, 32 heads, 4K, s = 0.5, 32 layer caches --------------------------------------------------
    def cfg1():
        if world == 1:
            cs = build(1, 32, 1, 4096, 0.5, 32, seed=1000)
            chain, nbytes = attend_us(cs, 1, True, K)
            # isolated: no PDL, a stream sync between launches (nothing of a predecessor to hide under)
            qs = [torch.randn(1, 32, HEAD_DIM, device=dev, dtype=torch.float16) for _ in cs]
            os_ = [torch.empty_like(q) for q in qs]
            ps = layer_params(cs, qs, os_, pdl=False)
            ev = []
            for rep in range(3):
                for p in ps:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    attn(C.byref(p), sp)
                    e1.record()
                    if rep:
                        ev.append((e0, e1))
            torch.cuda.synchronize()
            iso = sorted(a.elapsed_time(b) * 1e3 for a, b in ev)
            iso = iso[len(iso) // 2]
            # whole decode steps (fused append + attention of all 32 layers, the window growing): one FFI call per step vs
            # ONE CUDA-graph launch per step with the window lengths in device memory (attention.DecodeStepGraph)
            from mustafar_b200.attention import DecodeStepGraph, decode_step_layers
            qb = torch.randn(32, 1, 32, HEAD_DIM, device=dev, dtype=torch.float16)
            kb = torch.randn(32, 1, 32, HEAD_DIM, device=dev, dtype=torch.float16)
            ob = torch.empty_like(qb)
            n_steps = 40
            decode_step_layers(cs, qb, kb, kb, ob)
            ms_ffi = timed(lambda i: decode_step_layers(cs, qb, kb, kb, ob), n_steps, sample_clocks=False)
            graph = DecodeStepGraph(cs, qb, kb, kb, ob)
            graph.step()
            ms_graph = timed(lambda i: graph.step(), n_steps, sample_clocks=False)
            return {"workload": "configs[0]/[1] layer: batch 1 x 32 heads x 4K, s=0.5, 32 distinct layer caches (1.4 GB)",
                    "us_per_layer_step_one_ffi_call_per_step": ms_ffi * 1e3 / (n_steps * 32),
                    "us_per_layer_step_one_cuda_graph_per_step": ms_graph * 1e3 / (n_steps * 32),
                    "us_per_launch_in_pdl_chain": chain, "us_per_launch_isolated_cold": iso,
                    "frac_of_measured_hbm_in_chain": nbytes / (chain * 1e-6) / 1e9 / peak,
                    "frac_of_measured_hbm_isolated": nbytes / (iso * 1e-6) / 1e9 / peak, "algorithmic_bytes_per_launch": nbytes}
        # N > 1: batch 1 < N -> KV heads are partitioned, outputs gathered with ONE NCCL all-gather per layer
        p1 = make_partition(1, 32, world, rank)
        assert p1.mode == "head"
        cs = build(1, p1.local_kv_heads, 1, 4096, 0.5, 32, seed=1000 + rank)
        qs = [torch.randn(1, p1.local_kv_heads, 1, HEAD_DIM, device=dev, dtype=torch.float16) for _ in cs]
        os_ = [torch.empty_like(q) for q in qs]
        ps = layer_params(cs, qs, os_, pdl=True)
        full = []

        def step(_):
            full.clear()
            for p, o in zip(ps, os_):
                attn(C.byref(p), sp)
                full.append(gather_heads(p1, o))

        step(0)
        assert full[-1].shape == (1, 32, 1, HEAD_DIM)
        ms = timed(step, K, sample_clocks=False)
        launch_all(ps)
        ms_nc = timed(lambda i: launch_all(ps), K, sample_clocks=False)
        res = {"workload": f"configs[0]/[1] layer head-partitioned: batch 1, 32 heads / {world} ranks, 4K, s=0.5, 32 layers; the [1, 32, 1, 128] "
                           "output of a layer assembled (a) by ONE NCCL all-gather per layer, (b) by the fused launch's peer-to-peer "
                           "store epilogue into every rank's gathered buffer + one wait launch (no collective)",
               "us_per_layer_step_with_all_gather": ms * 1e3 / (K * 32), "us_per_layer_step_attention_only": ms_nc * 1e3 / (K * 32)}
        try:
            from mustafar_b200.partition import PeerOutput
            po = PeerOutput(p1, 1, 32, 1, dev)
            sid = [0]

            def step_peer(_):
                for p in ps:
                    p.peer = C.addressof(po.block(sid[0]))  # ps are copies of the caches' parameter blocks
                    attn(C.byref(p), sp)
                    po.wait(sid[0])
                    sid[0] += 1

            step_peer(0)
            torch.cuda.synchronize()
            want = gather_heads(p1, os_[-1])
            res["peer_stores_equal_all_gather"] = bool(torch.equal(po.gathered(sid[0] - 1), want))
            ms_p = timed(step_peer, K, sample_clocks=False)
            res["us_per_layer_step_with_peer_stores"] = ms_p * 1e3 / (K * 32)
            res["peer_wait_timed_out"] = po.timed_out()
            for p in ps:
                p.peer = None
            torch.cuda.synchronize()
            po.close()
        except Exception as e:  # noqa: BLE001
            res["peer_stores_error"] = f"{type(e).__name__}: {e}"[:200]
        return res
    guarded("cfg1_layer", cfg1)

    if world > 1 or rank != 0:
        return out

    # ---- configs[2]: Llama-3-8B GQA layer, batch 16 x 8K, s = 0.7 (4 distinct caches: 1 GB > L2) + comparison lines ----------
    def compare(b, hkv, g, T, s, n_caches, label, with_ref):
        cs = build(b, hkv, g, T, s, n_caches, seed=3000 + T)
        us, nbytes = attend_us(cs, g, True, max(5, K))
        res = {"workload": label, "us_per_launch": us, "GBps_on_compressed_bytes": nbytes / (us * 1e-6) / 1e9,
               "frac_of_measured_hbm": nbytes / (us * 1e-6) / 1e9 / peak, "algorithmic_bytes_per_launch": nbytes}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def cold(fn, iters=10):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            return sorted(ts)[len(ts) // 2]

        try:
            from flash_attn import flash_attn_with_kvcache
            gen = torch.Generator(device=dev).manual_seed(1)
            kd = torch.randn(b, T, hkv, HEAD_DIM, device=dev, dtype=torch.float16, generator=gen)
            vd = torch.randn(b, T, hkv, HEAD_DIM, device=dev, dtype=torch.float16, generator=gen)
            qd = torch.randn(b, 1, hkv * g, HEAD_DIM, device=dev, dtype=torch.float16, generator=gen)
            res["flash_attn_dense_us"] = cold(lambda: flash_attn_with_kvcache(qd, kd, vd))
            res["speedup_vs_flash_attn_dense"] = res["flash_attn_dense_us"] / us
            del kd, vd
        except Exception as e:  # noqa: BLE001
            res["flash_attn_dense_us"] = f"unavailable: {e}"[:120]
        if with_ref:
            try:
                from oracle import ref_cuda  # the reference's own CUDA kernels (baseline arm only, like cpu_baseline)
                if ref_cuda.available():
                    import torch.nn.functional as F
                    c = cs[0]
                    kc, kw, vc, vw, L, _ = c.as_reference_tuple()
                    nzk, nzv = ref_cuda.pad_nz(kc[2]), ref_cuda.pad_nz(vc[2])
                    bq = b * hkv * g
                    pq = F.pad(torch.randn(bq, 1, 128, device=dev, dtype=torch.float16), (0, 0, 0, 7)).contiguous()
                    pp = F.pad(torch.softmax(torch.randn(bq, 1, L, device=dev), -1).half(), (0, 0, 0, 7)).contiguous()
                    ik, iv = kc[1].reshape(-1), vc[1].reshape(-1)

                    def ref_kernels():
                        ref_cuda.key_formulation(kc[0], nzk, ik, kc[3], pq, L, 128, bq, g)
                        ref_cuda.value_formulation(vc[0], nzv, iv, vc[3], pp, 128, L, bq, g)

                    res["reference_cuda_kernels_us"] = cold(ref_kernels, iters=5)
                    res["speedup_vs_reference_cuda_kernels"] = res["reference_cuda_kernels_us"] / us
            except Exception as e:  # noqa: BLE001
                res["reference_cuda_kernels_us"] = f"unavailable: {e}"[:120]
        return res

    guarded("cfg3_layer", lambda: compare(16, 8, 4, 8192, 0.7, 4, "configs[2] layer: batch 16 x 8 KV heads (G=4) x 8K, s=0.7", True))

    def gqa_ab():
        """Same-run A/B of the two GQA contractions on the configs[2] shape (own processes: the library is fixed at first load):
        the shipped register-fragment HMMA path vs the tcgen05 / TMEM variant (`make tc`), cold-L2 medians of tools/prof_attn.py."""
        import re
        res = {}
        for name, lib_name in (("shipped_hmma_cold_us", "libmustafar_b200.so"), ("tcgen05_variant_cold_us", "libmustafar_b200_tc.so")):
            lib_path = os.path.join(ROOT, "mustafar_b200", lib_name)
            if not os.path.exists(lib_path):
                res[name] = "not built"
                continue
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_attn.py"), "cfg3", "5"], env=dict(os.environ, MFB200_LIB=lib_path),
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=180)
            m = re.search(r"cold median ([0-9.]+) us", r.stdout)
            res[name] = float(m.group(1)) if m else f"failed: {r.stderr[-120:]}"
        return res

    guarded("cfg3_gqa_ab", gqa_ab)
    guarded("cfg5_layer", lambda: compare(32, 8, 4, 32768, 0.5, 1, "configs[4] layer (the headline's kernel): batch 32 x 8 KV heads (G=4) x 32K, s=0.5", True))
    guarded("cfg1_compare", lambda: compare(1, 32, 1, 4096, 0.5, 8, "configs[0] layer, cold L2 rotation of 8 caches, vs baselines", True))
    # the MHA kernel where it is byte-bound rather than decode-bound: the time per 64-position tile does not depend on the
    # sparsity, so the HBM fraction is highest at s = 0.5 and a long context (configs[3]'s geometry at configs[0]'s sparsity)
    guarded("mha_s05_32k_layer", lambda: compare(8, 32, 1, 32768, 0.5, 1, "MHA layer: batch 8 x 32 heads x 32K, s=0.5 (configs[3] geometry at "
                                                 "configs[0] sparsity: the kernel's best HBM fraction)", True))

    def cfg2_model():
        """BASELINE configs[1] at model level (tools/model_bench.py, own process): stock transformers Llama-2-7B geometry,
        random init, 4096 prompt + 1024 generated.  Arms: this library with the whole decode step replayed as one CUDA graph
        (mustafar_b200.hf.GraphedDecoder, rotary embedding fused into the attention launch), the same through stock
        `generate()`, a dense StaticCache + SDPA step captured the same way, and a dense cache + FlashAttention-2 `generate()`."""
        tmp = os.path.join(ROOT, "gpurun_out", f"model_bench_{os.getpid()}.json")
        os.makedirs(os.path.dirname(tmp), exist_ok=True)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "model_bench.py"), "--arms", "mustafar_graph,sdpa_graph,mustafar,flash",
                            "--json", tmp], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=480)
        if r.returncode != 0:
            raise RuntimeError(r.stderr[-300:])
        res = json.load(open(tmp))
        os.remove(tmp)
        arms = {a["arm"]: a for a in res["arms"]}
        tps = {k: a["decode_tok_s_per_seq"] for k, a in arms.items() if "decode_tok_s_per_seq" in a}
        if "mustafar" in tps and "flash" in tps:
            res["decode_speedup_vs_dense_flash_attention_2"] = tps["mustafar"] / tps["flash"]
        if "mustafar_graph" in tps and "flash" in tps:
            res["graphed_decode_speedup_vs_dense_flash_attention_2"] = tps["mustafar_graph"] / tps["flash"]
        if "mustafar_graph" in tps and "sdpa_graph" in tps:
            res["graphed_decode_speedup_vs_graphed_dense_sdpa"] = tps["mustafar_graph"] / tps["sdpa_graph"]
        return res

    if world == 1 and not os.environ.get("MFB200_BENCH_NO_MODEL"):
        guarded("cfg2_model", cfg2_model)
    return out


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    # Libraries write to file descriptor 1 behind Python's back (NCCL prints "NCCL version ..." there when the first
    # communicator comes up): keep the original stdout for the JSON line alone and send everything else to stderr.
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
