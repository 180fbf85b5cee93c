"""Per-CTA timeline of the fused decode-attention kernel (debug build `make -C mustafar_b200/csrc trace`).

    MFB200_LIB=$PWD/mustafar_b200/libmustafar_b200_trace.so python tools/trace_attn.py [cfg]

Launches the 32-layer back-to-back sequence the bench times (PDL + early KV prefetch), then prints, for the
last two launches, when each phase of the compressed CTAs happened relative to the first CTA's start.
Slots: 0 entry, 1 idx+barriers ready, 2 PDL wait passed, 3 q staged, 4/5 K warp first/last block done,
6/7 V warp first/last block done, 8 roles joined, 9 partial published / ticket taken, 10 exit, 11 smid, 12 blocks, 13 merged.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from mustafar_b200 import _lib
from mustafar_b200.attention import MustafarKVCache
from tools.prof_attn import CFG


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    c = CFG[name]
    nl = 32 if name == "cfg1" else 4
    torch.manual_seed(0)
    caches = []
    for _ in range(nl):
        k = torch.randn(c["b"], c["hkv"], c["t"], 128, device="cuda", dtype=torch.float16)
        v = torch.randn(c["b"], c["hkv"], c["t"], 128, device="cuda", dtype=torch.float16)
        cache = MustafarKVCache(c["b"], c["hkv"], c["g"], c["t"], c["s"], c["s"])
        cache.prefill(k, v)
        caches.append(cache)
        del k, v
    q = torch.randn(c["b"], c["hkv"] * c["g"], 1, 128, device="cuda", dtype=torch.float16)
    out = torch.empty_like(q)
    lib = _lib.load()
    raw = C.CDLL(_lib.LIB_PATH)
    raw.mfb200_debug_trace.argtypes = [C.c_void_p, C.c_int]
    params = [cc.make_params(q.view(c["b"], -1, 128), out) for cc in caches]
    for p in params:
        p.flags |= _lib.F_PDL | _lib.F_PDL_EARLY_KV
    sp = _lib.stream_ptr()
    for _ in range(3):
        for p in params:
            lib.mfb200_sparse_decode_attention(C.byref(p), sp)
    torch.cuda.synchronize()
    # flag bit 0x100 picks the half of the trace buffer: the last launch writes half 1, all others half 0
    params[-1].flags |= 0x100
    for p in params:
        lib.mfb200_sparse_decode_attention(C.byref(p), sp)
    torch.cuda.synchronize()
    n = 4096
    buf = np.zeros((2, n, 32), dtype=np.uint64)
    slots = raw.mfb200_debug_trace(buf.ctypes.data, n)
    assert slots == 32
    names = ["entry", "idx+bars", "pdl wait", "q staged", "K first", "K last", "V first", "V last", "joined", "published", "exit"]
    t_first = None
    for half in range(2):
        t = buf[half].astype(np.int64)
        live = t[:, 0] > 0
        t = t[live]
        comp = t[:, 12] != -1
        tc, tw = t[comp], t[~comp]
        t0 = t[:, 0].min()
        if t_first is None:
            t_first = t0
        print(f"== launch {half}: {len(tc)} compressed CTAs (blocks {tc[:,12].min()}..{tc[:,12].max()}), {len(tw)} window CTAs; "
              f"first entry at {(t0 - t_first)/1e3:.2f} us after launch 0's; last exit {(t[:,10].max() - t0)/1e3:.2f} us after its own first entry")
        for k, nm in enumerate(names):
            d = (tc[:, k] - t0) / 1e3
            print(f"  compressed {nm:9s}: min {d.min():6.2f}  p50 {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us")
        if len(tw):
            for k, nm in ((0, "entry"), (10, "exit")):
                d = (tw[:, k] - t0) / 1e3
                print(f"  window     {nm:9s}: min {d.min():6.2f}  p50 {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us")
        merged = tc[tc[:, 13] == 1]
        if len(merged):
            d = (merged[:, 10] - merged[:, 9]) / 1e3
            print(f"  merge (published -> exit) of the {len(merged)} merging compressed CTAs: p50 {np.median(d):.2f} max {d.max():.2f} us")
        per_blk = (tc[:, 7] - tc[:, 3]) / 1e3 / np.maximum(tc[:, 12], 1)
        print(f"  (V last - q staged) / blocks: p50 {np.median(per_blk):.2f} us/block")
        sm = np.bincount(t[:, 11].astype(int), minlength=148)
        print(f"  CTAs per SM: min {sm.min()} max {sm.max()}")
        # balance: group the compressed CTAs by (own blocks, blocks of all compressed CTAs on the same SM)
        ids = np.nonzero(live)[0][comp]
        smid = tc[:, 11].astype(int)
        sm_blocks = np.bincount(smid, weights=tc[:, 12], minlength=148)
        sm_ctas = np.bincount(smid, minlength=148)
        key = [(int(tc[i, 12]), int(sm_ctas[smid[i]]), int(sm_blocks[smid[i]])) for i in range(len(tc))]
        for k in sorted(set(key)):
            sel = np.array([kk == k for kk in key])
            d = (tc[sel, 7] - t0) / 1e3
            print(f"  own blocks {k[0]}, {k[1]} compressed CTAs / {k[2]} blocks on its SM: {sel.sum():3d} CTAs, V last p50 {np.median(d):6.2f} max {d.max():6.2f} us")
        first_wave = ids < 148
        if first_wave.any():
            print(f"  ids 0..147 landed on {len(set(smid[first_wave]))} distinct SMs; ids 148..295 on {len(set(smid[(ids >= 148) & (ids < 296)]))}")


if __name__ == "__main__":
    main()
