"""Per-role cycle accounting of the tcgen05 GQA kernel (debug build `make -C mustafar_b200/csrc trace`).

    MFB200_LIB=$PWD/mustafar_b200/libmustafar_b200_trace.so python tools/trace_tc.py [cfg3|cfg5s]

Slots 16-31 of the per-CTA trace are SM-clock accumulators of one warp per role (see MFB_TACC in decode_attn.cu):
K decode warp 0 / V decode warp 4: wait TMA, records, wait dense buffer, decode; epilogue warp 8: wait S/O, load+max,
exp+publish; MMA-S: wait Kd, wait S free; MMA-O: wait Vd, wait p, wait O free.
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
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    c = CFG[name]
    torch.manual_seed(0)
    k = torch.randn(c["b"], c["hkv"], c["t"], 128, device="cuda", dtype=torch.float16)
    v = torch.randn(c["b"], c["hkv"], c["t"], 128, device="cuda", dtype=torch.float16)
    cache = MustafarKVCache(c["b"], c["hkv"], c["g"], c["t"], c["s"], c["s"])
    cache.prefill(k, v)
    del k, v
    q = torch.randn(c["b"], c["hkv"] * c["g"], 1, 128, device="cuda", dtype=torch.float16)
    out = torch.empty_like(q)
    lib = _lib.load()
    raw = C.CDLL(_lib.LIB_PATH)
    raw.mfb200_debug_trace.argtypes = [C.c_void_p, C.c_int]
    p = cache.make_params(q.view(c["b"], -1, 128), out)
    sp = _lib.stream_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        lib.mfb200_sparse_decode_attention(C.byref(p), sp)
    torch.cuda.synchronize()
    flush.zero_()
    p.flags |= 0x100
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.mfb200_sparse_decode_attention(C.byref(p), sp)
    e1.record()
    torch.cuda.synchronize()
    n = 4096
    buf = np.zeros((2, n, 32), dtype=np.uint64)
    slots = raw.mfb200_debug_trace(buf.ctypes.data, n)
    assert slots == 32, slots
    t = buf[1].astype(np.int64)
    t = t[t[:, 0] > 0]
    tc = t[t[:, 12] != -1]
    t0 = t[:, 0].min()
    print(f"{name}: {e0.elapsed_time(e1) * 1e3:.1f} us (trace build, cold L2), {len(tc)} compressed CTAs, blocks per CTA "
          f"{tc[:, 12].min()}..{tc[:, 12].max()}, last exit {(t[:, 10].max() - t0) / 1e3:.1f} us")
    nbk = np.maximum(tc[:, 12], 1)
    span = (tc[:, 8] - tc[:, 3]) * 1.965  # ns -> SM clocks at 1965 MHz (approximately)
    print(f"  loop span per block (q staged -> roles joined): p50 {np.median(span / nbk):.0f} clk")
    if os.environ.get("MFB_TRACE_TC"):  # tcgen05 variant (MFB_GQA_TC=1 builds)
        names = {16: "K wait TMA", 17: "K records", 18: "K wait dense", 19: "K decode", 20: "V wait TMA", 21: "V records", 22: "V wait dense",
                 23: "V decode", 24: "epi wait S/O", 25: "epi load+max", 26: "epi exp+publish", 27: "MMA-S wait Kd", 28: "MMA-S wait S free",
                 29: "MMA-O wait Vd", 30: "MMA-O wait p", 31: "MMA-O wait O free"}
    else:
        names = {16: "K wait TMA", 17: "K records", 18: "K wait score buf", 19: "K tiles+handoff", 20: "V wait TMA", 21: "V records",
                 22: "V wait p", 23: "V tiles+handoff"}
    for k_, nm in names.items():
        d = tc[:, k_] / nbk
        print(f"  {nm:18s}: p10 {np.percentile(d, 10):7.0f}  p50 {np.median(d):7.0f}  p90 {np.percentile(d, 90):7.0f} clk/block")


if __name__ == "__main__":
    main()
