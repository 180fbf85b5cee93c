"""Compiles the REFERENCE's own Triton compression kernels (kernel/compression.py:9-247, read where they lie under
/root/reference — never copied) into sm_100 cubins under oracle/_ref/triton/.  TEST INFRASTRUCTURE ONLY.

/root/reference does not exist on the GPU box, and Triton specialises these kernels on their tl.constexpr shape
arguments (total_elems, stride_batch, M, N), so the exact shape instances the GPU parity tests use are
cross-compiled here (no GPU needed: triton.compile with an explicit GPUTarget) and travel as binaries, like
oracle/_ref/libmustafar_ref.so.  `oracle/ref_triton.py` launches them through the CUDA driver API.

    python oracle/build_ref_triton.py            (also run by oracle/Makefile and __graft_entry__.build())
"""
import importlib.util
import json
import os
import sys

REF = "/root/reference/kernel/compression.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "triton")
# (B, M) instances of [B, M, 128]: SURVEY.md §7 minimum slice + one small shape
SHAPES = [(2, 256), (32, 3840), (128, 7936)]
N = 128

SIGS = {
    "calculate_bitmap_key_batched": {"input_ptr": "*fp16", "bitmaps_ptr": "*i64", "counts_ptr": "*i32", "total_elems": "constexpr",
                                     "shifts_ptr": "*i64", "stride_batch": "constexpr", "M": "constexpr", "N": "constexpr"},
    "compress_key_batched": {"input_ptr": "*fp16", "bitmaps_ptr": "*i64", "counts_ptr": "*i32", "packed_not_ptr": "*fp16",
                             "batch_offsets_ptr": "*i64", "total_elems": "constexpr", "stride_batch": "constexpr",
                             "M": "constexpr", "N": "constexpr"},
}
SIGS["calculate_bitmap_value_batched"] = SIGS["calculate_bitmap_key_batched"]
SIGS["compress_value_batched"] = SIGS["compress_key_batched"]


def main():
    if not os.path.exists(REF):
        print(f"{REF} not present: nothing to build (the prebuilt cubins travel with the repo snapshot)")
        return 0
    import triton
    from triton.backends.compiler import GPUTarget
    from triton.compiler import ASTSource
    spec = importlib.util.spec_from_file_location("ref_compression", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    os.makedirs(OUT, exist_ok=True)
    manifest = {"triton": triton.__version__, "source": REF, "kernels": {}}
    for b, m in SHAPES:
        consts = {"total_elems": b * m * N, "stride_batch": (m * N // 64) * 64, "M": m, "N": N}
        for name, sig in SIGS.items():
            k = triton.compile(ASTSource(fn=getattr(mod, name), signature=sig, constexprs=consts), target=GPUTarget("cuda", 100, 32))
            key = f"{name}_B{b}_M{m}"
            with open(os.path.join(OUT, key + ".cubin"), "wb") as f:
                f.write(k.asm["cubin"])
            md = k.metadata
            manifest["kernels"][key] = {"entry": md.name, "num_warps": md.num_warps, "shared": md.shared,
                                        "n_ptr_args": sum(1 for v in sig.values() if v.startswith("*")),
                                        "scratch_args": 2, "global_scratch_size": md.global_scratch_size}
            print("built", key, len(k.asm["cubin"]), "bytes")
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
