"""Generate golden vectors from the REFERENCE ITSELF (run in the build container only).

What runs:
  * /root/reference/kernel/compression.py — the reference Triton kernels, executed on CPU with
    ``TRITON_INTERPRET=1``.  The module source is read from /root/reference at run time and only
    its two device-placement lines are patched in memory (``assert inputs.is_cuda`` and
    ``device='cuda'``); nothing is copied into this repository.
  * /root/reference/models/llama_mustafar_kernel.py:77-113 — ``dh_prune_key`` is extracted from
    the file with ``ast`` (the file itself does not import under transformers 5.x) and executed
    verbatim as a plain function.

What is written: tests/golden/compress_*.npz, tests/golden/prune_*.npz (small, committed).
/root/reference does not exist on the GPU box; tests only read the committed .npz files.

Usage:  TRITON_INTERPRET=1 python oracle/make_golden.py
"""
import ast
import os
import sys
import types

os.environ.setdefault("TRITON_INTERPRET", "1")

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference_compression():
    src = open(os.path.join(REF, "kernel", "compression.py")).read()
    assert src.count("assert inputs.is_cuda") == 2 and src.count("device='cuda'") == 2
    src = src.replace("assert inputs.is_cuda", "pass").replace("device='cuda'", "device=inputs.device")
    mod = types.ModuleType("ref_compression")
    mod.__file__ = os.path.join(REF, "kernel", "compression.py")
    # triton's jit needs source lookup through inspect -> register in linecache
    import linecache
    fname = "<ref_compression_patched>"
    linecache.cache[fname] = (len(src), None, src.splitlines(True), fname)
    exec(compile(src, fname, "exec"), mod.__dict__)
    return mod


def load_reference_prune():
    src = open(os.path.join(REF, "models", "llama_mustafar_kernel.py")).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "dh_prune_key":
            fn = node
            break
    assert fn is not None
    modast = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch}
    exec(compile(modast, "<ref_dh_prune_key>", "exec"), ns)
    f = ns["dh_prune_key"]
    return lambda x, s: f(None, x, target_sparsity=s)


def make_inputs(seed, bk, m, kind):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(bk, m, 128, generator=g, dtype=torch.float32)
    if kind == "randn":
        pass
    elif kind == "ties":  # coarse grid → many equal magnitudes, exercises keep-all-ties
        x = torch.round(x * 4) / 4
    elif kind == "zeros":  # rows that are entirely zero and rows with few nonzeros
        x[:, ::3, :] = 0
        x[:, 1::7, 5:] = 0
    elif kind == "negzero":
        x = torch.round(x * 2) / 2
        x[x == 0] = -0.0
    elif kind == "dense":  # nothing pruned → every tile full (64 nonzeros, no padding)
        x = x.abs() + 0.5
    else:
        raise ValueError(kind)
    return x.to(torch.float16)


CASES = [
    # name, seed, Bk, M, kind, sparsity (None = compress the raw input without pruning)
    ("randn_s50", 42, 3, 128, "randn", 0.5),
    ("randn_s70", 43, 2, 128, "randn", 0.7),
    ("ties_s50", 44, 2, 64, "ties", 0.5),
    ("ties_s70", 45, 2, 64, "ties", 0.7),
    ("zeros_s50", 46, 2, 64, "zeros", 0.5),
    ("negzero_s70", 47, 1, 64, "negzero", 0.7),
    ("dense_noprune", 48, 1, 64, "dense", None),
    ("randn_s00", 49, 1, 64, "randn", 0.0),
    ("randn_s99", 50, 1, 64, "randn", 0.99),
    ("append256_s50", 51, 2, 256, "randn", 0.5),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    comp = load_reference_compression()
    prune = load_reference_prune()
    for name, seed, bk, m, kind, s in CASES:
        x = make_inputs(seed, bk, m, kind)
        if s is None:
            xp = x.clone()
        else:
            xp = prune(x.view(1, bk, m, 128), s).view(bk, m, 128).contiguous()
        out = {"x": x.numpy(), "pruned": xp.numpy(), "sparsity": np.float64(-1.0 if s is None else s)}
        for tag, fn in (("k", comp.convert_key_batched), ("v", comp.convert_value_batched)):
            bmp, acc, packed = fn(xp)
            out[f"{tag}_bitmaps"] = bmp.numpy()
            out[f"{tag}_accum"] = acc.numpy()
            out[f"{tag}_packed"] = np.concatenate([p.numpy() for p in packed]) if packed else np.zeros(0, np.float16)
            out[f"{tag}_packed_len"] = np.array([p.numel() for p in packed], dtype=np.int64)
        path = os.path.join(OUT, f"compress_{name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})

    # prune-only vectors over many rows (cheap): different sparsities and tie-heavy inputs
    for name, seed, kind, s in [("randn_s50", 7, "randn", 0.5), ("randn_s70", 8, "randn", 0.7),
                                ("ties_s50", 9, "ties", 0.5), ("ties_s70", 10, "ties", 0.7),
                                ("zeros_s70", 11, "zeros", 0.7), ("randn_s30", 12, "randn", 0.3),
                                ("randn_s001", 13, "randn", 0.001)]:
        x = make_inputs(seed, 4, 256, kind)
        y = prune(x.view(1, 4, 256, 128), s).view(4, 256, 128)
        path = os.path.join(OUT, f"prune_{name}.npz")
        np.savez_compressed(path, x=x.numpy(), y=y.numpy(), sparsity=np.float64(s))
        print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
