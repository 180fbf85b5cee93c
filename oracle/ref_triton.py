"""Runs the REFERENCE's Triton compression kernels on the GPU from the cubins `oracle/build_ref_triton.py` compiled
out of /root/reference/kernel/compression.py.  TEST INFRASTRUCTURE ONLY (tests/ import it; the product never does).

The device code is the reference's own (calculate_bitmap_{key,value}_batched, compress_{key,value}_batched); the
host glue between the two launches — cumsum, cat, zero-initialised packed buffer, per-head slicing — is restated
here from kernel/compression.py:255-335 (keys) and :348-428 (values) with the same torch calls.
"""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np
import torch

DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "triton")
_mods = {}


def available(b: int, m: int) -> bool:
    return os.path.exists(os.path.join(DIR, f"compress_key_batched_B{b}_M{m}.cubin"))


def _drv():
    from cuda.bindings import driver
    return driver


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _function(key: str):
    if key in _mods:
        return _mods[key]
    d = _drv()
    manifest = json.load(open(os.path.join(DIR, "manifest.json")))["kernels"][key]
    data = open(os.path.join(DIR, key + ".cubin"), "rb").read()
    torch.cuda.current_stream()  # make sure torch has created the primary context
    mod = _check(d.cuModuleLoadData(data))
    fn = _check(d.cuModuleGetFunction(mod, manifest["entry"].encode()))
    _mods[key] = (fn, manifest, mod)
    return _mods[key]


def _launch(key: str, grid, ptrs):
    d = _drv()
    fn, man, _ = _function(key)
    assert len(ptrs) == man["n_ptr_args"]
    vals = tuple(int(p) for p in ptrs) + (0,) * man["scratch_args"]  # trailing: Triton's global / profile scratch (unused)
    types = (ctypes.c_void_p,) * len(vals)
    stream = torch.cuda.current_stream().cuda_stream
    res = d.cuLaunchKernel(fn, grid[0], grid[1], 1, 32 * man["num_warps"], 1, 1, man["shared"], stream, (vals, types), 0)
    if int(res[0]) != 0:
        raise RuntimeError(f"cuLaunchKernel({key}) failed: {res[0]}")


def _convert(inputs: torch.Tensor, which: str):
    B, M, N = inputs.shape
    assert inputs.is_cuda and inputs.dtype == torch.float16 and N == 128 and M % 64 == 0
    # compression.py:255 (keys are transposed to [B, N, M]) / :348 (values stay [B, M, N])
    inputs_t = inputs.transpose(1, 2).contiguous() if which == "key" else inputs.contiguous()
    tiles = (M * N) // 64
    bitmaps = torch.empty((B, tiles), dtype=torch.int64, device=inputs.device)
    counts = torch.empty((B, tiles), dtype=torch.int32, device=inputs.device)
    shifts = torch.tensor(np.left_shift(np.int64(1), np.arange(63, -1, -1, dtype=np.int64)), device=inputs.device)  # :265-267
    grid = (tiles, B)                                                                                                 # :270
    sfx = f"_B{B}_M{M}"
    _launch(f"calculate_bitmap_{which}_batched" + sfx, grid, [inputs_t.data_ptr(), bitmaps.data_ptr(), counts.data_ptr(), shifts.data_ptr()])
    accum = torch.cumsum(counts, dim=1).to(torch.int32)                                                              # :294
    accum = torch.cat([torch.zeros((B, 1), dtype=counts.dtype, device=counts.device), accum], dim=1).contiguous()    # :295-298
    total = 2 * accum[:, -1]                                                                                         # :302
    offsets = torch.cumsum(total, dim=0)                                                                             # :303
    batch_offsets = torch.cat([torch.zeros(1, dtype=torch.int32, device=inputs.device), offsets[:-1]])               # :304
    assert batch_offsets.dtype == torch.int64
    total_packed = int(offsets[-1].item())                                                                           # :308
    packed = torch.zeros((total_packed,), dtype=torch.float16, device=inputs.device)                                 # :309
    _launch(f"compress_{which}_batched" + sfx, grid, [inputs_t.data_ptr(), bitmaps.data_ptr(), accum.data_ptr(), packed.data_ptr(),
                                                     batch_offsets.data_ptr()])
    torch.cuda.synchronize()
    return bitmaps, accum, packed, batch_offsets, offsets


def convert_key_batched(inputs: torch.Tensor):
    """-> (bitmaps [B, 2M] int64, accum_counts [B, 2M+1] int32, packed flat fp16, start offsets, end offsets)."""
    return _convert(inputs, "key")


def convert_value_batched(inputs: torch.Tensor):
    return _convert(inputs, "value")
