"""Pins oracle/torch_oracle.py (the restatement used for BASELINE-size parity checks on the GPU) to the numpy oracle,
which is itself pinned to the reference's golden vectors (tests/test_oracle_golden.py).  Runs on the CPU in the
`not gpu` suite and again on the GPU (same functions, CUDA tensors) in the `gpu` suite."""
import numpy as np
import pytest
import torch

from oracle import mustafar_oracle as O
from oracle import torch_oracle as TO


def _randn(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=torch.float32).to(torch.float16)


def _pin(device):
    # prune: bit-exact, incl. ties, zeros, negative zeros
    x = _randn((3, 5, 200, 128), 1)
    x[0, 0, :7] = 0.0
    x[0, 1, :9, ::2] = 0.5
    x[1, 2, 3, :] = -0.0
    x[2, 4, :11, 5:90] = x[2, 4, :11, 5:6]
    for s in (0.0, 0.3, 0.5, 0.7, 0.99):
        want = O.prune_rows(x.numpy(), s)
        got = TO.prune_rows(x.to(device), s).cpu().numpy()
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), s
    # attention: same rounding points; fp32 summation order differs (BLAS vs numpy), which can flip the fp16 rounding of
    # a score now and then -> compare at 1e-4, far below the 2e-3 parity budget
    for (b, hkv, g, T, s, masked) in [(2, 2, 1, 600, 0.5, False), (1, 4, 4, 1312, 0.7, False), (2, 1, 8, 832, 0.5, True)]:
        k, v, q = _randn((b, hkv, T, 128), T), _randn((b, hkv, T, 128), T + 1), _randn((b, hkv * g, 1, 128), T + 2)
        L = O.compressed_length(T)
        assert L == TO.compressed_length(T)
        kp, vp = k.numpy().copy(), v.numpy().copy()
        kp[:, :, :L] = O.prune_rows(kp[:, :, :L], s)
        vp[:, :, :L] = O.prune_rows(vp[:, :, :L], s)
        mask = None
        if masked:
            mask = torch.zeros(b, 1, 1, T, dtype=torch.float16)
            mask[:, :, :, 5:400:3] = torch.finfo(torch.float16).min
        want = O.masked_dense_attention(q.numpy(), kp, vp, None if mask is None else mask.numpy()).astype(np.float32)
        got = TO.masked_dense_attention(q.to(device), torch.from_numpy(kp).to(device), torch.from_numpy(vp).to(device),
                                        None if mask is None else mask.to(device)).float().cpu().numpy()
        assert np.abs(got - want).max() <= 1e-4, (b, hkv, g, T, np.abs(got - want).max())
        exact = TO.attention_f64(q.to(device), torch.from_numpy(kp).to(device), torch.from_numpy(vp).to(device)).cpu().numpy()
        if not masked:
            assert np.abs(exact - O.attention_exact_f64(q.numpy(), kp, vp)).max() <= 1e-9


def test_torch_oracle_matches_numpy_oracle_cpu():
    _pin("cpu")


@pytest.mark.gpu
def test_torch_oracle_matches_numpy_oracle_gpu():
    _pin("cuda")
