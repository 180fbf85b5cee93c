"""mustafar_b200 — B200-native (sm_100a) sparse-KV decode path of dhjoo98/mustafar.

Host side mirrors the reference operator surface:
  mustafar_b200.compression       ↔ kernel/compression.py           (convert_key_batched, convert_value_batched)
  mustafar_b200.mustafar_package  ↔ kernel/kernel_wrapper (pybind)   (mustafar_key_formulation, mustafar_value_formulation)
  mustafar_b200.pruning           ↔ models/llama_mustafar_kernel.py:77-153 (dh_prune_key / dh_prune_value)
  mustafar_b200.attention         — fused decode attention + the slab KV cache (new)
All compute is hand-written CUDA behind the C ABI in include/mustafar_b200.h; there is no fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["compression", "mustafar_package", "pruning", "attention"]
__version__ = "0.1.0"
