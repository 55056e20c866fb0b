"""Shared plumbing of the in-module ``benchmark_*`` / ``compare_with_*`` / ``validate_*`` / ``profile_memory_usage`` helpers
the reference ships next to each kernel (e.g. kernels/triton/flash_attention_kernels.py:1786-2060,
kernels/triton/mlp_kernels.py:810-1090): CUDA-event timing, peak-memory bracketing and the comparison tolerance.

The eager PyTorch functions these helpers run are COMPARATORS (what the reference calls "standard" / "PyTorch"
implementation in the same helpers) — nothing on the product path calls them."""
from __future__ import annotations

from typing import Callable, Tuple

import torch

# BASELINE north_star tolerance for 16-bit outputs (the reference's helpers use 1e-3 on fp32 tensors; these kernels
# compute in bf16 / fp16 with fp32 accumulation)
MAX_ABS_TOL = 2e-2


def cuda_ready(device: str) -> bool:
    """The reference's helpers return a zero-filled result instead of raising when asked for "cuda" without a GPU; the
    mirrors keep that (it is a report, not a compute fallback: nothing is computed on the CPU)."""
    return str(device).startswith("cuda") and torch.cuda.is_available()


def time_ms(fn: Callable[[], object], warmup: int, iterations: int) -> float:
    for _ in range(max(0, warmup)):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(max(1, iterations)):
        fn()
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / max(1, iterations)


def peak_mb(fn: Callable[[], object]) -> Tuple[float, object]:
    """Peak bytes allocated ABOVE what was live before the call, in MiB, and the call's result."""
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    before = torch.cuda.memory_allocated()
    out = fn()
    torch.cuda.synchronize()
    return (torch.cuda.max_memory_allocated() - before) / (1024 ** 2), out


def max_abs_diff(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.float() - b.float()).abs().max().item()
