"""Mirror of the functional API of ``kernels/triton/layernorm_kernels.py`` (reference :191-311)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from ... import ops
from .. import _measure as M

HAS_TRITON = True


def triton_layernorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, eps: float = 1e-5,
                     residual: Optional[torch.Tensor] = None, residual_alpha: float = 1.0) -> torch.Tensor:
    """``LayerNorm(x + residual_alpha * residual)``; x ``[B,S,h]`` or ``[B,h]`` (reference :191-277)."""
    return ops.layernorm(x, weight, bias, eps, residual, residual_alpha)


pytorch_layernorm = triton_layernorm  # same contract (reference :279-311); there is no eager fallback


# ------------------------------------------------------------------------------------------------------------------
# The reference's own measurement / validation helpers for this file (:318-600), same arguments and result keys.
# ------------------------------------------------------------------------------------------------------------------
def _ln_problem(batch_size, seq_len, hidden_size, device, dtype=torch.bfloat16):
    g = torch.Generator(device=device).manual_seed(0)
    r = lambda *s: torch.randn(*s, device=device, generator=g).to(dtype)
    return r(batch_size, seq_len, hidden_size), r(batch_size, seq_len, hidden_size), r(hidden_size), r(hidden_size)


def _torch_layernorm(x, w, b, residual=None, fp32: bool = False):
    """Comparator: ``torch.nn.functional.layer_norm`` (after an explicit residual add), in fp32 for validation."""
    if fp32:
        x, w, b, residual = x.float(), w.float(), b.float(), None if residual is None else residual.float()
    if residual is not None:
        x = x + residual
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def benchmark_layernorm(batch_size: int, seq_len: int, hidden_size: int, device: str = "cuda", iterations: int = 100,
                        warmup: int = 10) -> Dict[str, float]:
    """reference :318-425 — K5 next to torch's LayerNorm, without and with the residual add."""
    res = {"batch_size": batch_size, "seq_len": seq_len, "hidden_size": hidden_size, "triton_layernorm_ms": 0.0,
           "pytorch_layernorm_ms": 0.0, "speedup": 0.0}
    if not M.cuda_ready(device):
        return res
    x, r, w, b = _ln_problem(batch_size, seq_len, hidden_size, device)
    res["triton_layernorm_ms"] = M.time_ms(lambda: triton_layernorm(x, w, b), warmup, iterations)
    res["pytorch_layernorm_ms"] = M.time_ms(lambda: _torch_layernorm(x, w, b), warmup, iterations)
    res["triton_layernorm_residual_ms"] = M.time_ms(lambda: triton_layernorm(x, w, b, residual=r), warmup, iterations)
    res["pytorch_layernorm_residual_ms"] = M.time_ms(lambda: _torch_layernorm(x, w, b, r), warmup, iterations)
    res["speedup"] = res["pytorch_layernorm_ms"] / max(res["triton_layernorm_ms"], 1e-9)
    res["speedup_residual"] = res["pytorch_layernorm_residual_ms"] / max(res["triton_layernorm_residual_ms"], 1e-9)
    res["gbs"] = 2.0 * x.numel() * x.element_size() / max(res["triton_layernorm_ms"], 1e-9) / 1e6
    return res


def compare_with_torch_layernorm(batch_size: int, seq_len: int, hidden_size: int, device: str = "cuda") -> Dict[str, float]:
    """reference :428-498 — max difference to the fp32 LayerNorm of the same 16-bit inputs."""
    if not M.cuda_ready(device):
        return {"max_difference": 0.0, "is_correct": False}
    x, r, w, b = _ln_problem(batch_size, seq_len, hidden_size, device)
    d0 = M.max_abs_diff(triton_layernorm(x, w, b), _torch_layernorm(x, w, b, fp32=True))
    d1 = M.max_abs_diff(triton_layernorm(x, w, b, residual=r), _torch_layernorm(x, w, b, r, fp32=True))
    tol = 4 * M.MAX_ABS_TOL  # outputs reach |y| ~ 10 with unit-variance gamma/beta: one bf16 ulp there is 6e-2
    return {"max_difference": d0, "max_residual_difference": d1, "is_correct": d0 <= tol, "is_residual_correct": d1 <= tol,
            "batch_size": batch_size, "seq_len": seq_len, "hidden_size": hidden_size}


def profile_memory_usage(batch_size: int, seq_len: int, hidden_size: int, device: str = "cuda") -> Dict[str, float]:
    """reference :501-600 — peak memory of the residual form: torch materialises ``x + residual``, K5 does not."""
    if not M.cuda_ready(device):
        return {"torch_memory_mb": 0.0, "triton_memory_mb": 0.0, "memory_saving_percent": 0.0}
    x, r, w, b = _ln_problem(batch_size, seq_len, hidden_size, device)
    mem_t, _ = M.peak_mb(lambda: _torch_layernorm(x, w, b, r))
    mem_k, _ = M.peak_mb(lambda: triton_layernorm(x, w, b, residual=r))
    return {"torch_memory_mb": mem_t, "triton_memory_mb": mem_k, "memory_saving_mb": mem_t - mem_k,
            "memory_saving_percent": 100.0 * (mem_t - mem_k) / max(mem_t, 1e-6), "batch_size": batch_size, "seq_len": seq_len,
            "hidden_size": hidden_size}
