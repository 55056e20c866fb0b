"""Mirror of the functional API of ``kernels/triton/mlp_kernels.py`` (reference :648-803)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from ... import ops
from .. import _measure as M


def triton_fused_mlp(hidden_states: torch.Tensor, fc1_weight: torch.Tensor, fc1_bias: Optional[torch.Tensor],
                     fc2_weight: torch.Tensor, fc2_bias: Optional[torch.Tensor], activation: str = "gelu",
                     fc1_gate_weight: Optional[torch.Tensor] = None, fc1_gate_bias: Optional[torch.Tensor] = None
                     ) -> torch.Tensor:
    """``activation="gelu"`` is the tanh approximation, as in the reference's Triton kernel (:144-161)."""
    act = {"gelu": "gelu_tanh", "relu": "relu", "swiglu": "swiglu"}.get(activation)
    if act is None:
        raise ValueError(f"Unsupported activation function: {activation}")
    if act == "swiglu" and fc1_gate_weight is None:
        raise ValueError("SwiGLU activation requires gate weights")
    return ops.fused_mlp(hidden_states, fc1_weight, fc1_bias, fc2_weight, fc2_bias, act, fc1_gate_weight, fc1_gate_bias)


def pytorch_fused_mlp(hidden_states: torch.Tensor, fc1_weight: torch.Tensor, fc1_bias: Optional[torch.Tensor],
                      fc2_weight: torch.Tensor, fc2_bias: Optional[torch.Tensor], activation: str = "gelu",
                      fc1_gate_weight: Optional[torch.Tensor] = None, fc1_gate_bias: Optional[torch.Tensor] = None
                      ) -> torch.Tensor:
    """Same contract as the reference's eager form (:759-803): ``activation="gelu"`` is the EXACT erf GELU (:783)."""
    act = {"gelu": "gelu_erf", "relu": "relu", "swiglu": "swiglu"}.get(activation)
    if act is None:
        raise ValueError(f"Unsupported activation function: {activation}")
    if act == "swiglu" and fc1_gate_weight is None:
        raise ValueError("SwiGLU activation requires gate weights")
    return ops.fused_mlp(hidden_states, fc1_weight, fc1_bias, fc2_weight, fc2_bias, act, fc1_gate_weight, fc1_gate_bias)


# ------------------------------------------------------------------------------------------------------------------
# The reference's own measurement / validation helpers for this file (:810-1090), same arguments and result keys.
# ------------------------------------------------------------------------------------------------------------------
def _mlp_problem(batch_size, seq_len, hidden_size, intermediate_size, activation, device, dtype):
    if activation not in ("gelu", "relu", "swiglu"):
        raise ValueError(f"Unsupported activation function: {activation}")
    g = torch.Generator(device=device).manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, device=device, generator=g) * sc).to(dtype)
    x = r(batch_size, seq_len, hidden_size)
    w1, b1 = r(intermediate_size, hidden_size, sc=hidden_size ** -0.5), r(intermediate_size, sc=0.1)
    w2, b2 = r(hidden_size, intermediate_size, sc=intermediate_size ** -0.5), r(hidden_size, sc=0.1)
    wg, bg = (r(intermediate_size, hidden_size, sc=hidden_size ** -0.5), r(intermediate_size, sc=0.1)) if activation == "swiglu" else (None, None)
    return x, w1, b1, w2, b2, wg, bg


def _unfused_mlp(x, w1, b1, w2, b2, activation, wg, bg, fp32: bool):
    """Comparator: the unfused sequence of library GEMMs and elementwise kernels (fp32 math for validation, the tensors'
    own dtype — cuBLAS — for timing). ``gelu`` is the tanh form, as in triton_fused_mlp."""
    c = (lambda t: None if t is None else t.float()) if fp32 else (lambda t: t)
    h = F.linear(c(x), c(w1), c(b1))
    if activation == "swiglu":
        h = F.silu(F.linear(c(x), c(wg), c(bg))) * h
    elif activation == "gelu":
        h = F.gelu(h, approximate="tanh")
    else:
        h = F.relu(h)
    return F.linear(h, c(w2), c(b2))


def benchmark_fused_mlp(batch_size: int, seq_len: int, hidden_size: int, intermediate_size: int, activation: str = "gelu",
                        device: str = "cuda", dtype: torch.dtype = torch.bfloat16, num_warmup: int = 10,
                        num_iter: int = 100) -> Dict[str, float]:
    """reference :810-922 — K3 FusedMLP next to unfused cuBLAS + elementwise kernels in the same dtype."""
    res = {"batch_size": batch_size, "seq_len": seq_len, "hidden_size": hidden_size, "intermediate_size": intermediate_size,
           "activation": activation, "triton_time_ms": 0.0, "pytorch_time_ms": 0.0, "speedup": 0.0}
    if not M.cuda_ready(device):
        return res
    x, w1, b1, w2, b2, wg, bg = _mlp_problem(batch_size, seq_len, hidden_size, intermediate_size, activation, device, dtype)
    res["triton_time_ms"] = M.time_ms(lambda: triton_fused_mlp(x, w1, b1, w2, b2, activation, wg, bg), num_warmup, num_iter)
    res["pytorch_time_ms"] = M.time_ms(lambda: _unfused_mlp(x, w1, b1, w2, b2, activation, wg, bg, False), num_warmup, num_iter)
    res["speedup"] = res["pytorch_time_ms"] / max(res["triton_time_ms"], 1e-9)
    flops = 2.0 * batch_size * seq_len * hidden_size * intermediate_size * (3 if activation == "swiglu" else 2)
    res["tflops"] = flops / max(res["triton_time_ms"], 1e-9) / 1e9
    return res


def validate_fused_mlp(batch_size: int, seq_len: int, hidden_size: int, intermediate_size: int, activation: str = "gelu",
                       device: str = "cuda", dtype: torch.dtype = torch.bfloat16) -> Dict[str, float]:
    """reference :925-1000 — max difference against the unfused fp32 computation on the same 16-bit inputs."""
    if not M.cuda_ready(device):
        return {"is_correct": False, "max_diff": 0.0}
    x, w1, b1, w2, b2, wg, bg = _mlp_problem(batch_size, seq_len, hidden_size, intermediate_size, activation, device, dtype)
    diff = M.max_abs_diff(triton_fused_mlp(x, w1, b1, w2, b2, activation, wg, bg),
                          _unfused_mlp(x, w1, b1, w2, b2, activation, wg, bg, True))
    return {"is_correct": diff <= M.MAX_ABS_TOL, "max_diff": diff, "batch_size": batch_size, "seq_len": seq_len,
            "hidden_size": hidden_size, "intermediate_size": intermediate_size, "activation": activation}


def profile_memory_usage(batch_size: int, seq_len: int, hidden_size: int, intermediate_size: int, activation: str = "gelu",
                         device: str = "cuda") -> Dict[str, float]:
    """reference :1003-1090 — peak memory of the unfused sequence (up, gate and product tensors live together) and of K3
    (one ``[T, intermediate]`` workspace)."""
    if not M.cuda_ready(device):
        return {"pytorch_memory_mb": 0.0, "triton_memory_mb": 0.0, "memory_saving_mb": 0.0, "memory_saving_percent": 0.0}
    x, w1, b1, w2, b2, wg, bg = _mlp_problem(batch_size, seq_len, hidden_size, intermediate_size, activation, device, torch.bfloat16)
    mem_pt, _ = M.peak_mb(lambda: _unfused_mlp(x, w1, b1, w2, b2, activation, wg, bg, False))
    mem_k3, _ = M.peak_mb(lambda: triton_fused_mlp(x, w1, b1, w2, b2, activation, wg, bg))
    return {"pytorch_memory_mb": mem_pt, "triton_memory_mb": mem_k3, "memory_saving_mb": mem_pt - mem_k3,
            "memory_saving_percent": 100.0 * (mem_pt - mem_k3) / max(mem_pt, 1e-6)}
