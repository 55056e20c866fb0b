"""Mirror of the functional API of ``kernels/triton/mlp_kernels.py`` (reference :648-803)."""
from __future__ import annotations

from typing import Optional

import torch

from ... import ops


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
