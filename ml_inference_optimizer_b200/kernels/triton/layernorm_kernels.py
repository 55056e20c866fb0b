"""Mirror of the functional API of ``kernels/triton/layernorm_kernels.py`` (reference :191-311)."""
from __future__ import annotations

from typing import Optional

import torch

from ... import ops

HAS_TRITON = True


def triton_layernorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, eps: float = 1e-5,
                     residual: Optional[torch.Tensor] = None, residual_alpha: float = 1.0) -> torch.Tensor:
    """``LayerNorm(x + residual_alpha * residual)``; x ``[B,S,h]`` or ``[B,h]`` (reference :191-277)."""
    return ops.layernorm(x, weight, bias, eps, residual, residual_alpha)


pytorch_layernorm = triton_layernorm  # same contract (reference :279-311); there is no eager fallback
