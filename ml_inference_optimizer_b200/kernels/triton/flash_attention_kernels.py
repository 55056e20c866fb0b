"""Mirror of the functional API of ``kernels/triton/flash_attention_kernels.py``.

``triton_fused_attention`` / ``pytorch_fused_attention`` (reference :1361-1566, :1719-1779) run as K3 + K1 + K3.

``triton_flash_attention`` (reference :1150-1358) keeps its signature; it runs K1. ``pytorch_flash_attention``
(reference :1569-1700, which returns zeros — SURVEY.md F6) is kept as an alias of the same kernel: there is no
eager fallback."""
from __future__ import annotations

from typing import Optional, Tuple, Union

import torch

from ... import ops
from ..attention.flash_attention import key_padding_mask_to_lengths

TRITON_AVAILABLE = True  # callers gate on this flag; the CUDA path is always the one that runs


def triton_flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor] = None,
                           causal: bool = False, softmax_scale: Optional[float] = None, dropout_p: float = 0.0,
                           return_softmax: bool = False, block_size: int = 128
                           ) -> Union[torch.Tensor, Tuple[torch.Tensor, torch.Tensor]]:
    """q,k,v ``[B,S,H,D]`` -> ``[B,S,H,D]``. With ``return_softmax`` the second value is the log-sum-exp
    ``[B,H,S]`` (the reference stores the pair (l, m) it is made of, :305-314)."""
    if dropout_p > 0.0:
        raise NotImplementedError("attention dropout is not implemented on the inference path")
    del block_size
    kv_lens = key_padding_mask_to_lengths(mask, k.shape[1]) if mask is not None else None
    if return_softmax:
        return ops.flash_attn_fwd(q, k, v, causal=causal, softmax_scale=softmax_scale, kv_lens=kv_lens, return_lse=True)
    return ops.flash_attn_fwd(q, k, v, causal=causal, softmax_scale=softmax_scale, kv_lens=kv_lens)


pytorch_flash_attention = triton_flash_attention


def triton_fused_attention(hidden_states: torch.Tensor, qkv_weight: torch.Tensor, qkv_bias: Optional[torch.Tensor],
                           out_weight: torch.Tensor, out_bias: Optional[torch.Tensor], mask: Optional[torch.Tensor] = None,
                           causal: bool = False, num_heads: int = 8, head_dim: Optional[int] = None, dropout_p: float = 0.0,
                           softmax_scale: Optional[float] = None, block_size: int = 128) -> torch.Tensor:
    """QKV projection -> attention -> output projection (reference :1361-1566; its eager twin pytorch_fused_attention,
    :1719-1779, defines the layout: ``qkv.view(B, S, 3, H, D).unbind(2)``). The reference's intent is one kernel; here the
    projections are K3 GEMMs with the bias fused in the epilogue and the attention is K1 reading q, k, v as strided views
    of the fused projection (no copies, no ``[B, S, 3, H, D]`` split tensors): three launches, no elementwise kernels."""
    if dropout_p > 0.0:
        raise NotImplementedError("attention dropout is not implemented on the inference path")
    del block_size
    if hidden_states.dim() != 3:
        raise ValueError(f"hidden_states must be [batch, seq, hidden], got {tuple(hidden_states.shape)}")
    B, S, hidden = hidden_states.shape
    D = head_dim if head_dim is not None else hidden // num_heads
    if tuple(qkv_weight.shape) != (3 * num_heads * D, hidden):
        raise ValueError(f"qkv_weight must be [{3 * num_heads * D}, {hidden}], got {tuple(qkv_weight.shape)}")
    if out_weight.shape[1] != num_heads * D:
        raise ValueError(f"out_weight must have {num_heads * D} input columns, got {tuple(out_weight.shape)}")
    qkv = ops.linear_act(hidden_states.reshape(B * S, hidden), qkv_weight, qkv_bias, None).view(B, S, 3, num_heads, D)
    q, k, v = qkv.unbind(dim=2)
    kv_lens = key_padding_mask_to_lengths(mask, S) if mask is not None else None
    o = ops.flash_attn_fwd(q, k, v, causal=causal, softmax_scale=softmax_scale, kv_lens=kv_lens)
    return ops.linear_act(o.reshape(B * S, num_heads * D), out_weight, out_bias, None).view(B, S, out_weight.shape[0])


pytorch_fused_attention = triton_fused_attention   # (reference :1719-1779; there is no eager fallback here)
