"""Mirror of the functional API of ``kernels/triton/fused_layernorm_qkv.py`` (reference :422-611, wrappers :1073, :1118):
LayerNorm followed by the Q/K/V projections, returning head-major views ready for the attention kernels.

Here the op is the LayerNorm kernel (``b200_layernorm``) followed by ONE tcgen05 GEMM over the concatenated
[Wq; Wk; Wv] weight (``b200_linear_act``); the outputs are strided views of that single result — K1 consumes them
through its (batch, seq, head) strides without copies. (The reference fuses both into one Triton kernel that has never
run; projection fusion into the attention kernel itself is out of scope, §2.2.)"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ... import ops

HAS_TRITON = True


def _cat_bias(biases, weights):
    if all(b is None for b in biases):
        return None
    return torch.cat([b if b is not None else torch.zeros(w.shape[0], dtype=w.dtype, device=w.device)
                      for b, w in zip(biases, weights)])


def triton_fused_layernorm_qkv(hidden_states: torch.Tensor, layernorm_weight: torch.Tensor,
                               layernorm_bias: Optional[torch.Tensor], query_weight: torch.Tensor, key_weight: torch.Tensor,
                               value_weight: torch.Tensor, query_bias: Optional[torch.Tensor] = None,
                               key_bias: Optional[torch.Tensor] = None, value_bias: Optional[torch.Tensor] = None,
                               eps: float = 1e-5, num_heads: int = 0, num_kv_heads: Optional[int] = None
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns ``(q [B,S,H,Dq], k [B,S,Hkv,Dk], v [B,S,Hkv,Dk])`` with ``Dq = Wq.rows/H`` and ``Dk = Wk.rows/Hkv``."""
    B, S, hidden = hidden_states.shape
    if num_heads == 0:  # reference :470-480: try the common head sizes
        for hs in (64, 80, 128):
            if hidden % hs == 0:
                num_heads = hidden // hs
                break
        if num_heads == 0:
            num_heads = max(1, hidden // 64)
    if num_kv_heads is None:
        num_kv_heads = num_heads
    nq, nk, nv = query_weight.shape[0], key_weight.shape[0], value_weight.shape[0]
    if nq % num_heads or nk % num_kv_heads or nv % num_kv_heads:
        raise ValueError("projection widths must be divisible by the head counts")
    normed = ops.layernorm(hidden_states, layernorm_weight, layernorm_bias, eps)
    w = torch.cat([query_weight, key_weight, value_weight], dim=0)
    b = _cat_bias((query_bias, key_bias, value_bias), (query_weight, key_weight, value_weight))
    qkv = ops.linear_act(normed, w, b)
    q, k, v = qkv.split([nq, nk, nv], dim=-1)
    return (q.view(B, S, num_heads, nq // num_heads), k.view(B, S, num_kv_heads, nk // num_kv_heads),
            v.view(B, S, num_kv_heads, nv // num_kv_heads))


pytorch_fused_layernorm_qkv = triton_fused_layernorm_qkv


def flash_compatible_wrapper(hidden_states, layernorm_weight, layernorm_bias, qkv_weight, qkv_bias=None, eps: float = 1e-5,
                             num_heads: int = 0, num_kv_heads: Optional[int] = None):
    """reference :1073-1115 — combined ``[3*hidden, hidden]`` weight; outputs ``[B,S,H,D]`` for FlashAttention3."""
    hidden = hidden_states.shape[-1]
    wq, wk, wv = qkv_weight.split(hidden, dim=0) if qkv_weight.shape[0] == 3 * hidden else qkv_weight.chunk(3, dim=0)
    bq = bk = bv = None
    if qkv_bias is not None:
        bq, bk, bv = qkv_bias.split([wq.shape[0], wk.shape[0], wv.shape[0]])
    return triton_fused_layernorm_qkv(hidden_states, layernorm_weight, layernorm_bias, wq, wk, wv, bq, bk, bv, eps,
                                      num_heads, num_kv_heads)


def ring_compatible_wrapper(hidden_states, layernorm_weight, layernorm_bias, q_weight, k_weight, v_weight, q_bias=None,
                            k_bias=None, v_bias=None, eps: float = 1e-5, num_heads: int = 0,
                            num_kv_heads: Optional[int] = None):
    """reference :1118-1161 — outputs ``[B,H,S,D]`` (the ring / TP internal layout), as views."""
    q, k, v = triton_fused_layernorm_qkv(hidden_states, layernorm_weight, layernorm_bias, q_weight, k_weight, v_weight,
                                         q_bias, k_bias, v_bias, eps, num_heads, num_kv_heads)
    return q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
