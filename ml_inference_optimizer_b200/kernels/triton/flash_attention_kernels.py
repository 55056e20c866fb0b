"""Mirror of the functional API of ``kernels/triton/flash_attention_kernels.py``.

``triton_fused_attention`` / ``pytorch_fused_attention`` (reference :1361-1566, :1719-1779) run as K3 + K1 + K3.

``triton_flash_attention`` (reference :1150-1358) keeps its signature; it runs K1. ``pytorch_flash_attention``
(reference :1569-1700, which returns zeros — SURVEY.md F6) is kept as an alias of the same kernel: there is no
eager fallback."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple, Union

import torch

from ... import ops
from .. import _measure as M
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


# ------------------------------------------------------------------------------------------------------------------
# The reference's own measurement / validation helpers for this file (:1786-2060), same arguments and result keys.
# ------------------------------------------------------------------------------------------------------------------
def _standard_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """Comparator: materialised-score attention in fp32 (what the reference's helpers call "standard attention")."""
    qf, kf, vf = q.float(), k.float(), v.float()
    scores = torch.einsum("bshd,bkhd->bhsk", qf, kf) / math.sqrt(q.shape[-1])
    if causal:
        S, K = scores.shape[-2:]
        scores = scores.masked_fill(torch.ones(S, K, dtype=torch.bool, device=q.device).triu(1), float("-inf"))
    return torch.einsum("bhsk,bkhd->bshd", torch.softmax(scores, dim=-1), vf)


def _qkv(batch_size, seq_len, num_heads, head_dim, device, dtype=torch.bfloat16):
    g = torch.Generator(device=device).manual_seed(0)
    return tuple(torch.randn(batch_size, seq_len, num_heads, head_dim, device=device, dtype=dtype, generator=g) for _ in range(3))


def benchmark_flash_attention(seq_len: int, batch_size: int, num_heads: int, head_dim: int, device: str = "cuda",
                              causal: bool = False, iterations: int = 100, warmup: int = 10) -> Dict[str, float]:
    """reference :1786-1873 — K1 next to the materialised-score attention (CUDA events)."""
    res = {"sequence_length": seq_len, "batch_size": batch_size, "num_heads": num_heads, "head_dim": head_dim,
           "causal": causal, "flash_attention_ms": 0.0, "pytorch_attention_ms": 0.0, "speedup": 0.0}
    if not M.cuda_ready(device):
        return res
    q, k, v = _qkv(batch_size, seq_len, num_heads, head_dim, device)
    res["flash_attention_ms"] = M.time_ms(lambda: triton_flash_attention(q, k, v, causal=causal), warmup, iterations)
    res["pytorch_attention_ms"] = M.time_ms(lambda: _standard_attention(q, k, v, causal), min(warmup, 3), min(iterations, 10))
    res["speedup"] = res["pytorch_attention_ms"] / max(res["flash_attention_ms"], 1e-9)
    flops = 4.0 * batch_size * num_heads * seq_len * seq_len * head_dim * (0.5 if causal else 1.0)
    res["flash_attention_tflops"] = flops / max(res["flash_attention_ms"], 1e-9) / 1e9
    return res


def compare_with_standard_attention(seq_len: int, batch_size: int, num_heads: int, head_dim: int,
                                    device: str = "cuda") -> Dict[str, float]:
    """reference :1876-1963 — max difference and peak memory of both forms (the comparator holds the ``[B,H,S,S]`` scores)."""
    if not M.cuda_ready(device):
        return {"max_difference": 0.0, "is_correct": False, "memory_standard_mb": 0.0, "memory_flash_mb": 0.0,
                "memory_reduction": 0.0}
    q, k, v = _qkv(batch_size, seq_len, num_heads, head_dim, device)
    mem_std, ref = M.peak_mb(lambda: _standard_attention(q, k, v))
    mem_flash, out = M.peak_mb(lambda: triton_flash_attention(q, k, v))
    diff = M.max_abs_diff(out, ref)
    return {"max_difference": diff, "is_correct": diff <= M.MAX_ABS_TOL, "memory_standard_mb": mem_std,
            "memory_flash_mb": mem_flash, "memory_reduction": mem_std / max(mem_flash, 1e-6), "sequence_length": seq_len,
            "can_handle_longer_sequences": True}


def compare_with_xformers(seq_len: int, batch_size: int, num_heads: int, head_dim: int, device: str = "cuda") -> Dict[str, float]:
    """reference :1966-2060 — xFormers' memory-efficient attention next to K1 when that package is installed."""
    try:
        import xformers.ops as xops  # noqa: F401
    except Exception:
        return {"has_xformers": False, "flash_attention_ms": 0.0, "xformers_attention_ms": 0.0, "speedup_ratio": 0.0}
    if not M.cuda_ready(device):
        return {"has_xformers": True, "flash_attention_ms": 0.0, "xformers_attention_ms": 0.0, "speedup_ratio": 0.0}
    q, k, v = _qkv(batch_size, seq_len, num_heads, head_dim, device)
    t_flash = M.time_ms(lambda: triton_flash_attention(q, k, v), 5, 20)
    t_x = M.time_ms(lambda: xops.memory_efficient_attention(q, k, v), 5, 20)
    return {"has_xformers": True, "flash_attention_ms": t_flash, "xformers_attention_ms": t_x, "speedup_ratio": t_x / t_flash,
            "sequence_length": seq_len, "batch_size": batch_size}
