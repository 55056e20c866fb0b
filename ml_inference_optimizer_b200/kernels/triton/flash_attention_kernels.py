"""Mirror of the functional API of ``kernels/triton/flash_attention_kernels.py``.

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


def triton_fused_attention(*args, **kwargs):
    """reference :1361-1566 fuses the QKV / output projections into the attention kernel; out of scope here
    (SURVEY.md §2.2): projections stay cuBLAS GEMMs, use ``FlashSelfAttention``."""
    raise NotImplementedError("triton_fused_attention (projection-fused attention) is out of scope: use FlashSelfAttention")
