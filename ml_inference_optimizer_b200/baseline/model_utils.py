"""``baseline/model_utils.py`` surface on the hot path: ``add_paged_attention_to_model`` (reference
``baseline/model_utils.py:599-759``).

The reference walks ``named_modules()`` looking for attention layers and monkey-patches their ``forward`` so the K,V of
new tokens go through ``triton_reshape_and_cache`` and decode goes through ``triton_paged_attention_forward`` (both of
which never run there). Here the same entry point returns a deep copy whose attention layers are the
B200 shells (``ModelConverter``): prefill = K1 + block scatter, decode = ``b200_kv_append`` + K2 through the block tables.
The model keeps the reference's ``set_paged_kv_cache(cache)`` method and gains ``generate_paged(input_ids, n)``.
"""
from __future__ import annotations

import types
from copy import deepcopy
from typing import Optional

import torch
import torch.nn as nn

from ..kernels.attention import flash_attention as _fa
from . import inference as _inf


def add_paged_attention_to_model(model: nn.Module, config: Optional[_fa.FlashAttentionConfig] = None) -> nn.Module:
    """Return a copy of ``model`` whose attention layers use the paged KV cache kernels (reference :599-759).

    Raises ``ValueError`` when the model holds no attention layer the converter recognises — the reference logs a
    warning and returns the model untouched, which hides a silently unoptimised model."""
    model = deepcopy(model)
    if not any(isinstance(m, _fa._HFAttentionAdapter) for m in model.modules()):
        # compute in the model's own 16-bit dtype so the paged cache (allocated in that dtype) matches q
        dtype = next(model.parameters()).dtype
        prec = {torch.bfloat16: "bf16", torch.float16: "fp16"}.get(dtype, "bf16")
        causal_cfg = config or _fa.FlashAttentionConfig(causal=True, precision=prec)
        model = _fa.ModelConverter(causal_cfg).convert_model(model)
    if not any(isinstance(m, _fa._HFAttentionAdapter) for m in model.modules()):
        raise ValueError("add_paged_attention_to_model: no convertible attention layer found in the model")

    def set_paged_kv_cache(self, paged_kv_cache: "_inf.PagedKVCache") -> None:
        """Store the PagedKVCache used by ``generate_paged`` (reference :628-633)."""
        self._paged_kv_cache = paged_kv_cache

    def generate_paged(self, input_ids: torch.Tensor, max_new_tokens: int, block_size: int = 16) -> torch.Tensor:
        return _inf.generate_paged(self, input_ids, max_new_tokens, cache=getattr(self, "_paged_kv_cache", None),
                                   block_size=block_size)

    model._paged_kv_cache = None
    model.set_paged_kv_cache = types.MethodType(set_paged_kv_cache, model)
    model.generate_paged = types.MethodType(generate_paged, model)
    return model
