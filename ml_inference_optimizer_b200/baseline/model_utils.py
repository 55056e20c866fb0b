"""``baseline/model_utils.py`` surface on the hot path: ``add_paged_attention_to_model`` (reference
``baseline/model_utils.py:599-759``).

The reference walks ``named_modules()`` looking for attention layers and monkey-patches their ``forward`` so the K,V of
new tokens go through ``triton_reshape_and_cache`` and decode goes through ``triton_paged_attention_forward`` (both of
which never run there). Here the same entry point returns a deep copy whose attention layers are the
B200 shells (``ModelConverter``): prefill = K1 + block scatter, decode = ``b200_kv_append`` + K2 through the block tables.
The model keeps the reference's ``set_paged_kv_cache(cache)`` method and gains ``generate_paged(input_ids, n)``.
"""
from __future__ import annotations

import fnmatch
import re
import types
from copy import deepcopy
from typing import Dict, List, Optional, Tuple, Union

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


# ------------------------------------------------------------------------------------------------------------------
# Model inspection / preparation helpers the converters' callers use (reference :18-260, :455-598). The hook-based FLOP
# counters (:263-452, :470-522) are not mirrored: they count nn.Linear calls, which a converted model no longer makes —
# the FLOP models of the kernels live next to them (kernels/triton/attention_kernels.calculate_attention_theoretical_flops,
# kernels/attention/ring_attention.calculate_theoretical_flops, benchmarks/metrics).
# ------------------------------------------------------------------------------------------------------------------
_ACTS = (nn.ReLU, nn.GELU, nn.SiLU, nn.Tanh)


def get_model_size(model: nn.Module) -> Dict[str, Union[int, float, str]]:
    """reference :18-73 — parameter counts and the memory they take in the dtype of the first parameter; the "activation"
    entry is the reference's estimate (one output row per Linear / Conv layer)."""
    params = list(model.parameters())
    total = sum(p.numel() for p in params)
    trainable = sum(p.numel() for p in params if p.requires_grad)
    dtype = params[0].dtype if params else torch.float32
    width = torch.empty((), dtype=dtype).element_size()
    act = sum(m.weight.size(0) * width for m in model.modules() if isinstance(m, (nn.Linear, nn.Conv1d, nn.Conv2d, nn.Conv3d)))
    state = sum(t.numel() * width for t in model.state_dict().values())
    mb = 1024 * 1024
    return {"total_params": total, "trainable_params": trainable, "non_trainable_params": total - trainable,
            "param_memory_mb": total * width / mb, "activation_memory_mb": act / mb, "state_dict_memory_mb": state / mb,
            "total_estimated_memory_mb": (total * width + act) / mb, "param_dtype": str(dtype), "param_bytes_per_element": width}


def get_model_layers(model: nn.Module) -> List[nn.Module]:
    """reference :76-113 — the modules that compute: containers, Identity / Dropout / Flatten and parents without parameters
    of their own are left out."""
    skip = (nn.Sequential, nn.ModuleList, nn.ModuleDict, nn.Identity, nn.Dropout, nn.Flatten)
    out = []
    for m in model.modules():
        if isinstance(m, skip):
            continue
        if next(m.children(), None) is not None and next(m.parameters(recurse=False), None) is None:
            continue
        out.append(m)
    return out


def get_attention_modules(model: nn.Module) -> List[nn.Module]:
    """reference :116-151 — by class name, by ``q_proj/k_proj/v_proj`` attributes, or by ``num_heads`` + ``head_dim``."""
    words = ("attention", "attn", "mha", "multihead", "multi_head")
    return [m for m in model.modules()
            if any(w in type(m).__name__.lower() for w in words)
            or all(hasattr(m, a) for a in ("q_proj", "k_proj", "v_proj"))
            or isinstance(m, nn.MultiheadAttention) or (hasattr(m, "num_heads") and hasattr(m, "head_dim"))]


def get_mlp_modules(model: nn.Module) -> List[nn.Module]:
    """reference :154-209 — by module / class name (mlp, feedforward, ffn, fc), then by shape: at least two Linears and an
    activation among the direct children."""
    words = ("mlp", "feedforward", "feed_forward", "ffn", "fc")
    found: List[nn.Module] = []
    for name, m in model.named_modules():
        if any(w in name.lower() or w in type(m).__name__.lower() for w in words):
            found.append(m)
    for m in model.modules():
        if any(m is f for f in found):
            continue
        kids = list(m.children())
        if sum(isinstance(k, nn.Linear) for k in kids) >= 2 and any(isinstance(k, _ACTS) for k in kids):
            found.append(m)
    return found


def find_modules_by_type(model: nn.Module, module_type: Union[type, Tuple[type, ...]]) -> List[Tuple[str, nn.Module]]:
    """reference :455-467."""
    return [(name, m) for name, m in model.named_modules() if isinstance(m, module_type)]


def convert_precision(model: nn.Module, precision: str) -> nn.Module:
    """reference :212-241 — fp32 / fp16 / bf16, on the device the model is on."""
    dtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}.get(precision.lower())
    if dtype is None:
        raise ValueError(f"Unsupported precision: {precision}")
    return model.to(dtype=dtype)


def create_random_input(batch_size: int, seq_len: int, hidden_size: int, dtype: torch.dtype = torch.float32,
                        device: str = "cuda") -> torch.Tensor:
    """reference :244-260."""
    return torch.randn(batch_size, seq_len, hidden_size, dtype=dtype, device=device)


def load_partial_weights(model: nn.Module, state_dict: Dict[str, torch.Tensor], strict: bool = False):
    """reference :525-570 — loads the entries whose name AND shape match; ``strict`` raises with the three lists (missing,
    unexpected, shape-mismatched) after loading."""
    own = model.state_dict()
    missing = sorted(set(own) - set(state_dict))
    unexpected = sorted(set(state_dict) - set(own))
    mismatched = sorted(k for k in set(own) & set(state_dict) if own[k].shape != state_dict[k].shape)
    result = model.load_state_dict({k: v for k, v in state_dict.items() if k in own and k not in mismatched}, strict=False)
    if strict and (missing or unexpected or mismatched):
        parts = [f"{label}: {keys}" for label, keys in (("Missing keys", missing), ("Unexpected keys", unexpected),
                                                         ("Shape-mismatched keys", mismatched)) if keys]
        raise RuntimeError("Error(s) in loading state_dict: " + "\n".join(parts))
    return result


def freeze_layers(model: nn.Module, layer_names: List[str]) -> nn.Module:
    """reference :573-597 — ``requires_grad = False`` for every parameter whose name matches one of the glob patterns."""
    pats = [re.compile(fnmatch.translate(p)) for p in layer_names]
    for name, param in model.named_parameters():
        if any(p.match(name) for p in pats):
            param.requires_grad = False
    return model
