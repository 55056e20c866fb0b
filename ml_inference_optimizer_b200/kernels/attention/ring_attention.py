"""Host-side mirror of the reference's ``kernels/attention/ring_attention.py`` module API.

``RingSelfAttention`` / ``RingCrossAttention`` keep the constructor, attribute names (``qkv_proj`` or
``q_proj/k_proj/v_proj``, ``out_proj``) and ``forward`` signatures (reference :168-410, :413-669). The reference's
chunk loop is single-process and not attention (independent softmax per chunk, F7); here the attention is exact:
on one GPU one K1 launch covers all keys (the chunking lives inside the kernel's KV-tile loop); when a process group
is initialised and ``config.world_size > 1`` the hidden states are this rank's sequence shard and K4
(``parallelism.ring.ring_attention_forward``) circulates the KV blocks over NVLink.
Added keyword-only arguments (defaults keep the reference behaviour): ``causal``, ``partition``, ``group``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ... import ops
from ...parallelism.ring import ring_attention_forward

__all__ = ["RingAttentionConfig", "RingAttention", "RingSelfAttention", "RingCrossAttention", "ModelConverter",
           "benchmark_ring_attention", "calculate_theoretical_flops"]


@dataclass
class RingAttentionConfig:
    """reference :41-89."""
    world_size: int = 1
    chunk_size: Optional[int] = None
    fuse_qkv: bool = True
    use_flash_attention: bool = False
    use_triton: bool = True
    precision: str = "bf16"
    communication_dtype: torch.dtype = torch.bfloat16
    normalize_attention_scores: bool = True
    attention_dropout: float = 0.0

    def __post_init__(self):
        if self.world_size < 1:
            raise ValueError(f"world_size must be >= 1, got {self.world_size}")
        if self.chunk_size is not None and self.chunk_size <= 0:
            raise ValueError(f"chunk_size must be > 0 if specified, got {self.chunk_size}")
        if self.precision not in ["fp32", "fp16", "bf16"]:
            raise ValueError(f"precision must be one of ['fp32', 'fp16', 'bf16'], got {self.precision}")
        if self.attention_dropout < 0 or self.attention_dropout >= 1:
            raise ValueError(f"attention_dropout must be in [0, 1), got {self.attention_dropout}")
        # fp32 has no tensor-core path: storage is bf16, accumulation and softmax are fp32
        self.compute_dtype = {"fp32": torch.bfloat16, "fp16": torch.float16, "bf16": torch.bfloat16}[self.precision]


class RingAttention(nn.Module):
    """reference :92-165 (base class: sizes, scale, memory model)."""

    def __init__(self, hidden_size: int, num_attention_heads: int, config: RingAttentionConfig):
        super().__init__()
        self.hidden_size = hidden_size
        self.num_attention_heads = num_attention_heads
        self.config = config
        self.head_dim = hidden_size // num_attention_heads
        if self.head_dim * num_attention_heads != hidden_size:
            raise ValueError(f"hidden_size {hidden_size} not divisible by num_attention_heads {num_attention_heads}")
        self.scale = 1.0 / math.sqrt(self.head_dim)

    def get_effective_bytes_per_token(self) -> int:
        dtype_size = 4 if self.config.precision == "fp32" else 2
        return int(dtype_size * (4 * self.hidden_size + self.hidden_size / self.config.world_size))

    def calculate_theoretical_memory_savings(self, seq_len: int) -> float:
        standard_bytes = seq_len * seq_len * self.num_attention_heads * 2
        ring_bytes = self.get_effective_bytes_per_token() * seq_len
        return standard_bytes / ring_bytes

    def _attend(self, q, k, v, causal: bool, partition: str, group) -> torch.Tensor:
        """q [B,Sq,H,D], k/v [B,Sk,H,D] -> [B,Sq,H,D]; distributed when a group is live and world_size > 1."""
        scale = self.scale if self.config.normalize_attention_scores else 1.0
        if self.config.world_size > 1 and dist.is_initialized() and dist.get_world_size(group) > 1:
            return ring_attention_forward(q, k, v, causal=causal, softmax_scale=scale, group=group, partition=partition)
        return ops.flash_attn_fwd(q, k, v, causal=causal, softmax_scale=scale)


class RingSelfAttention(RingAttention):
    """reference :168-410."""

    def __init__(self, hidden_size: int, num_attention_heads: int, config: RingAttentionConfig, *, causal: bool = False,
                 partition: str = "contiguous", group=None):
        super().__init__(hidden_size, num_attention_heads, config)
        if config.fuse_qkv:
            self.qkv_proj = nn.Linear(hidden_size, 3 * hidden_size, bias=True)
        else:
            self.q_proj = nn.Linear(hidden_size, hidden_size, bias=True)
            self.k_proj = nn.Linear(hidden_size, hidden_size, bias=True)
            self.v_proj = nn.Linear(hidden_size, hidden_size, bias=True)
        self.out_proj = nn.Linear(hidden_size, hidden_size, bias=True)
        self.attention_dropout = nn.Dropout(config.attention_dropout)
        self.causal, self.partition, self.group = causal, partition, group

    def prepare_attention_inputs(self, hidden_states: torch.Tensor):
        """reference :237-273 — returns q,k,v as [B,S,H,D] views (the reference permutes to [B,H,S,D] and pre-scales q)."""
        B, S, _ = hidden_states.shape
        H, D = self.num_attention_heads, self.head_dim
        if self.config.fuse_qkv:
            q, k, v = self.qkv_proj(hidden_states).split(self.hidden_size, dim=-1)
        else:
            q, k, v = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        return q.view(B, S, H, D), k.view(B, S, H, D), v.view(B, S, H, D)

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("additive attention masks ([B,1,1,S]) are not supported on the CUDA path; use causal=True")
        if self.training and self.config.attention_dropout > 0:
            raise NotImplementedError("attention dropout is not implemented (inference path)")
        orig = hidden_states.dtype
        B, S, _ = hidden_states.shape
        q, k, v = self.prepare_attention_inputs(hidden_states)
        dt = self.config.compute_dtype
        if q.dtype != dt:
            q, k, v = q.to(dt), k.to(dt), v.to(dt)
        ctx = self._attend(q, k, v, self.causal, self.partition, self.group)
        return self.out_proj(ctx.reshape(B, S, self.hidden_size).to(orig))


class RingCrossAttention(RingAttention):
    """reference :413-669 — queries from ``query_states``, keys/values from ``key_value_states`` (never causal)."""

    def __init__(self, hidden_size: int, num_attention_heads: int, config: RingAttentionConfig, *, group=None):
        super().__init__(hidden_size, num_attention_heads, config)
        self.q_proj = nn.Linear(hidden_size, hidden_size, bias=True)
        self.k_proj = nn.Linear(hidden_size, hidden_size, bias=True)
        self.v_proj = nn.Linear(hidden_size, hidden_size, bias=True)
        self.out_proj = nn.Linear(hidden_size, hidden_size, bias=True)
        self.attention_dropout = nn.Dropout(config.attention_dropout)
        self.group = group

    def forward(self, query_states: torch.Tensor, key_value_states: torch.Tensor,
                attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("additive attention masks are not supported on the CUDA path")
        orig = query_states.dtype
        B, Sq, _ = query_states.shape
        Sk = key_value_states.shape[1]
        H, D = self.num_attention_heads, self.head_dim
        q = self.q_proj(query_states).view(B, Sq, H, D)
        k = self.k_proj(key_value_states).view(B, Sk, H, D)
        v = self.v_proj(key_value_states).view(B, Sk, H, D)
        dt = self.config.compute_dtype
        if q.dtype != dt:
            q, k, v = q.to(dt), k.to(dt), v.to(dt)
        ctx = self._attend(q, k, v, False, "contiguous", self.group)
        return self.out_proj(ctx.reshape(B, Sq, self.hidden_size).to(orig))


class ModelConverter:
    """reference :672-835 — replace self-attention modules exposing q/k/v/out projections by ``RingSelfAttention``,
    copying the weights."""

    def __init__(self, config: RingAttentionConfig, causal: bool = False):
        self.config = config
        self.causal = causal

    def convert_model(self, model: nn.Module) -> nn.Module:
        for name, sub in list(model.named_children()):
            new = self._convert(sub)
            if new is not None:
                setattr(model, name, new)
            else:
                self.convert_model(sub)
        return model

    def _convert(self, m: nn.Module) -> Optional[nn.Module]:
        if isinstance(m, RingAttention):
            return None
        out = getattr(m, "out_proj", None) or getattr(m, "o_proj", None)
        if out is None or not all(isinstance(getattr(m, a, None), nn.Linear) for a in ("q_proj", "k_proj", "v_proj")):
            return None
        heads = getattr(m, "num_heads", None) or getattr(m, "num_attention_heads", None)
        if heads is None or m.k_proj.out_features != m.q_proj.out_features:
            return None
        cfg = RingAttentionConfig(**{**{k: v for k, v in self.config.__dict__.items() if k != "compute_dtype"}, "fuse_qkv": True})
        new = RingSelfAttention(m.q_proj.in_features, heads, cfg, causal=self.causal)
        with torch.no_grad():
            new.qkv_proj.weight.copy_(torch.cat([m.q_proj.weight, m.k_proj.weight, m.v_proj.weight], dim=0))
            zeros = lambda l: torch.zeros(l.out_features, device=l.weight.device, dtype=l.weight.dtype)
            new.qkv_proj.bias.copy_(torch.cat([l.bias if l.bias is not None else zeros(l) for l in (m.q_proj, m.k_proj, m.v_proj)]))
            new.out_proj.weight.copy_(out.weight)
            new.out_proj.bias.copy_(out.bias if out.bias is not None else zeros(out))
        return new.to(device=out.weight.device, dtype=out.weight.dtype)


def calculate_theoretical_flops(seq_len: int, batch_size: int, hidden_size: int, num_heads: int) -> int:
    """reference :921-955 (multiply-accumulates counted once, as the reference does)."""
    head_dim = hidden_size // num_heads
    qkv = 3 * batch_size * seq_len * hidden_size * hidden_size
    scores = batch_size * num_heads * seq_len * seq_len * head_dim
    out = batch_size * num_heads * seq_len * seq_len * head_dim
    proj = batch_size * seq_len * hidden_size * hidden_size
    return qkv + scores + out + proj


def benchmark_ring_attention(batch_size: int = 1, seq_len: int = 8192, hidden_size: int = 4096, num_heads: int = 32,
                             causal: bool = True, num_iters: int = 10, warmup_iters: int = 3) -> Dict[str, float]:
    """reference :838-918 — time the module forward on this rank's shard (CUDA events)."""
    cfg = RingAttentionConfig(world_size=dist.get_world_size() if dist.is_initialized() else 1)
    mod = RingSelfAttention(hidden_size, num_heads, cfg, causal=causal).to("cuda", torch.bfloat16)
    x = torch.randn(batch_size, seq_len // cfg.world_size, hidden_size, device="cuda", dtype=torch.bfloat16)
    for _ in range(warmup_iters):
        mod(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(num_iters):
        mod(x)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / num_iters
    flops = 2 * calculate_theoretical_flops(seq_len, batch_size, hidden_size, num_heads) / cfg.world_size
    return {"ms": ms, "tflops_per_gpu": flops / ms / 1e9, "world_size": cfg.world_size}
