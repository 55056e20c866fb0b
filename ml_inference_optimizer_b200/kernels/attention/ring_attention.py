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
           "benchmark_ring_attention", "calculate_theoretical_flops", "compare_with_standard_attention"]


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


def _standard_self_attention(mod: "RingSelfAttention", x: torch.Tensor, fp32: bool) -> torch.Tensor:
    """Comparator of the helpers below: the same projections followed by materialised-score attention (``[B,H,S,S]``)."""
    if fp32:
        lin = lambda l, t: nn.functional.linear(t, l.weight.float(), l.bias.float())
        x = x.float()
    else:
        lin = lambda l, t: l(t)
    B, S, _ = x.shape
    H, D = mod.num_attention_heads, mod.head_dim
    if mod.config.fuse_qkv:
        q, k, v = lin(mod.qkv_proj, x).split(mod.hidden_size, dim=-1)
    else:
        q, k, v = lin(mod.q_proj, x), lin(mod.k_proj, x), lin(mod.v_proj, x)
    q, k, v = (t.view(B, S, H, D).transpose(1, 2) for t in (q, k, v))
    scores = q @ k.transpose(-1, -2) / math.sqrt(D)
    if mod.causal:
        scores = scores.masked_fill(torch.ones(S, S, dtype=torch.bool, device=x.device).triu(1), float("-inf"))
    ctx = (torch.softmax(scores.float(), dim=-1).to(v.dtype) @ v).transpose(1, 2).reshape(B, S, H * D)
    return lin(mod.out_proj, ctx)


def benchmark_ring_attention(seq_len: int = 8192, batch_size: int = 1, hidden_size: int = 4096, num_heads: int = 32,
                             causal: bool = False, num_iters: int = 10, warmup_iters: int = 3) -> Dict[str, float]:
    """reference :838-918 (same positional arguments and result keys) — the module forward on this rank's shard next to
    materialised-score attention with the same weights, CUDA events and peak memory. The comparator runs on a single
    process while its score matrix stays under 16 GB; beyond that its entries are NaN."""
    from .. import _measure as M

    world = dist.get_world_size() if dist.is_initialized() else 1
    cfg = RingAttentionConfig(world_size=world)
    mod = RingSelfAttention(hidden_size, num_heads, cfg, causal=causal).to("cuda", torch.bfloat16).eval()
    x = torch.randn(batch_size, seq_len // world, hidden_size, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        ring_mem, _ = M.peak_mb(lambda: mod(x))
        ms = M.time_ms(lambda: mod(x), warmup_iters, num_iters)
        std_ms = std_mem = float("nan")
        if world == 1 and batch_size * num_heads * seq_len * seq_len * 6 <= 16 * 2 ** 30:
            std_mem, _ = M.peak_mb(lambda: _standard_self_attention(mod, x, False))
            std_ms = M.time_ms(lambda: _standard_self_attention(mod, x, False), 1, max(1, min(num_iters, 5)))
    flops = 2 * calculate_theoretical_flops(seq_len, batch_size, hidden_size, num_heads) / world
    return {"standard_time_ms": std_ms, "ring_time_ms": ms, "speedup_factor": std_ms / ms, "standard_memory_mb": std_mem,
            "ring_memory_mb": ring_mem, "memory_savings_factor": std_mem / max(ring_mem, 1e-6),
            "ms": ms, "tflops_per_gpu": flops / ms / 1e9, "world_size": world}


def compare_with_standard_attention(seq_len: int, batch_size: int, hidden_size: int, num_heads: int) -> Dict[str, float]:
    """reference :958-1040 — RingSelfAttention (separate q/k/v projections, one process) against materialised-score
    attention computed in fp32 with the same weights, followed by the benchmark above."""
    cfg = RingAttentionConfig(world_size=1, fuse_qkv=False)
    torch.manual_seed(0)
    mod = RingSelfAttention(hidden_size, num_heads, cfg).to("cuda", torch.bfloat16).eval()
    x = torch.randn(batch_size, seq_len, hidden_size, device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        err = (mod(x).float() - _standard_self_attention(mod, x, True)).abs()
        ref_mean = _standard_self_attention(mod, x, True).abs().mean().item()
    return {"max_absolute_diff": err.max().item(), "mean_absolute_diff": err.mean().item(),
            "relative_error": err.mean().item() / max(ref_mean, 1e-12),
            **benchmark_ring_attention(seq_len, batch_size, hidden_size, num_heads)}
