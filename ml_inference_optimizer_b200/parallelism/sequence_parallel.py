"""Mirror of the ring / sequence-parallel slice of the reference's ``parallelism/sequence_parallel.py``.

``SequenceParallelAttention`` keeps the constructor, attribute names (``query``, ``key``, ``value``, ``output``) and
``forward(hidden_states[B,S/sp,h], attention_mask)`` of the reference (:345-640); ``attention_handling="ring"`` is the
exact, overlapped ring of ``parallelism/ring.py`` (the reference's is an approximation, F7), ``"local"`` attends to
the local shard only (:480-517) and ``"full"`` all-gathers K,V (:587-640).
Additions (keyword-only, defaults preserve the reference behaviour): ``causal`` and ``partition`` ("contiguous" |
"zigzag").
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import communication as comm
from .communication import get_rank
from .ring import CudaRingBackend, ring_attention_forward

__all__ = ["SequenceParallelConfig", "SequenceParallelAttention", "SequenceParallelMLP", "SequenceShardedModule",
           "SequenceParallelConverter", "partition_sequence", "gather_sequence", "create_sequence_parallel_attention_mask"]


@dataclass
class SequenceParallelConfig:
    """reference :22-85."""
    world_size: int = 1
    sp_size: int = 1
    overlap_communication: bool = True
    attention_handling: str = "ring"
    chunk_size: Optional[int] = None
    buffer_reuse: bool = True
    communication_dtype: torch.dtype = torch.float16

    def __post_init__(self):
        if self.world_size % self.sp_size != 0:
            raise ValueError(f"Sequence parallel size ({self.sp_size}) must divide world size ({self.world_size})")
        if self.attention_handling not in ["local", "ring", "full"]:
            raise ValueError(f"Attention handling strategy '{self.attention_handling}' not supported. "
                             f"Use 'local', 'ring', or 'full'.")
        if self.chunk_size is not None and self.chunk_size <= 0:
            raise ValueError(f"Chunk size must be positive, got {self.chunk_size}")

    def get_sp_group(self) -> Optional[dist.ProcessGroup]:
        if not dist.is_initialized() or self.sp_size == 1:
            return None
        return comm.setup_sequence_parallel_group(self.world_size, self.sp_size)  # cached, not re-created per call

    def get_dp_size(self) -> int:
        return self.world_size // self.sp_size

    def get_rank_info(self) -> Tuple[int, int]:
        rank = get_rank()
        return rank % self.sp_size, rank // self.sp_size


def partition_sequence(tensor: torch.Tensor, config: SequenceParallelConfig, partition: str = "contiguous") -> torch.Tensor:
    sp_rank, _ = config.get_rank_info()
    return comm.scatter_along_sequence_dim(tensor, config.sp_size, partition=partition, rank=sp_rank)


def gather_sequence(tensor: torch.Tensor, config: SequenceParallelConfig, partition: str = "contiguous") -> torch.Tensor:
    return comm.gather_along_sequence_dim(tensor, config.sp_size, partition=partition, group=config.get_sp_group())


def create_sequence_parallel_attention_mask(attention_mask: Optional[torch.Tensor], config: SequenceParallelConfig):
    """Key-padding masks are lowered to lengths inside the attention module; additive 4-D masks are not supported on the
    CUDA path, so this returns the mask untouched for the caller to pass on (reference :882-940 slices it per rank)."""
    return attention_mask


class SequenceParallelAttention(nn.Module):
    """reference :345-640."""

    def __init__(self, hidden_size: int, num_attention_heads: int, config: SequenceParallelConfig,
                 attention_dropout: float = 0.1, head_dim: Optional[int] = None, bias: bool = True, *,
                 causal: bool = False, partition: str = "contiguous", backend=None):
        super().__init__()
        self.config = config
        self.hidden_size = hidden_size
        self.num_attention_heads = num_attention_heads
        self.head_dim = head_dim if head_dim is not None else hidden_size // num_attention_heads
        self.all_head_size = self.num_attention_heads * self.head_dim
        self.query = nn.Linear(hidden_size, self.all_head_size, bias=bias)
        self.key = nn.Linear(hidden_size, self.all_head_size, bias=bias)
        self.value = nn.Linear(hidden_size, self.all_head_size, bias=bias)
        self.output = nn.Linear(self.all_head_size, hidden_size, bias=bias)
        self.dropout = nn.Dropout(attention_dropout)
        self.causal = causal
        self.partition = partition
        self.backend = backend or CudaRingBackend()
        self.sp_group = config.get_sp_group()
        self.sp_rank, self.dp_rank = config.get_rank_info()
        self.attention_impl = self._select_attention_impl()
        self._reset_parameters()

    def _reset_parameters(self):
        for lin in (self.query, self.key, self.value, self.output):
            nn.init.xavier_uniform_(lin.weight)
            if lin.bias is not None:
                nn.init.zeros_(lin.bias)

    def _select_attention_impl(self) -> Callable:
        return {"local": self._local_attention, "ring": self._ring_attention, "full": self._full_attention}[
            self.config.attention_handling]

    def _transpose_for_scores(self, x: torch.Tensor) -> torch.Tensor:
        """[B,S,all_head] -> [B,H,S,D] (a view; the kernels take the strides)."""
        return x.view(*x.size()[:-1], self.num_attention_heads, self.head_dim).permute(0, 2, 1, 3)

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("additive attention masks are not supported on the ring path; use causal=True")
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("attention dropout is not implemented (inference path)")
        q = self._transpose_for_scores(self.query(hidden_states))
        k = self._transpose_for_scores(self.key(hidden_states))
        v = self._transpose_for_scores(self.value(hidden_states))
        ctx = self.attention_impl(q=q, k=k, v=v, attention_mask=None)  # [B,H,S,D]
        ctx = ctx.permute(0, 2, 1, 3)
        ctx = ctx.reshape(*ctx.size()[:-2], self.all_head_size)
        return self.output(ctx)

    # all three take/return [B,H,S/sp,D] like the reference
    def _local_attention(self, q, k, v, attention_mask=None):
        o, _ = self.backend.attn(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), self.causal, None)
        return o.transpose(1, 2)

    def _ring_attention(self, q, k, v, attention_mask=None):
        if self.config.sp_size == 1:
            # no SP group: get_sp_group() is None, which the ring would read as WORLD and circulate K/V through the
            # data-parallel replicas, mixing unrelated batches. One rank's sequence is the whole sequence.
            return self._local_attention(q, k, v, attention_mask)
        o = ring_attention_forward(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=self.causal,
                                   group=self.sp_group, partition=self.partition, backend=self.backend,
                                   overlap=self.config.overlap_communication)
        return o.transpose(1, 2)

    def _full_attention(self, q, k, v, attention_mask=None):
        if self.config.sp_size == 1:
            return self._local_attention(q, k, v, attention_mask)
        kf = comm.gather_along_sequence_dim(k.transpose(1, 2).contiguous(), self.config.sp_size, self.partition, self.sp_group)
        vf = comm.gather_along_sequence_dim(v.transpose(1, 2).contiguous(), self.config.sp_size, self.partition, self.sp_group)
        if self.causal:
            if self.partition != "contiguous":
                raise NotImplementedError("full (all-gather) causal attention needs the contiguous partition")
            s_local = q.shape[2]
            o, _ = self._attn_offset(q.transpose(1, 2), kf, vf, self.sp_rank * s_local)
        else:
            o, _ = self.backend.attn(q.transpose(1, 2), kf, vf, False, None)
        return o.transpose(1, 2)

    def _attn_offset(self, q, k, v, offset):
        from .. import ops
        return ops.flash_attn_fwd(q, k, v, causal=True, causal_offset=offset, return_lse=True)


class SequenceParallelMLP(nn.Module):
    """reference :643-720 — token-wise, so every rank runs the fused MLP on its own sequence shard; no communication."""

    def __init__(self, hidden_size: int, intermediate_size: int, config: SequenceParallelConfig,
                 activation: Callable = F.gelu, dropout_prob: float = 0.1, bias: bool = True):
        super().__init__()
        self.config = config
        self.dense_h_to_4h = nn.Linear(hidden_size, intermediate_size, bias=bias)
        self.dense_4h_to_h = nn.Linear(intermediate_size, hidden_size, bias=bias)
        self.activation = activation
        self.dropout = nn.Dropout(dropout_prob)
        self.sp_group = config.get_sp_group()
        self.sp_rank, self.dp_rank = config.get_rank_info()
        self._local_mlp = None  # test hook (CPU/gloo host-logic tests)
        nn.init.xavier_uniform_(self.dense_h_to_4h.weight)
        nn.init.xavier_uniform_(self.dense_4h_to_h.weight)
        if bias:
            nn.init.zeros_(self.dense_h_to_4h.bias)
            nn.init.zeros_(self.dense_4h_to_h.bias)

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        from .tensor_parallel import activation_name
        from .. import ops
        if self.training and self.dropout.p > 0:
            raise NotImplementedError("dropout is not implemented (inference path)")
        if self._local_mlp is not None:
            return self._local_mlp(hidden_states, self.dense_h_to_4h.weight, self.dense_h_to_4h.bias, self.dense_4h_to_h.weight,
                                   self.dense_4h_to_h.bias, activation_name(self.activation))
        return ops.fused_mlp(hidden_states, self.dense_h_to_4h.weight, self.dense_h_to_4h.bias, self.dense_4h_to_h.weight,
                             self.dense_4h_to_h.bias, activation_name(self.activation))


class SequenceShardedModule(nn.Module):
    """reference :88-343 reduced to its data path: narrow ``[B,S,h]`` to this rank's shard, run the wrapped module,
    all-gather the result along the sequence on the SP group."""

    def __init__(self, module: nn.Module, config: SequenceParallelConfig, partition: str = "contiguous"):
        super().__init__()
        self.module = module
        self.config = config
        self.partition = partition

    def optimize_for_inference(self) -> None:
        """reference :326-342 — fp32 parameters are cast to the configured 16-bit communication dtype (these kernels compute
        in 16 bit anyway) and the wrapped module's own hook, if any, is called."""
        if self.config.communication_dtype in (torch.float16, torch.bfloat16):
            for p in self.parameters():
                if p.dtype == torch.float32:
                    p.data = p.data.to(self.config.communication_dtype)
        if hasattr(self.module, "optimize_for_inference"):
            self.module.optimize_for_inference()

    def forward(self, hidden_states: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        local = partition_sequence(hidden_states, self.config, self.partition)
        out = self.module(local.contiguous(), *args, **kwargs)
        if isinstance(out, tuple):
            out = out[0]
        return gather_sequence(out, self.config, self.partition)


class SequenceParallelConverter:
    """reference :723-879 — swaps attention modules that expose ``query/key/value/output`` or ``q_proj/k_proj/v_proj/
    o_proj`` Linears for ``SequenceParallelAttention`` and COPIES their weights (the reference creates fresh random
    modules, Appendix B)."""

    def __init__(self, config: SequenceParallelConfig, causal: bool = False, partition: str = "contiguous"):
        self.config = config
        self.causal = causal
        self.partition = partition

    def convert_model(self, model: nn.Module) -> nn.Module:
        """reference :734-755 — a deep copy with attention and MLP layers swapped, wrapped in ``SequenceShardedModule`` (narrow
        the input to this rank's shard, all-gather the output along the sequence)."""
        import copy

        converted = self.convert_mlp_layers(self.convert_attention_layers(copy.deepcopy(model)))
        return SequenceShardedModule(converted, self.config, self.partition)

    def convert_attention_layers(self, model: nn.Module) -> nn.Module:
        """reference :757-810 — in place; every attention block with four Linears becomes ``SequenceParallelAttention``."""
        for name, sub in list(model.named_children()):
            new = self._convert(sub)
            if new is not None:
                setattr(model, name, new)
            else:
                self.convert_attention_layers(sub)
        return model

    def convert_mlp_layers(self, model: nn.Module) -> nn.Module:
        """reference :812-878 — in place; two-Linear feed-forward blocks (``fc1/fc2``, ``dense_h_to_4h/dense_4h_to_h``, ``wi/wo``,
        GPT-2 ``c_fc/c_proj``) become ``SequenceParallelMLP`` with the weights copied and the block's own activation. Gated
        (three-Linear) blocks are token-wise too and run unchanged on the shard; ``MLPConverter`` fuses those."""
        for name, sub in list(model.named_children()):
            new = self._convert_mlp(sub)
            if new is not None:
                setattr(model, name, new)
            else:
                self.convert_mlp_layers(sub)
        return model

    def partition_input_data(self, inputs: Dict[str, torch.Tensor]) -> List[Dict[str, torch.Tensor]]:
        """reference :880-908 — one input dictionary per SP rank: the sequence tensors (``input_ids``, ``attention_mask``,
        ``token_type_ids``) narrowed to that rank's shard, everything else passed through."""
        seq_keys = ("input_ids", "attention_mask", "token_type_ids")
        return [{k: (comm.scatter_along_sequence_dim(v, self.config.sp_size, partition=self.partition, rank=r).contiguous()
                     if k in seq_keys else v) for k, v in inputs.items()} for r in range(self.config.sp_size)]

    def gather_output_data(self, outputs: List[torch.Tensor]) -> torch.Tensor:
        """reference :910-920 — inverse of the partition above for a list holding every rank's output."""
        return comm.merge_sequence_shards(list(outputs), partition=self.partition)

    def _convert_mlp(self, m: nn.Module) -> Optional[nn.Module]:
        if isinstance(m, (SequenceParallelMLP, SequenceParallelAttention, SequenceShardedModule)):
            return None
        if any(hasattr(m, a) for a in ("gate_proj", "fc1_gate", "w3")):
            return None
        pair, transposed = None, False
        for a, b in (("fc1", "fc2"), ("dense_h_to_4h", "dense_4h_to_h"), ("wi", "wo"), ("c_fc", "c_proj")):
            if hasattr(m, a) and hasattr(m, b):
                pair = (getattr(m, a), getattr(m, b))
                transposed = not isinstance(pair[0], nn.Linear)   # GPT-2 Conv1D keeps [in, out]
                break
        if pair is None or not all(hasattr(l, "weight") and l.weight.dim() == 2 for l in pair):
            return None
        w1 = pair[0].weight.t() if transposed else pair[0].weight
        w2 = pair[1].weight.t() if transposed else pair[1].weight
        if w2.shape != (w1.shape[1], w1.shape[0]):
            return None
        from ..kernels.mlp.fused_mlp import MLPConverter, resolve_activation
        act = MLPConverter._module_activation(m)
        resolve_activation(act)   # strict: a block without an activation attribute, or with one that has no fused epilogue, raises
        b1, b2 = getattr(pair[0], "bias", None), getattr(pair[1], "bias", None)
        new = SequenceParallelMLP(w1.shape[1], w1.shape[0], self.config, activation=act, dropout_prob=0.0,
                                  bias=b1 is not None and b2 is not None)
        with torch.no_grad():
            new.dense_h_to_4h.weight.copy_(w1)
            new.dense_4h_to_h.weight.copy_(w2)
            if new.dense_h_to_4h.bias is not None:
                new.dense_h_to_4h.bias.copy_(b1)
                new.dense_4h_to_h.bias.copy_(b2)
        return new.to(device=w1.device, dtype=w1.dtype)

    def _convert(self, m: nn.Module) -> Optional[nn.Module]:
        if isinstance(m, SequenceParallelAttention):
            return None
        names = None
        if all(hasattr(m, a) for a in ("query", "key", "value", "output")):
            names = ("query", "key", "value", "output")
        elif all(hasattr(m, a) for a in ("q_proj", "k_proj", "v_proj", "o_proj")):
            names = ("q_proj", "k_proj", "v_proj", "o_proj")
        if names is None or not all(isinstance(getattr(m, a), nn.Linear) for a in names):
            return None
        q = getattr(m, names[0])
        heads = getattr(m, "num_attention_heads", None) or getattr(m, "num_heads", None)
        if heads is None or getattr(m, names[1]).out_features != q.out_features:
            return None
        new = SequenceParallelAttention(q.in_features, heads, self.config, attention_dropout=0.0,
                                        head_dim=q.out_features // heads, bias=q.bias is not None, causal=self.causal,
                                        partition=self.partition)
        with torch.no_grad():
            for dst, src in zip((new.query, new.key, new.value, new.output), (getattr(m, a) for a in names)):
                dst.weight.copy_(src.weight)
                if src.bias is not None:
                    dst.bias.copy_(src.bias)
        return new.to(device=q.weight.device, dtype=q.weight.dtype)
