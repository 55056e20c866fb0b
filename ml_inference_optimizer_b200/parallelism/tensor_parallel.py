"""Mirror of the hot-path slice of the reference's ``parallelism/tensor_parallel.py`` (Megatron-style column / row
split): ``ColumnParallelLinear`` (:88-204), ``RowParallelLinear`` (:207-327), ``TensorParallelMLP`` (:330-400) and
``TensorParallelAttention`` (:403-599), on the sm_100a kernels.

``TensorParallelMLP`` runs K3 on the rank's shard — up/gate GEMM with the activation fused in the epilogue, then the
down GEMM over the rank's columns — followed by ONE all-reduce (NCCL; NVLS in-switch reduction on the NVSwitch box)
and the bias added once after the reduction (reference :302-308). ``TensorParallelAttention`` runs K1 on Hq/tp query
heads and Hkv/tp KV heads; its row-parallel output projection carries the only all-reduce.
Reference defects not reproduced (Appendix B): K/V projections aliased to Q (:481-483), converters that drop the
weights (:684-690), the ``num_partitions`` keyword mismatch (:293).
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import communication as comm
from . import parallel_utils as pu

__all__ = ["TensorParallelConfig", "ColumnParallelLinear", "RowParallelLinear", "TensorParallelMLP",
           "TensorParallelAttention", "ModelParallelConverter", "activation_name"]


def activation_name(fn) -> str:
    """Map the reference's activation callables (``F.gelu`` default, :336) to the fused-epilogue selector — strictly
    (``kernels.mlp.fused_mlp.resolve_activation``): an activation without a fused epilogue raises."""
    from ..kernels.mlp.fused_mlp import resolve_activation
    name = resolve_activation(fn)
    return "swiglu" if name == "silu" else name


class TensorParallelConfig:
    """reference :16-85. ``get_tp_group`` returns the group made by ``parallel_utils.initialize_tensor_parallel``
    (the reference returns None)."""

    def __init__(self, world_size: int = 1, tp_size: int = 1, dp_size: Optional[int] = None, parallel_dim: int = -1,
                 gather_output: bool = True, recompute_activation: bool = False,
                 communication_dtype: torch.dtype = torch.float16, sequence_parallel: bool = False,
                 gradient_accumulation_steps: int = 1, use_cpu_initialization: bool = False):
        self.world_size = world_size
        self.tp_size = tp_size
        if dp_size is None:
            assert world_size % tp_size == 0, "World size must be divisible by tensor parallel size"
            self.dp_size = world_size // tp_size
        else:
            self.dp_size = dp_size
            assert world_size == tp_size * dp_size, "World size must equal tp_size * dp_size"
        self.parallel_dim = parallel_dim
        self.gather_output = gather_output
        self.recompute_activation = recompute_activation
        self.communication_dtype = communication_dtype
        self.sequence_parallel = sequence_parallel
        self.gradient_accumulation_steps = gradient_accumulation_steps
        self.use_cpu_initialization = use_cpu_initialization

    def get_tp_group(self):
        if not dist.is_initialized() or self.tp_size == 1:
            return None
        g = pu.get_tensor_model_parallel_group()
        return g if g is not None else pu.initialize_tensor_parallel(self.tp_size)

    def get_dp_group(self):
        return None

    def tp_rank(self) -> int:
        g = self.get_tp_group()
        return dist.get_rank(g) if g is not None else 0


def _linear(x, w, b, local_linear=None):
    if local_linear is not None:
        return local_linear(x, w, b)
    from .. import ops
    return ops.linear_act(x, w, b)


def _reset_linear_shard(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> None:
    nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    if bias is not None:
        fan_in = weight.shape[1]
        nn.init.uniform_(bias, -1.0 / math.sqrt(fan_in), 1.0 / math.sqrt(fan_in))


class ColumnParallelLinear(nn.Module):
    """reference :88-204 — weight ``[out/tp, in]``; optional all-gather of the output."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, config: Optional[TensorParallelConfig] = None,
                 gather_output: Optional[bool] = None, stride: int = 1, skip_bias_add: bool = False):
        super().__init__()
        self.config = config or TensorParallelConfig()
        self.in_features, self.out_features = in_features, out_features
        self.gather_output = self.config.gather_output if gather_output is None else gather_output
        self.skip_bias_add = skip_bias_add
        self.tp_size = self.config.tp_size
        self.output_size_per_partition = pu.divide(out_features, self.tp_size)
        self.weight = nn.Parameter(torch.empty(self.output_size_per_partition, in_features))
        self.bias = nn.Parameter(torch.zeros(self.output_size_per_partition)) if bias else None
        nn.init.normal_(self.weight, std=0.02)
        self._local_linear = None  # test hook (CPU/gloo host-logic tests)

    def load_full(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None):
        r = self.config.tp_rank()
        s = self.output_size_per_partition
        with torch.no_grad():
            self.weight.copy_(weight[r * s:(r + 1) * s])
            if self.bias is not None and bias is not None:
                self.bias.copy_(bias[r * s:(r + 1) * s])

    def reset_parameters(self):
        """reference :152-159 — nn.Linear's default initialisation of the shard (Kaiming-uniform weight, fan-in bias)."""
        _reset_linear_shard(self.weight, self.bias)

    def get_master_weight(self) -> torch.Tensor:
        """reference :189-204 — the unsharded ``[out, in]`` weight (all-gather of the row blocks over the TP group)."""
        if self.tp_size == 1:
            return self.weight
        return pu.gather_tensor_along_dim(self.weight.detach(), 0, self.config.get_tp_group())

    def forward(self, x: torch.Tensor):
        b = None if self.skip_bias_add else self.bias
        y = _linear(x, self.weight, b, self._local_linear)
        if self.gather_output and self.tp_size > 1:
            y = pu.gather_tensor_along_dim(y, -1, self.config.get_tp_group())
        return (y, self.bias) if self.skip_bias_add else y


class RowParallelLinear(nn.Module):
    """reference :207-327 — weight ``[out, in/tp]``; all-reduce(sum) of the partial outputs, bias added after (:302-308)."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, config: Optional[TensorParallelConfig] = None,
                 input_is_parallel: bool = False, stride: int = 1, skip_bias_add: bool = False):
        super().__init__()
        self.config = config or TensorParallelConfig()
        self.in_features, self.out_features = in_features, out_features
        self.input_is_parallel = input_is_parallel
        self.skip_bias_add = skip_bias_add
        self.tp_size = self.config.tp_size
        self.input_size_per_partition = pu.divide(in_features, self.tp_size)
        self.weight = nn.Parameter(torch.empty(out_features, self.input_size_per_partition))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        nn.init.normal_(self.weight, std=0.02)
        self._local_linear = None

    def load_full(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None):
        r = self.config.tp_rank()
        s = self.input_size_per_partition
        with torch.no_grad():
            self.weight.copy_(weight[:, r * s:(r + 1) * s])
            if self.bias is not None and bias is not None:
                self.bias.copy_(bias)

    def reset_parameters(self):
        """reference :271-278."""
        _reset_linear_shard(self.weight, self.bias)

    def get_master_weight(self) -> torch.Tensor:
        """reference :312-327 — the unsharded ``[out, in]`` weight (all-gather of the column blocks over the TP group)."""
        if self.tp_size == 1:
            return self.weight
        return pu.gather_tensor_along_dim(self.weight.detach(), 1, self.config.get_tp_group())

    def forward(self, x: torch.Tensor):
        if not self.input_is_parallel and self.tp_size > 1:
            x = pu.split_tensor_along_dim(x, -1, num_partitions=self.tp_size)[self.config.tp_rank()].contiguous()
        y = _linear(x, self.weight, None, self._local_linear)
        if self.tp_size > 1:
            comm.all_reduce(y, group=self.config.get_tp_group())
        if self.skip_bias_add:
            return y, self.bias
        return y if self.bias is None else y + self.bias


class TensorParallelMLP(nn.Module):
    """reference :330-400 — ``dense_h_to_4h`` column-parallel (no gather) -> activation -> ``dense_4h_to_h`` row-parallel.
    ``gated=True`` adds the SwiGLU gate projection ``dense_h_to_4h_gate`` (Llama shapes, C4)."""

    def __init__(self, hidden_size: int, intermediate_size: int, config: Optional[TensorParallelConfig] = None,
                 activation: Callable = F.gelu, *, gated: bool = False):
        super().__init__()
        self.config = config or TensorParallelConfig()
        self.dense_h_to_4h = ColumnParallelLinear(hidden_size, intermediate_size, bias=True, config=self.config, gather_output=False)
        self.dense_4h_to_h = RowParallelLinear(intermediate_size, hidden_size, bias=True, config=self.config, input_is_parallel=True)
        self.dense_h_to_4h_gate = ColumnParallelLinear(hidden_size, intermediate_size, bias=True, config=self.config,
                                                       gather_output=False) if gated else None
        self.activation = activation
        self.gated = gated
        self._local_mlp = None  # test hook: fn(x, w_up, b_up, w_down, act, w_gate, b_gate) -> partial output
        # overlap of the all-reduce with the GEMMs of the next token chunk (prefill-sized inputs only)
        self.overlap_chunks = 4
        self.overlap_min_tokens = 8192
        self.comm_sms = 32  # SMs left free for the collective while chunks are in flight (measured at tp=8: 2.02 -> 1.90 ms)
        # "auto": K6 over symmetric memory when it can be set up, else NCCL; "symmetric": K6 or raise; "nccl": NCCL only
        self.reduce_impl = os.environ.get("B200_TP_REDUCE", "auto")
        # K6 CTAs are small enough (256 threads, <= 42 registers, no shared memory) to sit on an SM next to a resident GEMM
        # CTA: by default the GEMMs keep every SM (gemm_sm_reserve = 0) and K6 runs underneath them, one CTA per SM
        # (measured at tp=8, C3: 4 chunks x 64 CTAs 1.11 ms, x 148 CTAs 1.40 ms, NCCL 1.71 ms, GEMMs alone 0.78 ms)
        self.comm_ctas = int(os.environ.get("B200_TP_COMM_CTAS", "64"))           # 0 = one per SM
        self.comm_ctas_single = int(os.environ.get("B200_TP_COMM_CTAS_SINGLE", "0"))
        self.gemm_sm_reserve = int(os.environ.get("B200_TP_GEMM_SM_RESERVE", "0"))  # SMs withheld from the GEMMs while K6 runs
        # "copy": return a fresh tensor (safe default). "view": return the rows inside the symmetric buffer — valid
        # until the second-next TensorParallelMLP call on this group (two buffers rotate); saves one pass over [T, h].
        self.symmetric_output = "copy"
        self.last_reduce = "nccl"

    @classmethod
    def from_dense(cls, w_up, b_up, w_down, b_down, config: TensorParallelConfig, activation: Callable = F.gelu,
                   w_gate=None, b_gate=None) -> "TensorParallelMLP":
        """Shard full (replicated) weights: rows of W_up / W_gate, columns of W_down (tensor_parallel.py:130-135, :249-254)."""
        i, h = w_up.shape
        m = cls(h, i, config, activation, gated=w_gate is not None).to(device=w_up.device, dtype=w_up.dtype)
        m.dense_h_to_4h.load_full(w_up, b_up)
        m.dense_4h_to_h.load_full(w_down, b_down)
        if w_gate is not None:
            m.dense_h_to_4h_gate.load_full(w_gate, b_gate)
        return m

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        act = "swiglu" if self.gated else activation_name(self.activation)
        up, down, gate = self.dense_h_to_4h, self.dense_4h_to_h, self.dense_h_to_4h_gate
        gw, gb = (gate.weight, gate.bias) if gate is not None else (None, None)
        if self._local_mlp is not None:
            partial = self._local_mlp(hidden_states, up.weight, up.bias, down.weight, act, gw, gb)
        else:
            from .. import ops
            lead = hidden_states.shape[:-1]
            x2 = hidden_states.reshape(-1, hidden_states.shape[-1])
            T = x2.shape[0]
            if self.config.tp_size > 1 and dist.is_initialized() and x2.is_cuda:
                if self.reduce_impl in ("auto", "symmetric") and T > 0:
                    pool = self._symmetric_pool(T * down.out_features * x2.element_size(), x2.device)
                    if pool is not None:
                        return self._forward_symmetric(x2, act, gw, gb, pool).reshape(*lead, -1)
                if self.overlap_chunks > 1 and T >= self.overlap_min_tokens:
                    return self._forward_overlapped(x2, act, gw, gb).reshape(*lead, -1)
            # the down bias must be added once, after the reduction (reference :304-308)
            partial = ops.fused_mlp(hidden_states, up.weight, up.bias, down.weight, None, act, gw, gb)
        if self.config.tp_size > 1:
            comm.all_reduce(partial, group=self.config.get_tp_group())
        return partial if down.bias is None else partial + down.bias

    # ---- symmetric-memory path: K3 writes its partial rows into a peer-mapped buffer, K6 reduces them in the switch ----
    _POOLS: dict = {}

    def _symmetric_pool(self, nbytes: int, device):
        """Two rotating symmetric buffers shared by every TensorParallelMLP of the group (created collectively on first
        use, grown when a larger input arrives). None if symmetric memory cannot be set up (then NCCL reduces)."""
        from .symmetric import SymmetricBuffer, symmetric_available
        group = self.config.get_tp_group()
        key = (id(group), device.index)
        entry = TensorParallelMLP._POOLS.get(key)
        if entry is not None and entry["failed"]:
            if self.reduce_impl == "symmetric":
                raise RuntimeError(f"symmetric memory is not available: {entry['failed']}")
            return None
        if entry is None or entry["nbytes"] < nbytes:
            try:
                if not symmetric_available():
                    raise RuntimeError("torch.distributed._symmetric_memory is not importable")
                bufs = [SymmetricBuffer(nbytes, group, device) for _ in range(2)]
                entry = {"nbytes": nbytes, "bufs": bufs, "turn": 0, "failed": None, "stream": torch.cuda.Stream(device, priority=-1)}
            except Exception as e:  # noqa: BLE001 - no fabric / IPC support: the NCCL path stays correct
                entry = {"nbytes": 0, "bufs": [], "turn": 0, "failed": f"{type(e).__name__}: {e}", "stream": None}
                TensorParallelMLP._POOLS[key] = entry
                if self.reduce_impl == "symmetric":
                    raise
                return None
            TensorParallelMLP._POOLS[key] = entry
        return entry

    def _forward_symmetric(self, x2: torch.Tensor, act: str, gw, gb, pool) -> torch.Tensor:
        """Token-chunked pipeline on two streams: the GEMMs of chunk c+1 (K3, on all SMs but ``comm_ctas``) run while K6
        reduces chunk c through the NVSwitch (multimem.ld_reduce / multimem.st; bias fused). Same arithmetic as one fused
        call + one all-reduce + bias, with the partial sums accumulated in fp32 inside the switch."""
        from .. import ops
        up, down = self.dense_h_to_4h, self.dense_4h_to_h
        T, h_out = x2.shape[0], down.out_features
        buf = pool["bufs"][pool["turn"]]
        pool["turn"] ^= 1
        out = buf.view((T, h_out), x2.dtype)
        main = torch.cuda.current_stream(x2.device)
        side = pool["stream"]
        n = self.overlap_chunks if T >= self.overlap_min_tokens else 1
        rows = ((T + n - 1) // n + 255) // 256 * 256
        bias = down.bias
        reserve = self.gemm_sm_reserve if n > 1 else 0
        if reserve > 0:
            ops.set_sm_limit(max(2, ops.sm_count(x2.device) - reserve))
        try:
            for r0 in range(0, T, rows):
                r1 = min(T, r0 + rows)
                ops.fused_mlp(x2[r0:r1], up.weight, up.bias, down.weight, None, act, gw, gb, out=out[r0:r1])
                if n > 1:
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        buf.all_reduce_(out[r0:r1], bias, max_ctas=self.comm_ctas)
                else:
                    buf.all_reduce_(out[r0:r1], bias, max_ctas=self.comm_ctas_single)
        finally:
            if reserve > 0:
                ops.set_sm_limit(0)
        if n > 1:
            main.wait_stream(side)
        self.last_reduce = "symmetric-multicast" if buf.multicast else "symmetric-peer"
        return out if self.symmetric_output == "view" else out.clone()

    def _forward_overlapped(self, x2: torch.Tensor, act: str, gw, gb) -> torch.Tensor:
        """Token-chunked pipeline: while NCCL reduces chunk c (on its own high-priority stream, using the SMs the GEMMs
        leave free), the GEMMs of chunk c+1 run. Same arithmetic as one fused call followed by one all-reduce."""
        from .. import ops
        up, down = self.dense_h_to_4h, self.dense_4h_to_h
        T = x2.shape[0]
        out = torch.empty(T, down.out_features, dtype=x2.dtype, device=x2.device)
        group = self.config.get_tp_group()
        n = self.overlap_chunks
        rows = ((T + n - 1) // n + 127) // 128 * 128
        works = []
        ops.set_sm_limit(max(1, ops.sm_count(x2.device) - self.comm_sms))
        try:
            for r0 in range(0, T, rows):
                r1 = min(T, r0 + rows)
                ops.fused_mlp(x2[r0:r1], up.weight, up.bias, down.weight, None, act, gw, gb, out=out[r0:r1])
                works.append(dist.all_reduce(out[r0:r1], group=group, async_op=True))
        finally:
            ops.set_sm_limit(0)
        for w in works:
            w.wait()
        if down.bias is not None:
            out += down.bias
        return out


class TensorParallelAttention(nn.Module):
    """reference :403-599 — heads split across ranks (requires ``Hq % tp == 0``, :447-448), attention on the local heads,
    row-parallel output projection (one all-reduce). ``num_kv_heads`` adds GQA (tp=8 with 8 KV heads -> 1 per rank)."""

    def __init__(self, hidden_size: int, num_attention_heads: int, config: Optional[TensorParallelConfig] = None,
                 attention_dropout: float = 0.1, is_cross_attention: bool = False, head_dim: Optional[int] = None, *,
                 num_kv_heads: Optional[int] = None, causal: bool = False):
        super().__init__()
        self.config = config or TensorParallelConfig()
        tp = self.config.tp_size
        self.hidden_size = hidden_size
        self.num_attention_heads = num_attention_heads
        self.num_kv_heads = num_kv_heads or num_attention_heads
        self.head_dim = head_dim or hidden_size // num_attention_heads
        if num_attention_heads % tp != 0 or self.num_kv_heads % tp != 0:
            raise ValueError(f"num_attention_heads ({num_attention_heads}) and num_kv_heads ({self.num_kv_heads}) must be "
                             f"divisible by tp_size ({tp})")
        self.heads_per_rank = num_attention_heads // tp
        self.kv_heads_per_rank = self.num_kv_heads // tp
        self.is_cross_attention = is_cross_attention
        self.causal = causal
        self.query = ColumnParallelLinear(hidden_size, num_attention_heads * self.head_dim, config=self.config, gather_output=False)
        self.key = ColumnParallelLinear(hidden_size, self.num_kv_heads * self.head_dim, config=self.config, gather_output=False)
        self.value = ColumnParallelLinear(hidden_size, self.num_kv_heads * self.head_dim, config=self.config, gather_output=False)
        self.output = RowParallelLinear(num_attention_heads * self.head_dim, hidden_size, config=self.config, input_is_parallel=True)
        self.dropout = nn.Dropout(attention_dropout)
        self.communication_schedule_optimized = False
        self._local_attn = None  # test hook

    # reference attribute names of the per-rank head split (:449-451)
    @property
    def num_heads_per_partition(self) -> int:
        return self.heads_per_rank

    @property
    def attention_head_size(self) -> int:
        return self.head_dim

    def transpose_for_scores(self, x: torch.Tensor) -> torch.Tensor:
        """reference :497-509 — ``[B, S, heads_per_rank * D]`` -> ``[B, heads_per_rank, S, D]`` (a view; K1 itself reads the
        ``[B, S, H, D]`` form through strides, so ``forward`` does not need it)."""
        return x.view(*x.shape[:-1], -1, self.head_dim).permute(0, 2, 1, 3)

    def optimize_communication_schedule(self, inference_only: bool = True) -> None:
        """reference :601-614 (a flag there). The one collective of this block is the all-reduce behind the output
        projection; nothing is left to re-schedule for a single attention call, so this records the request only."""
        self.communication_schedule_optimized = True

    def forward(self, hidden_states: torch.Tensor, attention_mask: Optional[torch.Tensor] = None,
                encoder_hidden_states: Optional[torch.Tensor] = None) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("additive attention masks are not supported on the CUDA path; use causal=True")
        B, S, _ = hidden_states.shape
        kv_src = encoder_hidden_states if (self.is_cross_attention and encoder_hidden_states is not None) else hidden_states
        q = self.query(hidden_states).view(B, S, self.heads_per_rank, self.head_dim)
        k = self.key(kv_src).view(B, kv_src.shape[1], self.kv_heads_per_rank, self.head_dim)
        v = self.value(kv_src).view(B, kv_src.shape[1], self.kv_heads_per_rank, self.head_dim)
        if self._local_attn is not None:
            ctx = self._local_attn(q, k, v, self.causal)
        else:
            from .. import ops
            ctx = ops.flash_attn_fwd(q, k, v, causal=self.causal)
        return self.output(ctx.reshape(B, S, self.heads_per_rank * self.head_dim))


class ModelParallelConverter:
    """reference :617-798 — swaps ``*.mlp`` blocks made of two Linears (or gate/up/down) for ``TensorParallelMLP``,
    sharding and COPYING the weights."""

    def __init__(self, config: Optional[TensorParallelConfig] = None):
        self.config = config or TensorParallelConfig()

    def convert_to_column_parallel(self, linear: nn.Linear) -> ColumnParallelLinear:
        """reference :729-764 — this rank's row block of ``linear`` (weight and bias copied)."""
        col = ColumnParallelLinear(linear.in_features, linear.out_features, bias=linear.bias is not None, config=self.config)
        col = col.to(device=linear.weight.device, dtype=linear.weight.dtype)
        col.load_full(linear.weight.detach(), None if linear.bias is None else linear.bias.detach())
        return col

    def convert_to_row_parallel(self, linear: nn.Linear) -> RowParallelLinear:
        """reference :766-798 — this rank's column block of ``linear``; the bias stays whole (added after the all-reduce)."""
        row = RowParallelLinear(linear.in_features, linear.out_features, bias=linear.bias is not None, config=self.config)
        row = row.to(device=linear.weight.device, dtype=linear.weight.dtype)
        row.load_full(linear.weight.detach(), None if linear.bias is None else linear.bias.detach())
        return row

    def distribute_model(self, model: nn.Module) -> Dict[int, nn.Module]:
        """reference :800-816 — converts in place; one process per GPU, so every rank of the map is this process's model."""
        converted = self.convert_model(model)
        return {rank: converted for rank in range(self.config.world_size)}

    def convert_model(self, model: nn.Module) -> nn.Module:
        for name, sub in list(model.named_children()):
            new = self._convert_module(sub)
            if new is not None:
                setattr(model, name, new)
            else:
                self.convert_model(sub)
        return model

    def _convert_module(self, m: nn.Module) -> Optional[nn.Module]:
        if isinstance(m, (TensorParallelMLP, ColumnParallelLinear, RowParallelLinear)):
            return None
        if all(hasattr(m, a) for a in ("gate_proj", "up_proj", "down_proj")):
            from ..kernels.mlp.fused_mlp import MLPConverter, resolve_activation
            if resolve_activation(MLPConverter._module_activation(m)) != "silu":
                raise ValueError(f"{type(m).__name__}: only SiLU-gated (SwiGLU) MLPs have a fused tensor-parallel epilogue")
            return TensorParallelMLP.from_dense(m.up_proj.weight, m.up_proj.bias, m.down_proj.weight, m.down_proj.bias,
                                                self.config, F.silu, m.gate_proj.weight, m.gate_proj.bias)
        if hasattr(m, "fc1") and hasattr(m, "fc2") and isinstance(m.fc1, nn.Linear):
            from ..kernels.mlp.fused_mlp import MLPConverter
            act = MLPConverter._module_activation(m)
            activation_name(act)   # strict: no activation attribute / no fused epilogue raises instead of assuming GELU
            return TensorParallelMLP.from_dense(m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, self.config, act)
        if hasattr(m, "c_fc") and hasattr(m, "c_proj"):  # GPT-2 Conv1D, weight [in, out]
            from ..kernels.mlp.fused_mlp import MLPConverter
            act = MLPConverter._module_activation(m)
            activation_name(act)
            return TensorParallelMLP.from_dense(m.c_fc.weight.t().contiguous(), m.c_fc.bias, m.c_proj.weight.t().contiguous(),
                                                m.c_proj.bias, self.config, act)
        return None
