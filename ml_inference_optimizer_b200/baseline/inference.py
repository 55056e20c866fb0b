"""Mirror of the hot-path slice of the reference's ``baseline/inference.py``.

* ``BlockManager`` / ``SequenceMetadata`` / ``PagedKVCache`` — same classes, constructor arguments and method names as
  the reference (:1045-1302); the physical cache is ``[num_blocks, num_layers, block_size, num_kv_heads, head_dim]``
  (:1077-1084), the layout K2 (``b200_fa_decode``) and ``b200_kv_append`` read and write.
* ``InferenceRunner.run_inference`` — the reference's timing / memory harness with its metric keys (:653-713);
  ``create_inference_runner`` (:1779-1838) returns a runner whose ``_forward`` actually exists (the reference's
  ``TransformerInferenceRunner`` is abstract).
* ``generate_paged`` — a greedy decode loop that drives prefill (K1), KV append and paged decode attention (K2) through
  the attention modules installed by ``kernels.attention.flash_attention.ModelConverter``. Nothing in the reference
  drives its paged kernel; this is that missing loop.
"""
from __future__ import annotations

import logging
import math
import time
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from ..kernels.attention import flash_attention as _fa

__all__ = ["BlockManager", "SequenceMetadata", "PagedKVCache", "InferenceRunner", "BasicInferenceRunner",
           "create_inference_runner", "generate_paged"]


class BlockManager:
    """reference :1045-1127 — free-list allocator with reference counts over the physical blocks."""

    def __init__(self, num_blocks: int, block_size: int, num_layers: int, num_heads: int, head_dim: int, dtype: torch.dtype,
                 device: str):
        self.num_blocks, self.block_size, self.num_layers = num_blocks, block_size, num_layers
        self.num_heads, self.head_dim, self.dtype, self.device = num_heads, head_dim, dtype, device
        self.free_blocks = list(range(num_blocks))
        self.ref_counts = [0] * num_blocks  # host-side (the reference keeps them on the device and syncs on every access)
        shape = (num_blocks, num_layers, block_size, num_heads, head_dim)
        self.gpu_cache_k = torch.zeros(shape, dtype=dtype, device=device)
        self.gpu_cache_v = torch.zeros(shape, dtype=dtype, device=device)
        self.is_initialized = True

    def allocate_block(self) -> int:
        if not self.free_blocks:
            raise MemoryError("Out of memory: No free blocks available in KV cache.")
        idx = self.free_blocks.pop()
        self.ref_counts[idx] = 1
        return idx

    def free_block(self, block_idx: int) -> None:
        if self.ref_counts[block_idx] <= 0:
            logging.warning(f"Attempting to free block {block_idx} with ref count {self.ref_counts[block_idx]}.")
            return
        self.ref_counts[block_idx] -= 1
        if self.ref_counts[block_idx] == 0:
            self.free_blocks.append(block_idx)

    def increase_ref_count(self, block_idx: int) -> None:
        if self.ref_counts[block_idx] <= 0:
            raise ValueError(f"Cannot increase ref count for unallocated block {block_idx}.")
        self.ref_counts[block_idx] += 1

    def get_num_free_blocks(self) -> int:
        return len(self.free_blocks)

    def get_physical_block(self, block_idx: int, layer_idx: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.gpu_cache_k[block_idx, layer_idx], self.gpu_cache_v[block_idx, layer_idx]

    def get_physical_caches(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.gpu_cache_k, self.gpu_cache_v


class SequenceMetadata:
    """reference :1129-1147."""

    def __init__(self, seq_id: int):
        self.seq_id = seq_id
        self.block_table: List[int] = []
        self.logical_len = 0

    def append_block(self, block_idx: int):
        self.block_table.append(block_idx)

    def get_last_block_physical_idx(self) -> Optional[int]:
        return self.block_table[-1] if self.block_table else None

    def __len__(self) -> int:
        return len(self.block_table)


class PagedKVCache:
    """reference :1150-1302."""

    def __init__(self, num_blocks: int, block_size: int, num_layers: int, num_heads: int, head_dim: int,
                 dtype: torch.dtype = torch.float16, device: str = "cuda"):
        self.block_manager = BlockManager(num_blocks, block_size, num_layers, num_heads, head_dim, dtype, device)
        self.block_size, self.num_layers, self.num_heads, self.head_dim = block_size, num_layers, num_heads, head_dim
        self.dtype, self.device = dtype, device
        self.sequences: Dict[int, SequenceMetadata] = {}

    def _ensure_sequence_exists(self, seq_id: int):
        if seq_id not in self.sequences:
            self.sequences[seq_id] = SequenceMetadata(seq_id)

    def _get_logical_block_idx(self, token_pos: int) -> int:
        return token_pos // self.block_size

    def _get_block_offset(self, token_pos: int) -> int:
        return token_pos % self.block_size

    def allocate_blocks_for_sequence(self, seq_id: int, num_tokens: int):
        self._ensure_sequence_exists(seq_id)
        meta = self.sequences[seq_id]
        for _ in range(math.ceil(num_tokens / self.block_size) - len(meta)):
            meta.append_block(self.block_manager.allocate_block())
        meta.logical_len = max(meta.logical_len, num_tokens)

    def reserve_blocks(self, seq_id: int, num_tokens: int) -> None:
        """Allocate the blocks that ``num_tokens`` tokens will need without changing the sequence length, so the block
        table stays fixed while the tokens are appended (what a captured decode graph needs)."""
        self._ensure_sequence_exists(seq_id)
        meta = self.sequences[seq_id]
        for _ in range(math.ceil(num_tokens / self.block_size) - len(meta)):
            meta.append_block(self.block_manager.allocate_block())

    def append_token(self, seq_id: int) -> None:
        self._ensure_sequence_exists(seq_id)
        meta = self.sequences[seq_id]
        new_len = meta.logical_len + 1
        if math.ceil(new_len / self.block_size) > len(meta):
            try:
                meta.append_block(self.block_manager.allocate_block())
            except MemoryError:
                self.free_sequence(seq_id)
                raise
        meta.logical_len = new_len

    def get_block_table(self, seq_id: int) -> List[int]:
        if seq_id not in self.sequences:
            raise ValueError(f"Sequence {seq_id} not found in cache.")
        return self.sequences[seq_id].block_table

    def get_sequence_length(self, seq_id: int) -> int:
        return self.sequences[seq_id].logical_len if seq_id in self.sequences else 0

    def free_sequence(self, seq_id: int) -> None:
        meta = self.sequences.pop(seq_id, None)
        if meta is not None:
            for blk in meta.block_table:
                self.block_manager.free_block(blk)

    def get_physical_caches(self) -> Tuple[torch.Tensor, torch.Tensor]:
        return self.block_manager.get_physical_caches()

    def get_memory_usage(self) -> Dict[str, float]:
        k, v = self.get_physical_caches()
        total = (k.numel() + v.numel()) * k.element_size() / 2 ** 20
        used = self.block_manager.num_blocks - self.block_manager.get_num_free_blocks()
        return {"total_mb": total, "used_blocks": used, "free_blocks": self.block_manager.get_num_free_blocks(),
                "used_mb": total * used / max(1, self.block_manager.num_blocks)}

    # ---- device-side views the kernels consume ----
    def device_tables(self, seq_ids: List[int]) -> Tuple[torch.Tensor, torch.Tensor]:
        """int32 ``block_tables [B, max_blocks]`` and ``context_lengths [B]`` for a batch of sequences."""
        tables = [self.get_block_table(s) for s in seq_ids]
        width = max(1, max(len(t) for t in tables))
        bt = torch.zeros(len(seq_ids), width, dtype=torch.int32)
        for i, t in enumerate(tables):
            bt[i, :len(t)] = torch.tensor(t, dtype=torch.int32)
        lens = torch.tensor([self.get_sequence_length(s) for s in seq_ids], dtype=torch.int32)
        return bt.to(self.device), lens.to(self.device)

    def write_prefill(self, layer_idx: int, seq_ids: List[int], k: torch.Tensor, v: torch.Tensor) -> None:
        """Scatter the prompt's K,V ``[B,S,Hkv,D]`` into the blocks (prefill; the per-token path is ``b200_kv_append``)."""
        kc, vc = self.get_physical_caches()
        B, S = k.shape[:2]
        pos = torch.arange(S)
        for b, sid in enumerate(seq_ids):
            table = torch.tensor(self.get_block_table(sid), dtype=torch.long)
            blk = table[pos // self.block_size].to(k.device)
            off = (pos % self.block_size).to(k.device)
            kc[blk, layer_idx, off] = k[b].to(kc.dtype)
            vc[blk, layer_idx, off] = v[b].to(vc.dtype)


# ------------------------------------------------------------------------------------------------------------------
# runners
# ------------------------------------------------------------------------------------------------------------------
class InferenceRunner:
    """reference :377-789 reduced to its measurement harness: ``run_inference(inputs, **kw) -> (outputs, metrics)`` with
    the reference's metric keys (:653-713)."""

    def __init__(self, model: nn.Module, device: str = "cuda", precision: str = "fp16"):
        self.model, self.device, self.precision = model, device, precision

    def _forward(self, inputs: Any, **kwargs) -> Any:
        raise NotImplementedError

    def run_inference(self, inputs: Any, **kwargs) -> Tuple[Any, Dict[str, float]]:
        metrics: Dict[str, float] = {}
        cuda = torch.cuda.is_available() and str(self.device).startswith("cuda")
        if cuda:
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            metrics["memory_before_mb"] = torch.cuda.memory_allocated() / 2 ** 20
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record()
        t0 = time.perf_counter()
        with torch.no_grad():
            outputs = self._forward(inputs, **kwargs)
        if cuda:
            end.record()
            torch.cuda.synchronize()
            metrics["cuda_time_ms"] = start.elapsed_time(end)
        metrics["total_time_ms"] = (time.perf_counter() - t0) * 1e3
        if cuda:
            metrics["memory_after_mb"] = torch.cuda.memory_allocated() / 2 ** 20
            metrics["peak_memory_mb"] = torch.cuda.max_memory_allocated() / 2 ** 20
            metrics["memory_change_mb"] = metrics["memory_after_mb"] - metrics["memory_before_mb"]
        return outputs, metrics


class BasicInferenceRunner(InferenceRunner):
    """reference :1834-1838 plus the generation branch ``verify_baseline.py:277-288`` expects."""

    def _forward(self, inputs: Any, **kwargs) -> Any:
        gen_keys = {"max_new_tokens", "max_length", "do_sample", "num_beams", "temperature", "top_k", "top_p"}
        if isinstance(inputs, dict):
            if gen_keys & set(kwargs) and hasattr(self.model, "generate"):
                return self.model.generate(**inputs, **kwargs)
            return self.model(**inputs, **kwargs)
        if gen_keys & set(kwargs) and hasattr(self.model, "generate"):
            return self.model.generate(inputs, **kwargs)
        return self.model(inputs, **kwargs)


def create_inference_runner(model: nn.Module, device: str = "cuda", precision: str = "fp16", model_type: str = "transformer",
                            use_flash_attention: bool = False, use_kernel_fusion: bool = False, use_kv_cache: bool = True,
                            use_cuda_graph: bool = False) -> InferenceRunner:
    """reference :1779-1838. ``use_flash_attention`` / ``use_kernel_fusion`` swap in the B200 attention / FusedMLP modules
    (weights copied) before the runner is built."""
    del model_type, use_kv_cache, use_cuda_graph
    if use_flash_attention or use_kernel_fusion:
        from ..optimizer import Optimizer
        model = Optimizer(model).optimize(use_flash_attention=use_flash_attention, use_fused_mlp=use_kernel_fusion)
    return BasicInferenceRunner(model, device, precision)


# ------------------------------------------------------------------------------------------------------------------
# greedy generation over the paged cache
# ------------------------------------------------------------------------------------------------------------------
def _adapters(model: nn.Module) -> List[_fa._HFAttentionAdapter]:
    mods = [m for m in model.modules() if isinstance(m, _fa._HFAttentionAdapter)]
    if not mods:
        raise ValueError("generate_paged needs a model converted by kernels.attention.flash_attention.ModelConverter")
    return mods


def generate_paged(model: nn.Module, input_ids: torch.Tensor, max_new_tokens: int, cache: Optional[PagedKVCache] = None,
                   block_size: int = 16, use_cuda_graph: bool = False) -> torch.Tensor:
    """Greedy generation with the paged KV cache: prefill runs K1 and scatters the prompt's K,V into blocks; every decode
    step appends the new token's K,V (``b200_kv_append``) and attends through the block tables (K2). Returns
    ``[B, S + max_new_tokens]`` token ids.

    ``use_cuda_graph=True`` captures one decode step (all layers: projections, KV append, K2, MLP, LM head, argmax) in a
    CUDA graph and replays it: block tables are reserved up front, lengths / positions / the token are advanced on the
    device, so a step has no host work besides the replay (the loop is launch-bound otherwise)."""
    adapters = _adapters(model)
    first = adapters[0].inner
    B, S = input_ids.shape
    dev = input_ids.device
    if cache is None:
        blocks = B * math.ceil((S + max_new_tokens) / block_size) + 1
        dtype = first.flash_attention._compute_dtype(next(model.parameters()).dtype)
        cache = PagedKVCache(blocks, block_size, len(adapters), first.num_kv_heads, first.head_dim, dtype=dtype, device=str(dev))
    seq_ids = list(range(B))
    for sid in seq_ids:
        cache.allocate_blocks_for_sequence(sid, S)
    out = input_ids
    try:
        with torch.no_grad():
            _fa.set_paged_context({"mode": "prefill", "cache": cache, "seq_ids": seq_ids})
            logits = model(input_ids, use_cache=False).logits
            nxt = logits[:, -1].argmax(-1, keepdim=True)
            out = torch.cat([out, nxt], dim=1)
            if use_cuda_graph and max_new_tokens > 1:
                return torch.cat([out, _decode_with_graph(model, cache, seq_ids, nxt, S, max_new_tokens - 1)], dim=1)
            for step in range(1, max_new_tokens):
                for sid in seq_ids:
                    cache.append_token(sid)
                bt, lens = cache.device_tables(seq_ids)
                _fa.set_paged_context({"mode": "decode", "cache": cache, "seq_ids": seq_ids, "block_tables": bt,
                                       "context_lengths": lens})
                pos = torch.full((B, 1), S + step - 1, dtype=torch.long, device=dev)
                logits = model(nxt, position_ids=pos, use_cache=False).logits
                nxt = logits[:, -1].argmax(-1, keepdim=True)
                out = torch.cat([out, nxt], dim=1)
    finally:
        _fa.set_paged_context(None)
    return out


def _decode_with_graph(model: nn.Module, cache: PagedKVCache, seq_ids: List[int], first_token: torch.Tensor, prompt_len: int,
                       steps: int, eager_steps: int = 2) -> torch.Tensor:
    """``steps`` greedy decode steps after the prefill; returns the generated ids ``[B, steps]``. The first
    ``eager_steps`` run eagerly (they also size the kernel workspaces), then one step is captured and replayed."""
    B = first_token.shape[0]
    dev = first_token.device
    total = prompt_len + steps + 1
    for sid in seq_ids:
        cache.reserve_blocks(sid, total)
    bt, lens = cache.device_tables(seq_ids)            # fixed block tables; lens = prompt length
    tok = first_token.clone()
    pos = torch.full((B, 1), prompt_len, dtype=torch.long, device=dev)
    out_buf = torch.zeros(B, steps, dtype=torch.long, device=dev)
    idx = torch.zeros(1, 1, dtype=torch.long, device=dev)
    _fa.set_paged_context({"mode": "decode", "cache": cache, "seq_ids": seq_ids, "block_tables": bt, "context_lengths": lens,
                           "max_context_len": total})

    def step():
        lens.add_(1)                                   # the token appended in this step counts (attention_kernels.py:862-866)
        logits = model(tok, position_ids=pos, use_cache=False).logits
        n = logits[:, -1].argmax(-1, keepdim=True)
        tok.copy_(n)
        pos.add_(1)
        out_buf.scatter_(1, idx.expand(B, 1), n)
        idx.add_(1)

    # The attention blocks of a converted model mask inside the kernels; HF's "eager" mask builder would still create a
    # dense additive mask per step (with a host scalar -> device copy, which a capture forbids). The "sdpa" mask interface
    # returns None for a single un-padded query token, so it is selected for the duration of the decode loop.
    configs = {id(m.config): m.config for m in model.modules() if hasattr(getattr(m, "config", None), "_attn_implementation")}
    saved_impl = {k: c._attn_implementation for k, c in configs.items()}
    for c in configs.values():
        c._attn_implementation = "sdpa"
    try:
        return _decode_with_graph_body(step, steps, eager_steps, dev, cache, seq_ids, out_buf)
    finally:
        for k, c in configs.items():
            c._attn_implementation = saved_impl[k]


#: timing of the last captured decode loop (device time): {"replay_steps", "replay_ms", "capture_s"}
LAST_DECODE_STATS: Dict[str, float] = {}


def _decode_with_graph_body(step, steps, eager_steps, dev, cache, seq_ids, out_buf):
    done = 0
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):                      # warm-up on the stream the capture will use
        for _ in range(min(eager_steps, steps)):
            step()
            done += 1
    torch.cuda.current_stream(dev).wait_stream(side)
    if done < steps:
        t0 = time.perf_counter()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            step()
        capture_s = time.perf_counter() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps - done):
            graph.replay()
        e1.record()
        e1.synchronize()
        LAST_DECODE_STATS.update(replay_steps=steps - done, replay_ms=e0.elapsed_time(e1), capture_s=capture_s)
    for sid in seq_ids:                                # host bookkeeping of the lengths (blocks are already there)
        for _ in range(steps):
            cache.append_token(sid)
    return out_buf
