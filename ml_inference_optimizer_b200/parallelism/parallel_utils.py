"""Mirror of the helpers of ``parallelism/parallel_utils.py`` the TP / SP modules use (divide / split / gather and the
tensor-parallel group globals; reference :28-215, :882-1000)."""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.distributed as dist

__all__ = ["ensure_divisibility", "divide", "is_power_of_two", "split_tensor_along_dim", "gather_tensor_along_dim",
           "initialize_tensor_parallel", "get_tensor_model_parallel_group", "get_tensor_model_parallel_rank",
           "get_tensor_model_parallel_world_size", "get_partition_start_end", "create_attention_mask_for_tp",
           "split_tensor_into_1d_equal_chunks", "gather_1d_tensor_chunks", "set_tensor_model_parallel_attributes",
           "copy_tensor_model_parallel_attributes", "get_parallel_tensor_info"]

_TP_GROUP: Optional[dist.ProcessGroup] = None
_TP_SIZE: int = 1


def ensure_divisibility(numerator: int, denominator: int) -> None:
    if numerator % denominator != 0:
        raise ValueError(f"{numerator} is not divisible by {denominator}")


def divide(numerator: int, denominator: int) -> int:
    """reference :28-45."""
    ensure_divisibility(numerator, denominator)
    return numerator // denominator


def is_power_of_two(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


def split_tensor_along_dim(tensor: torch.Tensor, dim: int, num_partitions: Optional[int] = None,
                           contiguous_split_chunks: bool = False, world_size: Optional[int] = None) -> List[torch.Tensor]:
    """reference :137-174 (``num_partitions`` is the keyword the reference's own callers pass, tensor_parallel.py:293)."""
    n = num_partitions or world_size or get_tensor_model_parallel_world_size()
    size = divide(tensor.size(dim), n)
    chunks = torch.split(tensor, size, dim=dim)
    return [c.contiguous() for c in chunks] if contiguous_split_chunks else list(chunks)


def gather_tensor_along_dim(tensor: torch.Tensor, dim: int, group=None) -> torch.Tensor:
    """reference :176-215."""
    group = group if group is not None else get_tensor_model_parallel_group()
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    if ws == 1:
        return tensor
    parts = [torch.empty_like(tensor) for _ in range(ws)]
    dist.all_gather(parts, tensor.contiguous(), group=group)
    return torch.cat(parts, dim=dim)


def initialize_tensor_parallel(tp_size: int) -> Optional[dist.ProcessGroup]:
    """reference :882-1000 (grid of groups) reduced to what the hot path needs: consecutive ranks form a TP group."""
    global _TP_GROUP, _TP_SIZE
    if not dist.is_initialized() or tp_size == 1:
        _TP_GROUP, _TP_SIZE = None, 1
        return None
    ws = dist.get_world_size()
    ensure_divisibility(ws, tp_size)
    if tp_size == ws:
        _TP_GROUP = dist.group.WORLD
    else:
        me = dist.get_rank() // tp_size
        for i in range(ws // tp_size):
            g = dist.new_group(ranks=list(range(i * tp_size, (i + 1) * tp_size)))
            if i == me:
                _TP_GROUP = g
    _TP_SIZE = tp_size
    return _TP_GROUP


def get_tensor_model_parallel_group():
    return _TP_GROUP


def get_tensor_model_parallel_world_size() -> int:
    if not dist.is_initialized():
        return 1
    return dist.get_world_size(_TP_GROUP) if _TP_GROUP is not None else (_TP_SIZE if _TP_SIZE > 1 else 1)


def get_tensor_model_parallel_rank() -> int:
    if not dist.is_initialized() or _TP_GROUP is None:
        return 0
    return dist.get_rank(_TP_GROUP)


def get_partition_start_end(size: int, rank: int, world_size: int):
    per = divide(size, world_size)
    return rank * per, (rank + 1) * per


def create_attention_mask_for_tp(attention_mask: Optional[torch.Tensor], tp_size: int) -> Optional[torch.Tensor]:
    """reference :217-286 — this rank's block of the KEY (last) dimension of a mask, whatever its rank (the reference's three
    branches all slice the last dimension into ``tp_size`` equal parts)."""
    if tp_size == 1 or attention_mask is None:
        return attention_mask
    start, end = get_partition_start_end(attention_mask.size(-1), get_tensor_model_parallel_rank(), tp_size)
    return attention_mask[..., start:end]


def split_tensor_into_1d_equal_chunks(tensor: torch.Tensor, group_size: Optional[int] = None) -> torch.Tensor:
    """reference :426-452 — this rank's contiguous slice of the flattened tensor (a copy)."""
    group_size = get_tensor_model_parallel_world_size() if group_size is None else group_size
    start, end = get_partition_start_end(tensor.numel(), get_tensor_model_parallel_rank() if group_size > 1 else 0, group_size)
    return tensor.reshape(-1)[start:end].clone()


def gather_1d_tensor_chunks(tensor: torch.Tensor, tensor_shape: torch.Size, group_size: Optional[int] = None) -> torch.Tensor:
    """reference :455-488 — inverse of the split above: all-gather over the TP group, reshaped to ``tensor_shape``."""
    group_size = get_tensor_model_parallel_world_size() if group_size is None else group_size
    if group_size == 1:
        return tensor.reshape(tensor_shape)
    parts = [torch.empty_like(tensor) for _ in range(group_size)]
    dist.all_gather(parts, tensor.contiguous(), group=get_tensor_model_parallel_group())
    return torch.cat(parts, dim=0).reshape(tensor_shape)


_TP_ATTRS = ("is_tensor_parallel", "tensor_parallel_dim", "tensor_parallel_stride")


def set_tensor_model_parallel_attributes(tensor: torch.Tensor, is_parallel: bool, dim: int, stride: int) -> torch.Tensor:
    """reference :491-513 — tags a (parameter) tensor with how it is sharded."""
    for name, value in zip(_TP_ATTRS, (is_parallel, dim, stride)):
        setattr(tensor, name, value)
    return tensor


def copy_tensor_model_parallel_attributes(destination_tensor: torch.Tensor, source_tensor: torch.Tensor) -> None:
    """reference :516-533."""
    for name in _TP_ATTRS:
        if hasattr(source_tensor, name):
            setattr(destination_tensor, name, getattr(source_tensor, name))


def get_parallel_tensor_info(tensor: torch.Tensor) -> Dict[str, Any]:
    """reference :536-556."""
    return {"is_parallel": getattr(tensor, "is_tensor_parallel", False), "parallel_dim": getattr(tensor, "tensor_parallel_dim", None),
            "parallel_stride": getattr(tensor, "tensor_parallel_stride", None), "shape": tensor.shape, "dtype": tensor.dtype,
            "device": tensor.device}
