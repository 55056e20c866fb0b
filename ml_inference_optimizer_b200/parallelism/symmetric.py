"""Symmetric (peer-mapped) device buffers for the tensor-parallel all-reduce kernel K6 (``csrc/tp_allreduce.cu``).

The reference reduces the row-parallel partial outputs with ``torch.distributed.all_reduce``
(parallelism/tensor_parallel.py:296-302, parallelism/communication.py:37-209). Here the down projection writes its
partial output into a buffer that every rank of the TP group has mapped (unicast per peer, multicast through the
NVSwitch when available) and ``SymmetricBuffer.all_reduce_`` launches the in-switch two-shot reduction on it.

PyTorch is plumbing only: ``torch.distributed._symmetric_memory`` allocates the buffer and exchanges the handles
(cuMem + fabric/IPC handles + the multicast object); the reduction itself is this repo's kernel.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

__all__ = ["SymmetricBuffer", "symmetric_available"]


def symmetric_available() -> bool:
    try:
        import torch.distributed._symmetric_memory  # noqa: F401
        return dist.is_initialized() and torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


class SymmetricBuffer:
    """``nbytes`` of payload + the flag block of K6, allocated symmetrically on every rank of ``group``.

    Collective: every rank of the group must construct it (same size) and call ``all_reduce_`` in the same order with
    the same arguments."""

    def __init__(self, nbytes: int, group=None, device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm_mem
        from .. import _lib

        if not dist.is_initialized():
            raise RuntimeError("SymmetricBuffer needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if not 2 <= self.world <= 8:
            raise ValueError(f"the tensor-parallel all-reduce kernel supports 2..8 ranks, got {self.world}")
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        self.flag_bytes = int(lib.b200_tp_allreduce_flag_bytes())
        self.payload_bytes = (int(nbytes) + 255) & ~255
        self.flag_offset = self.payload_bytes
        total = self.payload_bytes + self.flag_bytes
        self.buffer = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous did not return a mapping for every peer")
        if ptrs[self.rank] != self.buffer.data_ptr():
            # the library may hand out a sub-allocation: all offsets below are relative to this rank's own mapping
            self._self_delta = self.buffer.data_ptr() - ptrs[self.rank]
        else:
            self._self_delta = 0
        self.peer_ptrs = (ctypes.c_void_p * self.world)(*[p + self._self_delta for p in ptrs])
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if os.environ.get("B200_TP_NO_MULTICAST") == "1":
            mc = 0
        self.multicast_ptr = mc + self._self_delta if mc else 0
        self.buffer.zero_()
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = 1
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # nobody signals before every rank's flags are zero

    @property
    def multicast(self) -> bool:
        return self.multicast_ptr != 0

    def view(self, shape, dtype: torch.dtype, byte_offset: int = 0) -> torch.Tensor:
        """A tensor over the payload (rows written here are what ``all_reduce_`` reduces)."""
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        if byte_offset % 16 or byte_offset + nbytes > self.payload_bytes:
            raise ValueError("view outside the symmetric payload (or not 16-byte aligned)")
        return self.buffer[byte_offset:byte_offset + nbytes].view(dtype).view(*shape)

    def all_reduce_(self, t: torch.Tensor, bias: Optional[torch.Tensor] = None, max_ctas: int = 0) -> torch.Tensor:
        """In-place sum over the group of ``t`` (a contiguous bf16/fp16 view of this buffer), ``bias`` added once to
        every row. Asynchronous on the current stream."""
        from .. import _lib
        from .._lib import DTYPE_BF16, DTYPE_FP16, check

        if not t.is_contiguous():
            raise ValueError("all_reduce_ needs a contiguous view of the symmetric buffer")
        off = t.data_ptr() - self.buffer.data_ptr()
        nbytes = t.numel() * t.element_size()
        if off < 0 or off + nbytes > self.payload_bytes:
            raise ValueError("tensor is not a view of this symmetric buffer")
        if t.dtype == torch.bfloat16:
            dt = DTYPE_BF16
        elif t.dtype == torch.float16:
            dt = DTYPE_FP16
        else:
            raise ValueError(f"all_reduce_ reduces bf16/fp16 partial sums, got {t.dtype}")
        ncols = int(t.shape[-1]) if bias is not None else 0
        if bias is not None and (bias.dtype != t.dtype or bias.numel() != ncols or not bias.is_contiguous()):
            raise ValueError("bias must be a contiguous [cols] tensor of the dtype of t")
        lib = _lib.load()
        with torch.cuda.device(self.device):
            rc = lib.b200_tp_allreduce(self.multicast_ptr or None, self.peer_ptrs, self.world, self.rank, off, nbytes,
                                       self.flag_offset, self.epoch, None if bias is None else bias.data_ptr(), ncols, dt,
                                       int(max_ctas), self.error_flag.data_ptr(),
                                       torch.cuda.current_stream(self.device).cuda_stream)
        check("b200_tp_allreduce", rc)
        self.epoch = (self.epoch + 2) & 0xFFFFFFFF
        return t

    def check(self) -> None:
        """Raise if a previous ``all_reduce_`` timed out waiting for a peer (synchronises)."""
        if int(self.error_flag.item()) != 0:
            raise RuntimeError("tensor-parallel all-reduce timed out waiting for a peer rank")
