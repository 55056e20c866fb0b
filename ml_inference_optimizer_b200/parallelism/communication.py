"""Mirror of the hot-path slice of the reference's ``parallelism/communication.py``: thin ``torch.distributed``
wrappers (NCCL over NVLink 5 / NVSwitch on the 8xB200 box; gloo on CPU for the host-logic tests) plus the ring
exchange used by ring attention. The topology probing / env tuning of the reference (:886-1630) is out of scope:
inside one NVSwitch domain every peer is one hop at full bandwidth, so there is nothing to search.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

__all__ = ["initialize_distributed", "get_rank", "get_world_size", "all_reduce", "all_gather", "reduce_scatter",
           "broadcast", "scatter", "barrier", "setup_sequence_parallel_group", "scatter_along_sequence_dim",
           "gather_along_sequence_dim", "ring_exchange", "RingExchange"]


def initialize_distributed(local_rank: int, world_size: int, backend: str = "nccl") -> None:
    """reference :12-27. One process per GPU; rendezvous through MASTER_ADDR/PORT (default 127.0.0.1)."""
    if dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    kwargs = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kwargs["device_id"] = torch.device("cuda", local_rank)
        # NCCL kernels must get SMs while an attention / GEMM kernel still has CTAs queued, otherwise the ring hop is
        # not overlapped but appended: run them on a high-priority stream
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        try:
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            kwargs["pg_options"] = opts
        except Exception:  # noqa: BLE001
            pass
    dist.init_process_group(backend=backend, world_size=world_size, rank=int(os.environ.get("RANK", local_rank)), **kwargs)


def get_rank(group: Optional[dist.ProcessGroup] = None) -> int:
    return dist.get_rank(group) if dist.is_initialized() else 0


def get_world_size(group: Optional[dist.ProcessGroup] = None) -> int:
    return dist.get_world_size(group) if dist.is_initialized() else 1


def all_reduce(tensor: torch.Tensor, op=dist.ReduceOp.SUM, async_op: bool = False, group=None, use_fp16: bool = False,
               use_bf16: bool = False, use_unbalanced: bool = False, stream: Optional[torch.cuda.Stream] = None):
    """reference :37-209. In place. ``use_fp16``/``use_bf16`` reduce in the narrower type and cast back (:60-90).
    ``use_unbalanced`` (the hand-rolled tree, :126-176) is accepted and ignored: NCCL picks NVLS/ring itself."""
    del use_unbalanced
    if not dist.is_initialized() or get_world_size(group) == 1:
        return tensor
    comm_dtype = torch.float16 if use_fp16 else (torch.bfloat16 if use_bf16 else None)

    def run():
        if comm_dtype is not None and tensor.dtype != comm_dtype:
            tmp = tensor.to(comm_dtype)
            work = dist.all_reduce(tmp, op=op, group=group, async_op=False)
            tensor.copy_(tmp)
            return work
        return dist.all_reduce(tensor, op=op, group=group, async_op=async_op)

    if stream is not None and tensor.is_cuda:
        stream.wait_stream(torch.cuda.current_stream(tensor.device))
        with torch.cuda.stream(stream):
            work = run()
        return work if async_op else tensor
    work = run()
    return work if async_op else tensor


def all_gather(tensor: torch.Tensor, dim: int = 0, async_op: bool = False, group=None) -> torch.Tensor:
    """reference :211-246 — gather along ``dim`` (list + cat)."""
    del async_op
    ws = get_world_size(group)
    if not dist.is_initialized() or ws == 1:
        return tensor
    parts = [torch.empty_like(tensor) for _ in range(ws)]
    dist.all_gather(parts, tensor.contiguous(), group=group)
    return torch.cat(parts, dim=dim)


def reduce_scatter(tensor: torch.Tensor, dim: int = 0, op=dist.ReduceOp.SUM, group=None) -> torch.Tensor:
    """reference :248-300 — sum across ranks, keep this rank's chunk along ``dim``."""
    ws = get_world_size(group)
    if not dist.is_initialized() or ws == 1:
        return tensor
    chunks = [c.contiguous() for c in tensor.chunk(ws, dim=dim)]
    out = torch.empty_like(chunks[0])
    if tensor.is_cuda:
        dist.reduce_scatter(out, chunks, op=op, group=group)
    else:  # gloo has no reduce_scatter: all-reduce and slice (CPU tests only)
        full = tensor.clone()
        dist.all_reduce(full, op=op, group=group)
        out = full.chunk(ws, dim=dim)[get_rank(group)].contiguous()
    return out


def broadcast(tensor: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    if dist.is_initialized() and get_world_size(group) > 1:
        dist.broadcast(tensor, src=src, group=group)
    return tensor


def scatter(tensor: torch.Tensor, dim: int = 0, group=None) -> torch.Tensor:
    """reference :330-370 — local chunk of a replicated tensor (no communication needed)."""
    ws = get_world_size(group)
    if not dist.is_initialized() or ws == 1:
        return tensor
    return tensor.chunk(ws, dim=dim)[get_rank(group)].contiguous()


def barrier(group=None) -> None:
    if dist.is_initialized():
        dist.barrier(group=group)


_SP_GROUPS: Dict[Tuple[int, int], List[dist.ProcessGroup]] = {}


def setup_sequence_parallel_group(world_size: int, sp_size: int) -> Optional[dist.ProcessGroup]:
    """reference :580-619 — ranks with the same DP index form an SP group. Groups are created ONCE and cached (the
    reference builds new groups on every call, Appendix B)."""
    if not dist.is_initialized():
        raise RuntimeError("Distributed environment not initialized. Call initialize_distributed first.")
    if world_size % sp_size != 0:
        raise ValueError(f"World size ({world_size}) must be divisible by sp_size ({sp_size})")
    key = (world_size, sp_size)
    if key not in _SP_GROUPS:
        if sp_size == world_size:
            _SP_GROUPS[key] = [dist.group.WORLD]
        else:
            _SP_GROUPS[key] = [dist.new_group(ranks=[i * sp_size + j for j in range(sp_size)])
                               for i in range(world_size // sp_size)]
    return _SP_GROUPS[key][get_rank() // sp_size]


def zigzag_chunk_ids(rank: int, sp_size: int) -> Tuple[int, int]:
    """Causal load balancing: the sequence is cut into 2*sp chunks and rank r owns chunks r and 2*sp-1-r."""
    return rank, 2 * sp_size - 1 - rank


def scatter_along_sequence_dim(tensor: torch.Tensor, sp_size: Optional[int] = None, partition: str = "contiguous",
                               rank: Optional[int] = None) -> torch.Tensor:
    """reference :621-661 (contiguous ``narrow``) plus the ``"zigzag"`` partition the exact causal ring uses."""
    ws = get_world_size() if sp_size is None else sp_size
    if ws == 1:
        return tensor
    r = (get_rank() if rank is None else rank) % ws
    S = tensor.size(1)
    if partition == "contiguous":
        if S % ws != 0:
            raise ValueError(f"Sequence length ({S}) must be divisible by sp_size ({ws})")
        return tensor.narrow(1, r * (S // ws), S // ws)
    if partition == "zigzag":
        if S % (2 * ws) != 0:
            raise ValueError(f"Sequence length ({S}) must be divisible by 2*sp_size ({2 * ws}) for the zigzag partition")
        c = S // (2 * ws)
        a, b = zigzag_chunk_ids(r, ws)
        return torch.cat([tensor.narrow(1, a * c, c), tensor.narrow(1, b * c, c)], dim=1)
    raise ValueError(f"unknown partition {partition!r}")


def gather_along_sequence_dim(tensor: torch.Tensor, sp_size: Optional[int] = None, partition: str = "contiguous",
                              group=None) -> torch.Tensor:
    """reference :663-698 — all-gather along dim 1, on the SP group (the reference gathers on WORLD, Appendix B)."""
    ws = get_world_size(group) if sp_size is None else sp_size
    if not dist.is_initialized() or ws == 1:
        return tensor
    parts = [torch.empty_like(tensor) for _ in range(ws)]
    dist.all_gather(parts, tensor.contiguous(), group=group)
    return merge_sequence_shards(parts, partition)


def merge_sequence_shards(parts: Sequence[torch.Tensor], partition: str = "contiguous") -> torch.Tensor:
    """Inverse of ``scatter_along_sequence_dim`` for a list holding every rank's shard (rank order)."""
    if partition == "contiguous":
        return torch.cat(list(parts), dim=1)
    if partition != "zigzag":
        raise ValueError(f"unknown partition {partition!r}")
    ws = len(parts)
    c = parts[0].size(1) // 2
    chunks = [None] * (2 * ws)
    for r, p in enumerate(parts):
        a, b = zigzag_chunk_ids(r, ws)
        chunks[a], chunks[b] = p.narrow(1, 0, c), p.narrow(1, c, c)
    return torch.cat(chunks, dim=1)


class RingExchange:
    """Double-buffered ring hop: post ``isend`` to rank+1 / ``irecv`` from rank-1 for a list of tensors (NCCL P2P over
    NVLink through the switch) on a side stream, so the transfer overlaps the attention tile of the current step.
    Replaces the blocking ``ring_exchange`` of the reference (communication.py:1694-1831)."""

    def __init__(self, group=None, use_side_stream: bool = True):
        self.group = group
        self.rank = get_rank(group)
        self.world = get_world_size(group)
        ranks = dist.get_process_group_ranks(group) if (dist.is_initialized() and group is not None) else None
        self._global = (lambda r: ranks[r]) if ranks is not None else (lambda r: r)
        self.stream = None
        self.use_side_stream = use_side_stream
        self._works = None
        self._event = None
        self.trace = None   # measurement hook: a list that receives one (start_event, end_event) pair per hop (side stream)

    def start(self, send: Sequence[torch.Tensor], recv: Sequence[torch.Tensor]) -> None:
        if self.world == 1:
            for s, r in zip(send, recv):
                r.copy_(s)
            return
        nxt, prv = self._global((self.rank + 1) % self.world), self._global((self.rank - 1) % self.world)
        cuda = send[0].is_cuda
        if cuda and dist.get_backend(self.group) == "gloo":
            # gloo has no device-to-device send/recv: stage the hop through host memory (transport only — used when
            # several ranks share one GPU, e.g. the single-GPU run of tests/multi_gpu_check.py; NCCL ranks never come here)
            host_send = [s.cpu() for s in send]
            self._staged = [(r, torch.empty(r.shape, dtype=r.dtype)) for r in recv]
            ops_ = []
            for s, (_, r) in zip(host_send, self._staged):
                ops_.append(dist.P2POp(dist.isend, s, nxt, group=self.group))
                ops_.append(dist.P2POp(dist.irecv, r, prv, group=self.group))
            self._works = dist.batch_isend_irecv(ops_)
            self._keep = host_send
            return
        if cuda and self.use_side_stream:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=send[0].device, priority=-1)
            self.stream.wait_stream(torch.cuda.current_stream(send[0].device))
            ctx = torch.cuda.stream(self.stream)
        else:
            import contextlib
            ctx = contextlib.nullcontext()
        with ctx:
            t0 = None
            if self.trace is not None and cuda and self.use_side_stream:
                t0 = torch.cuda.Event(enable_timing=True)
                t0.record(self.stream)
            ops_ = []
            for s, r in zip(send, recv):
                ops_.append(dist.P2POp(dist.isend, s, nxt, group=self.group))
                ops_.append(dist.P2POp(dist.irecv, r, prv, group=self.group))
            self._works = dist.batch_isend_irecv(ops_)
            if cuda and self.use_side_stream:
                for w in self._works:
                    w.wait()  # enqueues the completion on the side stream; does not block the host for NCCL
                self._event = torch.cuda.Event(enable_timing=t0 is not None)
                self._event.record(self.stream)
                if t0 is not None:
                    self.trace.append((t0, self._event))

    def wait(self) -> None:
        if self.world == 1 or self._works is None:
            return
        if self._event is not None:
            torch.cuda.current_stream().wait_event(self._event)
            self._event = None
        else:
            for w in self._works:
                w.wait()
        if getattr(self, "_staged", None):
            for dst, host in self._staged:
                dst.copy_(host)
            self._staged, self._keep = None, None
        self._works = None


def ring_exchange(*tensors: torch.Tensor, group=None, async_op: bool = False, use_fp16: bool = False,
                  use_nccl_collectives: bool = True) -> List[torch.Tensor]:
    """reference :1694-1831 (the later definition wins): send every tensor to rank+1 and return what rank-1 sent.
    ``None`` entries pass through (the reference crashes on them, Appendix B)."""
    del async_op, use_nccl_collectives
    live = [t for t in tensors if t is not None]
    if not dist.is_initialized() or get_world_size(group) == 1:
        return list(tensors)
    send = [t.to(torch.float16).contiguous() if use_fp16 and t.is_floating_point() else t.contiguous() for t in live]
    recv = [torch.empty_like(s) for s in send]
    ex = RingExchange(group, use_side_stream=False)
    ex.start(send, recv)
    ex.wait()
    it = iter(r.to(t.dtype) for r, t in zip(recv, live))
    return [None if t is None else next(it) for t in tensors]
