"""Exact ring attention across ranks (K4): K1 per ring step + LSE merge + double-buffered NCCL P2P of the KV blocks.

Rebuilds ``SequenceParallelAttention._ring_attention`` (reference parallelism/sequence_parallel.py:519-585), which
sums per-step softmaxes and divides by ``sp_size`` (not attention), as exact attention:
every rank keeps its query shard, the (K, V) shards travel around the ring, each step's partial result is merged
with the running (O fp32, LSE) by the log-sum-exp rule of kernels/triton/attention_kernels.py:1567-1585.

Overlap: the hop for step s+1 is posted (``isend`` to rank+1 / ``irecv`` from rank-1 on a side stream, NVLink P2P
through the switch) BEFORE the attention kernel of step s is launched, into the other half of a double buffer.

Causal load balance: with the reference's contiguous split (communication.py:651-659) rank r only has work in r+1
of the n steps. ``partition="zigzag"`` (rank r owns chunks r and 2n-1-r of 2n) gives every rank the same work in every
step: a full block at step 0 (plain local causal) and exactly half a block afterwards —
  src <  r : all local queries  x  first half of the visiting keys   (no mask)
  src >  r : second half of the local queries  x  all visiting keys  (no mask)
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from .communication import RingExchange, get_rank, get_world_size

__all__ = ["ring_attention_forward", "CudaRingBackend"]


class CudaRingBackend:
    """The product backend: sm_100a kernels through the C-ABI. (Tests of the host logic on CPU/gloo inject their own
    backend with the same three methods.)"""

    def attn(self, q, k, v, causal: bool, softmax_scale: Optional[float]):
        from .. import ops
        return ops.flash_attn_fwd(q, k, v, causal=causal, softmax_scale=softmax_scale, return_lse=True)

    def attn_accum(self, q, k, v, o_acc, lse_acc, init: bool, causal: bool, softmax_scale: Optional[float]) -> None:
        """One ring step fused: K1 with the log-sum-exp merge into (o_acc fp32, lse_acc) in its epilogue."""
        from .. import ops
        ops.flash_attn_fwd_accum(q, k, v, o_acc, lse_acc, init, causal=causal, softmax_scale=softmax_scale)

    def merge(self, o_acc, lse_acc, o_b, lse_b) -> None:
        from .. import ops
        ops.lse_merge(o_acc, lse_acc, o_b, lse_b)

    def finalize(self, o_acc, dtype):
        from .. import ops
        return ops.cast_out(o_acc, dtype)


class _Acc:
    """fp32 running output + LSE for a contiguous block of query rows."""

    def __init__(self, backend, B, S, H, D, device):
        self.backend = backend
        self.o = None
        self.lse = None
        self.shape = (B, S, H, D)
        self.device = device

    def update(self, o_b, lse_b):
        if self.o is None:
            self.o = o_b.float().contiguous()
            self.lse = lse_b.float().contiguous().clone()
        else:
            self.backend.merge(self.o, self.lse, o_b, lse_b.contiguous())

    def result(self, dtype):
        B, S, H, D = self.shape
        if self.o is None:  # no visible key at all
            return torch.zeros(B, S, H, D, dtype=dtype, device=self.device), \
                torch.full((B, H, S), float("-inf"), dtype=torch.float32, device=self.device)
        return self.backend.finalize(self.o, dtype), self.lse


def _ring_fused(backend, q, k, v, causal, softmax_scale, group, zigzag, overlap, return_lse, n, r, trace=None):
    """Ring with the merge fused into the attention kernel (``backend.attn_accum``): one launch per step, one fp32
    accumulator for all local query rows, no 16-bit partial outputs, no separate merge / slice / concatenate kernels.
    Step 0 is always the rank's own block (src == r), which touches every local query row, so it initialises the
    accumulator; every later step merges. Steps that only concern a part of the rows (zigzag, src > r) are given the
    views of that part."""
    B, S, Hq, D = q.shape
    half = S // 2
    kv = [torch.stack([k, v]).contiguous(), None]
    kv[1] = torch.empty_like(kv[0])
    exchange = RingExchange(group, use_side_stream=overlap)
    if trace is not None:   # measurement hook (tests/ring_timeline.py): CUDA events around every kernel and every hop
        exchange.trace = trace.setdefault("hops", [])
        trace["attn"] = []
    o_acc = torch.empty(B, S, Hq, D, dtype=torch.float32, device=q.device)
    lse_acc = torch.empty(B, Hq, S, dtype=torch.float32, device=q.device)
    for step in range(n):
        cur = kv[step % 2]
        if step + 1 < n:
            exchange.start([cur], [kv[(step + 1) % 2]])  # overlaps with the attention below
        src = (r - step) % n
        kc, vc = cur[0], cur[1]
        init = step == 0
        if trace is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), src)
            ev[0].record()
        if not causal:
            backend.attn_accum(q, kc, vc, o_acc, lse_acc, init, False, softmax_scale)
        elif src == r:
            backend.attn_accum(q, kc, vc, o_acc, lse_acc, init, True, softmax_scale)
        elif zigzag:
            if src < r:   # all local queries x first half of the visiting keys
                backend.attn_accum(q, kc[:, :half], vc[:, :half], o_acc, lse_acc, init, False, softmax_scale)
            else:         # second half of the local queries x all visiting keys
                backend.attn_accum(q[:, half:], kc, vc, o_acc[:, half:], lse_acc[:, :, half:], init, False, softmax_scale)
        elif src < r:     # causal, contiguous shards: earlier ranks are fully visible, later ranks not at all
            backend.attn_accum(q, kc, vc, o_acc, lse_acc, init, False, softmax_scale)
        if trace is not None:
            ev[1].record()
            trace["attn"].append(ev)
        if step + 1 < n:
            exchange.wait()
    out = backend.finalize(o_acc, q.dtype)
    return (out, lse_acc) if return_lse else out


def ring_attention_forward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False,
                           softmax_scale: Optional[float] = None, group=None, partition: str = "contiguous",
                           backend=None, overlap: bool = True, return_lse: bool = False, trace: Optional[dict] = None):
    """q ``[B,S_local,Hq,D]``, k/v ``[B,S_local,Hkv,D]`` — this rank's shard under ``partition``. Returns the local
    output shard ``[B,S_local,Hq,D]`` (and LSE ``[B,Hq,S_local]``)."""
    if partition not in ("contiguous", "zigzag"):
        raise ValueError(f"unknown partition {partition!r}")
    backend = backend or CudaRingBackend()
    # group=None means the WORLD group (as in torch.distributed). Callers with "no sequence-parallel group" must not get
    # here: SequenceParallelAttention routes sp_size == 1 to the local kernel.
    n = get_world_size(group) if dist.is_initialized() else 1
    r = get_rank(group) if dist.is_initialized() else 0
    B, S, Hq, D = q.shape
    if n == 1:
        o, lse = backend.attn(q, k, v, causal, softmax_scale)
        return (o, lse) if return_lse else o
    zigzag = causal and partition == "zigzag"
    if partition == "zigzag" and S % 2 != 0:
        raise ValueError("the zigzag partition needs an even local sequence length")
    half = S // 2
    if hasattr(backend, "attn_accum"):
        return _ring_fused(backend, q, k, v, causal, softmax_scale, group, zigzag, overlap, return_lse, n, r, trace)

    # KV double buffer: cur is attended to while nxt is being received
    kv = [torch.stack([k, v]).contiguous(), None]
    kv[1] = torch.empty_like(kv[0])
    exchange = RingExchange(group, use_side_stream=overlap)

    if zigzag:
        acc_lo = _Acc(backend, B, half, Hq, D, q.device)
        acc_hi = _Acc(backend, B, half, Hq, D, q.device)
        q_hi = q[:, half:]
    else:
        acc = _Acc(backend, B, S, Hq, D, q.device)

    for step in range(n):
        cur = kv[step % 2]
        if step + 1 < n:
            exchange.start([cur], [kv[(step + 1) % 2]])  # overlaps with the attention below
        src = (r - step) % n
        kc, vc = cur[0], cur[1]
        if not causal:
            o, lse = backend.attn(q, kc, vc, False, softmax_scale)
            acc.update(o, lse)
        elif zigzag:
            if src == r:
                o, lse = backend.attn(q, kc, vc, True, softmax_scale)
                acc_lo.update(o[:, :half], lse[:, :, :half])
                acc_hi.update(o[:, half:], lse[:, :, half:])
            elif src < r:
                o, lse = backend.attn(q, kc[:, :half], vc[:, :half], False, softmax_scale)
                acc_lo.update(o[:, :half], lse[:, :, :half])
                acc_hi.update(o[:, half:], lse[:, :, half:])
            else:
                o, lse = backend.attn(q_hi, kc, vc, False, softmax_scale)
                acc_hi.update(o, lse)
        else:  # causal, contiguous shards: earlier ranks are fully visible, later ranks not at all
            if src == r:
                o, lse = backend.attn(q, kc, vc, True, softmax_scale)
                acc.update(o, lse)
            elif src < r:
                o, lse = backend.attn(q, kc, vc, False, softmax_scale)
                acc.update(o, lse)
        if step + 1 < n:
            exchange.wait()

    if zigzag:
        o_lo, l_lo = acc_lo.result(q.dtype)
        o_hi, l_hi = acc_hi.result(q.dtype)
        out = torch.cat([o_lo, o_hi], dim=1)
        lse = torch.cat([l_lo, l_hi], dim=2)
    else:
        out, lse = acc.result(q.dtype)
    return (out, lse) if return_lse else out
