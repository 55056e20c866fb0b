"""Host logic of the multi-rank paths on CPU: world_size 2 and 4 over gloo (127.0.0.1). The CUDA kernels cannot run
here, so the per-step attention / local MLP are supplied by the ORACLE through the modules' backend hooks — what is
under test is the distributed bookkeeping: ring schedule, zigzag partition, LSE merging, P2P exchange, column/row
sharding, the single all-reduce and the bias-after-reduce rule."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import attn_mlp_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class OracleRingBackend:
    """Same three methods as parallelism.ring.CudaRingBackend, computed by the fp32 oracle."""

    def attn(self, q, k, v, causal, softmax_scale):
        return orc.attention_ref(q, k, v, causal=causal, softmax_scale=softmax_scale)

    def merge(self, o_acc, lse_acc, o_b, lse_b):
        o, l = orc.lse_merge_ref(o_acc, lse_acc, o_b, lse_b)
        o_acc.copy_(o)
        lse_acc.copy_(l)

    def finalize(self, o_acc, dtype):
        return o_acc.to(dtype)


class OracleFusedRingBackend(OracleRingBackend):
    """The fused-step protocol of CudaRingBackend (``attn_accum``: attention + merge into the accumulator views)."""

    def attn_accum(self, q, k, v, o_acc, lse_acc, init, causal, softmax_scale):
        o, lse = orc.attention_ref(q, k, v, causal=causal, softmax_scale=softmax_scale)
        if init:
            o_acc.copy_(o)
            lse_acc.copy_(lse)
        else:
            mo, ml = orc.lse_merge_ref(o_acc, lse_acc, o, lse)
            o_acc.copy_(mo)
            lse_acc.copy_(ml)


def _worker(rank, world, port, fn_name, return_dict):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        return_dict[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn_name, world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    assert len(ret) == world
    return [ret[r] for r in range(world)]


# ----------------------------------------------------------------------------------------------- workers
def _ring_case(rank, world):
    from parallelism import communication as comm
    from parallelism.ring import ring_attention_forward

    torch.manual_seed(0)  # every rank builds the same full tensors, then takes its shard
    B, S, Hq, Hkv, D = 2, 32 * world, 4, 2, 16
    q, k, v = torch.randn(B, S, Hq, D), torch.randn(B, S, Hkv, D), torch.randn(B, S, Hkv, D)
    errs = {}
    for causal in (False, True):
        full, lse_full = orc.attention_ref(q, k, v, causal=causal)
        for part in ("contiguous", "zigzag"):
            sh = lambda t: comm.scatter_along_sequence_dim(t, world, partition=part, rank=rank).contiguous()
            o, lse = ring_attention_forward(sh(q), sh(k), sh(v), causal=causal, partition=part, backend=OracleRingBackend(),
                                            overlap=False, return_lse=True)
            # the fused-step path (one accumulator, views for partial steps) must give the same shard
            o2, lse2 = ring_attention_forward(sh(q), sh(k), sh(v), causal=causal, partition=part,
                                              backend=OracleFusedRingBackend(), overlap=False, return_lse=True)
            errs[(causal, part, "fused")] = max(float((o2 - o).abs().max()), float((lse2 - lse).abs().max()))
            want = sh(full)
            want_lse = comm.scatter_along_sequence_dim(lse_full.transpose(1, 2), world, partition=part, rank=rank).transpose(1, 2)
            errs[(causal, part)] = (float((o - want).abs().max()), float((lse - want_lse).abs().max()))
            gathered = comm.gather_along_sequence_dim(o.contiguous(), world, partition=part)
            errs[(causal, part, "gather")] = float((gathered - full).abs().max())
    return errs


def _ring_exchange_case(rank, world):
    from parallelism import communication as comm

    a = torch.full((3,), float(rank))
    b = torch.full((2, 2), float(rank) * 10)
    ra, none, rb = comm.ring_exchange(a, None, b)
    prev = (rank - 1) % world
    return bool(none is None and torch.equal(ra, torch.full((3,), float(prev))) and torch.equal(rb, torch.full((2, 2), prev * 10.0)))


def _tp_mlp_case(rank, world):
    import torch.nn.functional as F
    from parallelism import parallel_utils as pu
    from parallelism.tensor_parallel import ColumnParallelLinear, RowParallelLinear, TensorParallelConfig, TensorParallelMLP

    pu.initialize_tensor_parallel(world)
    cfg = TensorParallelConfig(world_size=world, tp_size=world)
    torch.manual_seed(1)
    h, i, T = 32, 64 * world, 10
    x = torch.randn(T, h)
    wu, bu, wg, bg = torch.randn(i, h) * 0.1, torch.randn(i) * 0.1, torch.randn(i, h) * 0.1, torch.randn(i) * 0.1
    wd, bd = torch.randn(h, i) * 0.1, torch.randn(h) * 0.1

    def local_mlp(x_, w_up, b_up, w_down, act, w_gate, b_gate):
        return orc.mlp_ref(x_, w_up, b_up, w_down, None, act, w_gate, b_gate)

    out = {}
    for name, act, gate in (("gelu", F.gelu, None), ("swiglu", F.silu, (wg, bg))):
        m = TensorParallelMLP.from_dense(wu, bu, wd, bd, cfg, act, *(gate or (None, None)))
        assert m.dense_h_to_4h.weight.shape == (i // world, h) and m.dense_4h_to_h.weight.shape == (h, i // world)
        m._local_mlp = local_mlp
        y = m(x)
        ref = orc.mlp_ref(x, wu, bu, wd, bd, "swiglu" if gate else "gelu", *(gate or (None, None)))
        out[name] = float((y - ref).abs().max())
    # column (gathered) then row (scattered input) linear == dense chain
    lin = lambda x_, w, b: torch.nn.functional.linear(x_, w, b)
    col = ColumnParallelLinear(h, i, config=cfg, gather_output=True)
    col.load_full(wu, bu)
    col._local_linear = lin
    row = RowParallelLinear(i, h, config=cfg, input_is_parallel=False)
    row.load_full(wd, bd)
    row._local_linear = lin
    y = row(col(x))
    out["col_row"] = float((y - F.linear(F.linear(x, wu, bu), wd, bd)).abs().max())
    return out


def _sp_attention_module_case(rank, world):
    from parallelism import communication as comm
    from parallelism.sequence_parallel import SequenceParallelAttention, SequenceParallelConfig

    cfg = SequenceParallelConfig(world_size=world, sp_size=world, attention_handling="ring", overlap_communication=False)
    torch.manual_seed(3)
    B, S, h, H = 1, 16 * world, 32, 4
    x = torch.randn(B, S, h)
    out = {}
    for part in ("contiguous", "zigzag"):
        torch.manual_seed(4)
        mod = SequenceParallelAttention(h, H, cfg, attention_dropout=0.0, causal=True, partition=part, backend=OracleRingBackend()).eval()
        xs = comm.scatter_along_sequence_dim(x, world, partition=part, rank=rank).contiguous()
        y = comm.gather_along_sequence_dim(mod(xs).contiguous(), world, partition=part)
        # dense reference with the same weights
        q = mod.query(x).view(B, S, H, -1)
        k = mod.key(x).view(B, S, H, -1)
        v = mod.value(x).view(B, S, H, -1)
        ctx, _ = orc.attention_ref(q, k, v, causal=True)
        ref = mod.output(ctx.reshape(B, S, -1))
        out[part] = float((y - ref).abs().max())
    return out


def _sp_converter_case(rank, world):
    """SequenceParallelConverter.convert_model end to end: deep copy -> attention + MLP swapped (weights copied) -> wrapped in
    SequenceShardedModule, which narrows the full input to this rank's shard and all-gathers the output."""
    import torch.nn as nn

    from parallelism.sequence_parallel import (SequenceParallelAttention, SequenceParallelConfig, SequenceParallelConverter,
                                               SequenceParallelMLP)

    H, hid = 4, 32

    class Attn(nn.Module):
        def __init__(self):
            super().__init__()
            self.num_attention_heads = H
            self.q_proj, self.k_proj, self.v_proj, self.o_proj = (nn.Linear(hid, hid) for _ in range(4))

        def forward(self, x):
            B, S, _ = x.shape
            q, k, v = (l(x).view(B, S, H, hid // H) for l in (self.q_proj, self.k_proj, self.v_proj))
            ctx, _ = orc.attention_ref(q, k, v, causal=True)
            return self.o_proj(ctx.reshape(B, S, hid))

    class MLP(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2, self.act = nn.Linear(hid, 64), nn.Linear(64, hid), nn.ReLU()

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.attn, self.mlp = Attn(), MLP()

        def forward(self, x):
            x = x + self.attn(x)
            return x + self.mlp(x)

    torch.manual_seed(11)            # the same weights and input on every rank
    blk = Block().eval()
    x = torch.randn(2, 16 * world, hid)
    out = {}
    for part in ("contiguous", "zigzag"):
        cfg = SequenceParallelConfig(world_size=world, sp_size=world, attention_handling="ring", overlap_communication=False)
        sp = SequenceParallelConverter(cfg, causal=True, partition=part).convert_model(blk)
        assert isinstance(sp.module.attn, SequenceParallelAttention) and isinstance(sp.module.mlp, SequenceParallelMLP)
        sp.module.attn.backend = OracleRingBackend()
        sp.module.mlp._local_mlp = lambda t, w1, b1, w2, b2, act: orc.mlp_ref(t, w1, b1, w2, b2, act)
        with torch.no_grad():
            out[part] = float((sp(x) - blk(x)).abs().max())
    return out


# ----------------------------------------------------------------------------------------------- tests
@pytest.mark.parametrize("world", [2, 4])
def test_sequence_parallel_converter_end_to_end(world):
    for errs in _spawn("_sp_converter_case", world):
        for key, val in errs.items():
            assert val < 1e-4, (key, val)



@pytest.mark.parametrize("world", [2, 4])
def test_ring_attention_exact(world):
    for errs in _spawn("_ring_case", world):
        for key, val in errs.items():
            worst = max(val) if isinstance(val, tuple) else val
            assert worst < 2e-5, (key, val)


def test_ring_exchange_passes_tensors_and_none():
    assert all(_spawn("_ring_exchange_case", 2))


@pytest.mark.parametrize("world", [2, 4])
def test_tensor_parallel_mlp_and_linears(world):
    for out in _spawn("_tp_mlp_case", world):
        assert all(v < 1e-4 for v in out.values()), out


def test_sequence_parallel_attention_module():
    for out in _spawn("_sp_attention_module_case", 2):
        assert all(v < 1e-4 for v in out.values()), out
