"""Multi-GPU parity check, one rank per GPU over NCCL (launched by tests/test_multi_gpu.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Ring attention (contiguous + zigzag, causal + non-causal, GQA) and tensor-parallel MLP / attention against the fp32
oracle on the full problem."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # B200_MGC_SHARED_GPU=1: every rank runs its kernels on cuda:0 and the ranks talk over gloo (NCCL refuses two ranks
    # on one device) — the same checks on a box with ONE GPU: ring schedule + K1 accumulate steps + TP sharding with the
    # real kernels; only the transport differs (host-staged hops, gloo all-reduce).
    shared = os.environ.get("B200_MGC_SHARED_GPU", "") == "1"
    if shared:
        local = 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if shared:
        dist.init_process_group("gloo")
    else:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    from oracle import attn_mlp_oracle as orc
    from parallelism import communication as comm
    from parallelism import parallel_utils as pu
    from parallelism.ring import ring_attention_forward
    from parallelism.tensor_parallel import TensorParallelAttention, TensorParallelConfig, TensorParallelMLP

    ok = True
    g = torch.Generator().manual_seed(0)
    B, S, Hq, Hkv, D = 2, 256 * world, 8, 2, 128
    q, k, v = torch.randn(B, S, Hq, D, generator=g), torch.randn(B, S, Hkv, D, generator=g), torch.randn(B, S, Hkv, D, generator=g)
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    for causal in (False, True):
        full, lse_full = orc.attention_ref(q, k, v, causal=causal)
        for part in ("contiguous", "zigzag"):
            sh = lambda t: comm.scatter_along_sequence_dim(t, world, partition=part, rank=rank).contiguous().to(dev)
            o, lse = ring_attention_forward(sh(q), sh(k), sh(v), causal=causal, partition=part, return_lse=True)
            want = comm.scatter_along_sequence_dim(full, world, partition=part, rank=rank)
            want_lse = comm.scatter_along_sequence_dim(lse_full.transpose(1, 2), world, partition=part, rank=rank).transpose(1, 2)
            e_o = (o.float().cpu() - want).abs().max().item()
            e_l = (lse.cpu() - want_lse).abs().max().item()
            good = e_o <= 2e-2 and e_l <= 1e-2
            ok &= good
            print(f"[rank {rank}] ring causal={causal} {part}: max|dO|={e_o:.2e} max|dLSE|={e_l:.2e} {'OK' if good else 'FAIL'}", flush=True)
    # tensor-parallel MLP (SwiGLU, Llama-style) and attention
    pu.initialize_tensor_parallel(world)
    cfg = TensorParallelConfig(world_size=world, tp_size=world)
    g = torch.Generator().manual_seed(1)
    h, i, T = 512, 1024 * world, 300
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).bfloat16()
    x, wu, bu, wg, bg, wd, bd = r(T, h), r(i, h, sc=0.03), r(i, sc=0.1), r(i, h, sc=0.03), r(i, sc=0.1), r(h, i, sc=0.03), r(h, sc=0.1)
    c = lambda t: t.to(dev)
    for name, act, gate in (("gelu", F.gelu, (None, None)), ("swiglu", F.silu, (wg, bg))):
        m = TensorParallelMLP.from_dense(c(wu), c(bu), c(wd), c(bd), cfg, act, *(None if t is None else c(t) for t in gate))
        if shared:
            m.reduce_impl = "nccl"  # (= torch.distributed.all_reduce on the group's backend; no symmetric memory over gloo)
        y = m(c(x))
        ref, pabs = orc.tp_mlp_ref(x, wu, bu, wd, bd, "swiglu" if gate[0] is not None else "gelu", world, *gate,
                                   return_partial_abs_sum=True)
        e = (y.float().cpu() - ref).abs().max().item()
        # term-by-term bound (oracle.tp_mlp_tolerance): single-GPU bound + bf16 rounding of every rank's partial + the
        # switch's conversion of the reduced value; no factor fitted to the number of ranks
        good = e <= orc.tp_mlp_tolerance(ref, pabs, world, multicast="multicast" in m.last_reduce)
        ok &= good
        print(f"[rank {rank}] tp mlp {name} ({m.last_reduce}): max|dy|={e:.2e} |ref|max={ref.abs().max().item():.2f} {'OK' if good else 'FAIL'}", flush=True)
    # prefill-sized input: the chunked path that overlaps the all-reduce with the next chunk's GEMMs
    xl = r(8192 + 77, h)
    m = TensorParallelMLP.from_dense(c(wu), c(bu), c(wd), c(bd), cfg, F.silu, c(wg), c(bg))
    if shared:
        m.reduce_impl = "nccl"
    y = m(c(xl))
    ref, pabs = orc.tp_mlp_ref(xl, wu, bu, wd, bd, "swiglu", world, wg, bg, return_partial_abs_sum=True)
    e = (y.float().cpu() - ref).abs().max().item()
    m.overlap_chunks = 1
    y_plain = m(c(xl))  # one fused call + one all-reduce: the chunked pipeline must agree with it up to the reduction order
    e2 = (y.float() - y_plain.float()).abs().max().item()
    # partial sums are rounded to bf16 before the all-reduce (the reference reduces in the activation dtype too)
    # (so two all-reduce schedules differ by the bf16 rounding of `world` partial sums: same scale for both bounds)
    tol = orc.tp_mlp_tolerance(ref, pabs, world, multicast="multicast" in m.last_reduce)
    good = e <= tol and e2 <= tol
    ok &= good
    print(f"[rank {rank}] tp mlp swiglu overlapped (T={xl.shape[0]}): max|dy|={e:.2e} vs single all-reduce {e2:.2e} "
          f"{'OK' if good else 'FAIL'}", flush=True)
    torch.manual_seed(5)  # same weights on every rank, then sharded
    H, Hk, Dh, hid = 8, 2 * world if world <= 4 else 8, 64, 512
    attn = TensorParallelAttention(hid, H, cfg, attention_dropout=0.0, num_kv_heads=Hk, causal=True)
    wq, wk, wv, wo = r(H * Dh, hid, sc=0.03), r(Hk * Dh, hid, sc=0.03), r(Hk * Dh, hid, sc=0.03), r(hid, H * Dh, sc=0.03)
    attn = attn.to(dev, torch.bfloat16)
    attn.query.load_full(c(wq)); attn.key.load_full(c(wk)); attn.value.load_full(c(wv)); attn.output.load_full(c(wo))
    xs = r(2, 200, hid)
    y = attn(c(xs))
    qf, kf, vf = F.linear(xs.float(), wq.float()), F.linear(xs.float(), wk.float()), F.linear(xs.float(), wv.float())
    ctx, _ = orc.attention_ref(qf.view(2, 200, H, Dh), kf.view(2, 200, Hk, Dh), vf.view(2, 200, Hk, Dh), causal=True)
    ref = F.linear(ctx.reshape(2, 200, H * Dh), wo.float())
    e = (y.float().cpu() - ref).abs().max().item()
    good = e <= 3e-2
    ok &= good
    print(f"[rank {rank}] tp attention (Hq={H}, Hkv={Hk}): max|dy|={e:.2e} {'OK' if good else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
