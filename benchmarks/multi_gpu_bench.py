"""Multi-GPU benchmarks of the two places the hot path shards (one rank per GPU, NCCL over NVLink 5 / NVSwitch):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        benchmarks/multi_gpu_bench.py [ring] [tp] [--seq 131072]

  ring : BASELINE configs[4] — causal attention over one 128K-token sequence (32x128 heads, bf16), sequence-sharded,
         KV blocks circulating the ring overlapped with the local tile; zigzag and contiguous partitions;
         efficiency = T_1 / (N * T_N) needs the N=1 time (printed by running with N=1).
  tp   : BASELINE configs[3] — Llama-3-8B shapes (GQA 32q/8kv, SwiGLU 14336): column/row tensor-parallel FusedMLP with
         one all-reduce, and head-parallel attention, prefill T=32768 and decode T=64.
Prints one JSON line per measurement (rank 0); times are CUDA-event device times, max over ranks.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
import torch.nn.functional as F


def timed(fn, warmup, iters, dev):
    for _ in range(warmup):
        fn()
    if dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / iters], device=dev, dtype=torch.float64)
    if dist.is_initialized():
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["ring", "tp"])
    ap.add_argument("--seq", type=int, default=131072)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--quick", action="store_true", help="ring: zigzag + overlap only")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    from ml_inference_optimizer_b200 import ops
    from parallelism import parallel_utils as pu
    from parallelism.ring import ring_attention_forward
    from parallelism.tensor_parallel import TensorParallelConfig, TensorParallelMLP

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    if "ring" in args.what:
        S, H, D = args.seq, 32, 128
        Sl = S // world
        g = torch.Generator(device=dev).manual_seed(rank)
        q, k, v = (torch.randn(1, Sl, H, D, device=dev, dtype=torch.bfloat16, generator=g) for _ in range(3))
        flops = 4.0 * H * S * S * D * 0.5
        for part in (("zigzag",) if args.quick else ("zigzag", "contiguous")) if world > 1 else ("contiguous",):
            for overlap in ((True,) if args.quick else (True, False)) if world > 1 else (True,):
                ms = timed(lambda: ring_attention_forward(q, k, v, causal=True, partition=part, overlap=overlap), args.warmup,
                           args.iters, dev)
                emit({"bench": "ring_attention_causal", "seq": S, "heads": H, "head_dim": D, "n_gpus": world, "partition": part,
                      "overlap": overlap, "ms": ms, "tflops_total": flops / ms / 1e9, "tflops_per_gpu": flops / ms / 1e9 / world,
                      "kv_bytes_per_hop": 2 * Sl * H * D * 2})
    if "tp" in args.what:
        h, i, Hq, Hkv, D = 4096, 14336, 32, 8, 128
        if world > 1:
            pu.initialize_tensor_parallel(world)
        cfg = TensorParallelConfig(world_size=world, tp_size=world)
        mlp = TensorParallelMLP(h, i, cfg, F.silu, gated=True).to(dev, torch.bfloat16)
        if os.environ.get("B200_TP_COMM_SMS"):
            mlp.comm_sms = int(os.environ["B200_TP_COMM_SMS"])
        if os.environ.get("B200_TP_CHUNKS"):
            mlp.overlap_chunks = int(os.environ["B200_TP_CHUNKS"])
        for T, tag in ((32768, "prefill"),) if args.quick else ((32768, "prefill"), (64, "decode")):
            x = torch.randn(T, h, device=dev, dtype=torch.bfloat16)
            ms = timed(lambda: mlp(x), args.warmup, args.iters, dev)
            flops = 6.0 * T * h * i
            emit({"bench": f"tp_mlp_swiglu_{tag}", "T": T, "hidden": h, "intermediate": i, "n_gpus": world, "ms": ms,
                  "tflops_total": flops / ms / 1e9, "allreduce_bytes": T * h * 2, "overlap_chunks": mlp.overlap_chunks,
                  "comm_sms": mlp.comm_sms, "nccl_max_nchannels": os.environ.get("NCCL_MAX_NCHANNELS")})
        if args.quick:
            if world > 1:
                dist.destroy_process_group()
            return
        # head-parallel attention: Hq/tp query heads, Hkv/tp KV heads per rank (no collective inside attention)
        B, S = 4, 8192
        q = torch.randn(B, S, Hq // world, D, device=dev, dtype=torch.bfloat16)
        k = torch.randn(B, S, max(1, Hkv // world), D, device=dev, dtype=torch.bfloat16)
        v = torch.randn_like(k)
        ms = timed(lambda: ops.flash_attn_fwd(q, k, v, causal=True), args.warmup, args.iters, dev)
        flops = 4.0 * B * Hq * S * S * D * 0.5
        emit({"bench": "tp_attention_prefill", "B": B, "S": S, "Hq": Hq, "Hkv": Hkv, "n_gpus": world, "ms": ms,
              "tflops_total": flops / ms / 1e9})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
