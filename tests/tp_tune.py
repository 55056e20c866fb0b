"""Sweep of the tensor-parallel FusedMLP pipeline (token chunks x CTAs given to the in-switch all-reduce K6) at the C3 / C4
layer shapes; run under torchrun on N GPUs. Prints one JSON line per setting (rank 0; device time, max over ranks)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
import torch.nn.functional as F


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    from ml_inference_optimizer_b200 import ops
    from ml_inference_optimizer_b200.parallelism import parallel_utils as pu
    from ml_inference_optimizer_b200.parallelism.tensor_parallel import TensorParallelConfig, TensorParallelMLP

    pu.initialize_tensor_parallel(world)
    cfg = TensorParallelConfig(world_size=world, tp_size=world)

    def timed(fn, warmup=2, iters=6):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record(); torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for tag, (T, h, i) in {"c3": (32768, 4096, 11008), "c4": (32768, 4096, 14336)}.items():
        if i % (world * 8):
            continue
        mlp = TensorParallelMLP(h, i, cfg, F.silu, gated=True).to(dev, torch.bfloat16)
        mlp.symmetric_output = "view"
        x = torch.randn(T, h, device=dev, dtype=torch.bfloat16)
        flops = 6.0 * T * h * i
        up, down, gate = mlp.dense_h_to_4h, mlp.dense_4h_to_h, mlp.dense_h_to_4h_gate
        y = torch.empty(T, h, device=dev, dtype=torch.bfloat16)
        base = timed(lambda: ops.fused_mlp(x, up.weight, up.bias, down.weight, None, "swiglu", gate.weight, gate.bias, out=y))
        if rank == 0:
            print(json.dumps({"shape": tag, "n": world, "what": "gemms_alone", "ms": base, "tflops_total": flops / base / 1e9}), flush=True)
        mlp.reduce_impl = "nccl"
        for chunks in (1, 4):
            mlp.overlap_chunks = chunks
            ms = timed(lambda: mlp(x))
            if rank == 0:
                print(json.dumps({"shape": tag, "n": world, "what": "nccl", "chunks": chunks, "ms": ms, "tflops_total": flops / ms / 1e9}), flush=True)
        mlp.reduce_impl = "symmetric"
        for chunks in (1, 2, 4, 8):
            for ctas, reserve in (((0, 0),) if chunks == 1 else ((0, 0), (64, 0), (32, 16))):
                mlp.overlap_chunks, mlp.comm_ctas, mlp.comm_ctas_single, mlp.gemm_sm_reserve = chunks, ctas, ctas, reserve
                ms = timed(lambda: mlp(x))
                if rank == 0:
                    print(json.dumps({"shape": tag, "n": world, "what": "k6_" + mlp.last_reduce, "chunks": chunks, "comm_ctas": ctas or 148,
                                      "gemm_sm_reserve": reserve, "ms": ms, "tflops_total": flops / ms / 1e9}), flush=True)
        pool = mlp._symmetric_pool(1, dev)
        for b in pool["bufs"]:
            b.check()
        del mlp, x, y
        torch.cuda.empty_cache()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
