"""Probe of the symmetric-memory all-reduce kernel K6 (run under torchrun on >= 2 GPUs):
correctness against an fp32 sum of the per-rank partials, bandwidth against NCCL, CTA-count sweep.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/symm_probe.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from ml_inference_optimizer_b200.parallelism.symmetric import SymmetricBuffer

    T, h = 32768, 4096
    buf = SymmetricBuffer(T * h * 2)
    if rank == 0:
        print(json.dumps({"probe": "symmetric", "world": world, "multicast": buf.multicast,
                          "has_multicast_support": bool(getattr(buf.handle, "has_multicast_support", False))}), flush=True)
    ok = True
    for mode in (["multicast", "peer"] if buf.multicast else ["peer"]):
        mc_saved = buf.multicast_ptr
        if mode == "peer":
            buf.multicast_ptr = 0
        for rows, with_bias in ((300, False), (4096, True)):
            g = torch.Generator(device=dev).manual_seed(100 + rank)
            t = buf.view((rows, h), torch.bfloat16)
            part = torch.randn(rows, h, device=dev, dtype=torch.bfloat16, generator=g)
            bias = torch.randn(h, device=dev, dtype=torch.bfloat16, generator=torch.Generator(device=dev).manual_seed(5)) if with_bias else None
            gathered = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(gathered, part)
            want = sum(p.float() for p in gathered)
            if bias is not None:
                want = want + bias.float()
            t.copy_(part)
            buf.all_reduce_(t, bias)
            buf.check()
            err = (t.float() - want).abs().max().item()
            # fp32 accumulation of the bf16 partials; the switch's conversion back to bf16 is not round-to-nearest
            # (measured: up to one bf16 ulp, 2^-7 relative, against half an ulp for the unicast path) (+ bias rounding)
            tol = want.abs().max().item() * 2 ** -7 * (2.0 if with_bias else 1.0)
            good = err <= tol
            ok &= good
            print(f"[rank {rank}] {mode} rows={rows} bias={with_bias}: max|err|={err:.3e} (tol {tol:.3e}) {'OK' if good else 'FAIL'}", flush=True)
        # bandwidth
        t = buf.view((T, h), torch.bfloat16)
        for ctas in (16, 32, 64, 148):
            for _ in range(2):
                buf.all_reduce_(t, max_ctas=ctas)
            torch.cuda.synchronize(); dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(5):
                buf.all_reduce_(t, max_ctas=ctas)
            e.record(); torch.cuda.synchronize()
            ms = torch.tensor([s.elapsed_time(e) / 5], device=dev, dtype=torch.float64)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                bus = 2 * (world - 1) / world * T * h * 2
                print(json.dumps({"probe": "k6_allreduce", "mode": mode, "ctas": ctas, "bytes": T * h * 2, "ms": ms.item(),
                                  "busbw_gbs": bus / ms.item() / 1e6}), flush=True)
        buf.check()
        buf.multicast_ptr = mc_saved
    y = torch.randn(T, h, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        dist.all_reduce(y)
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        dist.all_reduce(y)
    e.record(); torch.cuda.synchronize()
    ms = torch.tensor([s.elapsed_time(e) / 5], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        bus = 2 * (world - 1) / world * T * h * 2
        print(json.dumps({"probe": "nccl_allreduce", "bytes": T * h * 2, "ms": ms.item(), "busbw_gbs": bus / ms.item() / 1e6}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
