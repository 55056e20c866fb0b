"""Developer tool (torchrun, one rank per GPU): timeline of one ring-attention pass at the C5 shape (128K causal tokens,
32 x 128 heads, zigzag): CUDA events around every attention launch (compute stream) and every K/V hop (side stream), all
relative to the start of the pass, plus the NVLink data counters of the GPU before / after (NVML field values) next to the
algorithmic bytes. Shows, rather than infers from an on/off A/B, that hop i+1 travels while block i is being attended to.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 tests/ring_timeline.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def nvlink_counters(index):
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        tx, rx = nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX
        vals = nv.nvmlDeviceGetFieldValues(h, [(tx, 0xFFFFFFFF), (rx, 0xFFFFFFFF)])   # scope: all links of the GPU
        if vals[0].nvmlReturn != 0 or vals[1].nvmlReturn != 0:
            return {"error": f"NVML field query returned {vals[0].nvmlReturn}/{vals[1].nvmlReturn}"}
        return {"tx_kib": int(vals[0].value.ullVal), "rx_kib": int(vals[1].value.ullVal)}
    except Exception as e:  # counters not exposed on this box
        return {"error": f"{type(e).__name__}: {e}"[:100]}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
    dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
    from parallelism.ring import ring_attention_forward

    S, H, D = int(os.environ.get("RING_SEQ", 131072)), 32, 128
    g = torch.Generator(device=dev).manual_seed(rank)
    q, k, v = (torch.randn(1, S // world, H, D, device=dev, dtype=torch.bfloat16, generator=g) for _ in range(3))
    for _ in range(2):
        ring_attention_forward(q, k, v, causal=True, partition="zigzag")
    nvlink_counters(local)            # (NVML initialisation takes a rank-dependent time: keep it out of the traced pass)
    torch.cuda.synchronize()
    before = nvlink_counters(local)
    dist.barrier()                    # the ranks enter the traced pass together
    torch.cuda.synchronize()
    trace = {}
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    ring_attention_forward(q, k, v, causal=True, partition="zigzag", trace=trace)
    t1.record()
    torch.cuda.synchronize()
    after = nvlink_counters(local)
    rel = lambda e: round(t0.elapsed_time(e), 3)
    rec = {"rank": rank, "world": world, "seq": S, "total_ms": rel(t1),
           "attention": [{"step": i, "kv_from_rank": src, "start_ms": rel(a), "end_ms": rel(b)} for i, (a, b, src) in enumerate(trace["attn"])],
           "hops": [{"hop": i, "start_ms": rel(a), "end_ms": rel(b), "ms": round(a.elapsed_time(b), 3)} for i, (a, b) in enumerate(trace["hops"])],
           "kv_bytes_per_hop": 2 * (S // world) * H * D * 2}
    for i, hop in enumerate(rec["hops"]):   # hop i is in flight during attention step i
        a = rec["attention"][i]
        hop["hidden_under_attention_step"] = bool(hop["end_ms"] <= a["end_ms"] + 0.05)
        hop["gbs"] = round(rec["kv_bytes_per_hop"] / hop["ms"] / 1e6, 1)
    if "tx_kib" in before and "tx_kib" in after:
        rec["nvlink"] = {"tx_bytes": (after["tx_kib"] - before["tx_kib"]) * 1024, "rx_bytes": (after["rx_kib"] - before["rx_kib"]) * 1024,
                         "algorithmic_bytes_each_way": rec["kv_bytes_per_hop"] * (world - 1)}
    else:
        rec["nvlink"] = {"before": before, "after": after}
    for r in range(world):
        if r == rank:
            print(json.dumps(rec), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
