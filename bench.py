#!/usr/bin/env python
"""bench.py — the contract benchmark of the attention + FusedMLP hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2]

A "step" is one pass of the hot path over one batch of synthetic input at the Llama-2-7B layer shapes of
BASELINE.json configs[2] (the configuration the metric's "SwiGLU-MLP" is quoted on and the largest single-GPU one):
causal FlashAttention forward (B=4, S=8192, 32x128 heads, bf16) followed by the SwiGLU FusedMLP (T=32768,
4096 -> 11008 -> 4096). ``--workload c2`` runs configs[1] (GPT-2 small: B=8, S=4096, 12x64 heads, GELU MLP).
At N>1 every rank runs the same step on its own batch (weak scaling, no data-path collective: the path shards by
sequence batch); the ring / tensor-parallel paths have their own benchmark (benchmarks/multi_gpu_bench.py).

Prints ONE JSON line (rank 0). ``value`` = algorithmic TFLOP/s of the whole job with inputs resident in HBM;
``e2e`` = the same through the public module API with pinned HOST buffers (H2D of q,k,v,x and D2H of both results
inside the timed region); ``roofline`` = the dominant kernel (the FusedMLP GEMM pair) against the measured bf16
peak; ``cpu_baseline`` = the fp32 oracle timed on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: B, S, Hq, Hkv, D, hidden, intermediate, activation
    "c3": dict(name="llama2-7b-layer prefill: causal attn B4 S8192 H32 D128 + SwiGLU MLP 4096->11008 (BASELINE configs[2])",
               B=4, S=8192, Hq=32, Hkv=32, D=128, h=4096, i=11008, act="swiglu"),
    "c2": dict(name="gpt2-small-layer prefill: causal attn B8 S4096 H12 D64 + GELU MLP 768->3072 (BASELINE configs[1])",
               B=8, S=4096, Hq=12, Hkv=12, D=64, h=768, i=3072, act="gelu_tanh"),
}


def flops_of(w, B=None, S=None):
    B = w["B"] if B is None else B
    S = w["S"] if S is None else S
    T = B * S
    attn = 4.0 * B * w["Hq"] * S * S * w["D"] * 0.5          # causal
    mlp = (6.0 if w["act"] == "swiglu" else 4.0) * T * w["h"] * w["i"]
    return attn, mlp


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return dict(tflops_burst=float(d["bf16_tflops"]), tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                        hbm_gbs=float(d["hbm_gbs"]), source="measured")
        except Exception:  # noqa: BLE001
            pass
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons. The sampler is started before the warm-up (nvidia-smi needs a moment to
    start); only samples whose arrival time falls inside [t_start, t_end] of the timed region are summarised."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_start=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()

        def parse(rows):
            sm, smax, reasons, power = [], [], set(), []
            for _, ln in rows:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            return sm, smax, reasons, power

        inside = [(t, l) for t, l in self.lines if t_start is None or (t_start <= t <= t_end + 0.03)]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to the whole run (warm-up + timed)
            inside, window = self.lines, "warm-up + timed region"
        sm, smax, reasons, power = parse(inside)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "window": window, "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the fp32 oracle on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(w, steps, warmup, sample_S=None, sample_B=1):
    import torch
    from oracle import attn_mlp_oracle as orc

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S = sample_S or min(w["S"], 2048)
    B = sample_B
    torch.manual_seed(0)
    q = torch.randn(B, S, w["Hq"], w["D"])
    k = torch.randn(B, S, w["Hkv"], w["D"])
    v = torch.randn(B, S, w["Hkv"], w["D"])
    x = torch.randn(B * S, w["h"])
    wu = torch.randn(w["i"], w["h"]) * 0.02
    wd = torch.randn(w["h"], w["i"]) * 0.02
    bu, bd = torch.zeros(w["i"]), torch.zeros(w["h"])
    wg = torch.randn(w["i"], w["h"]) * 0.02 if w["act"] == "swiglu" else None
    bg = torch.zeros(w["i"]) if w["act"] == "swiglu" else None

    def step():
        o, _ = orc.attention_ref(q, k, v, causal=True)
        y = orc.mlp_ref(x, wu, bu, wd, bd, w["act"], wg, bg)
        return o, y

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    fa, fm = flops_of(w, B=B, S=S)
    return dict(value=(fa + fm) / dt / 1e12, seconds_per_step=dt, cores=torch.get_num_threads(),
                sample=f"fp32 oracle (torch CPU eager), B={B} S={S} of the workload's B={w['B']} S={w['S']}, same heads/widths")


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    r = cpu_oracle_rate(w, steps=min(steps, 5), warmup=warmup)
    line = {
        "impl": "reference", "metric": "fwd attention + FusedMLP TFLOP/s (causal attn + MLP layer step)", "value": r["value"],
        "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": min(steps, 5), "warmup": warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "arm": "reference CPU eager path (oracle port of HF-eager attention + MLP)"},
        "cpu_baseline": {"value": r["value"], "unit": "TFLOP/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist

    from ml_inference_optimizer_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: there is no CPU fallback for the product path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert ops.arch_ok(), "needs an sm_100 device"

    B, S, Hq, Hkv, D, h, i, act = (w[k] for k in ("B", "S", "Hq", "Hkv", "D", "h", "i", "act"))
    T = B * S
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bf = torch.bfloat16
    q = torch.randn(B, S, Hq, D, device=dev, dtype=bf, generator=g)
    k = torch.randn(B, S, Hkv, D, device=dev, dtype=bf, generator=g)
    v = torch.randn(B, S, Hkv, D, device=dev, dtype=bf, generator=g)
    x = torch.randn(T, h, device=dev, dtype=bf, generator=g)
    wu = (torch.randn(i, h, device=dev, generator=g) * 0.02).to(bf)
    wd = (torch.randn(h, i, device=dev, generator=g) * 0.02).to(bf)
    bu = (torch.randn(i, device=dev, generator=g) * 0.02).to(bf)
    bd = (torch.randn(h, device=dev, generator=g) * 0.02).to(bf)
    wg = bg = None
    if act == "swiglu":
        wg = (torch.randn(i, h, device=dev, generator=g) * 0.02).to(bf)
        bg = (torch.randn(i, device=dev, generator=g) * 0.02).to(bf)
    o = torch.empty_like(q)
    y = torch.empty(T, h, device=dev, dtype=bf)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    mlp_events = []

    def step(record=False):
        ops.flash_attn_fwd(q, k, v, causal=True, out=o)
        if record:
            evs = (ev(), ev(), ev())
            ops.fused_mlp(x, wu, bu, wd, bd, act, wg, bg, out=y, timing_events=evs)
            mlp_events.append(evs)
        else:
            ops.fused_mlp(x, wu, bu, wd, bd, act, wg, bg, out=y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    start, end = ev(), ev()
    barrier()
    t_start = time.time()
    start.record()
    for _ in range(args.steps):
        step(record=True)
    end.record()
    barrier()
    t_end = time.time()
    elapsed_ms = start.elapsed_time(end)
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    gemm1_ms = statistics.mean(a.elapsed_time(b_) for a, b_, _ in mlp_events)
    gemm2_ms = statistics.mean(b_.elapsed_time(c_) for _, b_, c_ in mlp_events)
    attn_ms = elapsed_ms / args.steps - gemm1_ms - gemm2_ms

    # ---- e2e: host buffers through the public API, copies inside the timed region ----
    # Every step moves its inputs pinned-host -> device and both results device -> host. The three stages run on three
    # streams over a double-buffered set of device tensors (as a serving loop would): the H2D of step s+1 and the D2H of
    # step s-1 overlap the kernels of step s; PCIe is full duplex.
    pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
    hq, hk, hv, hx = pin(q), pin(k), pin(v), pin(x)
    ho = torch.empty(o.shape, dtype=bf, pin_memory=True)
    hy = torch.empty(y.shape, dtype=bf, pin_memory=True)
    sets = [(q, k, v, x, o, y), tuple(torch.empty_like(t) for t in (q, k, v, x, o, y))]
    s_h2d, s_cmp, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def e2e_run(n_steps):
        h2d_done, cmp_done, d2h_done = {}, {}, {}
        for st in range(n_steps):
            dq, dk, dv, dx, do, dy = sets[st % 2]
            with torch.cuda.stream(s_h2d):
                if st >= 2:
                    s_h2d.wait_event(cmp_done[st - 2])  # the kernels of step st-2 have consumed this input set
                dq.copy_(hq, non_blocking=True); dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True)
                dx.copy_(hx, non_blocking=True)
                h2d_done[st] = torch.cuda.Event(); h2d_done[st].record(s_h2d)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(h2d_done[st])
                if st >= 2:
                    s_cmp.wait_event(d2h_done[st - 2])  # the results of step st-2 have left this output set
                ops.flash_attn_fwd(dq, dk, dv, causal=True, out=do)
                ops.fused_mlp(dx, wu, bu, wd, bd, act, wg, bg, out=dy)
                cmp_done[st] = torch.cuda.Event(); cmp_done[st].record(s_cmp)
            with torch.cuda.stream(s_d2h):
                s_d2h.wait_event(cmp_done[st])
                ho.copy_(do, non_blocking=True); hy.copy_(dy, non_blocking=True)
                d2h_done[st] = torch.cuda.Event(); d2h_done[st].record(s_d2h)
        for st_ in (s_h2d, s_cmp, s_d2h):
            torch.cuda.current_stream(dev).wait_stream(st_)

    e2e_steps = max(4, min(args.steps, 10))
    e2e_run(2)
    barrier()
    s2, e2 = ev(), ev()
    for st_ in (s_h2d, s_cmp, s_d2h):
        st_.wait_stream(torch.cuda.current_stream(dev))
    s2.record()
    for st_ in (s_h2d, s_cmp, s_d2h):
        st_.wait_event(s2)
    e2e_run(e2e_steps)
    e2.record()
    barrier()
    e2e_ms = s2.elapsed_time(e2)
    h2d = sum(t.numel() * t.element_size() for t in (q, k, v, x))
    d2h = sum(t.numel() * t.element_size() for t in (o, y))

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = t.tolist()

    fa, fm = flops_of(w)
    total_flops = (fa + fm) * world
    ms_per_step = elapsed_ms / args.steps
    value = total_flops / (ms_per_step * 1e-3) / 1e12
    e2e_value = total_flops / (e2e_ms / e2e_steps * 1e-3) / 1e12

    if rank == 0:
        peaks = measured_peaks()
        peak = peaks["tflops_sustained"]  # the kernel is timed inside a long step
        gemm1_flops = (4.0 if act == "swiglu" else 2.0) * T * h * i   # up (+ gate) projection of the FusedMLP
        achieved = gemm1_flops / (gemm1_ms * 1e-3) / 1e12
        cpu = cpu_oracle_rate(w, steps=2, warmup=1) if world == 1 and not args.no_cpu_baseline else None
        line = {
            "metric": "fwd attention + FusedMLP TFLOP/s (causal attn + MLP layer step)", "value": value, "unit": "TFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": w["name"], "parallelism": f"dp{world} (independent batches per GPU, no collective)",
                       "l2": "inputs (q,k,v,x = %.0f MB per step) exceed the 126 MB L2" % (h2d / 1e6),
                       "attn_flops_per_step": fa, "mlp_flops_per_step": fm},
            "roofline": {"kernel": "gemm_act_pair_kernel<%s> (FusedMLP up%s GEMM on CTA pairs, activation fused in the epilogue)" %
                                   (act, "+gate" if act == "swiglu" else ""),
                         "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}): kernel timed inside a long step",
                         "ms_per_launch": gemm1_ms, "flops_per_launch": gemm1_flops,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r1_gemm1_v0_ncu.txt)
                         "traffic": 3.872766e9 if args.workload == "c3" else None,
                         "other_kernels_ms": {"fa_fwd_kernel": attn_ms, "gemm_act_pair_kernel<NONE> (down projection)": gemm2_ms}},
            "e2e": {"value": e2e_value, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "pipeline": "3 streams (H2D / kernels / D2H), double-buffered device tensors, pinned host buffers"},
            "gpu_launches": 3 * args.steps,
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": "TFLOP/s", "cores": cpu["cores"], "kind": "port",
                                    "sample": cpu["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
