"""Developer perf probe (run on a B200 via gpurun): kernel-level timings of the hot path at BASELINE shapes with
same-GPU comparators (SDPA / flash_attn / cuBLAS+elementwise). Not part of the pytest suite; bench.py is the
contract benchmark."""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ml_inference_optimizer_b200 import ops

L2_FLUSH = None


def flush_l2():
    global L2_FLUSH
    if L2_FLUSH is None:
        L2_FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    L2_FLUSH.zero_()


def timeit(fn, warmup=3, iters=10, flush=True):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2], times[0]


def attn(B, S, Hq, Hkv, D, causal, which):
    torch.manual_seed(0)
    q = torch.randn(B, S, Hq, D, device="cuda", dtype=torch.bfloat16)
    k = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    v = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    flops = 4.0 * B * Hq * S * S * D * (0.5 if causal else 1.0)
    res = {}
    if "ours" in which:
        med, best = timeit(lambda: ops.flash_attn_fwd(q, k, v, causal=causal))
        res["ours"] = (med, flops / med / 1e9)
    if "sdpa" in which:
        qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        if Hkv != Hq:
            kt = kt.repeat_interleave(Hq // Hkv, 1)
            vt = vt.repeat_interleave(Hq // Hkv, 1)
        from torch.nn.attention import SDPBackend, sdpa_kernel
        for name, be in (("sdpa_cudnn", SDPBackend.CUDNN_ATTENTION), ("sdpa_flash", SDPBackend.FLASH_ATTENTION)):
            try:
                with sdpa_kernel(be):
                    med, best = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal))
                res[name] = (med, flops / med / 1e9)
            except Exception as e:  # noqa: BLE001
                res[name] = ("err", str(e)[:80])
    if "fa2" in which:
        try:
            from flash_attn import flash_attn_func
            med, best = timeit(lambda: flash_attn_func(q, k, v, causal=causal))
            res["flash_attn2"] = (med, flops / med / 1e9)
        except Exception as e:  # noqa: BLE001
            res["flash_attn2"] = ("err", str(e)[:80])
    print(f"ATTN B{B} S{S} Hq{Hq} Hkv{Hkv} D{D} causal={causal}: " +
          "  ".join(f"{k}={v[0]:.3f}ms/{v[1]:.0f}TF" if not isinstance(v[0], str) else f"{k}=ERR {v[1]}" for k, v in res.items()),
          flush=True)
    return res


def decode(B, S, Hq, Hkv, D, splits=0):
    torch.manual_seed(0)
    q = torch.randn(B, Hq, D, device="cuda", dtype=torch.bfloat16)
    kc = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    vc = torch.randn(B, S, Hkv, D, device="cuda", dtype=torch.bfloat16)
    lens = torch.full((B,), S, device="cuda", dtype=torch.int32)
    nbytes = 2.0 * B * S * Hkv * D * 2
    med, best = timeit(lambda: ops.decode_attention(q, kc, vc, lens, num_splits=splits), flush=False)
    print(f"DECODE B{B} S{S} Hq{Hq} Hkv{Hkv} D{D} splits={splits}: {med:.3f} ms  {nbytes / med / 1e6:.0f} GB/s "
          f"(best {nbytes / best / 1e6:.0f})", flush=True)
    del kc, vc
    torch.cuda.empty_cache()


def mlp(T, h, i, act):
    torch.manual_seed(0)
    x = torch.randn(T, h, device="cuda", dtype=torch.bfloat16)
    wu = (torch.randn(i, h, device="cuda") * 0.02).to(torch.bfloat16)
    bu = torch.zeros(i, device="cuda", dtype=torch.bfloat16)
    wd = (torch.randn(h, i, device="cuda") * 0.02).to(torch.bfloat16)
    bd = torch.zeros(h, device="cuda", dtype=torch.bfloat16)
    wg = bg = None
    F = torch.nn.functional
    if act == "swiglu":
        wg = (torch.randn(i, h, device="cuda") * 0.02).to(torch.bfloat16)
        bg = torch.zeros(i, device="cuda", dtype=torch.bfloat16)
        flops = 6.0 * T * h * i
        ref = lambda: F.linear(F.silu(F.linear(x, wg, bg)) * F.linear(x, wu, bu), wd, bd)
    else:
        flops = 4.0 * T * h * i
        ref = lambda: F.linear(F.gelu(F.linear(x, wu, bu), approximate="tanh"), wd, bd)
    med, _ = timeit(lambda: ops.fused_mlp(x, wu, bu, wd, bd, act, wg, bg))
    medr, _ = timeit(ref)
    med1, _ = timeit(lambda: ops.linear_act(x, wu, bu, act, wg, bg))
    print(f"MLP T{T} h{h} i{i} {act}: ours={med:.3f}ms/{flops / med / 1e9:.0f}TF (gemm1 {med1:.3f}ms)  "
          f"cublas_unfused={medr:.3f}ms/{flops / medr / 1e9:.0f}TF  speedup={medr / med:.3f}", flush=True)


def graph_time(fn, iters=20):
    """device time of fn replayed from a CUDA graph (removes host launch overhead: what decode serving would see)"""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]


def mlp_decode(T, h, i):
    torch.manual_seed(0)
    bf = torch.bfloat16
    x = torch.randn(T, h, device="cuda", dtype=bf)
    wu, wg = ((torch.randn(i, h, device="cuda") * 0.02).to(bf) for _ in range(2))
    wd = (torch.randn(h, i, device="cuda") * 0.02).to(bf)
    F = torch.nn.functional
    y = torch.empty(T, h, device="cuda", dtype=bf)
    ours = graph_time(lambda: ops.fused_mlp(x, wu, None, wd, None, "swiglu", wg, None, out=y))
    ref = graph_time(lambda: F.linear(F.silu(F.linear(x, wg)) * F.linear(x, wu), wd))
    nbytes = 3.0 * h * i * 2
    print(f"MLP-decode (CUDA graph) T{T} h{h} i{i} swiglu: ours={ours * 1e3:.1f}us ({nbytes / ours / 1e6:.0f} GB/s of weights)  "
          f"cublas_unfused={ref * 1e3:.1f}us  speedup={ref / ours:.2f}", flush=True)


def gemm(T, K, N):
    x = torch.randn(T, K, device="cuda", dtype=torch.bfloat16)
    w = torch.randn(N, K, device="cuda", dtype=torch.bfloat16)
    flops = 2.0 * T * K * N
    med, _ = timeit(lambda: ops.linear_act(x, w))
    medr, _ = timeit(lambda: torch.nn.functional.linear(x, w))
    print(f"GEMM T{T} K{K} N{N}: ours={med:.3f}ms/{flops / med / 1e9:.0f}TF cublas={medr:.3f}ms/{flops / medr / 1e9:.0f}TF",
          flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["attn", "decode", "mlp", "gemm"]
    if "attn" in what:
        attn(4, 8192, 32, 32, 128, True, ("ours", "sdpa", "fa2"))
        attn(4, 8192, 32, 32, 128, False, ("ours", "sdpa"))
        attn(8, 4096, 12, 12, 64, True, ("ours", "sdpa", "fa2"))
        attn(4, 8192, 32, 8, 128, True, ("ours",))
    if "decode" in what:
        decode(64, 8192, 32, 32, 128)
        decode(64, 8192, 32, 32, 128, splits=2)
        decode(64, 8192, 32, 8, 128)
        decode(8, 8192, 32, 8, 128)
    if "gemm" in what:
        gemm(8192, 8192, 8192)
        gemm(32768, 4096, 4096)
    if "mlp" in what:
        mlp(32768, 4096, 11008, "swiglu")
        mlp(32768, 768, 3072, "gelu_tanh")
        mlp(32768, 4096, 14336, "swiglu")
        mlp(64, 4096, 11008, "swiglu")
    if "mlpdecode" in what or "mlp" in what:
        mlp_decode(64, 4096, 11008)
        mlp_decode(64, 4096, 14336)
        mlp_decode(8, 4096, 11008)
