"""Developer probe (B200 via gpurun; needs tests/experiments/r2_stream_k_skinny_gemm.patch applied — without it the
variants all run the shipped kernel): decode-sized linear layers and the T=64 FusedMLP replayed from CUDA graphs of 10 calls,
stream-K on / off and the producer's issue batch (B200_GEMM_STREAMK, B200_GEMM_ISSUE_BATCH, read per call = at capture time), next to cuBLAS."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from ml_inference_optimizer_b200 import ops

bf = torch.bfloat16


def graph_time(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) / n)
    return sorted(ts)[4] * 1e3


VARIANTS = {"base": {}, "streamk": {"B200_GEMM_STREAMK": "1"}, "streamk_nofix": {"B200_GEMM_STREAMK": "1x"}}


def ab(fn, rec):
    for _ in range(2):
        for name, env in VARIANTS.items():
            os.environ.update(env)
            rec.setdefault(name + "_us", []).append(round(graph_time(fn), 2))
            for k in env:
                os.environ.pop(k)


cases = (("up_gate_T64", 64, 4096, 11008, "swiglu"), ("up_gate_T8", 8, 4096, 11008, "swiglu"), ("up_gate_T128", 128, 4096, 11008, "swiglu"),
         ("c4_up_gate_T64", 64, 4096, 14336, "swiglu"), ("qkv_T64", 64, 4096, 12288, None), ("qkv_T8", 8, 4096, 12288, None),
         ("down_T64", 64, 11008, 4096, None))
for name, T, K, N, act in cases:
    x = torch.randn(T, K, device="cuda", dtype=bf)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(4)]
    wg = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(4)] if act == "swiglu" else None
    y = torch.empty(T, N, device="cuda", dtype=bf)
    it = {"i": 0}

    def ours():
        i = it["i"] = (it["i"] + 1) % 4
        ops.linear_act(x, ws[i], None, act, wg[i] if wg else None, None, out=y)

    def cublas():
        i = it["i"] = (it["i"] + 1) % 4
        if act == "swiglu":
            return F.silu(F.linear(x, wg[i])) * F.linear(x, ws[i])
        return F.linear(x, ws[i])

    rec = {"case": name, "weights_MB": round(N * K * 2 * (2 if act == "swiglu" else 1) / 1e6, 1)}
    ab(ours, rec)
    rec["cublas_us"] = round(graph_time(cublas), 2)
    print(json.dumps(rec), flush=True)

for T, h, i_ in ((64, 4096, 11008), (8, 4096, 11008), (64, 4096, 14336)):
    x = torch.randn(T, h, device="cuda", dtype=bf)
    W = [[(torch.randn(*s, device="cuda") * 0.02).to(bf) for s in ((i_, h), (i_, h), (h, i_))] for _ in range(3)]
    it = {"i": 0}

    def mlp():
        i = it["i"] = (it["i"] + 1) % 3
        return ops.fused_mlp(x, W[i][0], None, W[i][2], None, "swiglu", w_gate=W[i][1])

    def mlp_cublas():
        i = it["i"] = (it["i"] + 1) % 3
        return F.linear(F.silu(F.linear(x, W[i][1])) * F.linear(x, W[i][0]), W[i][2])

    rec = {"case": f"fused_mlp_swiglu_T{T}_{h}_{i_}"}
    ab(mlp, rec)
    rec["cublas_us"] = round(graph_time(mlp_cublas), 2)
    print(json.dumps(rec), flush=True)
    del W
