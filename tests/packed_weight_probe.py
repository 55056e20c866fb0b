"""Developer probe (B200 via gpurun; needs tests/experiments/r2_packed_weights.patch applied): decode-sized linear layers with
the weight repacked offline into [N/128][K/64][128][64] panels (one TMA box = one contiguous 16 KB read) against the
row-major weight the reference's API hands over (one box = 128 chunks of 128 bytes, a row apart) and cuBLAS; CUDA graphs of
10 calls, weights rotated so that they stream from HBM."""
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


def pack(w):
    N, K = w.shape
    return w.view(N // 128, 128, K // 64, 64).permute(0, 2, 1, 3).contiguous().view(N, K)


cases = (("up_gate_T64", 64, 4096, 11008, "swiglu"), ("up_gate_T8", 8, 4096, 11008, "swiglu"), ("c4_up_gate_T64", 64, 4096, 14336, "swiglu"),
         ("qkv_T64", 64, 4096, 12288, None), ("down_T64", 64, 11008, 4096, None), ("down_T8", 8, 11008, 4096, None))
for name, T, K, N, act in cases:
    x = torch.randn(T, K, device="cuda", dtype=bf)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(4)]
    wg = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(4)] if act == "swiglu" else None
    wsp = [pack(w) for w in ws]
    wgp = [pack(w) for w in wg] if wg else None
    y = torch.empty(T, N, device="cuda", dtype=bf)
    it = {"i": 0}

    def ours(packed):
        i = it["i"] = (it["i"] + 1) % 4
        W, G = (wsp, wgp) if packed else (ws, wg)
        ops.linear_act(x, W[i], None, act, G[i] if G else None, None, out=y)

    def cublas():
        i = it["i"] = (it["i"] + 1) % 4
        if act == "swiglu":
            return F.silu(F.linear(x, wg[i])) * F.linear(x, ws[i])
        return F.linear(x, ws[i])

    rec = {"case": name, "weights_MB": round(N * K * 2 * (2 if act == "swiglu" else 1) / 1e6, 1)}
    os.environ.pop("B200_GEMM_PACKED_B", None)
    it["i"] = 0; ours(False); ref = y.clone()
    os.environ["B200_GEMM_PACKED_B"] = "1"
    it["i"] = 0; ours(True); rec["bit_identical"] = bool(torch.equal(ref, y))
    for _ in range(2):
        os.environ.pop("B200_GEMM_PACKED_B", None)
        rec.setdefault("row_major_us", []).append(round(graph_time(lambda: ours(False)), 2))
        os.environ["B200_GEMM_PACKED_B"] = "1"
        rec.setdefault("packed_us", []).append(round(graph_time(lambda: ours(True)), 2))
    os.environ.pop("B200_GEMM_PACKED_B", None)
    rec["cublas_us"] = round(graph_time(cublas), 2)
    rec["packed_gbs"] = round(rec["weights_MB"] * 1e3 / min(rec["packed_us"]), 0)
    print(json.dumps(rec), flush=True)
    del ws, wg, wsp, wgp

for T, h, i_ in ((64, 4096, 11008), (8, 4096, 11008)):
    x = torch.randn(T, h, device="cuda", dtype=bf)
    W = [[(torch.randn(*s, device="cuda") * 0.02).to(bf) for s in ((i_, h), (i_, h), (h, i_))] for _ in range(3)]
    WP = [[pack(w) for w in ws] for ws in W]
    it = {"i": 0}

    def mlp(packed):
        i = it["i"] = (it["i"] + 1) % 3
        w = WP[i] if packed else W[i]
        return ops.fused_mlp(x, w[0], None, w[2], None, "swiglu", w_gate=w[1])

    def mlp_cublas():
        i = it["i"] = (it["i"] + 1) % 3
        return F.linear(F.silu(F.linear(x, W[i][1])) * F.linear(x, W[i][0]), W[i][2])

    rec = {"case": f"fused_mlp_swiglu_T{T}_{h}_{i_}"}
    it["i"] = 0; ref = mlp(False).clone()
    os.environ["B200_GEMM_PACKED_B"] = "1"
    it["i"] = 0; rec["bit_identical"] = bool(torch.equal(ref, mlp(True)))
    for _ in range(2):
        os.environ.pop("B200_GEMM_PACKED_B", None)
        rec.setdefault("row_major_us", []).append(round(graph_time(lambda: mlp(False)), 2))
        os.environ["B200_GEMM_PACKED_B"] = "1"
        rec.setdefault("packed_us", []).append(round(graph_time(lambda: mlp(True)), 2))
    os.environ.pop("B200_GEMM_PACKED_B", None)
    rec["cublas_us"] = round(graph_time(mlp_cublas), 2)
    print(json.dumps(rec), flush=True)
    del W, WP
