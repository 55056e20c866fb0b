"""Developer probe: skinny (decode-sized) linear layers — time vs forced split-K count (B200_GEMM_K_SPLITS), next to the cost
model's own choice (0) and cuBLAS, replayed from CUDA graphs of 10 calls with the weights (>> L2) streamed from HBM."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ml_inference_optimizer_b200 import ops
bf = torch.bfloat16

def graph_time(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(2): g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) / n)
    return sorted(ts)[3] * 1e3

for name, T, K, N, act in (("up_gate_T64", 64, 4096, 11008, "swiglu"), ("down_T64", 64, 11008, 4096, None), ("up_gate_T8", 8, 4096, 11008, "swiglu"),
                           ("down_T8", 8, 11008, 4096, None), ("qkv_T64", 64, 4096, 12288, None), ("up_gate_T256", 256, 4096, 11008, "swiglu"),
                           ("down_T256", 256, 11008, 4096, None), ("gpt2_fc1_T64", 64, 768, 3072, "gelu_tanh"), ("c4_down_T64", 64, 14336, 4096, None)):
    x = torch.randn(T, K, device="cuda", dtype=bf)
    # several weight copies so that consecutive calls in the graph stream different memory (weights >> L2 in aggregate)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(4)]
    wg = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(4)] if act == "swiglu" else None
    y = torch.empty(T, N, device="cuda", dtype=bf)
    it = {"i": 0}
    def ours():
        i = it["i"] = (it["i"] + 1) % 4
        ops.linear_act(x, ws[i], None, act, wg[i] if wg else None, None, out=y)
    def cublas():
        i = it["i"] = (it["i"] + 1) % 4
        if act == "swiglu": return F.silu(F.linear(x, wg[i])) * F.linear(x, ws[i])
        o = F.linear(x, ws[i])
        return F.gelu(o, approximate="tanh") if act == "gelu_tanh" else o
    res = {}
    for s in (0, 1, 2, 3, 4, 5, 6, 8, 9, 12, 16, 24, 32):
        os.environ["B200_GEMM_K_SPLITS"] = str(s) if s else ""
        if not s: os.environ.pop("B200_GEMM_K_SPLITS")
        try:
            res[s] = round(graph_time(ours), 1)
        except Exception as e:
            res[s] = str(e)[:30]
    os.environ.pop("B200_GEMM_K_SPLITS", None)
    cb = round(graph_time(cublas), 1)
    wbytes = N * K * 2 * (2 if act == "swiglu" else 1)
    best = min((v, k) for k, v in res.items() if isinstance(v, float))
    print(json.dumps({"case": name, "us_by_splits(0=model)": res, "best": best, "cublas_us": cb, "weights_MB": round(wbytes / 1e6, 1),
                      "model_gbs": round(wbytes / res[0] / 1e3, 0), "best_gbs": round(wbytes / best[0] / 1e3, 0)}), flush=True)
