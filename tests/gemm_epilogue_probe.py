"""Developer probe: is the small-K GEMM of the GPT-2 MLP epilogue-bound? fc1 (K=768) with and without the activation,
fc2 (K=3072), next to cuBLAS, L2 flushed."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ml_inference_optimizer_b200 import ops
bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]
T = 32768
for name, K, N, act in (("k512_none", 512, 3072, None), ("k256_none", 256, 3072, None), ("k512_relu", 512, 3072, "relu"), ("fc1_gelu", 768, 3072, "gelu_tanh"), ("fc1_relu", 768, 3072, "relu"), ("fc1_none", 768, 3072, None), ("fc2_none", 3072, 768, None),
                        ("k1536_gelu", 1536, 6144, "gelu_tanh"), ("k1536_none", 1536, 6144, None), ("llama_up_swiglu", 4096, 11008, "swiglu")):
    x = torch.randn(T, K, device="cuda", dtype=bf)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(bf); b = torch.zeros(N, device="cuda", dtype=bf)
    wg = (torch.randn(N, K, device="cuda") * 0.02).to(bf) if act == "swiglu" else None
    y = torch.empty(T, N, device="cuda", dtype=bf)
    ms = timeit(lambda: ops.linear_act(x, w, b, act, wg, b if wg is not None else None, out=y))
    k1 = ops.last_gemm_kernel()
    os.environ["B200_GEMM_EPI_WGS"] = "2" if "2wg" not in k1 else "1"
    ms_alt = timeit(lambda: ops.linear_act(x, w, b, act, wg, b if wg is not None else None, out=y))
    k_alt = ops.last_gemm_kernel()
    os.environ.pop("B200_GEMM_EPI_WGS")
    cb = timeit(lambda: F.linear(x, w, b))
    fl = 2.0 * T * K * N * (2 if act == "swiglu" else 1)
    print(json.dumps({"case": name, "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 0), "cublas_plain_gemm_ms": round(cb, 4), "kernel": k1, "other_epilogue_ms": round(ms_alt, 4), "other_kernel": k_alt}), flush=True)
