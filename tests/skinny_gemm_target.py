"""Developer tool: the decode-sized linear layers launched a few times each (target of an ncu launch-list pass: how the time
of a split-K call divides between the GEMM and the reduce kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ml_inference_optimizer_b200 import ops

bf = torch.bfloat16
for name, T, K, N, act in (("up_gate", 64, 4096, 11008, "swiglu"), ("down", 64, 11008, 4096, None), ("qkv", 64, 4096, 12288, None),
                           ("down_T8", 8, 11008, 4096, None)):
    x = torch.randn(T, K, device="cuda", dtype=bf)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(3)]
    wg = [(torch.randn(N, K, device="cuda") * 0.02).to(bf) for _ in range(3)] if act == "swiglu" else None
    y = torch.empty(T, N, device="cuda", dtype=bf)
    for i in range(6):
        ops.linear_act(x, ws[i % 3], None, act, wg[i % 3] if wg else None, None, out=y)
    torch.cuda.synchronize()
    print(name, ops.last_kernel(), flush=True)
