"""Tiny single-kernel driver for ncu captures (python tests/ncu_target.py fa|fa64|gemm1|gemm2|decode|decode_gqa)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ml_inference_optimizer_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "fa"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bf = torch.bfloat16
torch.manual_seed(0)
if what in ("fa", "fa64"):
    B, S, H, D = (4, 8192, 32, 128) if what == "fa" else (8, 4096, 12, 64)
    q, k, v = (torch.randn(B, S, H, D, device="cuda", dtype=bf) for _ in range(3))
    for _ in range(reps):
        ops.flash_attn_fwd(q, k, v, causal=True)
elif what in ("gemm1", "gemm2", "mlp", "mlp_c2"):
    T, h, i = (32768, 4096, 11008) if what != "mlp_c2" else (32768, 768, 3072)
    x = torch.randn(T, h, device="cuda", dtype=bf)
    wu, wg = (torch.randn(i, h, device="cuda", dtype=bf) * 0.02 for _ in range(2))
    wd = torch.randn(h, i, device="cuda", dtype=bf) * 0.02
    hmid = torch.randn(T, i, device="cuda", dtype=bf)
    for _ in range(reps):
        if what == "gemm1":
            ops.linear_act(x, wu, None, "swiglu", wg, None)
        elif what == "gemm2":
            ops.linear_act(hmid, wd, None, None)
        elif what == "mlp_c2":
            ops.fused_mlp(x, wu, None, wd, None, "gelu_tanh")
        else:
            ops.fused_mlp(x, wu, None, wd, None, "swiglu", wg, None)
elif what in ("decode", "decode_gqa"):
    B, S, Hq, Hkv, D = (64, 8192, 32, 32, 128) if what == "decode" else (64, 8192, 32, 8, 128)
    q = torch.randn(B, Hq, D, device="cuda", dtype=bf)
    kc, vc = (torch.randn(B, S, Hkv, D, device="cuda", dtype=bf) for _ in range(2))
    lens = torch.full((B,), S, device="cuda", dtype=torch.int32)
    for _ in range(reps):
        ops.decode_attention(q, kc, vc, lens)
elif what == "ln":
    x = torch.randn(32768, 4096, device="cuda", dtype=bf)
    r = torch.randn_like(x)
    w_, b_ = torch.randn(4096, device="cuda", dtype=bf), torch.randn(4096, device="cuda", dtype=bf)
    for _ in range(reps):
        ops.layernorm(x, w_, b_, 1e-5, residual=r)
torch.cuda.synchronize()
print("done", what)
if what == "mlp64":
    T, h, i = 64, 4096, 11008
    x = torch.randn(T, h, device="cuda", dtype=bf)
    wu, wg = (torch.randn(i, h, device="cuda", dtype=bf) * 0.02 for _ in range(2))
    wd = torch.randn(h, i, device="cuda", dtype=bf) * 0.02
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(reps):
        flush.zero_()
        ops.fused_mlp(x, wu, None, wd, None, "swiglu", wg, None)
        flush.zero_()
        torch.nn.functional.linear(torch.nn.functional.silu(torch.nn.functional.linear(x, wg)) * torch.nn.functional.linear(x, wu), wd)
    torch.cuda.synchronize()
