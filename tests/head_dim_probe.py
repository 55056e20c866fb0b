"""Developer probe (B200 via gpurun): prefill attention at head dims the kernels are not built for (they run in the next wider
build with TMA zero-fill) next to cuDNN SDPA, one JSON line per shape."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

from ml_inference_optimizer_b200 import ops


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for D in (32, 64, 80, 96, 128):
    B, S, H = 4, 8192, 32
    torch.manual_seed(0)
    q, k, v = [torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    flops = 4.0 * B * H * S * S * D * 0.5
    ms = timed(lambda: ops.flash_attn_fwd(q, k, v, causal=True))
    rec = {"probe": "head_dim", "D": D, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "kernel": ops.last_kernel()}
    try:
        def cudnn():
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                return torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=True)
        ref = cudnn()
        got = ops.flash_attn_fwd(q, k, v, causal=True)
        rec["max_abs_vs_cudnn"] = round((got.float() - ref.transpose(1, 2).float()).abs().max().item(), 5)
        cms = timed(cudnn)
        rec["cudnn_ms"] = round(cms, 4)
        rec["cudnn_tflops"] = round(flops / cms / 1e9, 1)
    except Exception as ex:  # cuDNN may not take this head dim
        rec["cudnn_error"] = str(ex)[:120]
    print(json.dumps(rec), flush=True)
