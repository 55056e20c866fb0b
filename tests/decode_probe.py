"""Developer probe (B200 via gpurun): decode attention bandwidth at the BASELINE shapes, L2 flushed between iterations.
B200_LIB_PATH selects another build of the library for A/B runs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ml_inference_optimizer_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bf = torch.bfloat16
tag = os.environ.get("TAG", os.path.basename(os.environ.get("B200_LIB_PATH", "default")))
for name, B, S, Hq, Hkv, D in (("c3_mha_b64", 64, 8192, 32, 32, 128), ("c4_gqa_b64", 64, 8192, 32, 8, 128), ("gqa_b8", 8, 8192, 32, 8, 128),
                               ("gqa_b64_s2k", 64, 2048, 32, 8, 128), ("gqa4_b64_d64", 64, 8192, 16, 4, 64), ("mqa_b32", 32, 8192, 16, 1, 128)):
    torch.manual_seed(0)
    q = torch.randn(B, Hq, D, device="cuda", dtype=bf)
    kc, vc = (torch.randn(B, S, Hkv, D, device="cuda", dtype=bf) for _ in range(2))
    lens = torch.full((B,), S, device="cuda", dtype=torch.int32)
    for _ in range(3):
        ops.decode_attention(q, kc, vc, lens)
    ts = []
    for _ in range(15):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.decode_attention(q, kc, vc, lens); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 2.0 * B * S * Hkv * D * 2
    print(json.dumps({"probe": "decode", "lib": tag, "case": name, "ms": round(ms, 4), "gbs": round(nbytes / ms / 1e6, 1)}), flush=True)
