"""Dev tool: fit t = waves * (a + b * kv_tiles) for the prefill kernel (non-causal, Sq fixed, Sk swept)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_inference_optimizer_b200 import ops

def t_ms(fn, iters=20):
    fn(); fn(); torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

B, H, D, Sq = 4, 32, 128, 8192
if len(sys.argv) > 1: D = int(sys.argv[1])
q = torch.randn(B, Sq, H, D, device="cuda", dtype=torch.bfloat16)
ctas = B * H * Sq // 256
sms = ops.sm_count()
for Sk in (128, 256, 512, 1024, 2048, 4096, 8192):
    k = torch.randn(B, Sk, H, D, device="cuda", dtype=torch.bfloat16); v = torch.randn_like(k)
    o = torch.empty_like(q)
    ms = t_ms(lambda: ops.flash_attn_fwd(q, k, v, causal=False, out=o))
    print(f"D={D} Sk={Sk:5d} tiles={Sk//128:3d} t={ms*1e3:8.1f} us  per-CTA-slot={ms*1e3/ (ctas/sms):7.2f} us  TF={4*B*H*Sq*Sk*D/ms/1e9:7.1f}", flush=True)
