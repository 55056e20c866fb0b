"""Developer probe (B200): where does the end-to-end step time go? Pure H2D / D2H / duplex PCIe rates with the bench's
buffers, then the 3-stream pipeline with (a) preallocated outputs through ops.* and (b) the nn.Module API."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_inference_optimizer_b200 import ops
from ml_inference_optimizer_b200.kernels.attention.flash_attention import FlashAttention3, FlashAttentionConfig
from ml_inference_optimizer_b200.kernels.mlp.fused_mlp import FusedMLPConfig, FusedMLPSwiGLU

dev = torch.device("cuda:0"); bf = torch.bfloat16
B, S, H, D, h, i = 4, 8192, 32, 128, 4096, 11008
T = B * S
q, k, v = (torch.randn(B, S, H, D, device=dev, dtype=bf) for _ in range(3))
x = torch.randn(T, h, device=dev, dtype=bf)
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
hq, hk, hv, hx = pin(q), pin(k), pin(v), pin(x)
ho = torch.empty(q.shape, dtype=bf, pin_memory=True); hy = torch.empty((T, h), dtype=bf, pin_memory=True)
s1, s2, s3 = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for s in (s1, s2, s3): s.wait_event(a)
    for _ in range(n): fn()
    for s in (s1, s2, s3): torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def h2d():
    with torch.cuda.stream(s1):
        q.copy_(hq, non_blocking=True); k.copy_(hk, non_blocking=True); v.copy_(hv, non_blocking=True); x.copy_(hx, non_blocking=True)
o = torch.empty_like(q); y = torch.empty_like(x)
def d2h():
    with torch.cuda.stream(s3):
        ho.copy_(o, non_blocking=True); hy.copy_(y, non_blocking=True)
def both():
    h2d(); d2h()
nb_in = sum(t.numel() * 2 for t in (q, k, v, x)); nb_out = sum(t.numel() * 2 for t in (o, y))
for name, fn, nb in (("h2d", h2d, nb_in), ("d2h", d2h, nb_out), ("duplex", both, nb_in + nb_out)):
    ms = timed(fn)
    print(json.dumps({"probe": "pcie", "case": name, "ms": round(ms, 3), "GBps": round(nb / ms / 1e6, 1)}), flush=True)

wu, wg = ((torch.randn(i, h, device=dev) * 0.02).to(bf) for _ in range(2)); wd = (torch.randn(h, i, device=dev) * 0.02).to(bf)
bu, bg = (torch.zeros(i, device=dev, dtype=bf) for _ in range(2)); bd = torch.zeros(h, device=dev, dtype=bf)
attn = FlashAttention3(FlashAttentionConfig(causal=True, precision="bf16"))
mlp = FusedMLPSwiGLU(h, i, FusedMLPConfig(activation_fn="gelu", precision="bf16")).to(dev, bf)
sets = [tuple(torch.empty_like(t) for t in (q, k, v, x, q, x)) for _ in range(2)]

def pipeline(n_steps, mode):
    h2d_done, cmp_done, d2h_done = {}, {}, {}
    outs = [None, None]
    for st in range(n_steps):
        dq, dk, dv, dx, do, dy = sets[st % 2]
        with torch.cuda.stream(s1):
            if st >= 2: s1.wait_event(cmp_done[st - 2])
            dq.copy_(hq, non_blocking=True); dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True); dx.copy_(hx, non_blocking=True)
            h2d_done[st] = torch.cuda.Event(); h2d_done[st].record(s1)
        with torch.cuda.stream(s2):
            s2.wait_event(h2d_done[st])
            if st >= 2: s2.wait_event(d2h_done[st - 2])
            if mode == "ops":
                ops.flash_attn_fwd(dq, dk, dv, causal=True, out=do)
                ops.fused_mlp(dx, wu, bu, wd, bd, "swiglu", wg, bg, out=dy)
            else:
                do = attn(dq, dk, dv); dy = mlp(dx)
                outs[st % 2] = (do, dy)
            cmp_done[st] = torch.cuda.Event(); cmp_done[st].record(s2)
        with torch.cuda.stream(s3):
            s3.wait_event(cmp_done[st])
            ho.copy_(do, non_blocking=True); hy.copy_(dy, non_blocking=True)
            d2h_done[st] = torch.cuda.Event(); d2h_done[st].record(s3)

for mode in ("ops", "module", "ops", "module"):
    t0 = time.time()
    ms = timed(lambda: pipeline(8, mode), n=1) / 8
    print(json.dumps({"probe": "e2e_pipeline", "mode": mode, "ms_per_step": round(ms, 3), "host_s": round(time.time() - t0, 3)}), flush=True)
