"""Developer probe: per-operation device times of ONE rank's ring-attention steps at the N=8, 128K shape (run on 1 GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ml_inference_optimizer_b200 import ops


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n


N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
S, H, D = 131072 // N, 32, 128
half = S // 2
bf = torch.bfloat16
q, k, v = (torch.randn(1, S, H, D, device="cuda", dtype=bf) for _ in range(3))
fl = lambda sq, sk, c: 4.0 * H * D * sq * sk * (0.5 if c else 1.0)
for name, fn, f in (
    ("local causal   q%d k%d" % (S, S), lambda: ops.flash_attn_fwd(q, k, v, causal=True, return_lse=True), fl(S, S, True)),
    ("src<r  full    q%d k%d" % (S, half), lambda: ops.flash_attn_fwd(q, k[:, :half], v[:, :half], return_lse=True), fl(S, half, False)),
    ("src>r  full    q%d k%d" % (half, S), lambda: ops.flash_attn_fwd(q[:, half:], k, v, return_lse=True), fl(half, S, False)),
):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms  {f / ms / 1e9:.0f} TF/s")
o, lse = ops.flash_attn_fwd(q, k, v, return_lse=True)
acc = o.float().contiguous(); lacc = lse.clone()
acc_h = o[:, :half].float().contiguous(); lacc_h = lse[:, :, :half].contiguous()
print(f"lse_merge {S} rows: {t(lambda: ops.lse_merge(acc, lacc, o, lse)):.3f} ms")
print(f"lse_merge {half} rows (strided halves): {t(lambda: ops.lse_merge(acc_h, lacc_h, o[:, half:], lse[:, :, half:].contiguous())):.3f} ms")
print(f"o.float().contiguous(): {t(lambda: o.float().contiguous()):.3f} ms")
print(f"cast_out: {t(lambda: ops.cast_out(acc, bf)):.3f} ms")
print(f"torch.stack([k,v]): {t(lambda: torch.stack([k, v]).contiguous()):.3f} ms")
print(f"torch.cat halves: {t(lambda: torch.cat([o[:, :half], o[:, half:]], dim=1)):.3f} ms")
