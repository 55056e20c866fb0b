"""Black-box comparator run for ncu: cuDNN SDPA at the C3 causal shape (aggregate counters only — launch configuration and
per-pipe instruction counts next to ours; no source / SASS page is read)."""
import torch
from torch.nn.attention import SDPBackend, sdpa_kernel
B, S, H, D = 4, 8192, 32, 128
q, k, v = (torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(2):
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
torch.cuda.synchronize()
print("done")
