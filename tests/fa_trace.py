"""Developer tool: per-iteration clock64 timeline of one CTA of the FA kernel (needs a library built with
-DB200_FA_TRACE: `B200_EXTRA_NVCC_FLAGS=-DB200_FA_TRACE python -m ml_inference_optimizer_b200.build --force`)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ml_inference_optimizer_b200 import _lib, ops

if os.environ.get("B200_TRACE_LIB"):  # a -DB200_FA_TRACE build kept next to the product library
    from pathlib import Path
    _lib.LIB_PATH = Path(os.environ["B200_TRACE_LIB"]).resolve()
lib = _lib.load()
B, S, H, D = (1, 8192, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 128)
causal = (sys.argv[2] != "0") if len(sys.argv) > 2 else False
q, k, v = (torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3))
ops.flash_attn_fwd(q, k, v, causal=causal)
buf = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
fn = lib.b200_debug_fa_trace
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.data_ptr()) == 0
ops.flash_attn_fwd(q, k, v, causal=causal)
torch.cuda.synchronize()
t = buf.cpu().view(4, 64, 8)
base = t[t > 0].min().item()
names = {0: "softmax0", 1: "softmax1", 2: "mma(t=0)", 3: "mma(t=1)"}
lab_s = ["s_ready", "ld_done", "max_done", "half0_exp", "pub0", "p_free", "exp+st", "published"]
lab_m = ["wait_p", "p_ready", "pv_issued", "qk_issued", "qk_wait"]
for slot in range(4):
    print("==", names[slot], lab_s if slot < 2 else lab_m)
    for j in range(8, 12):
        row = t[slot, j]
        vals = [(x.item() - base) if x.item() > 0 else None for x in row[: (8 if slot < 2 else 5)]]
        print(f"  j={j:2d} " + " ".join(f"{v:7d}" if v is not None else "      -" for v in vals))
for slot in (0, 1):
    d = (t[slot, 9:40, 0] - t[slot, 8:39, 0]).float()
    print(names[slot], "period mean", d.mean().item(), "min", d.min().item(), "max", d.max().item())
    for a, b in ((0, 1), (1, 3), (3, 2), (2, 5), (5, 4), (4, 6), (6, 7)):
        print(f"   {lab_s[a]:10s} -> {lab_s[b]:10s} {(t[slot, 8:40, b] - t[slot, 8:40, a]).float().mean().item():8.1f}")
    print(f"   pub1_done -> next s_ready {(t[slot, 9:40, 0] - t[slot, 8:39, 7]).float().mean().item():8.1f}")
for slot in (2, 3):
    for a, b, name in ((4, 3, "wait S free + issue QK(j)"), (0, 1, "wait P(j)"), (1, 2, "issue PV(j)")):
        print(f"   {names[slot]} {name:10s} {(t[slot, 8:40, b] - t[slot, 8:40, a]).float().mean().item():8.1f}")
