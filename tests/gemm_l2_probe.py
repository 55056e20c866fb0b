"""L2 raster / eviction-hint probe of the pair GEMM at the C3 shapes. Plain run: interleaved A/B timing. Under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.max` (NCU=1): one launch per variant, in the
order printed, for the DRAM bytes."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ml_inference_optimizer_b200 import ops

dev, bf = "cuda", torch.bfloat16
T, h, i = 32768, 4096, 11008
x = torch.randn(T, h, device=dev, dtype=bf)
wu, wg = (torch.randn(i, h, device=dev, dtype=bf) * 0.02 for _ in range(2))
wd = torch.randn(h, i, device=dev, dtype=bf) * 0.02
mid = torch.randn(T, i, device=dev, dtype=bf)
out1 = torch.empty(T, i, device=dev, dtype=bf)
out2 = torch.empty(T, h, device=dev, dtype=bf)
VARIANTS = [(hints, rows) for rows in (1024, 2048, 4096) for hints in ("000", "200", "210", "201", "211")]


def run(which, hints, rows):
    os.environ["B200_GEMM_L2_HINTS"] = hints
    ops.set_gemm_group_rows(rows)
    if which == 1:
        ops.linear_act(x, wu, None, "swiglu", wg, None, out=out1)
    else:
        ops.linear_act(mid, wd, None, None, out=out2)


if os.environ.get("NCU") == "1":
    for which in (1, 2):
        for hints, rows in VARIANTS:
            run(which, hints, rows)
    torch.cuda.synchronize()
    print(json.dumps({"order": [[w, hn, r] for w in (1, 2) for hn, r in VARIANTS]}))
    sys.exit(0)

for which in (1, 2):
    best = {v: float("inf") for v in VARIANTS}
    for v in VARIANTS:
        run(which, *v)
    for _ in range(4):
        for v in VARIANTS:
            run(which, *v); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(4):
                run(which, *v)
            e.record(); torch.cuda.synchronize()
            best[v] = min(best[v], s.elapsed_time(e) / 4)
    print(json.dumps({"gemm": which, "ms": {f"{hn}_r{r}": round(t, 4) for (hn, r), t in best.items()}}), flush=True)
