"""Dev tool: LayerNorm(+residual) bandwidth at the shapes the MLP/attention blocks see."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_inference_optimizer_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rows, cols in ((32768, 768), (32768, 1024), (32768, 2048), (32768, 4096), (32768, 8192)):
    x = torch.randn(rows, cols, device="cuda", dtype=torch.bfloat16); r = torch.randn_like(x)
    w = torch.randn(cols, device="cuda", dtype=torch.bfloat16); b = torch.randn_like(w)
    # what a plain device copy of the same footprint reaches (the size-matched roofline: MEASURED_PEAKS' 6555 GB/s is a 4 GB copy)
    yc = torch.empty_like(x); ts = []
    for _ in range(12):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); yc.copy_(x); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(f"COPY rows={rows} cols={cols}: {ms*1e3:7.1f} us  {rows * cols * 4 / ms / 1e6:7.0f} GB/s (torch copy_ of the same tensor, read + write)", flush=True)
    variants = [(None, "1", "", ""), (r, "1", "", "")]
    if cols > 1024:
        variants += [(rr, "1", nm, pc) for rr in (None, r) for nm in (["512"] if cols <= 2048 else [""]) for pc in ("", "2", "3", "4")]
    for res, pf, nm, pc in variants:
        os.environ["B200_LN_PREFETCH"] = pf
        os.environ["B200_LN_NARROW_MAX"] = nm
        os.environ["B200_LN_WIDE_CTAS_PER_SM"] = pc
        ts = []
        for _ in range(12):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); ops.layernorm(x, w, b, 1e-5, residual=res); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = sorted(ts)[len(ts) // 2]
        nbytes = rows * cols * 2 * (2 + (res is not None))
        print(f"LN rows={rows} cols={cols} residual={res is not None} prefetch={pf} narrow_max={nm or "dflt"} wide_ctas={pc or "occ"}: {ms*1e3:7.1f} us  {nbytes/ms/1e6:7.0f} GB/s", flush=True)
