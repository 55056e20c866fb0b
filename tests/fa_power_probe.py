"""Developer probe (B200 via gpurun): is the prefill attention kernel cycle-bound or power-bound on this pool?

Runs K1 and cuDNN SDPA at the C3 shape back to back for a few seconds each while NVML samples the SM clock and board
power, and prints TFLOP/s, median clock, mean power and the implied cycles per launch (time x clock). Equal clocks and
different times = a cycle gap; different clocks = an energy-per-FLOP gap under the power cap."""
from __future__ import annotations

import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ml_inference_optimizer_b200 import ops


class Sampler:
    def __init__(self, index=0, period=0.02):
        import pynvml

        self.nv = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.period = period
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def __enter__(self):
        self.rows = []
        self._stop.clear()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(self.period)

    def __exit__(self, *a):
        self._stop.set()
        self._t.join()

    def summary(self, skip=0.3):
        rows = self.rows[int(len(self.rows) * skip):] or self.rows
        if not rows:
            return {}
        clk = sorted(r[0] for r in rows)
        return {"sm_mhz_median": clk[len(clk) // 2], "sm_mhz_min": clk[0], "sm_mhz_max": clk[-1],
                "power_w_mean": round(sum(r[1] for r in rows) / len(rows), 1), "samples": len(rows)}


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    s.record()
    while time.time() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    from torch.nn.attention import SDPBackend, sdpa_kernel

    secs = float(os.environ.get("SECS", "3"))
    shapes = [("c3_causal", 4, 8192, 32, 128, True), ("c3_full", 4, 8192, 32, 128, False), ("c2_causal", 8, 4096, 12, 64, True)]
    samp = Sampler()
    for name, B, S, H, D, causal in shapes:
        torch.manual_seed(0)
        q, k, v = [torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
        qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
        flops = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)

        def ours():
            ops.flash_attn_fwd(q, k, v, causal=causal)

        def cudnn():
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal)

        for impl, fn in (("ours", ours), ("cudnn", cudnn), ("ours_again", ours)):
            time.sleep(1.0)  # let the board cool to the same starting point
            with samp:
                ms = sustained(fn, secs)
            rec = {"probe": "fa_power", "shape": name, "impl": impl, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1)}
            rec.update(samp.summary())
            if "sm_mhz_median" in rec:
                rec["mcycles_per_launch"] = round(ms * 1e-3 * rec["sm_mhz_median"], 1)
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
