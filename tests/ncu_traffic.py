"""Update profiles/kernel_traffic.json (what bench.py's roofline.traffic reads) from .ncu-rep files of `ncu --set full`:
    python tests/ncu_traffic.py gpurun_out/prof_gemm1_r2.ncu-rep [more.ncu-rep ...]
Per kernel (demangled name, template activation folded in): DRAM bytes read + written per launch (mean over the captured
launches), duration, tensor-pipe utilisation."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "kernel_traffic.json")
ACTS = {"0": "NONE", "1": "GELU_TANH", "2": "GELU_ERF", "3": "RELU", "4": "SWIGLU"}


def short_name(full):
    m = re.search(r"(\w+)<\(?(?:int\))?(\d+)", full)
    base = re.search(r"(?:\w+::)*(\w+)\s*[<(]", full)
    name = base.group(1) if base else full
    if name.startswith("gemm_act") or name.startswith("fused_mlp"):
        a = re.search(r"<\(int\)(\d+)|<(\d+)", full)
        act = (a.group(1) or a.group(2)) if a else None
        return f"{name}<{ACTS.get(act, act)}>"
    return name


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main(paths):
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for path in paths:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        per = {}
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = short_name(d.get("Kernel Name", "?"))
            rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
            wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            dur = float(d["gpu__time_duration.sum"].replace(",", ""))
            dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u["gpu__time_duration.sum"], 1e-6)
            tens = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
            per.setdefault(name, []).append((rd, wr, dur_ms, float(tens) if tens else None, d.get("launch__grid_size")))
        for name, ls in per.items():
            n = len(ls)
            table[name] = {"dram_bytes": sum(a + b for a, b, *_ in ls) / n, "dram_bytes_read": sum(a for a, *_ in ls) / n,
                           "dram_bytes_write": sum(b for _, b, *_ in ls) / n, "ms_under_ncu": sum(c for _, _, c, *_ in ls) / n,
                           "tensor_pipe_pct": (sum(t for *_, t, _ in ls if t is not None) / n) if ls[0][3] is not None else None,
                           "grid": ls[0][4], "launches_captured": n, "source": "profiles/" + os.path.basename(path).replace(".ncu-rep", ".txt")}
    json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(table, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1:])
