"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / profiles/ quote:
    python tests/ncu_summary.py gpurun_out/prof_fa_r1.ncu-rep > profiles/r1_fa_fwd_ncu.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fmalite.sum",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"kernel: {d.get('Kernel Name', '?')[:120]}")
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        tot = sum(int(r[isamp] or 0) for r in data)
        agg = sorted(((sum(int(r[i] or 0) for r in data), hdr[i]) for i in stall), reverse=True)
        print("  stall reasons (all instructions): " + ", ".join(f"{h}={n} ({100.0 * n / max(tot, 1):.1f}%)" for n, h in agg[:10]))
        print(f"  warp-stall samples: {tot}; top instructions:")
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:int(__import__('os').environ.get('NCU_TOP', '12'))]:
            st = sorted(((int(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:2]
            print(f"    {int(r[isamp]):7d} ({100.0 * int(r[isamp]) / max(tot, 1):4.1f}%)  {r[isrc][:64]:64s} {st}")


if __name__ == "__main__":
    main(sys.argv[1])
