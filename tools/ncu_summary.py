#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel headline metrics + hot SASS regions.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--regions kernel_regex]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return hdr, [dict(zip(hdr, r)) for r in rows[2:]]


def main():
    rep = sys.argv[1]
    hdr, kernels = raw(rep)
    for d in kernels:
        print("=====", d.get("Kernel Name"))
        for k in KEYS:
            if k in d:
                print(f"  {k:82s} {d[k]}")
        for k in hdr:
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v > 0.05:
                    print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:.3f}")
    if "--regions" in sys.argv:
        pat = sys.argv[sys.argv.index("--regions") + 1]
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h = rows[1]
        data = rows[2:]
        isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        tot = sum(int(r[iex]) for r in data if len(r) > iex and r[iex].isdigit())
        print(f"instructions {len(data)}  executed warp-instr {tot}")
        step = 64
        for b in range(0, len(data), step):
            blk = [r for r in data[b:b + step] if len(r) > iex and r[iex].isdigit()]
            ex = sum(int(r[iex]) for r in blk)
            sm = sum(int(r[ismp]) for r in blk)
            if ex / max(tot, 1) > 0.004:
                print(f"  {b:5d}  exec {ex / tot * 100:6.2f}%  samples {sm:8d}  {blk[0][isrc].strip()[:50]}")


if __name__ == "__main__":
    main()
