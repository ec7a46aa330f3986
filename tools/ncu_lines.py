#!/usr/bin/env python
"""Per-CUDA-source-line hot spots of one kernel in an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep trace_grid_kernel [top_n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    lines, fname, hdr = [], "", None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            d = dict(zip(hdr, r))
            ex, th = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])
            if ex:
                lines.append((ex, th, int(d["# Samples"]), fname, int(r[0]), r[1].strip()))
    tot = sum(l[0] for l in lines)
    tth = sum(l[1] for l in lines)
    smp = sum(l[2] for l in lines)
    print(f"warp-instr {tot}  thread-instr {tth}  avg active lanes {tth / max(tot, 1):.2f}  samples {smp}")
    for ex, th, sm, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{ex / tot * 100:6.2f}% instr {sm / max(smp, 1) * 100:6.2f}% smpl  lanes {th / ex:5.1f}  {f}:{ln}: {src[:100]}")


if __name__ == "__main__":
    main()
