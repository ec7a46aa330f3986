#!/usr/bin/env python
"""Hottest CUDA source lines of one kernel of an .ncu-rep captured with --import-source on (read here, no GPU needed).

    python tools/ncu_hotlines.py gpurun_out/prof.ncu-rep kernel_regex [top_n]

Per source line: share of the executed warp instructions, share of the stall samples, average active lanes.
"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr, lines = "", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            try:
                lines.append((fname, int(r[0]), r[1].strip(), int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
            except (KeyError, ValueError):
                pass
    ti = sum(x[3] for x in lines) or 1
    tt = sum(x[4] for x in lines)
    ts = sum(x[5] for x in lines) or 1
    print(f"warp-instr {ti}  thread-instr {tt}  avg active lanes {tt / ti:.2f}  samples {ts}")
    for f, ln, src, wi, th, sm in sorted(lines, key=lambda x: -x[3])[:top]:
        print(f"{wi / ti * 100:6.2f}% instr {sm / ts * 100:6.2f}% smpl  lanes {th / max(wi, 1):5.1f}  {f}:{ln}: {src[:110]}")


if __name__ == "__main__":
    main()
