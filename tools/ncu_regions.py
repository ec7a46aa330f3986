#!/usr/bin/env python
"""Executed instructions of one kernel of an .ncu-rep (--import-source on) by REGION of its main source file.

    python tools/ncu_regions.py prof.ncu-rep kernel_regex main_file.cu name:line,name:line,...

The SASS is walked in address order; an instruction belongs to the region of the most recent instruction attributed to
main_file (so inlined header code is charged to the call site's region). Regions start at the given lines.
"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, pat, main_file = sys.argv[1], sys.argv[2], sys.argv[3]
    regions = sorted(((int(x.split(":")[1]), x.split(":")[0]) for x in sys.argv[4].split(",")))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    fname, hdr, line, sass = "", None, 0, []
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            line = int(r[0])
        elif hdr and r[0] == "" and len(r) > 2 and r[2].startswith("0x"):
            d = dict(zip(hdr, r))
            try:
                sass.append((int(r[2], 16), fname, line, int(d["Instructions Executed"]), int(d["Thread Instructions Executed"]), int(d["# Samples"])))
            except (KeyError, ValueError):
                pass
    sass.sort(key=lambda x: (x[0], x[1] != main_file))    # (an inlined instruction is listed once per source line it belongs to)
    acc = {}
    cur = "(prologue)"
    seen = set()
    for addr, f, ln, wi, th, sm in sass:
        if addr in seen:
            continue
        seen.add(addr)
        if f == main_file:
            cur = "(prologue)"
            for start, name in regions:
                if ln >= start:
                    cur = name
        a = acc.setdefault(cur, [0, 0, 0])
        a[0] += wi; a[1] += th; a[2] += sm
    ti = sum(v[0] for v in acc.values()) or 1
    ts = sum(v[2] for v in acc.values()) or 1
    print(f"warp-instr {ti}")
    for name, (wi, th, sm) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"  {name:28s} {wi / ti * 100:6.2f}% instr  {sm / ts * 100:6.2f}% samples  lanes {th / max(wi, 1):5.1f}")


if __name__ == "__main__":
    main()
