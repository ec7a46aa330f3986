#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of a workload's frame (read here, no GPU needed):

    python tools/make_traffic.py gpurun_out/full.ncu-rep c3:1048576:1 [profiles/traffic.json]

Per kernel: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), issue-slot and lane utilisation. The entry
carries the SHA-1 of the kernel sources it was captured from; bench.py reports it only while that still matches.
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha1():
    h = hashlib.sha1()
    cs = os.path.join(ROOT, "audio-raytracer_b200", "csrc")
    for f in sorted(os.listdir(cs)):
        h.update(open(os.path.join(cs, f), "rb").read())
    return h.hexdigest()


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * scale.get(unit, 1)


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    traffic, util = {}, {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
        rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        traffic[name] = int(rd + wr)
        util[name] = {"gpu_time_ms": float(d["gpu__time_duration.sum"]) * {"usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u["gpu__time_duration.sum"], 1.0),
                      "issue_slots_busy_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                      "active_lanes_per_instruction": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                      "warp_instructions": float(d["smsp__inst_executed.sum"]),
                      "dram_read_bytes": int(rd), "dram_write_bytes": int(wr)}
    try:
        allt = json.load(open(out))
    except Exception:
        allt = {}
    allt = {k: v for k, v in allt.items() if isinstance(v, dict) and "csrc_sha1" in v}      # (drop entries of the old format)
    allt[key] = {"csrc_sha1": csrc_sha1(), "source": os.path.basename(rep), "traffic": traffic, "ncu": util}
    json.dump(allt, open(out, "w"), indent=1)
    print(json.dumps(allt[key], indent=1))


if __name__ == "__main__":
    main()
