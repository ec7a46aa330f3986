#!/usr/bin/env python
"""K1 launch-shape experiment: time the trace job of one shard of a W-way ray-sharded C3 frame for several warps-per-CTA
(ART_K1_WARPS, read per frame by trace_grid_plan).   python tools/exp_k1_warps.py [--workload c3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_raytracer_b200 import build, native, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--worlds", default="8,4,2,1")
    ap.add_argument("--warps", default="0,24,22,20,18,16,14,12")
    ap.add_argument("--frames", type=int, default=3)
    a = ap.parse_args()
    build.build()
    for world in [int(x) for x in a.worlds.split(",")]:
        s = scenes.make_config(a.workload, batch_count=world)
        with native.Context(0) as ctx:
            native.upload(ctx, s)
            flags = native.FRAME_NO_HOST_OUTPUTS
            if world > 1:
                ctx.set_ray_shard(0, world, 256)
                flags |= native.FRAME_PARTIALS_ONLY
            for w in [int(x) for x in a.warps.split(",")]:
                if w:
                    os.environ["ART_K1_WARPS"] = str(w)
                else:
                    os.environ.pop("ART_K1_WARPS", None)
                ms = []
                for i in range(a.frames):
                    c = ctx.run_frame(s, flags=flags, want=()).counters
                    ms.append((c["traceMs"], c["permeationMs"], c["deviceMs"]))
                best = min(ms)
                print(f"world {world} local rays {ctx.local_ray_count()} warps {w or 'auto'}: trace {best[0]:.3f} ms perm {best[1]:.3f} ms device {best[2]:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
