#!/usr/bin/env python
"""Run a few frames of one workload through the C ABI (for ncu / compute-sanitizer captures).

    python tools/prof_frame.py --workload c3 --rays 32768 --frames 2 [--jobs 7] [--counters]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from audio_raytracer_b200 import build, native, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--rays", type=int, default=None)
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--jobs", type=int, default=native.JOB_ALL)
    ap.add_argument("--counters", action="store_true")
    ap.add_argument("--seq", action="store_true")
    ap.add_argument("--force-grid", action="store_true")
    ap.add_argument("--brute", action="store_true")
    ap.add_argument("--muffle-dist", type=float, default=None, help="override MaxMuffleHitDistance (0 gates every muffle ray out)")
    ap.add_argument("--no-fans", action="store_true", help="grid kernels: every query walks the grid (no target fans)")
    ap.add_argument("--stats", action="store_true", help="ART_FRAME_GRID_STATS: tests / cells the grid kernels execute")
    ap.add_argument("--outputs", action="store_true", help="also produce hit points / ids / counts and copy everything back")
    ap.add_argument("--shard", default=None, help="I/W: trace only shard I of a W-way ray-sharded frame (what one rank of a W-GPU run executes)")
    a = ap.parse_args()
    build.build()
    shard = tuple(int(x) for x in a.shard.split("/")) if a.shard else None
    s = scenes.make_config(a.workload, n_rays=a.rays, batch_count=shard[1] if shard else 1)
    if a.muffle_dist is not None:
        s.max_muffle_hit_distance = a.muffle_dist
    flags = (0 if a.outputs else native.FRAME_NO_HOST_OUTPUTS) | (native.FRAME_COUNTERS if a.counters else 0) | (native.FRAME_REVERB_SEQ_FP32 if a.seq else 0)
    flags |= (native.FRAME_FORCE_GRID if a.force_grid else 0) | (native.FRAME_BRUTE_FORCE if a.brute else 0)
    flags |= (native.FRAME_NO_FANS if a.no_fans else 0) | (native.FRAME_GRID_STATS if a.stats else 0)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        if shard:
            ctx.set_ray_shard(shard[0], shard[1], 256)
            flags |= native.FRAME_PARTIALS_ONLY
        for i in range(a.frames):
            r = ctx.run_frame(s, jobs=a.jobs, flags=flags, want=("echo", "hit_points", "hit_counts", "hit_ids") if a.outputs else ())
            c = r.counters
            print(f"frame {i}: segments {c['segments']} trace {c['traceMs']:.3f} ms perm {c['permeationMs']:.3f} ms "
                  f"reduce {c['reduceMs']:.3f} ms launches {c['kernelLaunches']} gridUsed {c['gridUsed']}")
            if a.stats:
                print(f"   trace tests S/A/O {c['gridTraceTests']} cells {c['gridTraceCells']}; perm first {c['gridPermFirstTests']} "
                      f"loss {c['gridPermLossTests']} cells {c['gridPermCells']}")


if __name__ == "__main__":
    main()
