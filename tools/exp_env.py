#!/usr/bin/env python
"""Environment-knob experiment: time one workload's frame for several settings of per-launch knobs (and prebuilt library
variants via AUDIORT_LIB, one process per variant).

    python tools/exp_env.py --workload c3 --set ART_Q_SPAN0=1,2,4 --set ART_Q_SPAN_HOLD=0,2 [--shard 0/8] [--rays N]
"""
import argparse
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_raytracer_b200 import build, native, scenes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--rays", type=int, default=None)
    ap.add_argument("--set", action="append", default=[], help="NAME=v1,v2,... ('-' = unset)")
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--shard", default=None)
    ap.add_argument("--chunk", type=int, default=256, help="rays per interleaved shard chunk (0 = one contiguous slice)")
    a = ap.parse_args()
    build.build()
    shard = tuple(int(x) for x in a.shard.split("/")) if a.shard else None
    s = scenes.make_config(a.workload, n_rays=a.rays, batch_count=shard[1] if shard else 1)
    names = [x.split("=")[0] for x in a.set]
    values = [x.split("=")[1].split(",") for x in a.set]
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        flags = native.FRAME_NO_HOST_OUTPUTS
        if shard:
            ctx.set_ray_shard(shard[0], shard[1], a.chunk)
            flags |= native.FRAME_PARTIALS_ONLY
        for combo in itertools.product(*values) if values else [()]:
            for n, v in zip(names, combo):
                if v == "-":
                    os.environ.pop(n, None)
                else:
                    os.environ[n] = v
            ms = []
            for _ in range(a.frames):
                c = ctx.run_frame(s, flags=flags, want=()).counters
                ms.append((c["traceMs"], c["permeationMs"], c["deviceMs"]))
            best = min(ms)
            print(" ".join(f"{n}={v}" for n, v in zip(names, combo)) + f": trace {best[0]:.3f} ms (fan {c['fanBuildMs']:.3f} bounce {c['bounceMs']:.3f} "
                  f"queries {c['queryMs']:.3f}) perm {best[1]:.3f} ms device {best[2]:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
