#!/usr/bin/env python
"""Write a BASELINE config as a binary dump (scene_io.py format), or diff a harness output against the oracle.

    python tools/scene_dump.py write c1 /tmp/c1.artd [--rays N] [--batches T]
    python tools/scene_dump.py oracle /tmp/c1.artd /tmp/c1.oracle.arto      # reference outputs from the C oracle
    python tools/scene_dump.py diff /tmp/c1.artd /tmp/c1.csharp.arto        # harness (csharp/harness) vs oracle, bit for bit
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from audio_raytracer_b200 import scene_io, scenes  # noqa: E402


def main():
    cmd = sys.argv[1]
    if cmd == "write":
        name, path = sys.argv[2], sys.argv[3]
        rays = int(sys.argv[sys.argv.index("--rays") + 1]) if "--rays" in sys.argv else None
        T = int(sys.argv[sys.argv.index("--batches") + 1]) if "--batches" in sys.argv else 1
        s = scenes.make_config(name, batch_count=T, n_rays=rays)
        scene_io.write_dump(s, path)
        print(f"{path}: {len(s.aabbs)} AABB, {len(s.obbs)} OBB, {len(s.spheres)} spheres, {s.n_targets} targets, {s.n_rays} rays")
        return
    from oracle import oracle as orc
    s = scene_io.read_dump(sys.argv[2])
    f = orc.run_frame(s, threads=os.cpu_count() or 1)
    if cmd == "oracle":
        scene_io.write_outputs(sys.argv[3], f.echo, f.hit_points, f.hit_counts, f.muffle, f.permeation, f.settings)
        print("wrote", sys.argv[3])
    elif cmd == "diff":
        o = scene_io.read_outputs(sys.argv[3], s)
        bad = 0
        for k in ("echo", "hit_points", "hit_counts", "muffle"):
            n = int((np.asarray(getattr(f, k)).reshape(-1) != o[k].reshape(-1)).sum())
            print(f"{k:12s} {n} differing entries of {o[k].size}")
            bad += n
        n = int((f.permeation.view(np.uint32) != o["permeation"].view(np.uint32)).sum())
        print(f"{'permeation':12s} {n} differing entries of {o['permeation'].size}")
        bad += n
        n = int((f.settings.view(np.uint8) != o["settings"].view(np.uint8)).sum())
        print(f"{'settings':12s} {n} differing bytes")
        sys.exit(1 if bad + n else 0)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
