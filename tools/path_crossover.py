#!/usr/bin/env python
"""Where does the uniform-grid path overtake the brute-force scans? (sets ART_GRID_MIN_COLLIDERS)

    python tools/path_crossover.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from audio_raytracer_b200 import build, native, scenes  # noqa: E402


def main():
    build.build()
    with native.Context(0) as ctx:
        for n_rays in (314, 4096, 65536):
            for nt in (1, 16):
                for nc in (64, 128, 192, 256, 384, 512, 1024):
                    na, no, ns = nc * 4 // 7, max(nt, nc * 2 // 7), nc // 7
                    s = scenes.make_scene(n_aabb=max(6, na), n_obb=no, n_sphere=ns, n_targets=nt, seed=5000 + nc, n_rays=n_rays, max_hits=8)
                    native.upload(ctx, s)
                    out = []
                    for flags in (native.FRAME_BRUTE_FORCE, native.FRAME_FORCE_GRID):
                        best = 1e9
                        for _ in range(4):
                            c = ctx.run_frame(s, flags=flags | native.FRAME_NO_HOST_OUTPUTS, want=()).counters
                            best = min(best, c["deviceMs"])
                        out.append(best)
                    print(f"rays {n_rays:6d} targets {nt:3d} colliders {s.n_colliders:5d}: brute {out[0]:8.3f} ms  grid {out[1]:8.3f} ms  -> {'grid' if out[1] < out[0] else 'brute'}")


if __name__ == "__main__":
    main()
