#!/usr/bin/env python
"""Generate the committed golden fixtures (tests/golden/*.npz).

The reference (C#/Unity) cannot run in this environment and ships no vectors of its own, so these
fixtures are FROZEN ORACLE OUTPUTS: inputs in the reference's wire layout plus every output array the
oracle (oracle/audiort_oracle.c, pinned by the hand-derived vectors of tests/test_oracle_kat.py) produces
for them. They pin the oracle against regressions (CPU test) and give the CUDA path a fixture that does
not depend on the oracle being built on the GPU box (GPU test).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from audio_raytracer_b200 import scenes  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = {
    # name: (config, rays, batch count, overrides)
    "c2_n768_t3": ("c2", 768, 3, {}),
    "c3_n48_t2": ("c3", 48, 2, {}),
    "c4_n24_t1": ("c4", 24, 1, {}),
    "c2_n512_gated": ("c2", 512, 2, {"max_muffle_hit_distance": 20.0, "max_ray_life": 30.0}),
    # BASELINE config 1: the reference demo level with the demo's own parameters (tools/export_demo_scene.py)
    "c1_demo": ("c1", None, 1, {}),
    "c1_demo_1src_t3": ("c1_1src", None, 3, {}),
}


def make(name):
    cfg, n, T, over = CASES[name]
    s = scenes.make_config(cfg, batch_count=T, n_rays=n)
    for k, v in over.items():
        setattr(s, k, v)
    f = orc.run_frame(s)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        aabbs=s.aabbs.view(np.uint8), obbs=s.obbs.view(np.uint8), spheres=s.spheres.view(np.uint8),
        targets=s.targets, ray_directions=s.ray_directions, ray_origin=s.ray_origin,
        params=np.array([s.max_ray_life, s.max_hits_per_ray, s.max_muffle_hit_distance, s.permeation_strength_per_ray,
                         s.muffle_effectiveness, s.permeation_effectiveness, s.max_reverb_distance, s.batch_count],
                        dtype=np.float64),
        echo=f.echo, hit_points=f.hit_points, hit_counts=f.hit_counts, hit_ids=f.hit_ids, muffle=f.muffle,
        muffle_totals=f.muffle_totals, permeation=f.permeation, permeation_sum=f.permeation_sum,
        settings=f.settings.view(np.uint8), settings_fp64=f.settings_fp64.view(np.uint8),
        counters=np.array([f.counters["segments"], f.counters["segment_hits"]] + f.counters["trace_tests"]
                          + f.counters["echo_tests"] + f.counters["muffle_tests"], dtype=np.uint64))
    print(name, "segments", f.counters["segments"])


if __name__ == "__main__":
    for n in (sys.argv[1:] or CASES):
        make(n)
