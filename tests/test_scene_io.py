"""The binary scene dump (scene_io.py; BASELINE config 1 'exported to a binary dump') round-trips, and the oracle's
outputs survive the output container the C# harness writes."""
import numpy as np

from audio_raytracer_b200 import scene_io, scenes


def test_dump_round_trip(tmp_path):
    for name, n in (("c1", None), ("c2", 64), ("c3", 16)):
        s = scenes.make_config(name, batch_count=2, n_rays=n)
        p = str(tmp_path / f"{name}.artd")
        scene_io.write_dump(s, p)
        r = scene_io.read_dump(p)
        for k in ("aabbs", "obbs", "spheres"):
            assert getattr(s, k).tobytes() == getattr(r, k).tobytes()
        np.testing.assert_array_equal(s.targets, r.targets)
        np.testing.assert_array_equal(s.ray_directions, r.ray_directions)
        np.testing.assert_array_equal(s.ray_origin, r.ray_origin)
        assert (s.max_hits_per_ray, s.batch_count, s.max_ray_life, s.max_muffle_hit_distance) == \
               (r.max_hits_per_ray, r.batch_count, r.max_ray_life, r.max_muffle_hit_distance)


def test_outputs_round_trip(tmp_path, oracle):
    s = scenes.make_config("c1")
    f = oracle.run_frame(s)
    p = str(tmp_path / "c1.arto")
    scene_io.write_outputs(p, f.echo, f.hit_points, f.hit_counts, f.muffle, f.permeation, f.settings)
    o = scene_io.read_outputs(p, s)
    np.testing.assert_array_equal(o["echo"], f.echo)
    np.testing.assert_array_equal(o["hit_points"].reshape(-1), np.asarray(f.hit_points).reshape(-1))
    np.testing.assert_array_equal(o["settings"].view(np.uint8), f.settings.view(np.uint8))
