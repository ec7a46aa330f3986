"""BASELINE config 1: the reference demo level (tools/export_demo_scene.py -> audio-raytracer_b200/data/c1_demo_scene.npz)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from audio_raytracer_b200 import scenes  # noqa: E402
from audio_raytracer_b200.layouts import f16tof32  # noqa: E402

REFERENCE = "/root/reference"


def test_demo_scene_shape_and_params():
    s = scenes.make_config("c1")
    # run-time registration state of Sample Scene.unity ("Environment (Box)" root is inactive)
    assert (len(s.aabbs), len(s.obbs), len(s.spheres), s.n_targets) == (52, 38, 8, 2)
    # Prefabs/Player.prefab:223-234 + PrefabInstance overrides in the scene
    assert (s.n_rays, s.max_hits_per_ray, s.batch_count) == (314, 5, 1)
    assert (s.max_ray_life, s.max_muffle_hit_distance, s.max_reverb_distance) == (125.0, 250.0, 35.0)
    assert (s.muffle_effectiveness, s.permeation_effectiveness, s.permeation_strength_per_ray) == (1.0, 0.0, 1.0)
    np.testing.assert_array_equal(s.ray_origin, np.float32([15.51, -2.1, -3.11]) + np.float32([0, 0.65, 0]))
    # each MusicBox owns the OBB that shares its GameObject (AudioCollider.cs:29-36)
    owned = s.obbs[s.obbs["audioTargetId"] >= 0]
    assert sorted(owned["audioTargetId"].tolist()) == [0, 1]
    for o in owned:
        np.testing.assert_allclose(f16tof32(o["center"]), s.targets[o["audioTargetId"]], atol=2e-2)
    assert (s.aabbs["audioTargetId"] == -1).all() and (s.spheres["audioTargetId"] == -1).all()
    assert scenes.make_config("c1_1src").n_targets == 1


def test_exporter_quaternion_helpers():
    import export_demo_scene as E
    # Quaternion.Euler(0, 90, 0) = (0, sin45, 0, cos45); rotating +z by it gives +x
    q = E.qeuler([0, 90, 0])
    np.testing.assert_allclose(q, [0, np.sqrt(0.5), 0, np.sqrt(0.5)], atol=1e-6)
    np.testing.assert_allclose(E.qrot(q, [0, 0, 1]), [1, 0, 0], atol=1e-6)
    # Unity order z, x, y: Euler(90, 90, 0) == Euler(0, 90, 0) * Euler(90, 0, 0)
    np.testing.assert_allclose(E.qeuler([90, 90, 0]), E.qmul(E.qeuler([0, 90, 0]), E.qeuler([90, 0, 0])), atol=1e-6)
    # halfQuaternion round trip keeps w >= 0 and stays unit length; inverse flips xyz
    bits = E.half_quat_store(np.float32([0.1, -0.2, 0.3, -0.9273618]))
    ql = E.half_quat_load(bits)
    assert ql[3] > 0 and abs(np.linalg.norm(ql) - 1) < 1e-6
    np.testing.assert_allclose(ql[:3], [-0.1, 0.2, -0.3], atol=1e-3)
    np.testing.assert_allclose(E.qinverse(ql), ql * np.float32([-1, -1, -1, 1]), atol=1e-6)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not mounted (GPU box)")
def test_committed_dump_matches_a_fresh_export():
    import export_demo_scene as E
    e = E.export(REFERENCE)
    s = scenes.make_config("c1")
    for k in ("aabbs", "obbs", "spheres"):
        assert e[k].tobytes() == getattr(s, k).tobytes(), k
    np.testing.assert_array_equal(e["targets"], s.targets)
    np.testing.assert_array_equal(e["ray_origin"], s.ray_origin)


@pytest.mark.gpu
@pytest.mark.parametrize("name,T", [("c1", 1), ("c1", 4), ("c1_1src", 1)])
def test_cuda_matches_oracle_on_demo_scene(gpu_ctx, oracle, name, T):
    from audio_raytracer_b200 import native
    s = scenes.make_config(name, batch_count=T)
    f = oracle.run_frame(s)
    native.upload(gpu_ctx, s)
    r = gpu_ctx.run_frame(s, flags=native.FRAME_COUNTERS | native.FRAME_REVERB_SEQ_FP32)
    for k in ("echo", "hit_counts", "hit_ids", "muffle", "muffle_totals"):
        np.testing.assert_array_equal(getattr(r, k), getattr(f, k), err_msg=k)
    np.testing.assert_array_equal(r.permeation.view(np.uint32), f.permeation.view(np.uint32))
    np.testing.assert_array_equal(r.settings.view(np.uint8), f.settings.view(np.uint8))
    assert r.counters["segments"] == f.counters["segments"]
    for flags in (native.FRAME_REVERB_SEQ_FP32, native.FRAME_REVERB_SEQ_FP32 | native.FRAME_FORCE_GRID):
        g = gpu_ctx.run_frame(s, flags=flags)                  # default (brute force at 98 colliders) and forced grid
        for k in ("echo", "hit_counts", "hit_ids", "muffle"):
            np.testing.assert_array_equal(getattr(g, k), getattr(f, k), err_msg=k)
        np.testing.assert_array_equal(g.settings.view(np.uint8), f.settings.view(np.uint8))
