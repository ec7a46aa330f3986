"""Scenes that change between frames (`-m gpu`): AudioColliderManager.UpdateJobBatch re-bakes the dynamic colliders every
frame (Audio/AudioColliderManager.cs:115-122, double-buffered arrays of DataTypes/NativeJobBatch.cs:36-50). art_set_scene
diffs the new payload against the previous one, the next frame uploads only the structs that changed and rebuilds the grid's
cell lists on the device -- results must equal a fresh context that uploads everything."""
import time

import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes

pytestmark = pytest.mark.gpu
FLAGS = native.FRAME_FORCE_GRID


def assert_same(a, b, what, exact_sum=True):
    for k in ("hit_counts", "hit_ids", "echo", "hit_points", "muffle", "muffle_totals"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=f"{what}: {k}")
    if exact_sum:
        np.testing.assert_array_equal(a.permeation_sum, b.permeation_sum, err_msg=what)
    else:       # (the brute-force kernels evaluate the loss lines in another order: documented tolerance of the extension output)
        np.testing.assert_allclose(a.permeation_sum, b.permeation_sum, rtol=0, atol=1e-5 * len(a.hit_counts) ** 2, err_msg=what)
    np.testing.assert_array_equal(a.permeation.view(np.uint32), b.permeation.view(np.uint32), err_msg=what)
    np.testing.assert_array_equal(a.settings.view(np.uint8), b.settings.view(np.uint8), err_msg=what)


def fresh_frame(s):
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        return ctx.run_frame(s, flags=FLAGS)


def test_one_moving_collider_per_frame_equals_a_full_upload():
    s = scenes.make_config("c3", n_rays=20000)
    rng = np.random.default_rng(7)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        first = ctx.run_frame(s, flags=FLAGS)
        base_launches = None
        for frame in range(6):
            kind = frame % 3
            arr = (s.aabbs, s.obbs, s.spheres)[kind]
            i = int(rng.integers(6 if kind == 0 else 0, len(arr)))        # (not one of the room's walls)
            # nudge the collider: a few half ulps on its centre (the fields hold raw half bits)
            arr["center"][i] = (arr["center"][i].astype(np.int32) + np.array([3, -2, 5])).astype(np.uint16)
            ctx.set_scene(s.aabbs, s.obbs, s.spheres)                     # same layout, one struct differs
            got = ctx.run_frame(s, flags=FLAGS)
            assert_same(got, fresh_frame(s), f"frame {frame}: incremental upload vs fresh context")
            if base_launches is None:
                base_launches = got.counters["kernelLaunches"]
            assert got.counters["kernelLaunches"] == base_launches
        # an identical payload is recognised: no pack kernel, no grid build for the next frame
        ctx.set_scene(s.aabbs, s.obbs, s.spheres)
        same = ctx.run_frame(s, flags=FLAGS)
        assert same.counters["kernelLaunches"] == base_launches - 5
        assert_same(same, got, "unchanged scene")
    assert first.counters["segments"] > 0


def test_changed_layout_uploads_everything():
    a = scenes.make_config("c2", n_rays=8192)
    b = scenes.make_config("c3", n_rays=8192)
    with native.Context(0) as ctx:
        native.upload(ctx, a)
        ctx.run_frame(a, flags=FLAGS)
        native.upload(ctx, b)                                             # other collider counts
        got = ctx.run_frame(b, flags=FLAGS)
    assert_same(got, fresh_frame(b), "scene with other collider counts")


def test_grid_entry_buffer_overflow_reruns_on_brute_force(monkeypatch):
    s = scenes.make_config("c2", n_rays=16384)
    ref = fresh_frame(s)
    monkeypatch.setenv("ART_GRID_ENTRIES_PER_COLLIDER", "0")          # test knob: room for 16 entries only
    with native.Context(0) as ctx:
        monkeypatch.delenv("ART_GRID_ENTRIES_PER_COLLIDER")
        native.upload(ctx, s)
        got = ctx.run_frame(s, flags=FLAGS)
        assert got.counters["gridUsed"] & 8, "the grid build was expected to overflow its entry buffer"
        assert got.counters["gridUsed"] & 1 == 0, "the second pass runs on the brute-force kernels"
        assert_same(got, ref, "overflowed frame, re-run", exact_sum=False)
        ctx.set_scene(s.aabbs[::-1].copy(), s.obbs, s.spheres)       # (a changed scene: the lists are built again, into a larger buffer)
        ctx.set_scene(s.aabbs, s.obbs, s.spheres)
        again = ctx.run_frame(s, flags=FLAGS)
        assert again.counters["gridUsed"] & 9 == 1, "the grown buffer should fit"
        assert_same(again, ref, "frame after the buffer grew")


def test_schedule_latency_with_a_changed_scene():
    """art_trace_schedule with a changed 4,096-collider scene: no host cell walk, no stream synchronisation."""
    s = scenes.make_config("c3", n_rays=65536)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        ctx.run_frame(s, flags=native.FRAME_NO_HOST_OUTPUTS, want=())
        ts = []
        for k in range(20):
            s.aabbs["center"][100] = (s.aabbs["center"][100].astype(np.int32) + 1).astype(np.uint16)
            ctx.set_scene(s.aabbs, s.obbs, s.spheres)
            t0 = time.perf_counter()
            h = ctx.schedule(s, flags=native.FRAME_NO_HOST_OUTPUTS, want=())
            ts.append(time.perf_counter() - t0)
            ctx.complete(h)
        med = float(np.median(ts[2:]))
        print(f"schedule with a changed scene: median {med * 1e6:.0f} us (ctypes call included)")
        assert med < 2e-3
