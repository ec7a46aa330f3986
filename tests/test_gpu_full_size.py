"""Size-true parity of every BASELINE.json configuration (`-m gpu`): the default device path (uniform grid + target fans +
query kernel + binned loss lines) against the brute-force scans bit for bit at the configuration's full size, and against
the oracle on ray windows cut out of the full batch with the library's own shard map (so that every N-dependent constant --
PM:260's RayDirections.Length, the batch size of ART:161 -- stays the full configuration's)."""
import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes

pytestmark = pytest.mark.gpu


def window_frame(scene, k, n_windows, flags=0):
    """GPU frame of window k of n_windows contiguous ray slices of the FULL batch: per-ray outputs (local indexing) and the
    window's per-source totals (its partial blob finalised on its own)."""
    with native.Context(0) as ctx:
        native.upload(ctx, scene)
        ctx.set_ray_shard(k, n_windows, 0)                       # chunkRays = 0: one contiguous slice per shard (ART:161)
        r = ctx.run_frame(scene, flags=native.FRAME_PARTIALS_ONLY | flags)
        tot = native.finalize(ctx.get_partials(scene.n_targets, scene.batch_count), scene, scene.n_rays)
    return r, tot


def assert_window_equals_oracle(scene, oracle, k, n_windows, pm_rays, flags):
    N, H = scene.n_rays, scene.max_hits_per_ray
    per = (N + n_windows - 1) // n_windows
    first, count = k * per, min(per, N - k * per)
    g, tot = window_frame(scene, k, n_windows, flags)
    assert len(g.hit_counts) == count
    o = oracle.trace_range(scene, first, count, threads=16)
    sl = slice(first * H, (first + count) * H)
    np.testing.assert_array_equal(g.hit_ids, o.hit_ids[sl])
    np.testing.assert_array_equal(g.echo, o.echo[sl])
    np.testing.assert_array_equal(g.hit_points, o.hit_points[sl])
    np.testing.assert_array_equal(g.hit_counts, o.hit_counts[first:first + count])
    np.testing.assert_array_equal(tot.muffle_totals, o.muffle_totals)              # the window's muffle hits per source
    assert g.counters["segments"] == o.counters["segments"]
    return g, tot, first, count


@pytest.mark.parametrize("k", [0, 341, 682, 1023])
def test_c3_windows_of_1024_rays_equal_the_oracle(oracle, k):
    """C3 (64 sources, 1,048,576 rays x 12 hits, 4,096 colliders): 4 windows of 1,024 rays of the full batch on the default
    path (grid + fans forced for the small slice) -- hit ids, echo halves, hit points, counts, the window's muffle totals per
    source, and the permeation sums of its first 128 rays -- against the oracle."""
    s = scenes.make_config("c3")
    g, tot, first, count = assert_window_equals_oracle(s, oracle, k, 1024, 128, native.FRAME_FORCE_GRID)
    assert g.counters["gridUsed"] & 5 == 5
    # permeation: a narrower window (the loss lines have no early exit: 65 x 4,096 tests per ray on one host thread)
    n_pm = 128
    g2, tot2 = window_frame(s, k * 8, 1024 * 8, native.FRAME_FORCE_GRID)            # the first 128 rays of the same window
    sums = np.zeros(s.n_targets, np.float64)
    c = oracle.permeation_range(s, first, n_pm, threads=1, sums=sums)
    assert g2.counters["permHitRays"] == c["perm_hit_rays"]
    np.testing.assert_allclose(tot2.permeation_sum, sums, rtol=0, atol=1e-5 * s.n_rays * s.permeation_strength_per_ray * n_pm)


def test_c4_full_size_default_path_equals_brute_force_and_oracle_windows(gpu_ctx, oracle):
    """C4 (permeation / reverb stress: 256 sources, 65,536 rays, 4,096 colliders, full reduction) at full size: default path
    == brute-force scans bit for bit; two windows against the oracle."""
    s = scenes.make_config("c4")
    assert s.n_targets == 256 and s.n_rays == 65536
    native.upload(gpu_ctx, s)
    fast = gpu_ctx.run_frame(s)
    slow = gpu_ctx.run_frame(s, flags=native.FRAME_BRUTE_FORCE)
    assert fast.counters["gridUsed"] & 7 == 7 and slow.counters["gridUsed"] == 0
    for k in ("hit_counts", "hit_ids", "echo", "hit_points", "muffle", "muffle_totals"):
        np.testing.assert_array_equal(getattr(fast, k), getattr(slow, k), err_msg=k)
    np.testing.assert_array_equal(fast.permeation.view(np.uint32), slow.permeation.view(np.uint32))
    np.testing.assert_array_equal(fast.settings.view(np.uint8), slow.settings.view(np.uint8))
    np.testing.assert_allclose(fast.permeation_sum, slow.permeation_sum, rtol=1e-9, atol=1e-5 * s.n_rays * s.n_rays)
    for k in (3, 200):
        g, tot, first, count = assert_window_equals_oracle(s, oracle, k, 256, 0, native.FRAME_FORCE_GRID)   # 256 rays each
        sums = np.zeros(s.n_targets, np.float64)
        n_pm = 32
        g2, tot2 = window_frame(s, k * 8, 256 * 8, native.FRAME_FORCE_GRID)
        c = oracle.permeation_range(s, first, n_pm, threads=1, sums=sums)
        assert g2.counters["permHitRays"] == c["perm_hit_rays"]
        np.testing.assert_allclose(tot2.permeation_sum, sums, rtol=0, atol=1e-5 * s.n_rays * s.permeation_strength_per_ray * n_pm)


def test_c5_one_rank_shard_of_the_full_batch(oracle):
    """C5 (8 sources, 16,777,216 rays x 16 hits, 16,384 colliders -- the geometry does not fit shared memory): the shard one
    of 16 ranks traces (1,048,576 rays, interleaved chunks of 256) on the default path == brute-force scans bit for bit, and
    a 256-ray window of the full batch against the oracle."""
    s = scenes.make_config("c5")
    assert s.n_rays == 1 << 24 and s.n_colliders == 16384
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        ctx.set_ray_shard(5, 16, 256)
        assert ctx.local_ray_count() == 1 << 20
        fast = ctx.run_frame(s, flags=native.FRAME_PARTIALS_ONLY, want=("echo", "hit_counts", "hit_ids"))
        bf = native.finalize(ctx.get_partials(s.n_targets, s.batch_count), s, s.n_rays)
        slow = ctx.run_frame(s, flags=native.FRAME_PARTIALS_ONLY | native.FRAME_BRUTE_FORCE, want=("echo", "hit_counts", "hit_ids"))
        bs = native.finalize(ctx.get_partials(s.n_targets, s.batch_count), s, s.n_rays)
    assert fast.counters["gridUsed"] & 7 == 7 and slow.counters["gridUsed"] == 0
    assert fast.counters["segments"] == slow.counters["segments"] > 10_000_000
    for k in ("hit_counts", "hit_ids", "echo"):
        np.testing.assert_array_equal(getattr(fast, k), getattr(slow, k), err_msg=k)
    np.testing.assert_array_equal(bf.muffle_totals, bs.muffle_totals)
    np.testing.assert_array_equal(bf.permeation.view(np.uint32), bs.permeation.view(np.uint32))
    np.testing.assert_allclose(bf.permeation_sum, bs.permeation_sum, rtol=1e-9, atol=1e-5 * float(s.n_rays) * (1 << 20))
    assert_window_equals_oracle(s, oracle, 40000, 65536, 0, native.FRAME_FORCE_GRID)          # 256 rays of the full batch
