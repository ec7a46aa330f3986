"""GPU tests of the binned permeation path (k2_permeation_binned.cu): the loss lines sorted by (target, direction bin of
the target's fan) and evaluated 32 of one bin at a time. The per-line arithmetic is that of the fan path of
permeation_grid_kernel, so permeationSum must come out BIT-IDENTICAL to it (integer accumulation: order independent), the
canonical PermeationPowerRemains bit-identical to the brute-force kernels and the oracle."""
import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes

pytestmark = pytest.mark.gpu

F = native.FRAME_FORCE_GRID


@pytest.mark.parametrize("name,n_rays,T", [("c3", 3000, 1), ("c3", 4097, 3), ("c4", 700, 1), ("c2", 20000, 2), ("c5", 5000, 1)])
def test_binned_loss_lines_equal_the_per_line_path(monkeypatch, name, n_rays, T):
    s = scenes.make_config(name, n_rays=n_rays, batch_count=T)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        monkeypatch.setenv("ART_K2_BINNED", "0")
        a = ctx.run_frame(s, flags=F)
        monkeypatch.setenv("ART_K2_BINNED", "1")
        b = ctx.run_frame(s, flags=F)
        b2 = ctx.run_frame(s, flags=F)
        monkeypatch.delenv("ART_K2_BINNED")
        c = ctx.run_frame(s, flags=native.FRAME_BRUTE_FORCE)
    assert a.counters["gridUsed"] & 2 and not a.counters["gridUsed"] & 32
    assert b.counters["gridUsed"] & 32 and b2.counters["gridUsed"] & 32, "the binned path did not run"
    assert b.counters["debugViolations"] == 0
    for x in (b, b2):
        np.testing.assert_array_equal(x.permeation_sum, a.permeation_sum)            # same floats per line, integer sums
        np.testing.assert_array_equal(x.permeation.view(np.uint32), a.permeation.view(np.uint32))
        np.testing.assert_array_equal(x.settings.view(np.uint8), a.settings.view(np.uint8))
        assert x.counters["permHitRays"] == a.counters["permHitRays"]
    np.testing.assert_array_equal(b.permeation.view(np.uint32), c.permeation.view(np.uint32))
    scale = s.n_rays * s.permeation_strength_per_ray * max(1, c.counters["permHitRays"])
    np.testing.assert_allclose(b.permeation_sum, c.permeation_sum, rtol=0, atol=1e-5 * scale)


def test_binned_path_against_the_oracle(oracle, monkeypatch):
    s = scenes.make_config("c3", n_rays=1536, batch_count=2)
    o = oracle.run_frame(s, threads=8)
    monkeypatch.setenv("ART_K2_BINNED", "1")
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        g = ctx.run_frame(s, flags=F)
    assert g.counters["gridUsed"] & 32
    np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
    scale = s.n_rays * s.permeation_strength_per_ray * max(1, o.counters["perm_hit_rays"])
    np.testing.assert_allclose(g.permeation_sum, o.permeation_sum, rtol=0, atol=1e-5 * scale)
    assert g.counters["permHitRays"] == o.counters["perm_hit_rays"]
    np.testing.assert_array_equal(g.echo, o.echo)


def test_binned_sharded_frames_merge_to_the_unsharded_sums(monkeypatch):
    """the lines of a shard are binned per shard; merged partial sums == the single-context frame (integer sums)"""
    monkeypatch.setenv("ART_K2_BINNED", "1")
    s = scenes.make_config("c3", n_rays=4096, batch_count=2)
    with native.Context(0) as full:
        native.upload(full, s)
        ref = full.run_frame(s, flags=F)
    blobs = []
    for r in range(2):
        with native.Context(0) as ctx:
            native.upload(ctx, s)
            ctx.set_ray_shard(r, 2, 256)
            res = ctx.run_frame(s, flags=F | native.FRAME_PARTIALS_ONLY)
            assert res.counters["gridUsed"] & 32
            blobs.append(ctx.get_partials(s.n_targets, s.batch_count))
    merged = native.finalize(native.merge_partials(blobs), s, s.n_rays)
    np.testing.assert_array_equal(merged.permeation_sum, ref.permeation_sum)
    np.testing.assert_array_equal(merged.permeation.view(np.uint32), ref.permeation.view(np.uint32))
