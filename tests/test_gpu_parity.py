"""GPU parity: libaudiort_cuda (through the C ABI) against the CPU oracle on identical inputs.

Bar (BASELINE.json north_star / SURVEY 8c): hit collider indices, bounce counts, echo halves, hit points and
muffle counts BIT-EXACT (the kernels use the reference's operation order in un-fused IEEE FP32, so no
tolerance is needed and none is applied); canonical PermeationPowerRemains bit-exact; permeationSum within
1e-5 relative to N*S; ReverbStrength/ReverbVolume/MuffleStrength within 1e-6 absolute of the FP64 evaluation
(exact integer reduction) and bit-exact with ART_FRAME_REVERB_SEQ_FP32.
"""
import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes
from audio_raytracer_b200.layouts import TYPE_AABB, TYPE_OBB, TYPE_SPHERE
from helpers import aabb, hit_id, micro_scene, obb, sphere

pytestmark = pytest.mark.gpu

C = native.FRAME_COUNTERS


def assert_same_frame(a, b, what):
    """two GPU frames of the same inputs (e.g. uniform-grid traversal vs brute-force scans): bit-identical outputs."""
    for k in ("hit_counts", "hit_ids", "echo", "hit_points", "muffle", "muffle_totals"):
        x, y = getattr(a, k), getattr(b, k)
        if x is not None and y is not None:
            np.testing.assert_array_equal(x, y, err_msg=f"{what}: {k}")
    if a.permeation is not None and b.permeation is not None:
        np.testing.assert_array_equal(a.permeation.view(np.uint32), b.permeation.view(np.uint32), err_msg=f"{what}: permeation")
    if a.settings is not None and b.settings is not None:
        np.testing.assert_array_equal(a.settings.view(np.uint8), b.settings.view(np.uint8), err_msg=f"{what}: settings")
    assert a.counters["segments"] == b.counters["segments"] and a.counters["segmentHits"] == b.counters["segmentHits"]
    assert a.counters["debugViolations"] == 0 and b.counters["debugViolations"] == 0


def run_gpu(ctx, scene, jobs=native.JOB_ALL, flags=C):
    """The counting frame (brute-force kernels, oracle-equivalent work counters) is what the caller compares with the
    oracle; the same inputs are also run through the default path (uniform-grid traversal) and must match it bit for bit."""
    native.upload(ctx, scene)
    r = ctx.run_frame(scene, jobs=jobs, flags=flags)
    if flags & C:
        # default path: uniform grid for the bounce rays + target fans for the echo / muffle / permeation queries;
        # then the same frame with every query walking the grid (ART_FRAME_NO_FANS)
        for extra, what in ((0, "grid + fans"), (native.FRAME_NO_FANS, "grid walk")):
            fast = ctx.run_frame(scene, jobs=jobs, flags=(flags & ~C) | native.FRAME_FORCE_GRID | extra)
            if scene.name != "custom" and jobs & native.JOB_RAYTRACE:
                assert fast.counters["gridUsed"] & 1, "default path did not use the grid"
                assert bool(fast.counters["gridUsed"] & 4) == (extra == 0), "target fans used / not used as requested"
            assert not fast.counters["gridUsed"] & 8, "fan build overflowed"
            assert_same_frame(fast, r, what + " vs brute force")
            if r.permeation_sum is not None and fast.permeation_sum is not None:
                scale = scene.n_rays * scene.permeation_strength_per_ray * max(1, r.counters["permHitRays"])
                np.testing.assert_allclose(fast.permeation_sum, r.permeation_sum, rtol=0, atol=1e-5 * scale, err_msg=what)
    return r


def assert_rt_equal(g, o, scene, check_counters=True):
    np.testing.assert_array_equal(g.hit_counts, o.hit_counts, err_msg="RayHitResultCounts")
    np.testing.assert_array_equal(g.hit_ids, o.hit_ids, err_msg="hit collider ids")
    np.testing.assert_array_equal(g.echo, o.echo, err_msg="EchoRayDistances (half bits)")
    # hit points: equal as half bits; +0/-0 are the same point
    gp, op = g.hit_points.copy(), o.hit_points.copy()
    gp[gp == 0x8000] = 0
    op[op == 0x8000] = 0
    np.testing.assert_array_equal(gp, op, err_msg="RayHitResults.HitPoint")
    np.testing.assert_array_equal(g.muffle, o.muffle, err_msg="MuffleRayHits u16 table")
    np.testing.assert_array_equal(g.muffle_totals, o.muffle_totals, err_msg="muffle totals")
    assert g.counters["segments"] == o.counters["segments"]
    assert g.counters["segmentHits"] == o.counters["segment_hits"]
    if check_counters:
        assert g.counters["traceTests"] == o.counters["trace_tests"]
        assert g.counters["echoQueries"] == o.counters["echo_queries"]
        assert g.counters["muffleQueries"] == o.counters["muffle_queries"]
        assert g.counters["echoTests"] == o.counters["echo_tests"]
        assert g.counters["muffleTests"] == o.counters["muffle_tests"]


def assert_pm_equal(g, o, scene):
    np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32),
                                  err_msg="PermeationPowerRemains (canonical, bit-exact)")
    scale = scene.n_rays * scene.permeation_strength_per_ray * max(1, o.counters["perm_hit_rays"])
    np.testing.assert_allclose(g.permeation_sum, o.permeation_sum, rtol=0, atol=1e-5 * scale)
    for k in ("permRays", "permHitRays", "permPairs"):
        pass
    assert g.counters["permRays"] == o.counters["perm_rays"]
    assert g.counters["permHitRays"] == o.counters["perm_hit_rays"]
    assert g.counters["permPairs"] == o.counters["perm_pairs"]
    assert g.counters["permFirstTests"] == o.counters["perm_first_tests"]
    assert g.counters["permLossTests"] == o.counters["perm_loss_tests"]


def assert_settings(g, o, exact):
    if exact:
        for f in ("muffleStrength", "reverbStrength", "reverbVolume"):
            np.testing.assert_array_equal(g.settings[f].view(np.uint32), o.settings[f].view(np.uint32), err_msg=f)
    else:
        for f in ("muffleStrength", "reverbStrength", "reverbVolume"):
            np.testing.assert_allclose(g.settings[f], o.settings_fp64[f], rtol=0, atol=1e-6, err_msg=f)
    np.testing.assert_array_equal(g.settings["percievedAudioPosition"], o.settings["percievedAudioPosition"])


# ---------------------------------------------------------------- known-answer scenes through the CUDA path
def test_kat_scenes_on_gpu(gpu_ctx, oracle):
    cases = [
        micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1))]),
        micro_scene(spheres=[sphere((0, 0, 5), 1.0)]),
        micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1))], origin=(0, 0, 5), H=1),
        micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1))], spheres=[sphere((0, 0, 5), 1.0)], H=1),
        micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1)), aabb((0, 0, 5), (1, 1, 1))], H=1),
        micro_scene(obbs=[obb((0, 0, 5), (1, 1, 1))], H=1),
        micro_scene(obbs=[obb((0.9, 0, 5), (2.0, 1.0, 0.25), rot_xyz=(0.0, 0.38268, 0.0))], H=3, targets=((0, 0, 20),)),
        micro_scene(aabbs=[aabb((0, 0, 10), (1, 1, 1), target=0)], targets=((0, 0, 10),), H=1, max_muffle=100.0),
        micro_scene(aabbs=[aabb((0, 0, 10), (1, 1, 1), target=-1)], targets=((0, 0, 10),), H=1, max_muffle=100.0),
        micro_scene(aabbs=[aabb((0, 0, 5), (4, 4, 0.5)), aabb((0, 0, -5), (4, 4, 0.5))], H=4),
        micro_scene(aabbs=[aabb((0, 0, 5), (4, 4, 0.5), absorption=0.5), aabb((0, 0, -5), (4, 4, 0.5), absorption=0.5)], H=8),
        micro_scene(aabbs=[aabb((0, 0, 2), (5, 5, 0.125), density=0.0), aabb((0, 0, 5), (1, 1, 1), density=5.0),
                           aabb((0, 0, 14), (1, 1, 1), density=2.0)], targets=((0, 0, 10),), H=1),
        # a ray that hits nothing at all
        micro_scene(aabbs=[aabb((0, 0, -5), (1, 1, 1))], H=2),
    ]
    for i, s in enumerate(cases):
        g = run_gpu(gpu_ctx, s)
        o = oracle.run_frame(s)
        assert_rt_equal(g, o, s)
        assert_pm_equal(g, o, s)
        assert_settings(g, o, exact=False)


def test_first_kat_values_directly(gpu_ctx):
    """K1 of SURVEY Appendix C asserted on the GPU outputs themselves (not only against the oracle)."""
    s = micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1))])
    g = run_gpu(gpu_ctx, s)
    assert g.hit_counts[0] == 1 and g.hit_ids[0] == hit_id(TYPE_AABB, 0)
    assert list(g.hit_points[0]) == [0, 0, 0x4400] and g.echo[0] == 0x4400 and g.echo[1] == 0


# ---------------------------------------------------------------- synthetic rooms, scaled-down BASELINE configs
@pytest.mark.parametrize("name,n_rays,T", [("c2", 2048, 1), ("c2", 1500, 3), ("c3", 96, 2), ("c4", 48, 1), ("c5", 40, 4)])
def test_config_parity_scaled(gpu_ctx, oracle, name, n_rays, T):
    s = scenes.make_config(name, batch_count=T, n_rays=n_rays)
    g = run_gpu(gpu_ctx, s)
    o = oracle.run_frame(s, threads=8)
    assert_rt_equal(g, o, s)
    assert_pm_equal(g, o, s)
    assert_settings(g, o, exact=False)


@pytest.mark.parametrize("n_targets", [1, 3, 31, 32, 40, 70])
def test_target_group_boundaries(gpu_ctx, oracle, n_targets):
    """query slots are processed 32 at a time (slot 0 = echo): cover the group edges."""
    s = scenes.make_scene(n_aabb=60, n_obb=max(24, n_targets), n_sphere=20, n_targets=n_targets, seed=77 + n_targets,
                          n_rays=384, max_hits=5, batch_count=2)
    g = run_gpu(gpu_ctx, s)
    o = oracle.run_frame(s, threads=8)
    assert_rt_equal(g, o, s)
    assert_pm_equal(g, o, s)
    assert o.muffle_totals.sum() > 0          # the case exercises visible targets


@pytest.mark.parametrize("na,no,ns", [(6, 0, 0), (0, 3, 0), (0, 0, 5), (7, 1, 0), (130, 66, 129), (0, 0, 0)])
def test_ragged_and_empty_collider_lists(gpu_ctx, oracle, na, no, ns):
    rng = np.random.default_rng(5)
    base = scenes.make_scene(n_aabb=max(na, 6), n_obb=max(no, 1), n_sphere=max(ns, 1), n_targets=1, seed=9,
                             n_rays=256, max_hits=4)
    s = scenes.Scene(aabbs=base.aabbs[:na], obbs=base.obbs[:no], spheres=base.spheres[:ns], targets=base.targets,
                     ray_directions=base.ray_directions, max_hits_per_ray=4)
    g = run_gpu(gpu_ctx, s)
    o = oracle.run_frame(s)
    assert_rt_equal(g, o, s)
    assert_pm_equal(g, o, s)


def test_muffle_distance_gate_and_short_life(gpu_ctx, oracle):
    s = scenes.make_config("c2", n_rays=1024)
    s.max_muffle_hit_distance = 20.0       # RT:168 gate actually filters
    s.max_ray_life = 30.0                  # RT:179 life <= 0 ends rays early
    g = run_gpu(gpu_ctx, s)
    o = oracle.run_frame(s, threads=8)
    assert_rt_equal(g, o, s)
    assert 0 < o.counters["muffle_queries"] < o.counters["echo_queries"]
    assert np.bincount(o.hit_counts, minlength=9)[1:8].sum() > 0


def test_max_hits_255_and_1(gpu_ctx, oracle):
    for H in (1, 255):
        s = scenes.make_config("c2", n_rays=64)
        s.max_hits_per_ray = H
        g = run_gpu(gpu_ctx, s)
        o = oracle.run_frame(s, threads=8)
        assert_rt_equal(g, o, s)


def test_u16_muffle_wrap_q8(gpu_ctx, oracle):
    """An open scene with > 65535 visible muffle rays per slot: the u16 table wraps like C# unchecked ushort."""
    floor = aabb((0, -3, 0), (60, 0.5, 60))
    s = micro_scene(aabbs=[floor], targets=((5, 0, 5),), H=1, max_muffle=1e4, max_life=1e4)
    s.ray_directions = scenes.fibonacci_directions(150000)
    g = run_gpu(gpu_ctx, s)
    o = oracle.run_frame(s, threads=8)
    assert o.muffle_totals[0] > 65535
    assert_rt_equal(g, o, s)
    assert g.muffle[0] == o.muffle_totals[0] % 65536


def test_reverb_sequential_fp32_is_bit_exact(gpu_ctx, oracle):
    s = scenes.make_config("c2", n_rays=4096, batch_count=2)
    native.upload(gpu_ctx, s)
    g = gpu_ctx.run_frame(s, flags=native.FRAME_REVERB_SEQ_FP32)
    o = oracle.run_frame(s, threads=8)
    assert_settings(g, o, exact=True)
    g2 = gpu_ctx.run_frame(s, flags=0)
    assert_settings(g2, o, exact=False)


def test_determinism_and_no_counter_variant(gpu_ctx, oracle):
    s = scenes.make_config("c3", n_rays=512)
    native.upload(gpu_ctx, s)
    F = native.FRAME_FORCE_GRID                     # 512 rays: the library itself would pick the brute-force kernels
    a = gpu_ctx.run_frame(s, flags=F)
    b = gpu_ctx.run_frame(s, flags=F)
    c = gpu_ctx.run_frame(s, flags=C)
    for x in (b, c):
        np.testing.assert_array_equal(a.echo, x.echo)
        np.testing.assert_array_equal(a.hit_ids, x.hit_ids)
        np.testing.assert_array_equal(a.muffle, x.muffle)
        np.testing.assert_array_equal(a.permeation.view(np.uint32), x.permeation.view(np.uint32))
        np.testing.assert_array_equal(a.settings.view(np.uint8), x.settings.view(np.uint8))
    np.testing.assert_array_equal(a.permeation_sum, b.permeation_sum)              # deterministic reduction
    # the counting frame runs the brute-force kernels: same per-ray values up to the documented tolerance
    np.testing.assert_allclose(a.permeation_sum, c.permeation_sum, rtol=0, atol=1e-5 * s.n_rays * s.n_rays)
    d = gpu_ctx.run_frame(s, flags=native.FRAME_BRUTE_FORCE)
    assert d.counters["gridUsed"] == 0 and a.counters["gridUsed"] == 7
    np.testing.assert_array_equal(c.permeation_sum, d.permeation_sum)
    np.testing.assert_array_equal(a.echo, d.echo)


def test_sharded_contexts_merge_to_the_unsharded_result(art_lib, oracle):
    """G contexts each trace an interleaved slice; merged partials == single-context frame."""
    s = scenes.make_config("c2", n_rays=3000, batch_count=4)
    with native.Context(0) as full:
        native.upload(full, s)
        ref = full.run_frame(s, flags=C)
    blobs, parts = [], []
    G, chunk = 3, 128
    for r in range(G):
        with native.Context(0) as ctx:
            native.upload(ctx, s)
            ctx.set_ray_shard(r, G, chunk)
            res = ctx.run_frame(s, flags=native.FRAME_PARTIALS_ONLY | C)
            blobs.append(ctx.get_partials(s.n_targets, s.batch_count))
            parts.append((r, res))
    merged = native.finalize(native.merge_partials(blobs), s, s.n_rays)
    np.testing.assert_array_equal(merged.muffle, ref.muffle)
    np.testing.assert_array_equal(merged.muffle_totals, ref.muffle_totals)
    np.testing.assert_array_equal(merged.permeation.view(np.uint32), ref.permeation.view(np.uint32))
    np.testing.assert_array_equal(merged.permeation_sum, ref.permeation_sum)
    np.testing.assert_array_equal(merged.settings.view(np.uint8), ref.settings.view(np.uint8))
    # per-ray outputs: scatter the shards' local arrays back to global ray order
    H = s.max_hits_per_ray
    echo = np.zeros_like(ref.echo)
    for r, res in parts:
        n_local = len(res.hit_counts)
        j = np.arange(n_local)
        gidx = ((j // chunk) * G + r) * chunk + j % chunk
        echo.reshape(-1, H)[gidx] = res.echo.reshape(-1, H)
    np.testing.assert_array_equal(echo, ref.echo)


def test_async_handle_semantics(gpu_ctx):
    s = scenes.make_config("c2", n_rays=512)
    native.upload(gpu_ctx, s)
    h = gpu_ctx.schedule(s)
    with pytest.raises(native.ArtError) as e:       # one frame in flight per context (ART:95-97)
        gpu_ctx.schedule(s)
    assert e.value.code == native.ART_E_PENDING
    gpu_ctx.complete(h)
    assert gpu_ctx.is_completed(h)


def test_argument_errors(gpu_ctx):
    s = scenes.make_config("c2", n_rays=64)
    with pytest.raises(native.ArtError) as e:       # nothing uploaded yet
        gpu_ctx.run_frame(s)
    assert e.value.code in (native.ART_E_STATE,)
    native.upload(gpu_ctx, s)
    bad = scenes.make_config("c2", n_rays=64)
    bad.targets = np.zeros((0, 3), np.float32)      # TotalAudioTargets = 0: RT:63 divides by zero
    with pytest.raises(native.ArtError) as e:
        gpu_ctx.run_frame(bad)
    assert e.value.code == native.ART_E_ARG


def test_device_fibonacci_matches_host_generator(gpu_ctx, oracle):
    for n in (314, 65536):
        gpu_ctx.generate_fibonacci_rays(n)
        d = gpu_ctx.get_rays()
        ref = oracle.fibonacci_directions(n)
        mism = int((d != ref).any(axis=1).sum())
        assert mism <= n // 2000, f"{mism} of {n} directions differ (double cos/sin last-ulp differences only)"
        assert d[0, 1] == 0x3C00 and d[-1, 1] == 0xBC00


# ---------------------------------------------------------------- full BASELINE sizes
def test_c2_full_size_parity(gpu_ctx, oracle):
    s = scenes.make_config("c2")                    # 65,536 rays x 8 bounces vs 448 colliders
    g = run_gpu(gpu_ctx, s)
    o = oracle.run_frame(s, threads=8)
    assert_rt_equal(g, o, s)
    assert_pm_equal(g, o, s)
    assert_settings(g, o, exact=False)


def test_c3_full_size_properties(gpu_ctx, oracle):
    """C3 at full size (1M rays x 12 x 64 targets x 4096 colliders): size-independent properties + a sampled
    oracle comparison (the oracle needs hours for the whole frame)."""
    s = scenes.make_config("c3", batch_count=8)
    native.upload(gpu_ctx, s)
    g = gpu_ctx.run_frame(s, jobs=native.JOB_RAYTRACE, flags=0, want=("echo", "hit_counts", "hit_ids"))
    H = s.max_hits_per_ray
    assert g.counters["segmentHits"] == int(g.hit_counts.astype(np.int64).sum())
    assert g.hit_counts.max() <= H
    ids = g.hit_ids.reshape(-1, H)
    filled = (ids != 0)
    np.testing.assert_array_equal(filled.sum(axis=1), g.hit_counts)            # ids are written exactly per hit
    assert not (g.echo.reshape(-1, H)[~filled] != 0).any()                      # echo only where a hit exists
    assert int(g.muffle_totals.sum()) == int(g.muffle.astype(np.int64).sum()) or g.muffle_totals.max() > 65535
    # sampled oracle parity: 3 windows of 64 rays
    for first in (0, 524288, 1048576 - 64):
        o = oracle.trace_range(s, first, 64, threads=8)
        sl = slice(first * H, (first + 64) * H)
        np.testing.assert_array_equal(g.hit_ids[sl], o.hit_ids[sl])
        np.testing.assert_array_equal(g.echo[sl], o.echo[sl])
        np.testing.assert_array_equal(g.hit_counts[first:first + 64], o.hit_counts[first:first + 64])


def test_jobs_mirror_one_frame_latency(art_lib, oracle):
    """jobs.AudioRayTracer mirrors ART:92-238: results are consumed one frame after they were scheduled."""
    from audio_raytracer_b200 import jobs
    s = scenes.make_config("c2", n_rays=314)
    rt = jobs.AudioRayTracer(rayCount=314, maxBounces=4, maxRayLife=125.0, maxMuffleHitDistance=250.0,
                             maxReverbDistance=35.0, toUseThreadCount=1)
    rt.set_colliders(s.aabbs, s.obbs, s.spheres)
    assert rt.OnUpdate((0, 0, 0), s.targets) is None                 # first frame: nothing to consume yet
    res = None
    while res is None:
        res = rt.OnUpdate((0, 0, 0), s.targets)
    s.max_hits_per_ray, s.max_ray_life, s.max_muffle_hit_distance = 5, 125.0, 250.0
    s.ray_directions = rt._ctx.get_rays()                            # device-generated FIB:25-34 directions
    assert (s.ray_directions != oracle.fibonacci_directions(314)).any(axis=1).sum() <= 1
    o = oracle.run_frame(s)
    np.testing.assert_array_equal(res.hit_counts, o.hit_counts)
    np.testing.assert_array_equal(res.muffle, o.muffle)
    rt.OnDestroy()


# ---------------------------------------------------------------- uniform-grid path: edge cases of the traversal
def _grid_vs_oracle(ctx, oracle, s, expect_grid=True):
    o = oracle.run_frame(s, threads=8)
    native.upload(ctx, s)
    for extra in (0, native.FRAME_NO_FANS):                    # target fans / every query walking the grid
        g = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | extra)    # grid kernels whatever the scene size
        assert bool(g.counters["gridUsed"] & 1) == expect_grid
        assert bool(g.counters["gridUsed"] & 4) == (expect_grid and extra == 0) and not g.counters["gridUsed"] & 8
        for k in ("hit_counts", "hit_ids", "echo", "muffle", "muffle_totals"):
            np.testing.assert_array_equal(getattr(g, k), getattr(o, k), err_msg=f"{k} (flags {extra})")
        gp, op = g.hit_points.copy(), o.hit_points.copy()
        gp[gp == 0x8000] = 0
        op[op == 0x8000] = 0
        np.testing.assert_array_equal(gp, op)
        np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
        scale = max(s.n_rays * s.permeation_strength_per_ray, 50.0) * max(1, o.counters["perm_hit_rays"])
        np.testing.assert_allclose(g.permeation_sum, o.permeation_sum, rtol=0, atol=1e-5 * scale)
        assert g.counters["segments"] == o.counters["segments"]
    assert g.counters["debugViolations"] == 0
    return g, o


def test_grid_listener_and_targets_outside_the_scene(gpu_ctx, oracle):
    """rays that start outside the grid (clip to the grid first) and targets far outside it"""
    s = scenes.make_scene(n_aabb=40, n_obb=20, n_sphere=12, n_targets=5, seed=901, n_rays=600, max_hits=4, batch_count=2,
                          room_half=(6.0, 3.0, 6.0))
    s.ray_origin = np.array([14.0, 1.0, -11.0], dtype=np.float32)           # outside the room, inside the listener range
    s.targets[0] = [40.0, 2.0, 3.0]
    s.targets[1] = [-25.0, -30.0, 0.5]
    _grid_vs_oracle(gpu_ctx, oracle, s)
    s.ray_origin = np.array([400.0, 0.0, 0.0], dtype=np.float32)            # beyond the listener range: brute-force fallback
    _grid_vs_oracle(gpu_ctx, oracle, s, expect_grid=False)


def test_grid_owned_colliders_of_every_type(gpu_ctx, oracle):
    """the owner skip (RT:413/426/439, PM:235/245/255) inside the traversal, for spheres, AABBs and OBBs"""
    s = scenes.make_scene(n_aabb=60, n_obb=30, n_sphere=20, n_targets=6, seed=902, n_rays=700, max_hits=5, batch_count=1,
                          room_half=(8.0, 4.0, 8.0))
    from audio_raytracer_b200.layouts import f32tof16
    for t in range(6):                                                      # wrap every target in owned colliders
        s.spheres["center"][t] = f32tof16(s.targets[t]); s.spheres["radius"][t] = f32tof16(np.float32([1.2]))[0]
        s.spheres["audioTargetId"][t] = t
        s.aabbs["center"][6 + t] = f32tof16(s.targets[t]); s.aabbs["size"][6 + t] = f32tof16(np.float32([0.9, 0.7, 0.8]))
        s.aabbs["audioTargetId"][6 + t] = t if t % 2 else (t + 1) % 6         # even: owned by a DIFFERENT target (blocks), odd: own
    g, o = _grid_vs_oracle(gpu_ctx, oracle, s)
    assert o.muffle_totals.sum() > 0


def test_grid_large_coordinates_and_mixed_sizes(gpu_ctx, oracle):
    """a 20x larger room (half precision spacing up to 0.5 at these coordinates), walls spanning the whole grid next to small boxes"""
    s = scenes.make_scene(n_aabb=80, n_obb=40, n_sphere=30, n_targets=4, seed=903, n_rays=500, max_hits=6, batch_count=1,
                          room_half=(640.0, 160.0, 640.0), size_range=(2.0, 60.0))
    s.ray_origin = np.array([3.0, 13.0, -7.0], dtype=np.float32)
    s.max_ray_life = 1e6
    s.max_muffle_hit_distance = 1e6
    _grid_vs_oracle(gpu_ctx, oracle, s)


def test_grid_rebuilds_when_the_scene_changes(gpu_ctx, oracle):
    a = scenes.make_config("c2", n_rays=300)
    b = scenes.make_scene(n_aabb=30, n_obb=10, n_sphere=50, n_targets=2, seed=904, n_rays=300, max_hits=3)
    for s in (a, b, a):
        _grid_vs_oracle(gpu_ctx, oracle, s)


def test_small_scenes_default_to_the_brute_force_kernels(gpu_ctx, oracle):
    """demo scene (314 rays): the library picks the low-latency brute-force kernels unless told otherwise"""
    s = scenes.make_config("c1")
    native.upload(gpu_ctx, s)
    a = gpu_ctx.run_frame(s)
    b = gpu_ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)
    assert a.counters["gridUsed"] == 0 and b.counters["gridUsed"] == 7
    assert_same_frame(a, b, "default vs forced grid")
    big = scenes.make_config("c2")                  # 65,536 rays: the grid kernels
    native.upload(gpu_ctx, big)
    assert gpu_ctx.run_frame(big).counters["gridUsed"] == 7
    small = scenes.make_config("c2", n_rays=2048)
    native.upload(gpu_ctx, small)
    assert gpu_ctx.run_frame(small).counters["gridUsed"] == 0


def test_grid_stats_are_reported(gpu_ctx):
    s = scenes.make_config("c3", n_rays=256)
    native.upload(gpu_ctx, s)
    c = gpu_ctx.run_frame(s, flags=native.FRAME_GRID_STATS | native.FRAME_FORCE_GRID).counters
    full = gpu_ctx.run_frame(s, flags=C).counters
    assert c["gridUsed"] == 7 and c["gridTraceCells"] > 0 and c["gridPermCells"] > 0
    executed = sum(c["gridTraceTests"])
    scanned = sum(full["traceTests"]) + sum(full["echoTests"]) + sum(full["muffleTests"])
    assert 0 < executed < scanned / 10                     # the traversal runs a small fraction of the full scans' tests
    assert 0 < sum(c["gridPermLossTests"]) < sum(full["permLossTests"]) / 5


def test_c3_full_size_grid_equals_brute_force(gpu_ctx):
    """BASELINE config 3 at full size (1,048,576 rays, 12.4 M segments): the default (uniform-grid) path and the brute-force
    scans must agree on every per-ray and per-source output, bit for bit."""
    s = scenes.make_config("c3")
    native.upload(gpu_ctx, s)
    fast = gpu_ctx.run_frame(s)
    slow = gpu_ctx.run_frame(s, flags=native.FRAME_BRUTE_FORCE)
    # grid (1) + grid permeation (2) + target fans (4) + loss lines binned by (target, direction bin) (32)
    assert fast.counters["gridUsed"] == 39 and slow.counters["gridUsed"] == 0
    assert fast.counters["segments"] == slow.counters["segments"] == 12360709
    for k in ("hit_counts", "hit_ids", "echo", "hit_points", "muffle", "muffle_totals"):
        assert np.array_equal(getattr(fast, k), getattr(slow, k)), k
    np.testing.assert_array_equal(fast.permeation.view(np.uint32), slow.permeation.view(np.uint32))
    np.testing.assert_array_equal(fast.settings.view(np.uint8), slow.settings.view(np.uint8))
    np.testing.assert_allclose(fast.permeation_sum, slow.permeation_sum, rtol=1e-9, atol=1e-5 * s.n_rays * s.n_rays)
    # a second frame of the same context (buffers reused, fans rebuilt): same outputs
    again = gpu_ctx.run_frame(s)
    assert again.counters["gridUsed"] == 39
    assert_same_frame(again, fast, "full-size C3, second frame")
    # the grid walk (no fans): the previous frames showed rays living nearly all their 12 bounces, so it rotates its ray
    # groups through the warps (gridUsed bit 4, k1_trace_grid.cu): same outputs
    walk = gpu_ctx.run_frame(s, flags=native.FRAME_NO_FANS)
    assert walk.counters["gridUsed"] == 19
    assert_same_frame(walk, fast, "full-size C3, grid walk with group rotation vs fans")


@pytest.mark.parametrize("seed", range(16))
def test_random_scenes_and_parameters(gpu_ctx, oracle, seed):
    """fuzz: random room sizes, collider mixes (including empty types), target counts around the 15-target switch between
    the one-pass and the two-stage occlusion pool, ray lives / muffle distances that cut rays short, several batch counts"""
    rng = np.random.default_rng(7000 + seed)
    na, no, ns = int(rng.integers(6, 200)), int(rng.integers(0, 90)), int(rng.integers(0, 60))
    nt = int(rng.choice([1, 2, 7, 14, 15, 16, 33, 50]))
    no = max(no, nt)
    room = rng.uniform(3.0, 40.0, size=3)
    s = scenes.make_scene(n_aabb=na, n_obb=no, n_sphere=ns, n_targets=nt, seed=8000 + seed, n_rays=int(rng.integers(100, 900)),
                          max_hits=int(rng.integers(1, 10)), batch_count=int(rng.integers(1, 6)), room_half=room,
                          size_range=(0.1, float(rng.uniform(0.5, 4.0))))
    s.max_ray_life = float(rng.choice([1000.0, 60.0, 15.0]))
    s.max_muffle_hit_distance = float(rng.choice([1000.0, 30.0, 8.0]))
    s.ray_origin = (rng.uniform(-0.5, 0.5, size=3) * room).astype(np.float32)
    _grid_vs_oracle(gpu_ctx, oracle, s)


def test_small_frames_in_a_loop(gpu_ctx, oracle):
    """the reference's own frame sizes, frame after frame on one context (two jobs side by side on two streams, per-ray
    outputs copied back on a third): a moving listener, moving targets, a scene re-uploaded every frame, another frame shape
    in between -- every frame must still give the oracle's results"""
    s = scenes.make_config("c1")
    native.upload(gpu_ctx, s)
    for step in range(6):
        s.ray_origin = (np.float32([15.51, -1.45, -3.11]) + np.float32([0.37 * step, 0.05 * step, -0.21 * step])).astype(np.float32)
        s.targets[0] = s.targets[0] + np.float32([0.0, 0.1 * step, 0.0])
        g = gpu_ctx.run_frame(s, flags=native.FRAME_REVERB_SEQ_FP32)
        o = oracle.run_frame(s)
        for k in ("echo", "hit_counts", "hit_ids", "muffle", "muffle_totals"):
            np.testing.assert_array_equal(getattr(g, k), getattr(o, k), err_msg=f"step {step}: {k}")
        np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
        np.testing.assert_array_equal(g.settings.view(np.uint8), o.settings.view(np.uint8))
    # dynamic scene: art_set_scene every frame (colliders re-baked)
    for step in range(5):
        s.aabbs["center"][10, 1] += 8          # move one collider (raw half bits: a small change of the value)
        native.upload(gpu_ctx, s)
        g = gpu_ctx.run_frame(s, flags=native.FRAME_REVERB_SEQ_FP32)
        o = oracle.run_frame(s)
        np.testing.assert_array_equal(g.echo, o.echo)
        np.testing.assert_array_equal(g.muffle, o.muffle)
        np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
    # another frame shape in between
    t = scenes.make_config("c2", n_rays=700)
    native.upload(gpu_ctx, t)
    g = gpu_ctx.run_frame(t)
    o = oracle.run_frame(t, threads=8)
    np.testing.assert_array_equal(g.echo, o.echo)
    native.upload(gpu_ctx, s)
    g = gpu_ctx.run_frame(s, flags=native.FRAME_REVERB_SEQ_FP32)
    o = oracle.run_frame(s)
    np.testing.assert_array_equal(g.echo, o.echo)
    np.testing.assert_array_equal(g.settings.view(np.uint8), o.settings.view(np.uint8))


def test_fan_overflow_reruns_the_frame_on_the_grid_walk(oracle, monkeypatch):
    """a fan entry buffer that is too small is detected on the device; art_complete re-runs the frame without fans (same
    results), reports it in gridUsed bit 3 and grows the buffer so that the next frame fits"""
    monkeypatch.setenv("ART_FAN_ENTRIES_PER_PAIR", "1")
    s = scenes.make_config("c3", n_rays=512)
    o = oracle.run_frame(s, threads=8)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        seen = []
        for _ in range(4):
            g = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)
            seen.append(g.counters["gridUsed"])
            np.testing.assert_array_equal(g.hit_ids, o.hit_ids)
            np.testing.assert_array_equal(g.echo, o.echo)
            np.testing.assert_array_equal(g.muffle, o.muffle)
            np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
        assert seen[0] & 8 and seen[-1] == 7, seen


def test_fan_buffers_that_cannot_be_allocated_fall_back_to_the_grid_walk(oracle, monkeypatch):
    """a frame whose fan buffers cannot be allocated (ART_FAN_ALLOC_FAIL stands in for cudaMalloc failing at tens of thousands of
    targets) walks the grid instead -- same results, no error -- and the next frame builds its fans again"""
    s = scenes.make_config("c3", n_rays=512)
    o = oracle.run_frame(s, threads=8)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        seen = []
        for fail in ("0", "1", "0"):
            monkeypatch.setenv("ART_FAN_ALLOC_FAIL", fail)
            g = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)
            seen.append(g.counters["gridUsed"])
            np.testing.assert_array_equal(g.hit_ids, o.hit_ids)
            np.testing.assert_array_equal(g.echo, o.echo)
            np.testing.assert_array_equal(g.muffle, o.muffle)
            np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
        assert seen[0] & 4 and not seen[1] & 4 and seen[1] & 1 and seen[2] & 4, seen


@pytest.mark.parametrize("name,n_rays,flags", [("c3", 6000, 0), ("c3", 6000, native.FRAME_NO_FANS), ("c2", 20000, 0)])
def test_goal_tables_in_shared_memory_do_not_change_results(monkeypatch, name, n_rays, flags):
    """The grid kernels keep the listener + target positions of the query pool in shared memory when they fit
    (k1_trace_grid.cu, launch_trace_grid); with very many targets they read targetOrder -> targets from global memory
    instead. ART_K1_NO_GOAL_TABLES=1 forces that fallback: every output must stay the same, and equal the brute-force scans."""
    s = scenes.make_config(name, n_rays=n_rays)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        a = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | flags)
        monkeypatch.setenv("ART_K1_NO_GOAL_TABLES", "1")
        b = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | flags)
        monkeypatch.delenv("ART_K1_NO_GOAL_TABLES")
        c = ctx.run_frame(s, flags=native.FRAME_BRUTE_FORCE)
    assert a.counters["gridUsed"] & 1 and b.counters["gridUsed"] & 1 and c.counters["gridUsed"] == 0
    for x in (b, c):
        for k in ("hit_counts", "hit_ids", "echo", "hit_points", "muffle", "muffle_totals"):
            np.testing.assert_array_equal(getattr(a, k), getattr(x, k), err_msg=k)
        np.testing.assert_array_equal(a.permeation.view(np.uint32), x.permeation.view(np.uint32))
        np.testing.assert_array_equal(a.settings.view(np.uint8), x.settings.view(np.uint8))


@pytest.mark.parametrize("name,n_rays,T,warps,life", [("c2", 65536, 1, 8, 1000.0), ("c3", 40000, 3, 4, 1000.0),
                                                      ("c3", 30001, 2, 4, 70.0), ("c5", 60000, 1, 6, 150.0)])
def test_group_rotation_is_bit_identical(monkeypatch, name, n_rays, T, warps, life):
    """Small batches (one shard of a ray-sharded frame) rotate their groups of 32 rays through the warps after every
    bounce round (k1_trace_grid.cu, gridUsed bit 4). Which warp traces a ray must not change any output: the rotated
    frame equals the same frame without rotation and the brute-force scans, bit for bit -- including rays that die at
    different bounces (short MaxRayLife) and a last group that is not full. ART_K1_WARPS shrinks the launch so that
    test-sized batches give every warp between one and four groups."""
    s = scenes.make_config(name, batch_count=T, n_rays=n_rays)
    s.max_ray_life = life
    monkeypatch.setenv("ART_K1_WARPS", str(warps))
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        slow = ctx.run_frame(s, flags=native.FRAME_BRUTE_FORCE)
        # (the rotation belongs to the grid walk; with the target fans the queries run in their own kernel, k1_query_fan.cu)
        rot = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | native.FRAME_NO_FANS)
        rot2 = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | native.FRAME_NO_FANS)
        monkeypatch.setenv("ART_K1_ROTATE", "0")
        plain = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | native.FRAME_NO_FANS)
        fans = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)
    assert rot.counters["gridUsed"] & 16 and rot2.counters["gridUsed"] & 16, "rotation was not used"
    assert not plain.counters["gridUsed"] & 16 and slow.counters["gridUsed"] == 0
    assert 0 < rot.hit_counts.min() or life < 1000.0
    if life < 1000.0:
        assert len(np.unique(rot.hit_counts)) > 3, "rays should die at different bounces in this case"
    assert fans.counters["gridUsed"] & 4 and not fans.counters["gridUsed"] & 16
    for other, what in ((plain, "rotation vs no rotation"), (slow, "rotation vs brute force"), (rot2, "rotation twice"), (fans, "rotation vs fans")):
        assert_same_frame(rot, other, what)
