"""Multi-GPU reached through the C ABI alone (`-m gpu`): ONE context over several devices of one process
(ArtConfig.nDevices / devices[]) and one context per process joined by the library's own communicator (art_comm_init).
Both must reproduce the single-device frame bit for bit (the merge is exact: integer sums + a max-by-ray-index select)."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_count():
    import torch
    return torch.cuda.device_count()


def assert_same(a, b, what):
    for k in ("hit_counts", "hit_ids", "echo", "hit_points", "muffle", "muffle_totals", "permeation_sum"):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=f"{what}: {k}")
    np.testing.assert_array_equal(a.permeation.view(np.uint32), b.permeation.view(np.uint32), err_msg=what)
    np.testing.assert_array_equal(a.settings.view(np.uint8), b.settings.view(np.uint8), err_msg=what)


def run_multi(devices, name, n_rays, T, chunk=0, flags=native.FRAME_FORCE_GRID):
    s = scenes.make_config(name, n_rays=n_rays, batch_count=T)
    with native.Context(0) as one:
        native.upload(one, s)
        ref = one.run_frame(s, flags=flags)
    with native.Context(devices=devices, shard_chunk_rays=chunk) as multi:
        native.upload(multi, s)
        assert multi.local_ray_count() == s.n_rays
        with pytest.raises(native.ArtError):
            multi.set_ray_shard(0, 2, 0)                         # the library owns the shard map
        h = multi.schedule(s, flags=flags)
        while not multi.is_completed(h):
            pass
        got = multi.complete(h)
        again = multi.run_frame(s, flags=flags)
    assert got.counters["devicesUsed"] == len(devices)
    assert got.counters["segments"] == ref.counters["segments"]
    assert_same(got, ref, f"{len(devices)} device contexts vs one")
    assert_same(again, ref, "second frame")


@pytest.mark.parametrize("name,n_rays,T,n,chunk", [("c2", 3000, 4, 3, 128), ("c3", 40000, 2, 4, 0), ("c5", 9001, 1, 2, 256)])
def test_multi_device_context_on_one_gpu(name, n_rays, T, n, chunk):
    """nDevices = n with every entry device 0: exercises the library-owned shard map, the threaded schedule / complete, the
    scatter of the per-ray outputs to global ray positions and the exact merge -- on a single-GPU box."""
    run_multi([0] * n, name, n_rays, T, chunk)


def test_multi_device_context_on_distinct_gpus():
    if device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(device_count(), 8)
    run_multi(list(range(n)), "c3", 200000, n, 0, flags=0)


def test_multi_process_communicator_merges_inside_the_library():
    """One rank per GPU (what torchrun / mpirun launch): the ranks' partial blobs are all-gathered on the device by the
    library and every rank's art_complete returns the merged per-source outputs == the single-device frame."""
    world = min(device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    s = scenes.make_config("c3", n_rays=40000, batch_count=world)
    with native.Context(0) as one:
        native.upload(one, s)
        ref = one.run_frame(s, flags=native.FRAME_FORCE_GRID)
    with tempfile.TemporaryDirectory() as td:
        idfile = os.path.join(td, "id")
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_comm_worker.py"), str(r), str(world), idfile,
                                   os.path.join(td, f"out{r}.npz")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                 for r in range(world)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o
        seg = 0
        for r in range(world):
            z = np.load(os.path.join(td, f"out{r}.npz"))
            np.testing.assert_array_equal(z["muffle"], ref.muffle)
            np.testing.assert_array_equal(z["muffle2"], ref.muffle)
            np.testing.assert_array_equal(z["muffle_totals"], ref.muffle_totals)
            np.testing.assert_array_equal(z["permeation"].view(np.uint32), ref.permeation.view(np.uint32))
            np.testing.assert_array_equal(z["permeation_sum"], ref.permeation_sum)
            np.testing.assert_array_equal(z["settings"], ref.settings.view(np.uint8))
            seg += int(z["segments"])
            assert 0.0 < float(z["exchange_ms"]) < 5.0
        assert seg == ref.counters["segments"]


def test_a_sharded_context_keeps_only_its_own_directions():
    """art_set_rays on a context that already has a shard map stages only that shard's directions (0.8 MB instead of 6.3 MB
    per rank of 8 at C3); the frame equals the one of a context that was given the whole batch first and the shard map
    afterwards. Changing the shard map then requires the rays again."""
    s = scenes.make_config("c2", n_rays=30000, batch_count=3)
    with native.Context(0) as a, native.Context(0) as b:
        a.set_scene(s.aabbs, s.obbs, s.spheres)
        a.set_rays(s.ray_directions)                     # whole batch, then the shard map
        a.set_ray_shard(1, 3, 512)
        b.set_scene(s.aabbs, s.obbs, s.spheres)
        b.set_ray_shard(1, 3, 512)                       # shard map first: only the shard is kept
        b.set_rays(s.ray_directions)
        ra = a.run_frame(s, flags=native.FRAME_PARTIALS_ONLY | native.FRAME_FORCE_GRID)
        rb = b.run_frame(s, flags=native.FRAME_PARTIALS_ONLY | native.FRAME_FORCE_GRID)
        for k in ("hit_counts", "hit_ids", "echo", "hit_points"):
            np.testing.assert_array_equal(getattr(ra, k), getattr(rb, k), err_msg=k)
        np.testing.assert_array_equal(a.get_partials(s.n_targets, s.batch_count), b.get_partials(s.n_targets, s.batch_count))
        assert a.get_rays().shape == (30000, 3)
        with pytest.raises(native.ArtError):
            b.get_rays()
        b.set_ray_shard(2, 3, 512)
        with pytest.raises(native.ArtError) as e:
            b.run_frame(s, flags=native.FRAME_PARTIALS_ONLY)
        assert e.value.code == native.ART_E_STATE
        b.set_rays(s.ray_directions)
        b.run_frame(s, flags=native.FRAME_PARTIALS_ONLY)
