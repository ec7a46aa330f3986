"""The C++ mirror of the reference's scheduling seam (include/audiort_jobs.hpp) compiles against the C ABI and behaves:
without a CUDA device the plugin refuses to start (exit code 3, no CPU fallback); with one, a frame of the reference's demo
level scheduled through the job structs equals the oracle bit for bit."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_raytracer_b200 import build, scene_io, scenes  # noqa: E402


@pytest.fixture(scope="module")
def demo_exe(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    lib = build.build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_jobs_demo")
    libdir = os.path.dirname(lib)
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "cpp", "host_jobs_demo.cpp"),
                           "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-laudiort_cuda", "-Wl,-rpath," + libdir, "-o", out])
    return out


def _dump(tmp_path, name="c1", **kw):
    s = scenes.make_config(name, **kw)
    path = str(tmp_path / f"{name}.artd")
    scene_io.write_dump(s, path)
    return s, path


def test_cpp_host_refuses_to_run_without_a_device(demo_exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _, path = _dump(tmp_path)
    r = subprocess.run([demo_exe, path, str(tmp_path / "out.arto")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "out.arto")


def test_cpp_host_rejects_a_bad_dump(demo_exe, tmp_path):
    bad = tmp_path / "bad.artd"
    bad.write_bytes(b"nope")
    r = subprocess.run([demo_exe, str(bad), str(tmp_path / "out.arto")], capture_output=True, text=True)
    assert r.returncode == 2


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", [("c1", {}), ("c2", {"n_rays": 2048, "batch_count": 2})])
def test_cpp_host_frame_equals_the_oracle(demo_exe, tmp_path, oracle, name, kw):
    s, path = _dump(tmp_path, name, **kw)
    out = str(tmp_path / "out.arto")
    r = subprocess.run([demo_exe, path, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    g = scene_io.read_outputs(out, s)
    o = oracle.run_frame(s, threads=os.cpu_count() or 1)
    for k in ("echo", "hit_counts", "muffle"):
        np.testing.assert_array_equal(g[k].reshape(-1), np.asarray(getattr(o, k)).reshape(-1), err_msg=k)
    gp, op = g["hit_points"].copy().reshape(-1), np.asarray(o.hit_points).copy().reshape(-1)
    gp[gp == 0x8000] = 0
    op[op == 0x8000] = 0
    np.testing.assert_array_equal(gp, op, err_msg="hit points")
    np.testing.assert_array_equal(g["permeation"].view(np.uint32), o.permeation.view(np.uint32))
    for f in ("muffleStrength", "reverbStrength", "reverbVolume"):
        np.testing.assert_allclose(g["settings"][f], o.settings_fp64[f], rtol=0, atol=1e-6)
