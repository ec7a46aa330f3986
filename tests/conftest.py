import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def art_lib():
    """libaudiort_cuda.so, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from audio_raytracer_b200 import build, native
    build.build()
    return native.load_library()


@pytest.fixture()
def gpu_ctx(art_lib):
    from audio_raytracer_b200 import native
    ctx = native.Context(device=0)   # raises if there is no CUDA device: no CPU fallback
    yield ctx
    ctx.close()
