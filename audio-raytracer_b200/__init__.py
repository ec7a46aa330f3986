"""audio-raytracer_b200: B200-native (sm_100a) implementation of the batched acoustic
hot path of FirePixel8422/Audio-Raytracer (AudioRaytracerJobBatched +
AudioPermeationJobBatched + ProcessAudioDataJob) behind a C ABI (libaudiort_cuda).

Python here is host-side plumbing only: the ctypes binding of include/audiort.h
(``native``), a mirror of the reference's job-scheduling seam (``jobs``), wire
layouts (``layouts``) and synthetic workload generation (``scenes``).
There is no CPU fallback: anything that computes calls into the CUDA library.
"""
from . import layouts, scenes  # noqa: F401

__all__ = ["layouts", "scenes", "native", "jobs"]


def __getattr__(name):
    if name in ("native", "jobs", "build"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
