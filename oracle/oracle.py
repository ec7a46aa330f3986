"""ctypes wrapper of the CPU ORACLE (oracle/audiort_oracle.c) -- test infrastructure.

PARITY UNPINNED (see oracle/audiort_oracle.h): the reference ships no golden
vectors and cannot run here. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libaudiort_oracle.so")

JOB_RT, JOB_PM, JOB_PA = 1, 2, 4


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "audiort_oracle.c")
    hdr = os.path.join(_HERE, "audiort_oracle.h")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaudiort_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class OrCounters(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("segment_hits", C.c_uint64), ("trace_tests", C.c_uint64 * 3),
                ("echo_queries", C.c_uint64), ("echo_tests", C.c_uint64 * 3),
                ("muffle_queries", C.c_uint64), ("muffle_tests", C.c_uint64 * 3),
                ("perm_rays", C.c_uint64), ("perm_hit_rays", C.c_uint64), ("perm_first_tests", C.c_uint64 * 3),
                ("perm_pairs", C.c_uint64), ("perm_loss_tests", C.c_uint64 * 3)]

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else int(v)
        return out


class OrScene(C.Structure):
    _fields_ = [("rayOrigin", C.c_float * 3), ("rayDirections", C.c_void_p), ("rayCount", C.c_int32),
                ("aabbs", C.c_void_p), ("nAABB", C.c_int32), ("obbs", C.c_void_p), ("nOBB", C.c_int32),
                ("spheres", C.c_void_p), ("nSphere", C.c_int32),
                ("targetPositions", C.c_void_p), ("nTargets", C.c_int32),
                ("maxRayLife", C.c_float), ("maxHitsPerRay", C.c_uint8), ("maxMuffleHitDistance", C.c_float),
                ("permeationStrengthPerRay", C.c_float), ("muffleEffectiveness", C.c_float),
                ("permeationEffectiveness", C.c_float), ("maxReverbDistance", C.c_float),
                ("batchCount", C.c_int32)]


class OrOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "echoRayDistances", "rayHitResults", "rayHitResultCounts", "muffleRayHits", "permeationPowerRemains",
        "settings", "hitColliderIds", "hitDistances", "muffleTotals", "permeationSum", "settingsFp64")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.or_f32tof16.restype = C.c_uint16
        _lib.or_f32tof16.argtypes = [C.c_float]
        _lib.or_f16tof32.restype = C.c_float
        _lib.or_f16tof32.argtypes = [C.c_uint16]
        _lib.or_batch_size.restype = C.c_int32
        _lib.or_batch_size.argtypes = [C.c_int32, C.c_int32]
        _lib.or_fibonacci_directions.restype = None
        _lib.or_fibonacci_directions.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        for fn in (_lib.or_run_frame,):
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(OrScene), C.POINTER(OrOutputs), C.c_int, C.c_int, C.POINTER(OrCounters)]
        _lib.or_run_frame_faithful_q1.restype = C.c_int
        _lib.or_run_frame_faithful_q1.argtypes = [C.POINTER(OrScene), C.POINTER(OrOutputs), C.POINTER(OrCounters)]
        for fn in (_lib.or_trace_range, _lib.or_permeation_range):
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(OrScene), C.POINTER(OrOutputs), C.c_int32, C.c_int32, C.c_int, C.POINTER(OrCounters)]
    return _lib


def f32tof16(x: float) -> int:
    return int(lib().or_f32tof16(float(np.float32(x))))


def f16tof32(h: int) -> float:
    return float(lib().or_f16tof32(int(h)))


def fibonacci_directions(n_rays: int, first: int = 0, count: Optional[int] = None) -> np.ndarray:
    if count is None:
        count = n_rays - first
    out = np.zeros((count, 3), dtype=np.uint16)
    lib().or_fibonacci_directions(n_rays, first, count, out.ctypes.data)
    return out


def batch_size(n_rays: int, batch_count: int) -> int:
    return int(lib().or_batch_size(n_rays, batch_count))


@dataclass
class Frame:
    """All oracle outputs for one frame (numpy arrays) + work counters."""
    echo: np.ndarray            # uint16 [N*H]   EchoRayDistances (half bits)
    hit_points: np.ndarray      # uint16 [N*H,3] RayHitResults.HitPoint
    hit_counts: np.ndarray      # uint8  [N]     RayHitResultCounts
    muffle: np.ndarray          # uint16 [T*Na]  MuffleRayHits
    permeation: np.ndarray      # float32 [T*Na] PermeationPowerRemains
    settings: np.ndarray        # SETTINGS_DT [Na]
    hit_ids: np.ndarray         # uint32 [N*H]
    hit_dist: np.ndarray        # float32 [N*H]
    muffle_totals: np.ndarray   # uint32 [Na]
    permeation_sum: np.ndarray  # float64 [Na]
    settings_fp64: np.ndarray   # SETTINGS_DT [Na]
    counters: dict


def _scene_struct(scene, keep):
    from audio_raytracer_b200.layouts import AABB_DT, OBB_DT, SPHERE_DT
    s = OrScene()
    ro = np.asarray(scene.ray_origin, dtype=np.float32)
    s.rayOrigin = (C.c_float * 3)(*[float(v) for v in ro])
    dirs = np.ascontiguousarray(scene.ray_directions, dtype=np.uint16)
    a = np.ascontiguousarray(scene.aabbs, dtype=AABB_DT)
    o = np.ascontiguousarray(scene.obbs, dtype=OBB_DT)
    sp = np.ascontiguousarray(scene.spheres, dtype=SPHERE_DT)
    t = np.ascontiguousarray(scene.targets, dtype=np.float32)
    keep.extend([dirs, a, o, sp, t])
    s.rayDirections, s.rayCount = dirs.ctypes.data, dirs.shape[0]
    s.aabbs, s.nAABB = a.ctypes.data, len(a)
    s.obbs, s.nOBB = o.ctypes.data, len(o)
    s.spheres, s.nSphere = sp.ctypes.data, len(sp)
    s.targetPositions, s.nTargets = t.ctypes.data, t.shape[0]
    s.maxRayLife = scene.max_ray_life
    s.maxHitsPerRay = scene.max_hits_per_ray
    s.maxMuffleHitDistance = scene.max_muffle_hit_distance
    s.permeationStrengthPerRay = scene.permeation_strength_per_ray
    s.muffleEffectiveness = scene.muffle_effectiveness
    s.permeationEffectiveness = scene.permeation_effectiveness
    s.maxReverbDistance = scene.max_reverb_distance
    s.batchCount = scene.batch_count
    return s


def _alloc(scene):
    from audio_raytracer_b200.layouts import SETTINGS_DT
    N, H, Na, T = scene.n_rays, scene.max_hits_per_ray, scene.n_targets, scene.batch_count
    fr = Frame(echo=np.zeros(N * H, np.uint16), hit_points=np.zeros((N * H, 3), np.uint16),
               hit_counts=np.zeros(N, np.uint8), muffle=np.zeros(T * Na, np.uint16),
               permeation=np.zeros(T * Na, np.float32), settings=np.zeros(Na, SETTINGS_DT),
               hit_ids=np.zeros(N * H, np.uint32), hit_dist=np.zeros(N * H, np.float32),
               muffle_totals=np.zeros(Na, np.uint32), permeation_sum=np.zeros(Na, np.float64),
               settings_fp64=np.zeros(Na, SETTINGS_DT), counters={})
    o = OrOutputs()
    o.echoRayDistances = fr.echo.ctypes.data
    o.rayHitResults = fr.hit_points.ctypes.data
    o.rayHitResultCounts = fr.hit_counts.ctypes.data
    o.muffleRayHits = fr.muffle.ctypes.data
    o.permeationPowerRemains = fr.permeation.ctypes.data
    o.settings = fr.settings.ctypes.data
    o.hitColliderIds = fr.hit_ids.ctypes.data
    o.hitDistances = fr.hit_dist.ctypes.data
    o.muffleTotals = fr.muffle_totals.ctypes.data
    o.permeationSum = fr.permeation_sum.ctypes.data
    o.settingsFp64 = fr.settings_fp64.ctypes.data
    return fr, o


def run_frame(scene, jobs: int = JOB_RT | JOB_PM | JOB_PA, threads: int = 1) -> Frame:
    """Canonical evaluation of the three jobs on ``scene`` (a scenes.Scene)."""
    keep = []
    s = _scene_struct(scene, keep)
    fr, o = _alloc(scene)
    c = OrCounters()
    rc = lib().or_run_frame(C.byref(s), C.byref(o), jobs, threads, C.byref(c))
    if rc != 0:
        raise ValueError("oracle: invalid arguments")
    fr.counters = c.as_dict()
    return fr


def run_frame_faithful_q1(scene) -> Frame:
    keep = []
    s = _scene_struct(scene, keep)
    fr, o = _alloc(scene)
    c = OrCounters()
    if lib().or_run_frame_faithful_q1(C.byref(s), C.byref(o), C.byref(c)) != 0:
        raise ValueError("oracle: invalid arguments")
    fr.counters = c.as_dict()
    return fr


def trace_range(scene, first: int, count: int, threads: int = 1, with_outputs: bool = True):
    """RT over rays [first, first+count) of the full-N scene (bounded samples).

    ``with_outputs=False`` allocates nothing and returns only the counters dict
    (CPU-baseline timing of huge configs)."""
    keep = []
    s = _scene_struct(scene, keep)
    c = OrCounters()
    if not with_outputs:
        o = OrOutputs()
        if lib().or_trace_range(C.byref(s), C.byref(o), first, count, threads, C.byref(c)) != 0:
            raise ValueError("oracle: invalid arguments")
        return c.as_dict()
    fr, o = _alloc(scene)
    if lib().or_trace_range(C.byref(s), C.byref(o), first, count, threads, C.byref(c)) != 0:
        raise ValueError("oracle: invalid arguments")
    fr.counters = c.as_dict()
    return fr


def permeation_range(scene, first: int, count: int, threads: int = 1, sums: Optional[np.ndarray] = None) -> dict:
    """PM work over rays [first, first+count). ``sums`` (float64 [Na], single thread only): receives the window's per-target
    sums of the PM:260 values (the permeationSum extension)."""
    keep = []
    s = _scene_struct(scene, keep)
    o = OrOutputs()
    if sums is not None:
        if threads > 1:
            raise ValueError("permeation sums need threads=1 (order-dependent double accumulation)")
        sums[:] = 0
        o.permeationSum = sums.ctypes.data
    c = OrCounters()
    if lib().or_permeation_range(C.byref(s), C.byref(o), first, count, threads, C.byref(c)) != 0:
        raise ValueError("oracle: invalid arguments")
    return c.as_dict()
