"""ctypes binding of include/audiort.h (libaudiort_cuda.so).

This is the Python twin of the C# P/Invoke layer shown in INTEGRATION.md. It only moves
pointers and PODs across the C ABI; all computation happens in the CUDA library. If the library
is missing or there is no CUDA device, everything here raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .layouts import AABB_DT, OBB_DT, SPHERE_DT, SETTINGS_DT

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AUDIORT_LIB") or os.path.join(_HERE, "libaudiort_cuda.so")

ART_ABI_VERSION = 3
ART_MAX_DEVICES = 16
ART_OK, ART_E_ARG, ART_E_CUDA, ART_E_PENDING, ART_E_NO_DEVICE, ART_E_STATE, ART_E_NOMEM = 0, -1, -2, -3, -4, -5, -6
JOB_RAYTRACE, JOB_PERMEATION, JOB_PROCESS, JOB_ALL = 1, 2, 4, 7
FRAME_COUNTERS, FRAME_REVERB_SEQ_FP32, FRAME_NO_HOST_OUTPUTS, FRAME_PARTIALS_ONLY, FRAME_BRUTE_FORCE, FRAME_GRID_STATS, FRAME_FORCE_GRID = 1, 2, 4, 8, 16, 32, 64
FRAME_NO_FANS = 128

EXPORTS = [
    "art_create", "art_destroy", "art_set_scene", "art_set_rays", "art_generate_fibonacci_rays", "art_get_rays",
    "art_set_ray_shard", "art_local_ray_count", "art_trace_schedule", "art_is_completed", "art_complete",
    "art_get_counters", "art_last_error", "art_partials_size", "art_get_partials", "art_partials_merge",
    "art_finalize", "art_microbench", "art_grid_build_host", "art_debug_get_fans", "art_debug_get_fan_cover", "art_comm_unique_id", "art_comm_init",
]


class ArtGridInfo(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("g0", C.c_float * 3), ("g1", C.c_float * 3),
                ("cellSize", C.c_float * 3), ("margin", C.c_float), ("nCells", C.c_int64), ("nEntries", C.c_int64)]


class ArtFanInfo(C.Structure):
    _fields_ = [("nFans", C.c_int32), ("binsPerFace", C.c_int32), ("cellsPerFan", C.c_int32), ("nearDist", C.c_float),
                ("nCells", C.c_int64), ("nEntries", C.c_int64)]


class ArtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libaudiort_cuda error {code}: {msg}")
        self.code = code


class ArtConfig(C.Structure):
    _fields_ = [("abiVersion", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("nDevices", C.c_int32),
                ("devices", C.c_int32 * ART_MAX_DEVICES), ("shardChunkRays", C.c_int32), ("reserved", C.c_int32 * 3)]


class ArtParams(C.Structure):
    _fields_ = [("rayOrigin", C.c_float * 3), ("audioTargetPositions", C.c_void_p), ("totalAudioTargets", C.c_int32),
                ("maxRayLife", C.c_float), ("maxHitsPerRay", C.c_uint8), ("maxMuffleHitDistance", C.c_float),
                ("permeationStrengthPerRay", C.c_float), ("muffleEffectiveness", C.c_float),
                ("permeationEffectiveness", C.c_float), ("maxReverbDistance", C.c_float), ("batchCount", C.c_int32),
                ("jobs", C.c_uint32), ("flags", C.c_uint32)]


class ArtOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "echoRayDistances", "rayHitResults", "rayHitResultCounts", "muffleRayHits", "permeationPowerRemains",
        "audioTargetSettings", "hitColliderIds", "muffleTotals", "permeationSum")]


class ArtCounters(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("segmentHits", C.c_uint64), ("traceTests", C.c_uint64 * 3),
                ("echoQueries", C.c_uint64), ("echoTests", C.c_uint64 * 3),
                ("muffleQueries", C.c_uint64), ("muffleTests", C.c_uint64 * 3),
                ("permRays", C.c_uint64), ("permHitRays", C.c_uint64), ("permFirstTests", C.c_uint64 * 3),
                ("permPairs", C.c_uint64), ("permLossTests", C.c_uint64 * 3),
                ("traceMs", C.c_float), ("permeationMs", C.c_float), ("reduceMs", C.c_float), ("deviceMs", C.c_float),
                ("h2dMs", C.c_float), ("d2hMs", C.c_float), ("kernelLaunches", C.c_uint32), ("gridUsed", C.c_uint32),
                ("gridTraceTests", C.c_uint64 * 3), ("gridPermFirstTests", C.c_uint64 * 3), ("gridPermLossTests", C.c_uint64 * 3),
                ("gridTraceCells", C.c_uint64), ("gridPermCells", C.c_uint64), ("debugViolations", C.c_uint64),
                ("fanBuildMs", C.c_float), ("bounceMs", C.c_float), ("queryMs", C.c_float), ("exchangeMs", C.c_float),
                ("devicesUsed", C.c_uint32), ("reserved0", C.c_uint32), ("gridQueryTests", C.c_uint64 * 3), ("gridQueryLists", C.c_uint64)]

    def as_dict(self) -> dict:
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else (float(v) if isinstance(v, float) else int(v))
        return out


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libaudiort_cuda.so and declare every prototype of include/audiort.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ArtError(ART_E_NO_DEVICE, f"{p} not found: build it with `python -m audio_raytracer_b200.build` "
                                        "(there is no CPU fallback)")
    lib = C.CDLL(p)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.art_create.restype = i32
    lib.art_create.argtypes = [C.POINTER(ArtConfig), C.POINTER(vp)]
    lib.art_destroy.restype = None
    lib.art_destroy.argtypes = [vp]
    lib.art_set_scene.restype = i32
    lib.art_set_scene.argtypes = [vp, vp, i32, vp, i32, vp, i32]
    lib.art_set_rays.restype = i32
    lib.art_set_rays.argtypes = [vp, vp, i32]
    lib.art_generate_fibonacci_rays.restype = i32
    lib.art_generate_fibonacci_rays.argtypes = [vp, i32]
    lib.art_get_rays.restype = i32
    lib.art_get_rays.argtypes = [vp, vp, i32]
    lib.art_set_ray_shard.restype = i32
    lib.art_set_ray_shard.argtypes = [vp, i32, i32, i32]
    lib.art_local_ray_count.restype = i32
    lib.art_local_ray_count.argtypes = [vp]
    lib.art_trace_schedule.restype = i32
    lib.art_trace_schedule.argtypes = [vp, C.POINTER(ArtParams), C.POINTER(ArtOutputs), C.POINTER(i32)]
    lib.art_is_completed.restype = i32
    lib.art_is_completed.argtypes = [vp, i32]
    lib.art_complete.restype = i32
    lib.art_complete.argtypes = [vp, i32]
    lib.art_get_counters.restype = i32
    lib.art_get_counters.argtypes = [vp, i32, C.POINTER(ArtCounters)]
    lib.art_last_error.restype = C.c_char_p
    lib.art_last_error.argtypes = [vp]
    lib.art_grid_build_host.restype = i32
    lib.art_grid_build_host.argtypes = [vp, i32, vp, i32, vp, i32, C.c_float, C.POINTER(ArtGridInfo), vp, i64, vp, i64]
    lib.art_debug_get_fans.restype = i32
    lib.art_debug_get_fans.argtypes = [vp, C.POINTER(ArtFanInfo), vp, i64, vp, i64]
    lib.art_debug_get_fan_cover.restype = i32
    lib.art_debug_get_fan_cover.argtypes = [vp, vp, i64, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.art_partials_size.restype = i64
    lib.art_partials_size.argtypes = [i32, i32]
    lib.art_get_partials.restype = i32
    lib.art_get_partials.argtypes = [vp, i32, vp, i64]
    lib.art_partials_merge.restype = i32
    lib.art_partials_merge.argtypes = [vp, vp, i64]
    lib.art_finalize.restype = i32
    lib.art_finalize.argtypes = [vp, i64, C.POINTER(ArtParams), i32, C.POINTER(ArtOutputs)]
    lib.art_microbench.restype = i32
    lib.art_microbench.argtypes = [vp, i32, C.POINTER(C.c_double)]
    lib.art_comm_unique_id.restype = i32
    lib.art_comm_unique_id.argtypes = [vp]
    lib.art_comm_init.restype = i32
    lib.art_comm_init.argtypes = [vp, vp, i32, i32, i32]
    if path is None:
        _lib = lib
    return lib


@dataclass
class FrameResult:
    """Outputs of one frame, local ray indexing (see ArtOutputs)."""
    echo: Optional[np.ndarray] = None            # uint16 [n*H]
    hit_points: Optional[np.ndarray] = None      # uint16 [n*H, 3]
    hit_counts: Optional[np.ndarray] = None      # uint8  [n]
    hit_ids: Optional[np.ndarray] = None         # uint32 [n*H]
    muffle: Optional[np.ndarray] = None          # uint16 [T*Na]
    permeation: Optional[np.ndarray] = None      # float32 [T*Na]
    settings: Optional[np.ndarray] = None        # SETTINGS_DT [Na]
    muffle_totals: Optional[np.ndarray] = None   # uint32 [Na]
    permeation_sum: Optional[np.ndarray] = None  # float64 [Na]
    counters: dict = field(default_factory=dict)


class Context:
    """One ArtCtx: the plugin-side state of one AudioRayTracer (ART:53-87, 241-254)."""

    def __init__(self, device: int = 0, devices=None, shard_chunk_rays: int = 0):
        """devices: several CUDA ordinals -> ONE context over those GPUs (ArtConfig.nDevices / devices[])."""
        self._lib = load_library()
        self._ctx = C.c_void_p()
        cfg = ArtConfig(abiVersion=ART_ABI_VERSION, device=device, flags=0)
        if devices is not None:
            cfg.nDevices = len(devices)
            for i, d in enumerate(devices):
                cfg.devices[i] = int(d)
            cfg.shardChunkRays = shard_chunk_rays
        rc = self._lib.art_create(C.byref(cfg), C.byref(self._ctx))
        if rc != ART_OK:
            raise ArtError(rc, (self._lib.art_last_error(None) or b"").decode())
        self._keep = []
        self._handle = C.c_int32(0)
        self._pending = None
        self.n_rays = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.art_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc < 0:
            raise ArtError(rc, (self._lib.art_last_error(self._ctx) or b"").decode())
        return rc

    # -- inputs -----------------------------------------------------------------------------
    def set_scene(self, aabbs: np.ndarray, obbs: np.ndarray, spheres: np.ndarray):
        a = np.ascontiguousarray(aabbs, dtype=AABB_DT)
        o = np.ascontiguousarray(obbs, dtype=OBB_DT)
        s = np.ascontiguousarray(spheres, dtype=SPHERE_DT)
        self._check(self._lib.art_set_scene(self._ctx, a.ctypes.data if len(a) else None, len(a),
                                            o.ctypes.data if len(o) else None, len(o),
                                            s.ctypes.data if len(s) else None, len(s)))

    def set_rays(self, half3_dirs: np.ndarray):
        d = np.ascontiguousarray(half3_dirs, dtype=np.uint16).reshape(-1, 3)
        self._check(self._lib.art_set_rays(self._ctx, d.ctypes.data, d.shape[0]))
        self.n_rays = d.shape[0]

    def generate_fibonacci_rays(self, n_rays: int):
        self._check(self._lib.art_generate_fibonacci_rays(self._ctx, n_rays))
        self.n_rays = n_rays

    def get_rays(self) -> np.ndarray:
        out = np.zeros((self.n_rays, 3), dtype=np.uint16)
        self._check(self._lib.art_get_rays(self._ctx, out.ctypes.data, self.n_rays))
        return out

    def set_ray_shard(self, shard_index: int, shard_count: int, chunk_rays: int = 0):
        self._check(self._lib.art_set_ray_shard(self._ctx, shard_index, shard_count, chunk_rays))

    def local_ray_count(self) -> int:
        return self._check(self._lib.art_local_ray_count(self._ctx))

    def comm_init(self, unique_id: bytes, rank: int, world: int, chunk_rays: int = 0):
        """Join a multi-process communicator (one rank per GPU); frames are then merged inside the library."""
        buf = (C.c_ubyte * 128).from_buffer_copy(bytes(unique_id)[:128].ljust(128, b"\0"))
        self._check(self._lib.art_comm_init(self._ctx, buf, rank, world, chunk_rays))

    # -- frames -----------------------------------------------------------------------------
    @staticmethod
    def make_params(scene, jobs: int = JOB_ALL, flags: int = 0, keep: Optional[list] = None) -> ArtParams:
        t = np.ascontiguousarray(scene.targets, dtype=np.float32)
        if keep is not None:
            keep.append(t)
        p = ArtParams()
        ro = np.asarray(scene.ray_origin, dtype=np.float32)
        p.rayOrigin = (C.c_float * 3)(float(ro[0]), float(ro[1]), float(ro[2]))
        p.audioTargetPositions = t.ctypes.data
        p.totalAudioTargets = t.shape[0]
        p.maxRayLife = scene.max_ray_life
        p.maxHitsPerRay = scene.max_hits_per_ray
        p.maxMuffleHitDistance = scene.max_muffle_hit_distance
        p.permeationStrengthPerRay = scene.permeation_strength_per_ray
        p.muffleEffectiveness = scene.muffle_effectiveness
        p.permeationEffectiveness = scene.permeation_effectiveness
        p.maxReverbDistance = scene.max_reverb_distance
        p.batchCount = scene.batch_count
        p.jobs = jobs
        p.flags = flags
        return p

    def schedule(self, scene, jobs: int = JOB_ALL, flags: int = 0, want=("echo", "hit_points", "hit_counts", "hit_ids"),
                 result: Optional[FrameResult] = None) -> int:
        """≙ the three Schedule() calls (ART:191-237). Returns the frame handle."""
        n = self.local_ray_count()
        H, Na, T = scene.max_hits_per_ray, scene.n_targets, scene.batch_count
        keep = []
        params = self.make_params(scene, jobs, flags, keep)
        r = result or FrameResult()
        host = not (flags & FRAME_NO_HOST_OUTPUTS)
        if jobs & JOB_RAYTRACE and host:
            if "echo" in want and r.echo is None:
                r.echo = np.zeros(n * H, np.uint16)
            if "hit_points" in want and r.hit_points is None:
                r.hit_points = np.zeros((n * H, 3), np.uint16)
            if "hit_counts" in want and r.hit_counts is None:
                r.hit_counts = np.zeros(n, np.uint8)
            if "hit_ids" in want and r.hit_ids is None:
                r.hit_ids = np.zeros(n * H, np.uint32)
        if r.muffle is None:
            r.muffle = np.zeros(T * Na, np.uint16)
        if r.permeation is None:
            r.permeation = np.zeros(T * Na, np.float32)
        if r.settings is None:
            r.settings = np.zeros(Na, SETTINGS_DT)
        if r.muffle_totals is None:
            r.muffle_totals = np.zeros(Na, np.uint32)
        if r.permeation_sum is None:
            r.permeation_sum = np.zeros(Na, np.float64)
        o = ArtOutputs()
        o.echoRayDistances = r.echo.ctypes.data if r.echo is not None else None
        o.rayHitResults = r.hit_points.ctypes.data if r.hit_points is not None else None
        o.rayHitResultCounts = r.hit_counts.ctypes.data if r.hit_counts is not None else None
        o.hitColliderIds = r.hit_ids.ctypes.data if r.hit_ids is not None else None
        o.muffleRayHits = r.muffle.ctypes.data
        o.permeationPowerRemains = r.permeation.ctypes.data
        o.audioTargetSettings = r.settings.ctypes.data
        o.muffleTotals = r.muffle_totals.ctypes.data
        o.permeationSum = r.permeation_sum.ctypes.data
        self._check(self._lib.art_trace_schedule(self._ctx, C.byref(params), C.byref(o), C.byref(self._handle)))
        self._pending = (r, keep, params, o)
        return int(self._handle.value)

    def is_completed(self, handle: Optional[int] = None) -> bool:
        """≙ JobHandle.IsCompleted (ART:95)."""
        return bool(self._check(self._lib.art_is_completed(self._ctx, handle or self._handle.value)))

    def complete(self, handle: Optional[int] = None) -> FrameResult:
        """≙ JobHandle.Complete() (ART:97)."""
        self._check(self._lib.art_complete(self._ctx, handle or self._handle.value))
        r = self._pending[0]
        c = ArtCounters()
        self._check(self._lib.art_get_counters(self._ctx, handle or self._handle.value, C.byref(c)))
        r.counters = c.as_dict()
        return r

    def run_frame(self, scene, jobs: int = JOB_ALL, flags: int = 0, **kw) -> FrameResult:
        self.schedule(scene, jobs, flags, **kw)
        return self.complete()

    # -- sharded frames -----------------------------------------------------------------------
    def get_partials(self, n_targets: int, batch_count: int) -> np.ndarray:
        size = int(self._lib.art_partials_size(n_targets, batch_count))
        blob = np.zeros(size, np.uint8)
        self._check(self._lib.art_get_partials(self._ctx, self._handle.value, blob.ctypes.data, size))
        return blob

    def get_fans(self):
        """(info, cells uint32 [nFans, cellsPerFan, 2], entries uint16) of the target fans the last frame built."""
        info = ArtFanInfo()
        self._check(self._lib.art_debug_get_fans(self._ctx, C.byref(info), None, 0, None, 0))
        cells = np.zeros(2 * info.nCells, dtype=np.uint32)
        entries = np.zeros(max(1, info.nEntries), dtype=np.uint16)
        self._check(self._lib.art_debug_get_fans(self._ctx, C.byref(info), cells.ctypes.data, cells.size, entries.ctypes.data, entries.size))
        return info, cells.reshape(info.nFans, info.cellsPerFan, 2), entries

    def get_fan_cover(self, info):
        """(codes uint8 [nFans, cellsPerFan, 4], logS, logK): covering-depth codes per sub-bin (sb * 2 + sa) of the last frame's
        fans; a code c < 255 stands for the depth 2 ** ((c - logK) / logS), 255 for none."""
        codes = np.zeros(info.nCells, dtype=np.uint32)
        s, k = C.c_float(0), C.c_float(0)
        self._check(self._lib.art_debug_get_fan_cover(self._ctx, codes.ctypes.data, codes.size, C.byref(s), C.byref(k)))
        return codes.view(np.uint8).reshape(info.nFans, info.cellsPerFan, 4), float(s.value), float(k.value)

    def microbench(self, kind: int) -> float:
        g = C.c_double(0)
        self._check(self._lib.art_microbench(self._ctx, kind, C.byref(g)))
        return float(g.value)


def comm_unique_id() -> bytes:
    """128-byte communicator id (rank 0 creates it, the caller distributes it to all ranks)."""
    lib = load_library()
    buf = (C.c_ubyte * 128)()
    rc = lib.art_comm_unique_id(buf)
    if rc != ART_OK:
        raise ArtError(rc, "art_comm_unique_id: libnccl.so.2 not available")
    return bytes(buf)


def merge_partials(blobs) -> np.ndarray:
    lib = load_library()
    acc = np.array(blobs[0], dtype=np.uint8, copy=True)
    for b in blobs[1:]:
        b = np.ascontiguousarray(b, dtype=np.uint8)
        rc = lib.art_partials_merge(acc.ctypes.data, b.ctypes.data, acc.size)
        if rc != ART_OK:
            raise ArtError(rc, "art_partials_merge: blobs do not match")
    return acc


def finalize(blob: np.ndarray, scene, n_rays: int, jobs: int = JOB_ALL, flags: int = 0) -> FrameResult:
    """art_finalize on a (merged) partial blob -> per-target outputs."""
    lib = load_library()
    keep = []
    params = Context.make_params(scene, jobs, flags, keep)
    Na, T = scene.n_targets, scene.batch_count
    r = FrameResult(muffle=np.zeros(T * Na, np.uint16), permeation=np.zeros(T * Na, np.float32),
                    settings=np.zeros(Na, SETTINGS_DT), muffle_totals=np.zeros(Na, np.uint32),
                    permeation_sum=np.zeros(Na, np.float64))
    o = ArtOutputs()
    o.muffleRayHits = r.muffle.ctypes.data
    o.permeationPowerRemains = r.permeation.ctypes.data
    o.audioTargetSettings = r.settings.ctypes.data
    o.muffleTotals = r.muffle_totals.ctypes.data
    o.permeationSum = r.permeation_sum.ctypes.data
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    rc = lib.art_finalize(blob.ctypes.data, blob.size, C.byref(params), n_rays, C.byref(o))
    if rc != ART_OK:
        raise ArtError(rc, "art_finalize failed")
    return r


def upload(ctx: Context, scene):
    """set_scene + set_rays for a scenes.Scene."""
    ctx.set_scene(scene.aabbs, scene.obbs, scene.spheres)
    ctx.set_rays(scene.ray_directions)


def build_grid_host(scene, cell_scale: float = 1.1):
    """The uniform grid art_set_scene would build for ``scene`` (host only, no GPU): (ArtGridInfo, cells [nCells, 2]
    uint32 = {first entry, nS | nA << 10 | nO << 21}, entries uint16), or None when the scene cannot be gridded."""
    lib = load_library()
    a, o, s = (np.ascontiguousarray(x) for x in (scene.aabbs, scene.obbs, scene.spheres))
    ptr = lambda x: x.ctypes.data_as(C.c_void_p) if len(x) else None
    info = ArtGridInfo()
    rc = lib.art_grid_build_host(ptr(a), len(a), ptr(o), len(o), ptr(s), len(s), cell_scale, C.byref(info), None, 0, None, 0)
    if rc == ART_E_STATE:
        return None
    if rc != ART_OK:
        raise ArtError(rc, "art_grid_build_host")
    cells = np.zeros((info.nCells, 2), dtype=np.uint32)
    entries = np.zeros(info.nEntries, dtype=np.uint16)
    rc = lib.art_grid_build_host(ptr(a), len(a), ptr(o), len(o), ptr(s), len(s), cell_scale, C.byref(info),
                                 cells.ctypes.data_as(C.c_void_p), cells.size, entries.ctypes.data_as(C.c_void_p), entries.size)
    if rc != ART_OK:
        raise ArtError(rc, "art_grid_build_host")
    return info, cells, entries
