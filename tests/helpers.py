"""Shared builders for hand-made micro scenes (tests only)."""
import numpy as np

from audio_raytracer_b200.layouts import AABB_DT, OBB_DT, SPHERE_DT, f32tof16
from audio_raytracer_b200.scenes import Scene


def h(x):
    return f32tof16(np.asarray(x, dtype=np.float32))


def aabb(center, half, absorption=0.0, density=1.0, echo=1.0, target=-1):
    a = np.zeros(1, AABB_DT)
    a["center"], a["size"] = h(center), h(half)
    a["absorption"], a["density"], a["echo"] = h(absorption), h(density), h(echo)
    a["audioTargetId"] = target
    return a


def obb(center, half, rot_xyz=(0, 0, 0), absorption=0.0, density=1.0, echo=1.0, target=-1):
    a = np.zeros(1, OBB_DT)
    a["center"], a["size"], a["rot"] = h(center), h(half), h(rot_xyz)
    a["absorption"], a["density"], a["echo"] = h(absorption), h(density), h(echo)
    a["audioTargetId"] = target
    return a


def sphere(center, radius, absorption=0.0, density=1.0, echo=1.0, target=-1):
    a = np.zeros(1, SPHERE_DT)
    a["center"], a["radius"] = h(center), h(radius)
    a["absorption"], a["density"], a["echo"] = h(absorption), h(density), h(echo)
    a["audioTargetId"] = target
    return a


def cat(dt, *items):
    return np.concatenate(items).astype(dt) if items else np.zeros(0, dt)


def micro_scene(aabbs=(), obbs=(), spheres=(), dirs=((0, 0, 1),), origin=(0, 0, 0), targets=((100, 100, 100),),
                H=2, max_life=100.0, max_muffle=1.0, T=1, **kw):
    return Scene(aabbs=cat(AABB_DT, *aabbs), obbs=cat(OBB_DT, *obbs), spheres=cat(SPHERE_DT, *spheres),
                 targets=np.asarray(targets, dtype=np.float32).reshape(-1, 3),
                 ray_directions=h(np.asarray(dirs, dtype=np.float32).reshape(-1, 3)),
                 ray_origin=np.asarray(origin, dtype=np.float32), max_ray_life=max_life, max_hits_per_ray=H,
                 max_muffle_hit_distance=max_muffle, batch_count=T, **kw)


def hit_id(type_code, index):
    return (type_code << 30) | index


# ---------------------------------------------------------------------------------------------
# golden fixtures and partial-result blobs
# ---------------------------------------------------------------------------------------------
def load_golden(name):
    """tests/golden/<name>.npz -> (Scene, dict of frozen oracle outputs)."""
    import os
    from audio_raytracer_b200.layouts import SETTINGS_DT
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    p = z["params"]
    s = Scene(aabbs=z["aabbs"].view(AABB_DT), obbs=z["obbs"].view(OBB_DT), spheres=z["spheres"].view(SPHERE_DT),
              targets=z["targets"], ray_directions=z["ray_directions"], ray_origin=z["ray_origin"],
              max_ray_life=float(p[0]), max_hits_per_ray=int(p[1]), max_muffle_hit_distance=float(p[2]),
              permeation_strength_per_ray=float(p[3]), muffle_effectiveness=float(p[4]),
              permeation_effectiveness=float(p[5]), max_reverb_distance=float(p[6]), batch_count=int(p[7]), name=name)
    out = {k: z[k] for k in z.files}
    out["settings"] = z["settings"].view(SETTINGS_DT)
    out["settings_fp64"] = z["settings_fp64"].view(SETTINGS_DT)
    return s, out


N_COUNTERS = 38          # scene_dev.cuh CounterIdx::C_COUNT
BLOB_MAGIC = 0x41525442  # "ARTB"


def blob_dtype(n_targets, batch_count):
    """numpy mirror of the partial-result blob (include/audiort.h, 'sharded frames')."""
    def pad8(n):
        return (n + 7) & ~7
    T, Na = batch_count, n_targets
    fields = [("magic", "<u4"), ("nTargets", "<i4"), ("batchCount", "<i4"), ("shards", "<i4"),
              ("fixedLo", "<i8"), ("fixedHi", "<i8"), ("zeros", "<u8"), ("entries", "<u8"),
              ("posInf", "<u8"), ("negInf", "<u8"), ("nan", "<u8"),
              ("seqTotal", "<f4"), ("seqZeros", "<f4"), ("seqValid", "<u4"), ("pad", "<u4"),
              ("counters", "<u8", (N_COUNTERS,)),
              ("lastHitRay", "<i4", (T,)), ("padA", "u1", (pad8(4 * T) - 4 * T,)),
              ("muffleCounts", "<u4", (T * Na,)), ("padB", "u1", (pad8(4 * T * Na) - 4 * T * Na,)),
              ("permLast", "<f4", (T * Na,)), ("padC", "u1", (pad8(4 * T * Na) - 4 * T * Na,)),
              ("permSumInt", "<i8", (Na,)), ("permSumFrac", "<i8", (Na,))]
    return np.dtype(fields)


def echo_fixed_sums(echo_half_bits):
    """exact sum of half values * 2^24, split like echo_stats_kernel (hi = >>20, lo = & 0xFFFFF)."""
    h = np.asarray(echo_half_bits, dtype=np.uint16).astype(np.int64)
    mag = h & 0x7FFF
    e, m = mag >> 10, mag & 1023
    fx = np.where(e == 0, m, (1024 + m) << np.maximum(e - 1, 0))
    fx = np.where(mag == 0, 0, fx)
    sign = np.where(h & 0x8000, -1, 1)
    return int((sign * (fx & 0xFFFFF)).sum()), int((sign * (fx >> 20)).sum()), int((mag == 0).sum())
