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
