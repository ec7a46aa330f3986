"""Binary scene dump shared by the Python tools, the CUDA path and the headless C# harness (csharp/harness/).

BASELINE config 1 asks for the demo scene "exported to a binary dump"; the same container carries any frame's job
inputs in the reference's wire layouts (layouts.py), little endian:

    char[4] "ARTD" | int32 version (1)
    int32 nAABB, nOBB, nSphere, nTargets, nRays, maxHitsPerRay, batchCount
    float32 rayOrigin[3], maxRayLife, maxMuffleHitDistance, permeationStrengthPerRay,
            muffleEffectiveness, permeationEffectiveness, maxReverbDistance
    ColliderAABBStruct[nAABB] (20 B) | ColliderOBBStruct[nOBB] (26 B) | ColliderSphereStruct[nSphere] (16 B)
    float3 AudioTargetPositions[nTargets] | half3 RayDirections[nRays]

and the outputs of one frame ("ARTO"): half EchoRayDistances[N*H] | half3 RayHitResults[N*H] | byte RayHitResultCounts[N] |
ushort MuffleRayHits[T*Na] | float PermeationPowerRemains[T*Na] | AudioTargetRTSettings[Na] (24 B).
"""
from __future__ import annotations

import struct

import numpy as np

from .layouts import AABB_DT, OBB_DT, SETTINGS_DT, SPHERE_DT
from .scenes import Scene

MAGIC_IN, MAGIC_OUT, VERSION = b"ARTD", b"ARTO", 1
_HDR = "<4si7i9f"


def write_dump(scene: Scene, path: str) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack(_HDR, MAGIC_IN, VERSION, len(scene.aabbs), len(scene.obbs), len(scene.spheres), scene.n_targets,
                            scene.n_rays, scene.max_hits_per_ray, scene.batch_count, *[float(v) for v in scene.ray_origin],
                            scene.max_ray_life, scene.max_muffle_hit_distance, scene.permeation_strength_per_ray,
                            scene.muffle_effectiveness, scene.permeation_effectiveness, scene.max_reverb_distance))
        for arr, dt in ((scene.aabbs, AABB_DT), (scene.obbs, OBB_DT), (scene.spheres, SPHERE_DT)):
            f.write(np.ascontiguousarray(arr, dtype=dt).tobytes())
        f.write(np.ascontiguousarray(scene.targets, dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(scene.ray_directions, dtype="<u2").tobytes())


def read_dump(path: str) -> Scene:
    with open(path, "rb") as f:
        raw = f.read()
    n = struct.calcsize(_HDR)
    (magic, ver, na, no, ns, nt, nr, H, T, ox, oy, oz, life, muffle, strength, meff, peff, reverb) = struct.unpack(_HDR, raw[:n])
    if magic != MAGIC_IN or ver != VERSION:
        raise ValueError(f"{path}: not an ARTD v{VERSION} dump")
    off = n

    def take(dt, count, shape=None):
        nonlocal off
        nbytes = np.dtype(dt).itemsize * count
        a = np.frombuffer(raw, dtype=dt, count=count, offset=off).copy()
        off += nbytes
        return a.reshape(shape) if shape else a
    aabbs, obbs, spheres = take(AABB_DT, na), take(OBB_DT, no), take(SPHERE_DT, ns)
    targets = take("<f4", 3 * nt, (nt, 3))
    dirs = take("<u2", 3 * nr, (nr, 3))
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} trailing bytes")
    return Scene(aabbs=aabbs, obbs=obbs, spheres=spheres, targets=targets, ray_directions=dirs,
                 ray_origin=np.array([ox, oy, oz], dtype=np.float32), max_ray_life=life, max_hits_per_ray=H,
                 max_muffle_hit_distance=muffle, permeation_strength_per_ray=strength, muffle_effectiveness=meff,
                 permeation_effectiveness=peff, max_reverb_distance=reverb, batch_count=T, name="dump")


def write_outputs(path: str, echo, hit_points, hit_counts, muffle, permeation, settings) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<4si", MAGIC_OUT, VERSION))
        for a, dt in ((echo, "<u2"), (hit_points, "<u2"), (hit_counts, "u1"), (muffle, "<u2"), (permeation, "<f4")):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
        f.write(np.ascontiguousarray(settings, dtype=SETTINGS_DT).tobytes())


def read_outputs(path: str, scene: Scene) -> dict:
    raw = open(path, "rb").read()
    magic, ver = struct.unpack("<4si", raw[:8])
    if magic != MAGIC_OUT or ver != VERSION:
        raise ValueError(f"{path}: not an ARTO v{VERSION} file")
    N, H, Na, T = scene.n_rays, scene.max_hits_per_ray, scene.n_targets, scene.batch_count
    off, out = 8, {}
    for name, dt, count, shape in (("echo", "<u2", N * H, None), ("hit_points", "<u2", N * H * 3, (N * H, 3)),
                                   ("hit_counts", "u1", N, None), ("muffle", "<u2", T * Na, None),
                                   ("permeation", "<f4", T * Na, None), ("settings", SETTINGS_DT, Na, None)):
        a = np.frombuffer(raw, dtype=dt, count=count, offset=off).copy()
        off += np.dtype(dt).itemsize * count
        out[name] = a.reshape(shape) if shape else a
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} trailing bytes")
    return out
