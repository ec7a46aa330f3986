"""Seeded synthetic workloads for the BASELINE.json configs.

Everything here is *input generation*: Fibonacci ray directions exactly as
Assets/C# Scripts/Jobs/FibonacciDirectionsJobParallel.cs:25-34 produces them, and
collider scenes in the reference's baked struct layouts (layouts.py). The OBB
rotation is stored inverted and squeezed through halfQuaternion exactly like
Assets/C# Scripts/Audio/Colliders/AudioOBBCollider.cs:59 +
DataTypes/halfQuaternion.cs:47-61 do.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Dict

import numpy as np

from .layouts import AABB_DT, OBB_DT, SPHERE_DT, f32tof16

# ScriptableObjects/AudioMaterials/{Concrete,Echo,Steel,Wood}.asset (raw half bits:
# absorption, density, echo) and the mix weights of SURVEY 8d.
MATERIALS = np.array([[0x3400, 0x3C00, 0x3C00],   # Concrete 0.25 / 1 / 1
                      [0x0000, 0x4500, 0x4200],   # Echo     0    / 5 / 3
                      [0x0000, 0x3C00, 0x3C00],   # Steel    0    / 1 / 1
                      [0x0000, 0x4500, 0x3C00]],  # Wood     0    / 5 / 1
                     dtype=np.uint16)
MATERIAL_WEIGHTS = np.array([0.1, 0.1, 0.2, 0.6])

ROOM_HALF = np.array([32.0, 8.0, 32.0], dtype=np.float32)


def fibonacci_directions(n_rays: int, first: int = 0, count: int | None = None) -> np.ndarray:
    """half3 directions [count, 3] (uint16 bits) of an ``n_rays`` Fibonacci sphere.

    FIB:25-34 in binary32; ``math.cos(float)`` is ``(float)Math.Cos((double)x)``.
    """
    if count is None:
        count = n_rays - first
    f = np.float32
    i = np.arange(first, first + count, dtype=np.int64)
    fi = i.astype(np.float32)
    phi = f(3.14159274) * (f(3.0) - np.sqrt(f(5.0)))
    with np.errstate(invalid="ignore", divide="ignore"):
        y = f(1.0) - (fi / f(n_rays - 1)) * f(2.0)
    radius = np.sqrt(f(1.0) - y * y)
    theta = (phi * fi).astype(np.float32)
    x = np.cos(theta.astype(np.float64)).astype(np.float32) * radius
    z = np.sin(theta.astype(np.float64)).astype(np.float32) * radius
    return f32tof16(np.stack([x, y, z], axis=1).astype(np.float32))


@dataclass
class Scene:
    """One frame's worth of job inputs (the fields ART:163-234 copies into the job structs)."""
    aabbs: np.ndarray
    obbs: np.ndarray
    spheres: np.ndarray
    targets: np.ndarray                     # float32 [Na, 3]  AudioTargetPositions
    ray_directions: np.ndarray              # uint16 [N, 3]    RayDirections (half3)
    ray_origin: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.65, 0.0], dtype=np.float32))
    max_ray_life: float = 1000.0
    max_hits_per_ray: int = 8               # MaxHitsPerRay = maxBounces + 1 (ART:16)
    max_muffle_hit_distance: float = 1000.0
    permeation_strength_per_ray: float = 1.0
    muffle_effectiveness: float = 1.0
    permeation_effectiveness: float = 0.5
    max_reverb_distance: float = 35.0
    batch_count: int = 1                    # T = ToUseThreadCount (ARM:19)
    name: str = "custom"

    @property
    def n_rays(self) -> int:
        return int(self.ray_directions.shape[0])

    @property
    def n_targets(self) -> int:
        return int(self.targets.shape[0])

    @property
    def n_colliders(self) -> int:
        return len(self.aabbs) + len(self.obbs) + len(self.spheres)

    def with_rays(self, first: int, count: int) -> "Scene":
        """Same scene, but RayDirections = rays [first, first+count) of the full sphere."""
        return replace(self, ray_directions=np.ascontiguousarray(self.ray_directions[first:first + count]))


def _materials(rng: np.random.Generator, n: int) -> np.ndarray:
    return MATERIALS[rng.choice(4, size=n, p=MATERIAL_WEIGHTS)]


def _store_rotation(q_world: np.ndarray) -> np.ndarray:
    """world rotation quaternions [n,4] (x,y,z,w) -> stored halfQuaternion bits [n,3].

    AudioOBBCollider.cs:59 stores math.inverse(rotation) (conjugate of a unit
    quaternion); halfQuaternion's setter flips the sign so that w >= 0 and keeps x,y,z.
    """
    q = q_world.astype(np.float32).copy()
    q[:, :3] *= np.float32(-1.0)
    neg = q[:, 3] < 0
    q[neg, :3] *= np.float32(-1.0)
    return f32tof16(q[:, :3])


def make_scene(n_aabb: int, n_obb: int, n_sphere: int, n_targets: int, seed: int,
               n_rays: int, max_hits: int, batch_count: int = 1, name: str = "custom",
               room_half=ROOM_HALF, size_range=(0.25, 1.5)) -> Scene:
    """The SURVEY 8d "room": 6 AABB walls, interior colliders, one owned OBB per target."""
    rng = np.random.default_rng(seed)
    room = np.asarray(room_half, dtype=np.float32)
    origin = np.array([0.0, 0.65, 0.0], dtype=np.float32)
    assert n_aabb >= 6 and n_obb >= n_targets

    def draw_centres(n, bound):
        """centres uniform in the room, redrawn while closer than bound+1 to the ray origin."""
        c = rng.uniform(-room, room, size=(n, 3)).astype(np.float32)
        for _ in range(64):
            bad = np.linalg.norm(c - origin, axis=1) < bound + 1.0
            if not bad.any():
                break
            c[bad] = rng.uniform(-room, room, size=(int(bad.sum()), 3)).astype(np.float32)
        return c

    # ---- AABBs: 6 walls (thickness 1, half 0.5) + interior -------------------
    aabbs = np.zeros(n_aabb, dtype=AABB_DT)
    wall_c, wall_h = [], []
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            c = np.zeros(3, dtype=np.float32)
            c[axis] = sgn * (room[axis] + 0.5)
            h = room + 1.0
            h[axis] = 0.5
            wall_c.append(c)
            wall_h.append(h)
    n_int = n_aabb - 6
    h_int = rng.uniform(*size_range, size=(n_int, 3)).astype(np.float32)
    c_int = draw_centres(n_int, np.linalg.norm(h_int, axis=1))
    aabbs["center"] = f32tof16(np.concatenate([np.array(wall_c, dtype=np.float32), c_int]))
    aabbs["size"] = f32tof16(np.concatenate([np.array(wall_h, dtype=np.float32), h_int]))
    mats = _materials(rng, n_aabb)
    aabbs["absorption"], aabbs["density"], aabbs["echo"] = mats[:, 0], mats[:, 1], mats[:, 2]
    aabbs["audioTargetId"] = -1

    # ---- targets + their owned OBBs, then interior OBBs ----------------------
    tlim = room.copy()
    tlim[1] = 6.0
    targets = rng.uniform(-tlim, tlim, size=(n_targets, 3)).astype(np.float32)
    obbs = np.zeros(n_obb, dtype=OBB_DT)
    n_int = n_obb - n_targets
    h_int = rng.uniform(*size_range, size=(n_int, 3)).astype(np.float32)
    c_int = draw_centres(n_int, np.linalg.norm(h_int, axis=1))
    q = rng.normal(size=(n_obb, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype(np.float32)
    h_own = np.tile(np.array([0.5, 1.0, 0.5], dtype=np.float32), (n_targets, 1))
    obbs["center"] = f32tof16(np.concatenate([targets, c_int]))
    obbs["size"] = f32tof16(np.concatenate([h_own, h_int]))
    obbs["rot"] = _store_rotation(q)
    mats = _materials(rng, n_obb)
    obbs["absorption"], obbs["density"], obbs["echo"] = mats[:, 0], mats[:, 1], mats[:, 2]
    obbs["audioTargetId"] = -1
    obbs["audioTargetId"][:n_targets] = np.arange(n_targets, dtype=np.int16)

    # ---- spheres -------------------------------------------------------------
    spheres = np.zeros(n_sphere, dtype=SPHERE_DT)
    r = rng.uniform(*size_range, size=n_sphere).astype(np.float32)
    spheres["center"] = f32tof16(draw_centres(n_sphere, r))
    spheres["radius"] = f32tof16(r)
    mats = _materials(rng, n_sphere)
    spheres["absorption"], spheres["density"], spheres["echo"] = mats[:, 0], mats[:, 1], mats[:, 2]
    spheres["audioTargetId"] = -1

    return Scene(aabbs=aabbs, obbs=obbs, spheres=spheres, targets=targets,
                 ray_directions=fibonacci_directions(n_rays), ray_origin=origin,
                 max_hits_per_ray=max_hits, batch_count=batch_count, name=name)


# BASELINE.json configs -> sizes. Seeds: 0xA0D10 + k.
CONFIGS: Dict[str, dict] = {
    # C2: 65,536 rays x 8 bounces vs 256 AABB + 128 OBB + 64 spheres, 1 source
    "c2": dict(n_rays=65536, max_hits=8, n_targets=1, n_aabb=256, n_obb=128, n_sphere=64, seed=0xA0D10 + 2),
    # C3: 64 sources, 1M rays x 12 bounces vs 4,096 mixed colliders (4:2:1)
    "c3": dict(n_rays=1 << 20, max_hits=12, n_targets=64, n_aabb=2340, n_obb=1170, n_sphere=586, seed=0xA0D10 + 3),
    # C4: permeation/reverb stress, 256 listener-source pairs through 4,096 colliders
    "c4": dict(n_rays=65536, max_hits=1, n_targets=256, n_aabb=2340, n_obb=1170, n_sphere=586, seed=0xA0D10 + 4),
    # C5: 8 sources x 16M rays x 16 bounces vs 16,384 mixed colliders
    "c5": dict(n_rays=1 << 24, max_hits=16, n_targets=8, n_aabb=9362, n_obb=4681, n_sphere=2341, seed=0xA0D10 + 5),
}


def load_demo_scene(n_targets: int | None = None, batch_count: int | None = None, n_rays: int | None = None) -> Scene:
    """BASELINE config 1: the reference's demo level (Assets/Scenes/Sample Scene.unity) as the job structs see it on
    the first frame, with the demo's own parameters (Prefabs/Player.prefab:223-234 + scene overrides: 314 rays,
    maxBounces 4, maxRayLife 125, maxMuffleHitDistance 250, permeation effectiveness 0, 1 worker thread).

    ``data/c1_demo_scene.npz`` is written by tools/export_demo_scene.py, which restates the collider bake of
    Audio/Colliders/*.cs offline; at run time the scene has 52 AABB + 38 OBB + 8 sphere colliders and 2 audio targets
    (the third MusicBox and 13 more colliders sit under the inactive "Environment (Box)" root and never register).
    ``n_targets=1`` keeps only the first MusicBox, as BASELINE.json words the config ("1 audio source").
    """
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "c1_demo_scene.npz"))
    p = z["params"]
    targets = z["targets"] if n_targets is None else z["targets"][:n_targets]
    n = int(z["ray_count"][0]) if n_rays is None else n_rays
    return Scene(aabbs=z["aabbs"].view(AABB_DT), obbs=z["obbs"].view(OBB_DT), spheres=z["spheres"].view(SPHERE_DT),
                 targets=np.ascontiguousarray(targets, dtype=np.float32), ray_directions=fibonacci_directions(n),
                 ray_origin=z["ray_origin"].astype(np.float32), max_ray_life=float(p[0]), max_hits_per_ray=int(p[1]),
                 max_muffle_hit_distance=float(p[2]), permeation_strength_per_ray=float(p[3]),
                 muffle_effectiveness=float(p[4]), permeation_effectiveness=float(p[5]), max_reverb_distance=float(p[6]),
                 batch_count=int(p[7]) if batch_count is None else batch_count, name="c1")


def make_config(name: str, batch_count: int = 1, n_rays: int | None = None) -> Scene:
    """Build a BASELINE config. ``n_rays`` overrides the ray count (scaled-down parity cases)."""
    if name == "c1":
        return load_demo_scene(batch_count=batch_count, n_rays=n_rays)
    if name == "c1_1src":
        return load_demo_scene(n_targets=1, batch_count=batch_count, n_rays=n_rays)
    cfg = dict(CONFIGS[name])
    if n_rays is not None:
        cfg["n_rays"] = n_rays
    return make_scene(batch_count=batch_count, name=name, **cfg)
