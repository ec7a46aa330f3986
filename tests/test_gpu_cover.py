"""Covering-depth cull of the query kernel (k4_fan_build.cu "covering depth", k1_query_fan.cu "cull") on geometry built to
sit on its margins: goals on, just off and inside box faces, thin plates, boxes whose faces touch or coincide with cube-map
bin edges as seen from a goal, nested boxes, axis-aligned rays (zero direction components, hit points that share a
coordinate with their goal) and large coordinate offsets. Every frame is compared bit for bit with the oracle on the fan
path (the cull) and on the grid walk (no fans)."""
import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes
from helpers import aabb, micro_scene, obb, sphere

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, s):
    o = oracle.run_frame(s, threads=8)
    native.upload(ctx, s)
    for extra in (0, native.FRAME_NO_FANS):
        g = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID | extra)
        assert g.counters["gridUsed"] & 1 and bool(g.counters["gridUsed"] & 4) == (extra == 0) and not g.counters["gridUsed"] & 8
        for k in ("hit_counts", "hit_ids", "echo", "muffle", "muffle_totals"):
            np.testing.assert_array_equal(getattr(g, k), getattr(o, k), err_msg=f"{k} (flags {extra})")
        np.testing.assert_array_equal(g.permeation.view(np.uint32), o.permeation.view(np.uint32))
        assert g.counters["debugViolations"] == 0           # (-DART_Q_VERIFY builds: a culled query that turned out visible)
    return g, o


def _rays(rng, n):
    d = scenes.fibonacci_directions(n).view(np.float16).astype(np.float32).reshape(-1, 3)
    axis = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                     [1, 1, 0], [1, -1, 0], [0, 1, 1], [0, -1, 1], [1, 0, 1], [-1, 0, 1], [1, 1, 1], [-1, 1, -1]], dtype=np.float32)
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    return np.concatenate([axis, d, rng.normal(size=(32, 3)).astype(np.float32)])


def _room(half, offset=(0.0, 0.0, 0.0)):
    walls = []
    for k in range(3):
        for sgn in (-1.0, 1.0):
            c = np.zeros(3)
            c[k] = sgn * (half[k] + 0.5)
            hh = np.array(half) + 1.0
            hh[k] = 0.5
            walls.append(aabb(c + np.asarray(offset), hh, echo=0.7))
    return walls


@pytest.mark.parametrize("seed", range(20))
def test_cull_on_adversarial_boxes_and_goals(gpu_ctx, oracle, seed):
    rng = np.random.default_rng(4200 + seed)
    half = (8.0, 4.0, 8.0)
    offset = np.zeros(3) if seed % 4 else np.array([64.0, -32.0, 96.0])      # every fourth scene far from the origin
    boxes, centres, halves = [], [], []
    for _ in range(int(rng.integers(25, 70))):
        kind = rng.integers(0, 4)
        hh = rng.choice([0.125, 0.25, 0.5, 1.0, 2.0], size=3)
        if kind == 0:
            hh[rng.integers(0, 3)] = rng.choice([0.001, 0.004, 0.016, 0.0625])    # a thin plate
        elif kind == 1:
            hh[rng.integers(0, 3)] = 4.0                                          # a long beam
        c = np.round(rng.uniform(-np.array(half) + 1, np.array(half) - 1) * 4) / 4   # on a 0.25 lattice: faces touch and coincide
        centres.append(c)
        halves.append(hh)
        boxes.append(aabb(c + offset, hh, absorption=0.1, density=float(rng.uniform(0.2, 2.0)), echo=float(rng.uniform(0.2, 1.0))))
    if seed % 3 == 0:                                                             # nested boxes
        for i in range(0, 6):
            boxes.append(aabb(centres[i] + offset, halves[i] * 0.5))
    goals = []
    for i in range(7):                                                            # listener + 6 targets
        j = int(rng.integers(0, len(centres)))
        k = int(rng.integers(0, 3))
        sgn = rng.choice([-1.0, 1.0])
        mode = (seed + i) % 6
        if mode == 5:
            g = rng.uniform(-np.array(half) + 0.5, np.array(half) - 0.5)          # anywhere (maybe inside a box)
        else:
            d = [0.0, 1e-3, 0.01, 0.05, 0.3][mode]                                # on / just off / a little off a face
            g = centres[j].copy()
            g[k] += sgn * (halves[j][k] + d)
            if i % 2:
                g[(k + 1) % 3] += halves[j][(k + 1) % 3]                          # ... at an edge of that face
        goals.append(np.clip(g, -np.array(half) + 0.05, np.array(half) - 0.05) + offset)
    extra_o = [obb(rng.uniform(-4, 4, 3) + offset, rng.uniform(0.2, 1.0, 3), rot_xyz=rng.uniform(-1, 1, 3)) for _ in range(6)]
    extra_s = [sphere(rng.uniform(-5, 5, 3) + offset, float(rng.uniform(0.2, 1.0))) for _ in range(5)]
    s = micro_scene(aabbs=_room(half, offset) + boxes, obbs=extra_o, spheres=extra_s, dirs=_rays(rng, 700), origin=goals[0], targets=goals[1:], H=6,
                    max_life=200.0, max_muffle=float(rng.choice([1000.0, 12.0])), T=int(rng.integers(1, 4)))
    _check(gpu_ctx, oracle, s)


def test_cull_bins_aligned_with_box_edges(gpu_ctx, oracle):
    """a goal at the origin and plates whose edges lie exactly on cube-map bin boundaries (x / z = k / 16) and sub-bin
    boundaries (k / 32): the covering interval of a border cell must not count"""
    plates = []
    for z, m in ((2.0, 16), (4.0, 32), (8.0, 16)):
        for k in range(-3, 4, 2):
            x0, x1 = z * k / m, z * (k + 1) / m
            plates.append(aabb(((x0 + x1) / 2, 0.0, z + 0.125), ((x1 - x0) / 2, z / m, 0.125)))
            plates.append(aabb((0.0, (x0 + x1) / 2, -(z + 0.125)), (z / m, (x1 - x0) / 2, 0.125)))
    rng = np.random.default_rng(5)
    s = micro_scene(aabbs=_room((12.0, 12.0, 12.0)) + plates, dirs=_rays(rng, 1500), origin=(0.0, 0.0, 0.0),
                    targets=[(0.0, 0.0, 0.0), (0.5, 0.25, -0.125), (-3.0, 2.0, 1.0)], H=5, max_life=300.0, max_muffle=1000.0, T=2)
    _check(gpu_ctx, oracle, s)
