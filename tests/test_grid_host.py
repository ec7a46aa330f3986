"""Host-side logic of the acceleration structure (csrc/grid_host.h through art_grid_build_host; no GPU needed):
every collider must be listed in every cell its true bounds touch -- the property the traversal kernels rely on."""
import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes
from audio_raytracer_b200.layouts import f16tof32
from helpers import aabb, micro_scene, obb, sphere


def _cell_lists(info, cells, entries):
    """-> dict type -> list of sets (one per cell)"""
    nS = cells[:, 1] & 1023
    nA = (cells[:, 1] >> 10) & 2047
    nO = cells[:, 1] >> 21
    out = {"s": [], "a": [], "o": []}
    for c in range(len(cells)):
        b = int(cells[c, 0])
        out["s"].append(set(entries[b:b + nS[c]].tolist()))
        out["a"].append(set(entries[b + nS[c]:b + nS[c] + nA[c]].tolist()))
        out["o"].append(set(entries[b + nS[c] + nA[c]:b + nS[c] + nA[c] + nO[c]].tolist()))
    return out


def _cells_touching(info, lo, hi):
    g0 = np.array(list(info.g0)); cs = np.array(list(info.cellSize)); n = np.array([info.nx, info.ny, info.nz])
    i0 = np.clip(np.floor((lo - g0) / cs).astype(int), 0, n - 1)
    i1 = np.clip(np.floor((hi - g0) / cs).astype(int), 0, n - 1)
    for z in range(i0[2], i1[2] + 1):
        for y in range(i0[1], i1[1] + 1):
            for x in range(i0[0], i1[0] + 1):
                yield (z * info.ny + y) * info.nx + x


def _quat_matrix(xyz):
    x, y, z = xyz
    w = np.sqrt(max(0.0, 1.0 - (x * x + y * y + z * z)))
    q = np.array([x, y, z, w]); q /= np.linalg.norm(q)
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


@pytest.mark.parametrize("name,kw", [("c2", {}), ("c3", {}), ("c1", {}), ("c5", {})])
def test_every_collider_is_listed_wherever_it_reaches(art_lib, name, kw):
    s = scenes.make_config(name, n_rays=8)
    info, cells, entries = native.build_grid_host(s)
    assert info.nx * info.ny * info.nz == len(cells) and int(cells[-1, 0]) <= len(entries)
    lists = _cell_lists(info, cells, entries)
    # lists are sorted by canonical index and hold every collider at least once
    for t, arr in (("s", s.spheres), ("a", s.aabbs), ("o", s.obbs)):
        assert set().union(*lists[t]) == set(range(len(arr)))
    rng = np.random.default_rng(1)
    for i in rng.choice(len(s.aabbs), size=min(200, len(s.aabbs)), replace=False):
        c, h = f16tof32(s.aabbs["center"][i]).astype(np.float64), np.abs(f16tof32(s.aabbs["size"][i]).astype(np.float64))
        for cell in _cells_touching(info, c - h, c + h):
            assert int(i) in lists["a"][cell]
    for i in rng.choice(len(s.spheres), size=min(200, len(s.spheres)), replace=False):
        c, r = f16tof32(s.spheres["center"][i]).astype(np.float64), abs(float(f16tof32(s.spheres["radius"][i:i + 1])[0]))
        for cell in _cells_touching(info, c - r, c + r):
            assert int(i) in lists["s"][cell]
    for i in rng.choice(len(s.obbs), size=min(200, len(s.obbs)), replace=False):
        c, h = f16tof32(s.obbs["center"][i]).astype(np.float64), np.abs(f16tof32(s.obbs["size"][i]).astype(np.float64))
        R = _quat_matrix(f16tof32(s.obbs["rot"][i]).astype(np.float64))
        corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]) * h
        for M in (R, R.T):                       # both rotation senses (RT:314-320 vs PM:172-179, quirk Q4)
            w = corners @ M.T
            for cell in _cells_touching(info, c + w.min(0), c + w.max(0)):
                assert int(i) in lists["o"][cell]


def test_grid_of_a_single_collider_and_of_nothing(art_lib):
    s = micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1))])
    info, cells, entries = native.build_grid_host(s)
    assert len(cells) >= 1 and ((cells[:, 1] >> 10) & 2047).max() == 1
    assert info.g0[2] < 4.0 and info.g1[2] > 6.0 and info.margin > 0
    assert native.build_grid_host(micro_scene()) is None            # empty scene: brute-force kernels handle it


def test_degenerate_scenes_are_refused_not_mis_gridded(art_lib):
    # a collider with an infinite half extent cannot be bounded: the library must fall back to brute force
    s = micro_scene(aabbs=[aabb((0, 0, 5), (1, 1, 1))])
    s.aabbs["size"][0, 0] = 0x7C00
    assert native.build_grid_host(s) is None
    # 1,100 coincident spheres exceed the per-cell list limit
    many = [sphere((1, 2, 3), 0.5) for _ in range(1100)]
    assert native.build_grid_host(micro_scene(spheres=many)) is None
