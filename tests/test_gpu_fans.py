"""Structure of the target fans (csrc/fan_dev.cuh, k4_fan_build.cu): the per-frame, per-goal direction-binned collider
lists the echo / muffle / permeation queries test instead of walking the grid. Parity of the RESULTS is covered by
test_gpu_parity.py (every case runs with and without fans); here the lists themselves are downloaded and checked against
plain geometry in numpy: nothing a ray from the goal can hit may be missing from near list + direction bin."""
import numpy as np
import pytest

from audio_raytracer_b200 import native, scenes
from audio_raytracer_b200.layouts import f16tof32

pytestmark = pytest.mark.gpu


def _bin_of(v, B):
    """cell index (within a fan) of direction v -- the mapping of fan_bin() in csrc/fan_dev.cuh"""
    a = np.abs(v)
    k = 0 if (a[0] >= a[1] and a[0] >= a[2]) else (1 if a[1] >= a[2] else 2)
    w = a[k]
    p, q = v[(k + 1) % 3] / w, v[(k + 2) % 3] / w
    ia = int(min(B - 1, max(0, np.floor((p + 1.0) * 0.5 * B))))
    ib = int(min(B - 1, max(0, np.floor((q + 1.0) * 0.5 * B))))
    return (2 * k + (1 if v[k] < 0 else 0)) * B * B + ib * B + ia


def _lists(cells, entries, fan, cell):
    first, packed = int(cells[fan, cell, 0]), int(cells[fan, cell, 1])
    nS, nA, nO = packed & 1023, (packed >> 10) & 2047, packed >> 21
    e = entries[first:first + nS + nA + nO].astype(np.int64)
    return e[:nS], e[nS:nS + nA], e[nS + nA:]


def _aabb_bounds(s):
    """min / max of every AABB as the reference forms them: fl32(center -+ size) of the half-precision fields (RT:291-292)"""
    c, h = f16tof32(s.aabbs["center"]).astype(np.float32), f16tof32(s.aabbs["size"]).astype(np.float32)
    return (c - h).astype(np.float64), (c + h).astype(np.float64)


@pytest.fixture(scope="module")
def fans():
    s = scenes.make_config("c3", n_rays=512)
    with native.Context(0) as ctx:
        native.upload(ctx, s)
        r = ctx.run_frame(s, flags=native.FRAME_FORCE_GRID)
        assert r.counters["gridUsed"] == 7
        info, cells, entries = ctx.get_fans()
        cover = ctx.get_fan_cover(info)
    return s, info, cells, entries, cover


def test_fan_layout(fans):
    s, info, cells, entries, _ = fans
    B = info.binsPerFace
    assert info.nFans == s.n_targets + 1 and info.cellsPerFan == 6 * B * B + 1
    first, packed = cells[..., 0].astype(np.int64), cells[..., 1]
    n = (packed & 1023).astype(np.int64) + ((packed >> 10) & 2047) + (packed >> 21)
    assert (first + n <= info.nEntries).all()
    assert n.sum() == info.nEntries                       # the lists tile the entry array exactly
    order = np.argsort(first.ravel(), kind="stable")
    f, m = first.ravel()[order], n.ravel()[order]
    nz = m > 0
    assert (f[nz][1:] == (f[nz] + m[nz])[:-1]).all()      # ... without gaps or overlaps
    # a collider owned by target a (its own music box) is in none of fan a's lists (RT:439, PM:255)
    own = s.obbs["audioTargetId"].astype(np.int64)
    for a in range(s.n_targets):
        mine = set(np.nonzero(own == a)[0].tolist())
        assert mine
        for cell in range(0, info.cellsPerFan, 97):
            _, _, eO = _lists(cells, entries, a, cell)
            assert not mine & set(eO.tolist())


def test_fans_are_conservative_for_spheres_and_aabbs(fans):
    """rays from a goal in random directions: every sphere / AABB the ray really meets (plain FP64 geometry on the
    un-inflated shapes) is listed in the goal's near list or in the direction's bin"""
    s, info, cells, entries, _ = fans
    B = info.binsPerFace
    rng = np.random.default_rng(7)
    sc, sr = f16tof32(s.spheres["center"]).astype(np.float64), np.abs(f16tof32(s.spheres["radius"]).astype(np.float64))
    ac, ah = f16tof32(s.aabbs["center"]).astype(np.float64), np.abs(f16tof32(s.aabbs["size"]).astype(np.float64))
    goals = list(range(0, s.n_targets, 9)) + [s.n_targets]            # a few targets + the listener
    checked = 0
    for fan in goals:
        G = (s.targets[fan] if fan < s.n_targets else s.ray_origin).astype(np.float64)
        nS0, nA0, _ = _lists(cells, entries, fan, 6 * B * B)
        for _ in range(200):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d)
            bS, bA, _ = _lists(cells, entries, fan, _bin_of(d, B))
            candS, candA = set(nS0.tolist()) | set(bS.tolist()), set(nA0.tolist()) | set(bA.tolist())
            # spheres: |oc x d| < R and the sphere not entirely behind the goal
            oc = sc - G
            t = oc @ d
            miss2 = (oc * oc).sum(1) - t * t
            hitS = np.nonzero((miss2 < sr * sr * 0.999) & (t + sr > 0))[0]
            assert set(hitS.tolist()) <= candS
            # AABBs: slab test, slightly deflated boxes
            with np.errstate(divide="ignore", invalid="ignore"):
                t0, t1 = (ac - ah * 0.999 - G) / d, (ac + ah * 0.999 - G) / d
            tn, tf = np.minimum(t0, t1).max(1), np.maximum(t0, t1).min(1)
            hitA = np.nonzero((tn <= tf) & (tf > 0))[0]
            assert set(hitA.tolist()) <= candA
            checked += len(hitS) + len(hitA)
    assert checked > 1000


def test_fans_list_every_obb_towards_its_centre_nearest_first(fans):
    s, info, cells, entries, _ = fans
    B = info.binsPerFace
    oc = f16tof32(s.obbs["center"]).astype(np.float64)
    own = s.obbs["audioTargetId"].astype(np.int64)
    for fan in (0, 17, s.n_targets):
        G = (s.targets[fan] if fan < s.n_targets else s.ray_origin).astype(np.float64)
        _, _, nearO = _lists(cells, entries, fan, 6 * B * B)
        for i in range(0, len(oc), 7):
            if own[i] == fan:
                continue
            v = oc[i] - G
            if np.abs(v).max() < 1e-6:
                continue
            _, _, eO = _lists(cells, entries, fan, _bin_of(v, B))
            assert i in set(eO.tolist()) | set(nearO.tolist()), (fan, i)
    # AABB lists are ordered by the distance of the (conservative) box from the goal, up to the inflation margin
    ac, ah = f16tof32(s.aabbs["center"]).astype(np.float64), np.abs(f16tof32(s.aabbs["size"]).astype(np.float64))
    G = s.targets[3].astype(np.float64)
    for cell in range(5, 6 * B * B, 211):
        _, eA, _ = _lists(cells, entries, 3, cell)
        if len(eA) < 2:
            continue
        dist = np.linalg.norm(np.maximum(np.maximum(ac[eA] - ah[eA] - G, G - ac[eA] - ah[eA]), 0.0), axis=1)
        assert (np.diff(dist) > -0.25).all(), (cell, dist)


def test_covering_depths_are_backed_by_one_box(fans):
    """query_fan_kernel drops every query that lies beyond the covering depth of its sub-bin (k4_fan_build.cu). The claim
    behind that: from the covering depth on (and for at least 0.4 % of it) the WHOLE sub-bin, seen from the goal, lies inside
    one AABB. Checked in FP64 against the reference's own box bounds: corners, edge midpoints, centre and random directions of
    sampled sub-bins must all be inside ONE box (not the fan's own) at a depth within one code step below the decoded one."""
    s, info, cells, entries, (codes, logS, logK) = fans
    B = info.binsPerFace
    lo, hi = _aabb_bounds(s)
    own = s.aabbs["audioTargetId"].astype(np.int64)
    goals = np.concatenate([s.targets.astype(np.float64), s.ray_origin.astype(np.float64)[None]])
    rng = np.random.default_rng(11)
    covered = codes[:, :6 * B * B, :] < 255
    assert covered.mean() > 0.5                               # C3: most sub-bins have a covering depth
    checked = 0
    for fan in list(range(0, info.nFans, 7)) + [info.nFans - 1]:
        G = goals[fan]
        usable = np.nonzero(own != fan)[0] if fan < s.n_targets else np.arange(len(lo))
        for cell in rng.choice(6 * B * B, size=160, replace=False):
            face, rb = divmod(int(cell), B * B)
            ib, ia = divmod(rb, B)
            k, neg = face >> 1, face & 1
            for sub in range(4):
                c = int(codes[fan, cell, sub])
                if c == 255:
                    continue
                depth = 2.0 ** ((c - logK) / logS)
                assert depth >= info.nearDist * 0.999
                sa, sb = sub & 1, sub >> 1
                a0, a1 = -1 + (2 * ia + sa) / B, -1 + (2 * ia + sa + 1) / B
                b0, b1 = -1 + (2 * ib + sb) / B, -1 + (2 * ib + sb + 1) / B
                ab = [(a, b) for a in (a0, (a0 + a1) / 2, a1) for b in (b0, (b0 + b1) / 2, b1)]
                ab += list(zip(rng.uniform(a0, a1, 8), rng.uniform(b0, b1, 8)))
                # the code is the threshold w1 * (1 + 2^-9) rounded UP to the next of 253 log-spaced steps, so w1 lies within one
                # step below the decoded depth; the box holds the sub-bin from w1 on for at least 0.4 % of it
                step = 2.0 ** (1.0 / logS)
                ok = False
                for z in depth / step ** np.linspace(0.0, 1.05, 22):
                    pts = []
                    for a, b in ab:
                        v = np.zeros(3)
                        v[k], v[(k + 1) % 3], v[(k + 2) % 3] = (-1.0 if neg else 1.0), a, b
                        pts.append(G + z * v)
                    pts = np.array(pts)
                    inside = ((pts[:, None, :] >= lo[None, usable]) & (pts[:, None, :] <= hi[None, usable])).all(axis=2)   # [point, box]
                    if inside.all(axis=0).any():
                        ok = True
                        break
                assert ok, f"fan {fan} cell {cell} sub-bin {sub}: no single AABB holds the sub-bin near depth {depth}"
                checked += 1
    assert checked > 500
