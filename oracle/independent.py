"""independent.py -- a SECOND, independently written restatement of the reference's hot path, in numpy float32.

TEST INFRASTRUCTURE ONLY (tier rule 3): nothing in the product imports this. Purpose: the reference cannot run in this image
or on the GPU box (C#/Unity; `dotnet`, `mono`, `mcs` probed on both: profiles/r02_dotnet_probe.txt), so the C oracle
(oracle/audiort_oracle.c) cannot be pinned against it. This module removes the transcription risk instead: it was written
from the reference sources alone, shares no code with the C oracle (nor its structure: the C oracle walks ray by ray,
collider by collider; this one evaluates whole (ray x collider) planes at once), and tests/test_independent_restatement.py
requires both to agree bit for bit on the committed goldens. What it cannot remove is a shared misreading of
Unity.Mathematics 1.3.2 (not in the reference tree, Packages/packages-lock.json:57-58): those semantics are restated from
the package's published source in SURVEY.md Appendix A -- "parity unpinned" stands.

Reference lines followed (Assets/C# Scripts/...):
  Jobs/AudioRaytracerJobBatched.cs:61-215  Execute            -> trace()
  :225-280 ShootRayCast, :284-355 intersection tests, :365-449 CanRaySeePoint / CanRaySeeAudioTarget, :456-532 ReflectRay
  Jobs/AudioPermeationJobBatched.cs:34-91, 101-141, 172-179, 225-328 -> permeation()
  DataTypes/halfQuaternion.cs:34-46 (rotation getter), Utility/HalfDataTypesUtility.cs:86-90 (Half.Multiply)
Every arithmetic step is one numpy float32 ufunc call, i.e. one IEEE binary32 rounding, in the reference's order.
"""
from __future__ import annotations

import numpy as np

F = np.float32
EPS = F(0.0001)                      # RT:57 / PM:30


# ---- Unity.Mathematics 1.3.2 primitives ---------------------------------------------------------------
def f16_to_f32(h):
    """math.f16tof32: exact."""
    return np.asarray(h, np.uint16).view(np.float16).astype(F)


def f32_to_f16(x):
    """math.f32tof16: drop 12 mantissa bits, scale by 2^-112 (subnormals kept), clamp, +0x1000, >>13 = nearest, ties away."""
    x = np.asarray(x, F)
    ux = x.view(np.uint32)
    uux = ux & np.uint32(0x7FFFF000)
    with np.errstate(all="ignore"):
        sb = (uux.view(F) * F(1.92592994e-34)).view(np.uint32)
    sb = np.minimum(sb, np.uint32(0x0F7FF000))
    h = (sb + np.uint32(0x1000)) >> np.uint32(13)
    inf = np.uint32(255 << 23)
    h = np.where(uux >= inf, np.where(uux > inf, np.uint32(0x7E00), np.uint32(0x7C00)), h)
    return (h | ((ux & ~np.uint32(0x7FFFF000)) >> np.uint32(16))).astype(np.uint16)


def um_min(x, y):
    return np.where(np.isnan(y) | (x < y), x, y)


def um_max(x, y):
    return np.where(np.isnan(y) | (x > y), x, y)


def dot3(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross(a, b):
    # (a * b.yzx - a.yzx * b).yzx
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def qmul(q, v):
    """math.mul(quaternion, float3): t = 2 * cross(q.xyz, v); v + q.w * t + cross(q.xyz, t)."""
    t = tuple(F(2) * c for c in cross(q[:3], v))
    c2 = cross(q[:3], t)
    return tuple((v[k] + q[3] * t[k]) + c2[k] for k in range(3))


def qinverse(q):
    """math.inverse(quaternion): rcp(dot(q, q)) * q * (-1, -1, -1, 1)."""
    with np.errstate(all="ignore"):
        r = F(1) / (((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3])
    return ((r * q[0]) * F(-1), (r * q[1]) * F(-1), (r * q[2]) * F(-1), (r * q[3]) * F(1))


def rotation_getter(hx, hy, hz):
    """halfQuaternion.QuaternionValue get (DataTypes/halfQuaternion.cs:34-46)."""
    xx, yy, zz = f16_to_f32(hx), f16_to_f32(hy), f16_to_f32(hz)
    w2 = F(1) - ((xx * xx + yy * yy) + zz * zz)
    with np.errstate(all="ignore"):
        w = np.where(w2 > 0, np.sqrt(np.maximum(w2, F(0))), F(0)).astype(F)
        rs = F(1) / np.sqrt(((xx * xx + yy * yy) + zz * zz) + w * w)      # math.normalize = rsqrt(dot(q, q)) * q
    return (rs * xx, rs * yy, rs * zz, rs * w)


def normalize3(v):
    with np.errstate(all="ignore"):
        rs = F(1) / np.sqrt(dot3(v, v))
    return (rs * v[0], rs * v[1], rs * v[2])


def sign(x):
    return (x > 0).astype(F) - (x < 0).astype(F)


# ---- scene decoding ---------------------------------------------------------------------------------------------------------
class Colliders:
    def __init__(self, scene):
        s, a, o = scene.spheres, scene.aabbs, scene.obbs
        self.ns, self.na, self.no = len(s), len(a), len(o)
        self.sC = tuple(f16_to_f32(s["center"][:, k]) for k in range(3))
        self.sR = f16_to_f32(s["radius"])
        self.aC = tuple(f16_to_f32(a["center"][:, k]) for k in range(3))
        self.aH = tuple(f16_to_f32(a["size"][:, k]) for k in range(3))
        self.oC = tuple(f16_to_f32(o["center"][:, k]) for k in range(3))
        self.oH = tuple(f16_to_f32(o["size"][:, k]) for k in range(3))
        self.oQ = rotation_getter(o["rot"][:, 0], o["rot"][:, 1], o["rot"][:, 2])
        self.oQinv = qinverse(self.oQ)
        mat = ("absorption", "density", "echo")                                          # CS/AudioMaterialProperties.cs:7-16
        self.sMat = {n: f16_to_f32(s[n]) for n in mat}
        self.aMat = {n: f16_to_f32(a[n]) for n in mat}
        self.oMat = {n: f16_to_f32(o[n]) for n in mat}
        self.sOwn, self.aOwn, self.oOwn = (x["audioTargetId"].astype(np.int32) for x in (s, a, o))


def _col(x):      # per-collider values along axis 1
    return x[None, :]


def _row(x):      # per-ray values along axis 0
    return x[:, None]


# ---- intersection tests on (ray x collider) planes: (hit mask, distance) ------------------------------------------------------
def slab(o, d, lo, hi):
    """RT:284-308 given min / max planes."""
    with np.errstate(all="ignore"):
        inv = tuple(F(1) / d[k] for k in range(3))
        t0 = tuple((lo[k] - o[k]) * inv[k] for k in range(3))
        t1 = tuple((hi[k] - o[k]) * inv[k] for k in range(3))
        tmin = tuple(um_min(t0[k], t1[k]) for k in range(3))
        tmax = tuple(um_max(t0[k], t1[k]) for k in range(3))
        tnear = um_max(um_max(tmin[0], tmin[1]), tmin[2])
        tfar = um_min(um_min(tmax[0], tmax[1]), tmax[2])
        miss = (tnear > tfar) | (tfar < 0)
        dist = np.where(tnear > 0, tnear, tfar)
    return ~miss, dist


def hit_aabb(C, o, d):
    oo = tuple(_row(o[k]) for k in range(3)); dd = tuple(_row(d[k]) for k in range(3))
    lo = tuple(_col(C.aC[k] - C.aH[k]) for k in range(3))
    hi = tuple(_col(C.aC[k] + C.aH[k]) for k in range(3))
    return slab(oo, dd, lo, hi)


def hit_obb(C, o, d, q):
    """RT:314-320 with rotation q per collider (RT passes Rotation, PM:172-179 its inverse)."""
    qq = tuple(_col(q[k]) for k in range(4))
    rel = tuple(_row(o[k]) - _col(C.oC[k]) for k in range(3))
    lo_ = qmul(qq, rel)
    ld = qmul(qq, tuple(np.broadcast_to(_row(d[k]), rel[0].shape) for k in range(3)))
    zero = F(0)
    lo = tuple(zero - _col(C.oH[k]) for k in range(3))
    hi = tuple(zero + _col(C.oH[k]) for k in range(3))
    return slab(lo_, ld, lo, hi)


def hit_sphere(C, o, d):
    """RT:323-355."""
    with np.errstate(all="ignore"):
        oc = tuple(_row(o[k]) - _col(C.sC[k]) for k in range(3))
        dd = tuple(_row(d[k]) for k in range(3))
        a = dot3(dd, dd)
        b = F(2) * dot3(oc, dd)
        c = dot3(oc, oc) - _col(C.sR * C.sR)
        disc = b * b - (F(4) * a) * c
        sq = np.sqrt(np.maximum(disc, F(0)))
        t0 = (-b - sq) / (F(2) * a)
        t1 = (-b + sq) / (F(2) * a)
        ok = disc >= 0
        hit = ok & ((t0 >= 0) | (t1 >= 0))
        dist = np.where(t0 >= 0, t0, t1)
    return hit, dist


def can_see(C, o, d, limit, skip_owner=None):
    """RT:365-397 (skip_owner None) / RT:405-449: True where NO collider reports a distance < limit."""
    blocked = np.zeros(o[0].shape[0], bool)
    for (hit, dist), own in ((hit_sphere(C, o, d), C.sOwn), (hit_aabb(C, o, d), C.aOwn), (hit_obb(C, o, d, C.oQ), C.oOwn)):
        m = hit & (dist < _row(limit))
        if skip_owner is not None:
            m &= _col(own != skip_owner)
        blocked |= m.any(axis=1)
    return ~blocked


# ---- AudioRaytracerJobBatched.Execute (canonical arrays: zeroed, then written) ---------------------------------------
def trace(scene):
    C = Colliders(scene)
    N, H, Na, T = scene.n_rays, scene.max_hits_per_ray, scene.n_targets, scene.batch_count
    dirs = scene.ray_directions.reshape(-1, 3)
    d = [f16_to_f32(dirs[:, k]) for k in range(3)]                                   # RT:94
    RO = [F(scene.ray_origin[k]) for k in range(3)]
    o = [np.full(N, RO[k], F) for k in range(3)]                                     # RT:95
    life = np.full(N, F(scene.max_ray_life), F)
    hits = np.zeros(N, np.int32)
    alive = np.ones(N, bool)
    echo = np.zeros(N * H, np.uint16)
    hit_ids = np.zeros(N * H, np.uint32)
    hit_points = np.zeros((N * H, 3), np.uint16)
    muffle_rows = np.zeros((T, Na), np.int64)
    bsz = int(max(1.0, np.ceil(F(N) / F(T))))                                        # ART:161
    batch_of = np.arange(N) // bsz
    row_of = (batch_of.astype(np.int64) * bsz * T) // N                              # RT:63-64 (no int32 wrap at golden sizes)
    targets = np.asarray(scene.targets, F).reshape(-1, 3)
    segments = 0
    for _ in range(H + 1):
        idx = np.nonzero(alive)[0]
        if idx.size == 0:
            break
        segments += idx.size
        oo = tuple(o[k][idx] for k in range(3)); dd = tuple(d[k][idx] for k in range(3))
        # ShootRayCast: spheres, then AABBs, then OBBs, strict '<' -> first minimum in that order (RT:244, 257, 270)
        hs, ds = hit_sphere(C, oo, dd); ha, da = hit_aabb(C, oo, dd); ho, do_ = hit_obb(C, oo, dd, C.oQ)
        allh = np.concatenate([hs, ha, ho], axis=1)
        alld = np.where(allh, np.concatenate([ds, da, do_], axis=1), F(np.inf))
        alld = np.where(np.isnan(alld), F(np.inf), alld)
        win = np.argmin(alld, axis=1)                                                # first index among equal minima
        best = alld[np.arange(idx.size), win]
        got = allh[np.arange(idx.size), win] & (best < F(3.402823466e+38))           # closestDist = float.MaxValue, strict '<'
        alive[idx[~got]] = False                                                     # ray left the scene (RT:201-207)
        idx, win, best = idx[got], win[got], best[got]
        if idx.size == 0:
            continue
        for k in range(3):
            o[k][idx] = o[k][idx] + d[k][idx] * best                                 # RT:111
        life[idx] = life[idx] - best                                                 # RT:112
        hits[idx] += 1                                                               # RT:113
        rid = idx * H + hits[idx] - 1                                                # RT:115
        typ = np.where(win < C.ns, 3, np.where(win < C.ns + C.na, 1, 2)).astype(np.uint32)   # Enums/ColliderType.cs
        loc = np.where(win < C.ns, win, np.where(win < C.ns + C.na, win - C.ns, win - C.ns - C.na))
        hit_ids[rid] = (typ << np.uint32(30)) | loc.astype(np.uint32)
        for k in range(3):
            hit_points[rid, k] = f32_to_f16(o[k][idx])                               # RT:118
        oo = tuple(o[k][idx] for k in range(3)); dd = tuple(d[k][idx] for k in range(3))
        off = tuple(oo[k] - dd[k] * EPS for k in range(3))                           # RT:124
        back = normalize3(tuple(RO[k] - off[k] for k in range(3)))                   # RT:127
        w = tuple(oo[k] - RO[k] for k in range(3))
        dist0 = np.sqrt(dot3(w, w))                                                  # RT:130 math.distance(RayOrigin, cRayOrigin)
        sees = can_see(C, off, back, dist0)
        mat_echo = np.empty(idx.size, F)
        for t_, tab in ((3, C.sMat), (1, C.aMat), (2, C.oMat)):
            m = typ == t_
            mat_echo[m] = tab["echo"][loc[m]]
        echo[rid[sees]] = f32_to_f16(dist0[sees] * mat_echo[sees])                   # RT:135-144, Half.Multiply(float, float)
        for a in range(Na):                                                          # RT:153-173
            v = tuple(targets[a, k] - off[k] for k in range(3))
            dirT = normalize3(v)
            distT = np.sqrt(dot3(v, v))                                              # math.distance(offsetted, target) = length(target - offsetted)
            gate = distT < F(scene.max_muffle_hit_distance)
            if gate.any():
                vis = can_see(C, tuple(x[gate] for x in off), tuple(x[gate] for x in dirT), distT[gate], skip_owner=a)
                np.add.at(muffle_rows[:, a], row_of[idx[gate][vis]], 1)
        # termination / reflection (RT:178-193)
        done = (hits[idx] >= H) | (life[idx] <= 0)
        alive[idx[done]] = False
        ridx = idx[~done]
        if ridx.size:
            t2, l2 = typ[~done], loc[~done]
            n = [np.zeros(ridx.size, F) for _ in range(3)]
            absorption = np.zeros(ridx.size, F)
            po = tuple(o[k][ridx] for k in range(3))
            m = t2 == 1                                                              # AABB: RT:464-484
            if m.any():
                lp = tuple(po[k][m] - C.aC[k][l2[m]] for k in range(3))
                e = tuple(C.aH[k][l2[m]] - np.abs(lp[k]) for k in range(3))
                cx = (e[0] < e[1]) & (e[0] < e[2]); cy = ~cx & (e[1] < e[0]) & (e[1] < e[2]); cz = ~cx & ~cy
                n[0][m] = np.where(cx, sign(lp[0]), F(0)); n[1][m] = np.where(cy, sign(lp[1]), F(0)); n[2][m] = np.where(cz, sign(lp[2]), F(0))
                absorption[m] = C.aMat["absorption"][l2[m]]
            m = t2 == 2                                                              # OBB: RT:487-512 (quirk Q3: inverse for the point, Rotation for the normal)
            if m.any():
                qi = tuple(C.oQinv[k][l2[m]] for k in range(4)); q = tuple(C.oQ[k][l2[m]] for k in range(4))
                lh = qmul(qi, tuple(po[k][m] - C.oC[k][l2[m]] for k in range(3)))
                e = tuple(C.oH[k][l2[m]] - np.abs(lh[k]) for k in range(3))
                cx = (e[0] < e[1]) & (e[0] < e[2]); cy = ~cx & (e[1] < e[0]) & (e[1] < e[2]); cz = ~cx & ~cy
                ln = (np.where(cx, sign(lh[0]), F(0)), np.where(cy, sign(lh[1]), F(0)), np.where(cz, sign(lh[2]), F(0)))
                wn = qmul(q, ln)
                for k in range(3):
                    n[k][m] = wn[k]
                absorption[m] = C.oMat["absorption"][l2[m]]
            m = t2 == 3                                                              # sphere: RT:515-518
            if m.any():
                wn = normalize3(tuple(po[k][m] - C.sC[k][l2[m]] for k in range(3)))
                for k in range(3):
                    n[k][m] = wn[k]
                absorption[m] = C.sMat["absorption"][l2[m]]
            dr = tuple(d[k][ridx] for k in range(3))
            dn = dot3(dr, n)                                                         # math.reflect(i, n) = i - 2 * n * dot(i, n)
            for k in range(3):
                d[k][ridx] = dr[k] - (F(2) * n[k]) * dn
                o[k][ridx] = po[k] + d[k][ridx] * EPS                                # RT:528
            life[ridx] = life[ridx] - F(scene.max_ray_life) * absorption             # RT:531
            alive[ridx[life[ridx] < 0]] = False                                      # RT:189
    return dict(echo=echo, hit_ids=hit_ids, hit_points=hit_points, hit_counts=hits.astype(np.uint8),
                muffle=(muffle_rows % 65536).astype(np.uint16).reshape(-1), muffle_totals=muffle_rows.sum(axis=0).astype(np.uint32),
                segments=segments)


# ---- AudioPermeationJobBatched.Execute: per-ray values (the reference keeps only the last writer) -----------------------
def permeation(scene):
    """Returns (first-hit mask [N], values [N, Na]) with values[r, a] = N * S - loss (PM:260) for every ray with a first hit."""
    C = Colliders(scene)
    N, Na = scene.n_rays, scene.n_targets
    dirs = scene.ray_directions.reshape(-1, 3)
    d = tuple(f16_to_f32(dirs[:, k]) for k in range(3))
    o = tuple(np.full(N, F(scene.ray_origin[k]), F) for k in range(3))
    # PM:101-141 nearest distance; PM:172-179 applies math.inverse to the stored rotation (quirk Q4)
    hs, ds = hit_sphere(C, o, d); ha, da = hit_aabb(C, o, d); ho, do_ = hit_obb(C, o, d, C.oQinv)
    alld = np.where(np.concatenate([hs, ha, ho], axis=1), np.concatenate([ds, da, do_], axis=1), F(np.inf))
    alld = np.where(np.isnan(alld), F(np.inf), alld)
    t = alld.min(axis=1)
    hit = t < F(np.inf)
    P = tuple((o[k] + d[k] * t)[hit] for k in range(3))                              # PM:61
    dh = tuple(d[k][hit] for k in range(3))
    off = tuple(P[k] - dh[k] * EPS for k in range(3))                                # PM:73
    targets = np.asarray(scene.targets, F).reshape(-1, 3)
    NS = F(F(N) * F(scene.permeation_strength_per_ray))
    vals = np.zeros((int(hit.sum()), Na), F)
    for a in range(Na):
        dirT = normalize3(tuple(targets[a, k] - off[k] for k in range(3)))          # PM:76
        loss = np.zeros(off[0].shape[0], F)
        # PM:303-328 spheres (unit direction assumed): b = dot(oc, d), disc = b*b - c
        with np.errstate(all="ignore"):
            oc = tuple(_row(off[k]) - _col(C.sC[k]) for k in range(3)); dd = tuple(_row(dirT[k]) for k in range(3))
            b = dot3(oc, dd)
            c = dot3(oc, oc) - _col(C.sR * C.sR)
            disc = b * b - c
            sq = np.sqrt(np.maximum(disc, F(0)))
            tin, tout = -b - sq, -b + sq
            seg = np.where((disc >= 0) & ~(tout < 0), um_max(F(0), tout - um_max(tin, F(0))) * _col(C.sMat["density"]), F(0))
        contrib = [np.where(_col(C.sOwn != a), seg, F(0))]
        # PM:265-288 AABBs / PM:294-300 OBBs (stored rotation as is)
        for (okh, tin, tout), dens, own in ((perm_slab_aabb(C, off, dirT), C.aMat["density"], C.aOwn),
                                            (perm_slab_obb(C, off, dirT), C.oMat["density"], C.oOwn)):
            with np.errstate(all="ignore"):
                seg = np.where(okh, um_max(F(0), tout - um_max(tin, F(0))) * _col(dens), F(0))
            contrib.append(np.where(_col(own != a), seg, F(0)))
        allc = np.concatenate(contrib, axis=1)
        skip = np.concatenate([C.sOwn == a, C.aOwn == a, C.oOwn == a])
        for j in range(allc.shape[1]):                                               # sequential FP32 sum in collider order (PM:233-258)
            if not skip[j]:
                loss = loss + allc[:, j]
        vals[:, a] = NS - loss                                                       # PM:260
    return hit, vals


def perm_slab_aabb(C, o, d):
    with np.errstate(all="ignore"):
        oo = tuple(_row(o[k]) for k in range(3)); inv = tuple(F(1) / _row(d[k]) for k in range(3))
        t0 = tuple((_col(C.aC[k] - C.aH[k]) - oo[k]) * inv[k] for k in range(3))
        t1 = tuple((_col(C.aC[k] + C.aH[k]) - oo[k]) * inv[k] for k in range(3))
        tin = um_max(um_max(um_min(t0[0], t1[0]), um_min(t0[1], t1[1])), um_min(t0[2], t1[2]))
        tout = um_min(um_min(um_max(t0[0], t1[0]), um_max(t0[1], t1[1])), um_max(t0[2], t1[2]))
    return ~((tin > tout) | (tout < 0)), tin, tout


def perm_slab_obb(C, o, d):
    with np.errstate(all="ignore"):
        q = tuple(_col(C.oQ[k]) for k in range(4))
        rel = tuple(_row(o[k]) - _col(C.oC[k]) for k in range(3))
        lo_ = qmul(q, rel)
        ld = qmul(q, tuple(np.broadcast_to(_row(d[k]), rel[0].shape) for k in range(3)))
        inv = tuple(F(1) / ld[k] for k in range(3))
        t0 = tuple(((F(0) - _col(C.oH[k])) - lo_[k]) * inv[k] for k in range(3))
        t1 = tuple(((F(0) + _col(C.oH[k])) - lo_[k]) * inv[k] for k in range(3))
        tin = um_max(um_max(um_min(t0[0], t1[0]), um_min(t0[1], t1[1])), um_min(t0[2], t1[2]))
        tout = um_min(um_min(um_max(t0[0], t1[0]), um_max(t0[1], t1[1])), um_max(t0[2], t1[2]))
    return ~((tin > tout) | (tout < 0)), tin, tout
