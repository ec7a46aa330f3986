"""The covering-depth cull (csrc/k4_fan_build.cu "covering depth", csrc/k1_query_fan.cu "cull") drops an echo / muffle query
without evaluating anything when its hit point lies beyond the covering depth of its direction sub-bin. That is only legal if
the reference's OWN FP32 evaluation (RT:284-308 + the caller's `dist < limit`) reports a blocker for every such query. The
GPU tests check that on whole frames; this CPU test attacks the margins directly: boxes that cover a sub-bin as barely as the
build's rules allow (thin, near the goal, tangent extents tight to the sub-bin), hit points just beyond the threshold, at the
sub-bin's corners, far away, and with direction components that are exactly zero -- each evaluated with the float32
restatement of the reference's slab test (oracle/independent.py)."""
import numpy as np
import pytest

from oracle import independent as ind

F = np.float32
B, SUB = 32, 2
FINE = B * SUB
TAN = 1e-3            # kFanCoverTan


def cover_interval(lo, hi, G, k, neg, fa, fb, near_dist, min_thick):
    """FP64 restatement of fan_cover_params / fan_cover_clip for sub-bin (fa, fb) of face (k, neg): (threshold, w2) or None"""
    lo, hi, G = lo.astype(np.float64), hi.astype(np.float64), G.astype(np.float64)
    e = 4e-6 * (np.abs(G) + np.abs(lo) + np.abs(hi)) + 1e-6
    rl, rh = lo - G + e, hi - G - e
    i, j = (k + 1) % 3, (k + 2) % 3
    zl, zh = (-rh[k], -rl[k]) if neg else (rl[k], rh[k])
    if not (zl >= near_dist and zl <= zh and rl[i] <= rh[i] and rl[j] <= rh[j]):
        return None
    a0, a1 = -1 + fa * 2 / FINE - TAN, -1 + (fa + 1) * 2 / FINE + TAN
    b0, b1 = -1 + fb * 2 / FINE - TAN, -1 + (fb + 1) * 2 / FINE + TAN
    w1, w2 = zl, zh
    for coef, bound in ((a1, rh[i]), (-a0, -rl[i]), (b1, rh[j]), (-b0, -rl[j])):
        if coef > 0:
            w2 = min(w2, bound / coef)
        elif coef < 0:
            w1 = max(w1, bound / coef)
        elif bound < 0:
            return None
    if w2 >= w1 * 1.00390625 and w2 - w1 >= min_thick:
        return w1 * 1.001953125, w2
    return None


def fine_bin(v):
    """fan_bin_w in FP32 (face, fa, fb, depth) of direction v (goal -> hit point)"""
    a = np.abs(v)
    k = 0 if (a[0] >= a[1] and a[0] >= a[2]) else (1 if a[1] >= a[2] else 2)
    w = a[k]
    r = F(1) / w
    p, q = v[(k + 1) % 3] * r, v[(k + 2) % 3] * r
    fa = int(min(FINE - 1, max(0, np.floor(F(p * F(FINE / 2) + F(FINE / 2))))))
    fb = int(min(FINE - 1, max(0, np.floor(F(q * F(FINE / 2) + F(FINE / 2))))))
    return 2 * k + (1 if v[k] < 0 else 0), fa, fb, w


def reference_blocks(P, G, lo, hi, limit):
    """RT:127/162 + RT:284-308 + the caller's compare, all in float32"""
    v = tuple(F(G[c] - P[c]) for c in range(3))
    d = ind.normalize3(v)
    o = tuple(np.array([P[c]], F) for c in range(3))
    dd = tuple(np.array([d[c]], F) for c in range(3))
    hit, dist = ind.slab(o, dd, tuple(np.array([lo[c]], F) for c in range(3)), tuple(np.array([hi[c]], F) for c in range(3)))
    return bool(hit[0] and dist[0] < limit), float(dist[0])


@pytest.mark.parametrize("seed", range(4))
def test_reference_reports_the_covering_box_for_every_culled_query(seed):
    rng = np.random.default_rng(900 + seed)
    checked = 0
    for trial in range(1000):
        D = float(rng.choice([40.0, 300.0, 3000.0]))                  # errScale
        near_dist, min_thick = 1e-3 * D, 1e-4 * D
        G = (rng.uniform(-1, 1, 3) * D * rng.choice([0.0, 0.05, 0.25])).astype(F)
        k, neg = int(rng.integers(0, 3)), bool(rng.integers(0, 2))
        i, j = (k + 1) % 3, (k + 2) % 3
        fa, fb = (int(x) for x in rng.choice([0, 1, 15, 31, 32, 33, 47, 62, 63], 2))
        # a box that holds the (widened) sub-bin between the depths w1t and w2t, as tightly as the slack allows
        w1t = float(np.exp(rng.uniform(np.log(near_dist * 1.01), np.log(D / 3))))
        w2t = w1t * (1 + float(rng.choice([0.0045, 0.01, 0.1, 1.0, 6.0]))) + 1.3 * min_thick
        slack = float(rng.choice([3e-5, 1e-3, 0.05])) * (1 + np.abs(G).max() + w2t)
        a0, a1 = -1 + fa * 2 / FINE - TAN, -1 + (fa + 1) * 2 / FINE + TAN
        b0, b1 = -1 + fb * 2 / FINE - TAN, -1 + (fb + 1) * 2 / FINE + TAN
        lo64, hi64 = np.zeros(3), np.zeros(3)
        lo64[k], hi64[k] = (-w2t - slack, -w1t + slack) if neg else (w1t - slack, w2t + slack)
        lo64[i], hi64[i] = min(a0 * w1t, a0 * w2t) - slack, max(a1 * w1t, a1 * w2t) + slack
        lo64[j], hi64[j] = min(b0 * w1t, b0 * w2t) - slack, max(b1 * w1t, b1 * w2t) + slack
        lo, hi = (lo64 + G).astype(F), (hi64 + G).astype(F)
        cov = cover_interval(lo, hi, G, k, neg, fa, fb, near_dist, min_thick)
        if cov is None:
            continue
        thr, w2 = cov
        ea0, ea1 = -1 + fa * 2 / FINE, -1 + (fa + 1) * 2 / FINE
        eb0, eb1 = -1 + fb * 2 / FINE, -1 + (fb + 1) * 2 / FINE
        dirs = [(a, b) for a in (ea0, (ea0 + ea1) / 2, ea1) for b in (eb0, (eb0 + eb1) / 2, eb1)]
        if ea0 <= 0 <= ea1:
            dirs += [(0.0, eb0), (0.0, (eb0 + eb1) / 2)]                 # direction components that are exactly zero
        if eb0 <= 0 <= eb1:
            dirs += [(ea1, 0.0)]
        dirs += list(zip(rng.uniform(ea0, ea1, 4), rng.uniform(eb0, eb1, 4)))
        depths = [thr * (1 - 1e-6), thr * 1.0001, w2, float(rng.uniform(thr, min(D, 40 * thr))), min(D, thr * 900)]
        for a, b in dirs:
            for wP in depths:
                v = np.zeros(3)
                v[k], v[i], v[j] = (-wP if neg else wP), a * wP, b * wP
                P = (G.astype(np.float64) + v).astype(F)
                if a == 0.0:
                    P[i] = G[i]
                if b == 0.0:
                    P[j] = G[j]
                face, qa, qb, w = fine_bin((P - G).astype(F))
                if (face, qa, qb) != (2 * k + int(neg), fa, fb) or not (w > thr * (1 - 2e-6)) or w > D:
                    continue                                             # (rounding moved the point into another sub-bin: not this cull)
                length = float(np.sqrt(np.sum((G.astype(np.float64) - P.astype(np.float64)) ** 2)))
                for limit in (F(length), F(length - 1.01e-4)):           # muffle: |goal - P|; echo: distance(RayOrigin, hit) >= |goal - P| - epsilon
                    blocked, dist = reference_blocks(P, G, lo, hi, limit)
                    assert blocked, (seed, trial, D, k, neg, fa, fb, a, b, wP, thr, w2, dist, float(limit))
                checked += 1
    assert checked > 2000, checked
