// intersect.cuh -- ray / collider primitive tests shared by K1 (trace) and K2 (permeation).
// RT = Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs, PM = .../AudioPermeationJobBatched.cs.
//
// Every test is split into a FAST part that runs for all lanes (a handful of un-fused FP32 ops
// that decide "certainly a miss" with exactly the reference's own comparison), and an EXACT part
// (IEEE sqrt / divisions / quaternion rotations in the reference's operation order) kept out of
// line (__noinline__) so that the hot loops stay small enough for the instruction caches; the
// exact part only runs for the few lanes whose ray actually comes near the collider.
#pragma once
#include "um_math.cuh"

namespace art {

__device__ __forceinline__ float quiet_nan() { return __int_as_float(0x7FC00000); }
__device__ __forceinline__ float pos_inf() { return __int_as_float(0x7F800000); }

// ---- slab kernels -------------------------------------------------------------------------------
// CLS 0..7: bit k set <=> invDir component k is negative (near plane = max). CLS 8: generic form
// with min/max, used whenever a component of invDir is +-Inf (0*Inf NaN rule).
template <int CLS>
__device__ __forceinline__ void slab(float lox, float loy, float loz, float hix, float hiy, float hiz,
                                     float ix, float iy, float iz, float& tNear, float& tFar)
{
    if (CLS == 8) {
        float t0x = mulr(lox, ix), t0y = mulr(loy, iy), t0z = mulr(loz, iz);
        float t1x = mulr(hix, ix), t1y = mulr(hiy, iy), t1z = mulr(hiz, iz);
        tNear = max3f(um_min(t0x, t1x), um_min(t0y, t1y), um_min(t0z, t1z));   // RT:294, 297
        tFar = min3f(um_max(t0x, t1x), um_max(t0y, t1y), um_max(t0z, t1z));    // RT:295, 298
    } else {
        float nx = (CLS & 1) ? hix : lox, fx = (CLS & 1) ? lox : hix;
        float ny = (CLS & 2) ? hiy : loy, fy = (CLS & 2) ? loy : hiy;
        float nz = (CLS & 4) ? hiz : loz, fz = (CLS & 4) ? loz : hiz;
        tNear = max3f(mulr(nx, ix), mulr(ny, iy), mulr(nz, iz));
        tFar = min3f(mulr(fx, ix), mulr(fy, iy), mulr(fz, iz));
    }
}

__device__ __forceinline__ int slab_class(float ix, float iy, float iz)
{
    if (isinf(ix) || isinf(iy) || isinf(iz)) return 8;
    return (int)((__float_as_uint(ix) >> 31) | ((__float_as_uint(iy) >> 31) << 1) | ((__float_as_uint(iz) >> 31) << 2));
}

// RT:300-307: miss if tNear > tFar || tFar < 0; distance = tNear > 0 ? tNear : tFar
__device__ __forceinline__ bool slab_hit(float tNear, float tFar, float& dist)
{
    dist = tNear > 0.0f ? tNear : tFar;
    return !(tNear > tFar || tFar < 0.0f);
}

// ---- sphere (RT:323-355 == PM:182-214) ----------------------------------------------------------
// Reference: b = 2*dot(oc,d); disc = b*b - 4*a*c; miss if disc < 0.
// Scaling by 2 and 4 is exact, so  disc < 0  <=>  dot*dot < a*c  (both products rounded once, exactly
// as in the reference up to the exact power-of-two scale): that comparison is the fast reject.
__device__ __forceinline__ bool sphere_fast_miss(f3 oc, float cc, f3 d, float a)
{
    const float dt = dot3(oc, d);
    return mulr(dt, dt) < mulr(a, cc);
}
// Conservative variant for the occlusion loops (the exact path re-evaluates survivors, so this only has
// to be SAFE): the dot product is formed with FMAs (1 FMUL + 2 FFMA instead of 3 FMUL + 2 FADD); its
// difference from the reference's rounding is at most 3*2^-23*|oc||d|, hence dot^2 moves by less than
// 2^-20*|oc|^2*a; ccm = cc - 2^-19*|oc|^2 (sphere_cull_c) absorbs that with a 2x margin.
__device__ __forceinline__ float sphere_cull_c(float cc, float rr) { return fmaf(-1.9073486328125e-06f, cc + rr, cc); }
__device__ __forceinline__ bool sphere_sure_miss(f3 oc, float ccm, f3 d, float a)
{
    const float dt = fmaf(oc.z, d.z, fmaf(oc.y, d.y, oc.x * d.x));
    return dt * dt < a * ccm;
}
// Exact distance in the reference's operation order, or NaN for a miss (NaN fails every `dist < x`).
static __device__ __noinline__ float sphere_dist_exact(float ocx, float ocy, float ocz, float cc, float dx, float dy, float dz, float a)
{
    const float b = mulr(2.0f, dot3(mk3(ocx, ocy, ocz), mk3(dx, dy, dz)));     // RT:327
    const float disc = subr(mulr(b, b), mulr(mulr(4.0f, a), cc));              // RT:329
    if (disc < 0.0f) return quiet_nan();
    const float sq = sqrtr(disc);
    const float twoA = mulr(2.0f, a);
    const float t0 = divr(subr(-b, sq), twoA);                                 // RT:338
    if (t0 >= 0.0f) return t0;
    const float t1 = divr(addr(-b, sq), twoA);                                 // RT:339
    if (t1 >= 0.0f) return t1;
    return quiet_nan();
}

// ---- OBB (RT:314-320) ----------------------------------------------------------------------------
// Conservative rejection by the bounding sphere: true only if the exact test is certain to report a
// miss. pc = o - C, cB = |pc|^2 - Rb^2 (Rb inflated), dd = dot(d,d). The ray's line misses the sphere
// when (pc.d)^2 - dd*cB < 0; the tolerance term covers the FP32 cancellation error of that
// difference; a sphere behind an outside origin is missed as well. FMAs are fine here: the outcome
// only decides whether the exact test runs, and it is skipped only when it must fail.
__device__ __forceinline__ bool obb_sure_miss(f3 pc, float cB, f3 d, float dd)
{
    float bq = fmaf(pc.z, d.z, fmaf(pc.y, d.y, pc.x * d.x));
    float ac = dd * cB;
    float b2 = bq * bq;
    float disc = b2 - ac;
    return (disc < -1e-4f * (b2 + fabsf(ac))) || (bq > 0.0f && cB > 0.0f);
}
__device__ __forceinline__ float obb_cull_c(f3 pc, f3 h)
{
    float rb2 = fmaf(h.z, h.z, fmaf(h.y, h.y, h.x * h.x));
    float pp = fmaf(pc.z, pc.z, fmaf(pc.y, pc.y, pc.x * pc.x));
    return pp - (rb2 * 1.05f + 1e-3f) - 1e-4f * pp;
}
// Exact OBB distance (or NaN): local = q*(o-C) is computed here so that callers keep only pc live.
static __device__ __noinline__ float obb_dist_exact(float qx, float qy, float qz, float qw, float pcx, float pcy, float pcz,
                                             float hx, float hy, float hz, float dx, float dy, float dz)
{
    f4 q; q.x = qx; q.y = qy; q.z = qz; q.w = qw;
    const f3 lo = qmul3(q, mk3(pcx, pcy, pcz));                                // RT:316
    const f3 ld = qmul3(q, mk3(dx, dy, dz));                                   // RT:317
    const float ix = rcpr(ld.x), iy = rcpr(ld.y), iz = rcpr(ld.z);             // RT:289
    float tNear, tFar, dist;
    slab<8>(subr(-hx, lo.x), subr(-hy, lo.y), subr(-hz, lo.z), subr(hx, lo.x), subr(hy, lo.y), subr(hz, lo.z),
            ix, iy, iz, tNear, tFar);
    return slab_hit(tNear, tFar, dist) ? dist : quiet_nan();
}

// Conservative OBB pre-test for the grid kernels: a full slab test in cheap arithmetic (FMA rotations,
// approximate reciprocals) against the box inflated by eta. Returns false only when the exact test
// (obb_dist_exact / obb_loss_exact with this q) certainly reports a miss.
// Why it is safe: the cheap and the exact evaluation of local = q*(o-C), ld = q*d differ by less than
// 4e-6*|pc| and 4e-6*|d| (a dozen roundings of magnitude <= |v|), and the exact FP32 slab comparison can
// accept a graze that misses by at most ~6e-7 of the distance travelled. Any parameter t* >= 0 at which
// the exact evaluation is inside the box is therefore inside the inflated box (eta = 2e-5*(|pc|_1 +
// errScale), errScale >= distance travelled, > 4x the bound above) for the cheap ray as well, with room
// that dwarfs the rounding of the cheap slab itself. NaNs (0*Inf) drop out of fminf/fmaxf, i.e. an axis
// that cannot be evaluated imposes no constraint.
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_fast(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ f3 qrot_fast(float4 q, f3 v)
{
    const float cx = fmaf(q.y, v.z, -(q.z * v.y)), cy = fmaf(q.z, v.x, -(q.x * v.z)), cz = fmaf(q.x, v.y, -(q.y * v.x));
    const float ex = fmaf(q.y, cz, -(q.z * cy)), ey = fmaf(q.z, cx, -(q.x * cz)), ez = fmaf(q.x, cy, -(q.y * cx));
    return mk3(fmaf(2.0f, fmaf(q.w, cx, ex), v.x), fmaf(2.0f, fmaf(q.w, cy, ey), v.y), fmaf(2.0f, fmaf(q.w, cz, ez), v.z));
}
// Cheap slab intervals of the ray against the box inflated by eta (tnI, tfI) and, if DEFL, deflated by eta
// (tnD, tfD; tnD > tfD when the deflated box is empty).
template <bool DEFL>
__device__ __forceinline__ void obb_pretest_local(f3 lo, f3 ld, f3 h, float eta, float& tnI, float& tfI, float& tnD, float& tfD)
{
    const float rx = rcp_fast(ld.x), ry = rcp_fast(ld.y), rz = rcp_fast(ld.z);
    {
        const float hx = h.x + eta, hy = h.y + eta, hz = h.z + eta;
        const float ax = (-hx - lo.x) * rx, bx = (hx - lo.x) * rx;
        const float ay = (-hy - lo.y) * ry, by = (hy - lo.y) * ry;
        const float az = (-hz - lo.z) * rz, bz = (hz - lo.z) * rz;
        tnI = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        tfI = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    }
    if (DEFL) {
        const float hx = h.x - eta, hy = h.y - eta, hz = h.z - eta;
        const float ax = (-hx - lo.x) * rx, bx = (hx - lo.x) * rx;
        const float ay = (-hy - lo.y) * ry, by = (hy - lo.y) * ry;
        const float az = (-hz - lo.z) * rz, bz = (hz - lo.z) * rz;
        tnD = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
        tfD = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
        if (!(hx > 0.0f && hy > 0.0f && hz > 0.0f)) { tnD = pos_inf(); tfD = -pos_inf(); }
    }
}
template <bool DEFL>
__device__ __forceinline__ void obb_pretest(float4 q, f3 pc, f3 h, f3 d, float errScale, float& tnI, float& tfI, float& tnD, float& tfD)
{
    const float eta = 2e-5f * (fabsf(pc.x) + fabsf(pc.y) + fabsf(pc.z) + errScale);
    obb_pretest_local<DEFL>(qrot_fast(q, pc), qrot_fast(q, d), h, eta, tnI, tfI, tnD, tfD);
}
// Nearest-hit use: false only if the exact test certainly misses or its distance certainly exceeds `best`
// (the exact distance is >= the entry into the inflated box up to rounding far below eta).
__device__ __forceinline__ bool obb_maybe_nearer(float4 q, f3 pc, f3 h, f3 d, float errScale, float best)
{
    float tnI, tfI, tnD, tfD;
    obb_pretest<false>(q, pc, h, d, errScale, tnI, tfI, tnD, tfD);
    return !(tnI > tfI) && !(tfI < 0.0f) && !(tnI > best);
}
// Any-hit use against a distance limit: 0 = the exact test certainly misses, 1 = it certainly reports a hit
// with distance < limit, 2 = undecided (run the exact test).
// "Certainly hits": the origin is outside the inflated box (tnI > 0), so the exact evaluation's origin is outside
// the real box and its distance is its entry parameter tNear; the cheap ray is inside the DEFLATED box at tnD, so
// the exact evaluation is inside the real box there with ~eta to spare (the FP32 slab comparison cannot miss
// that), hence tNear <= tnD * (1 + 2e-6) < limit when tnD < limit * (1 - 1e-4) ...
__device__ __forceinline__ int obb_classify(float4 q, f3 pc, f3 h, f3 d, float errScale, float limit)
{
    float tnI, tfI, tnD, tfD;
    obb_pretest<true>(q, pc, h, d, errScale, tnI, tfI, tnD, tfD);
    if ((tnI > tfI) || (tfI < 0.0f)) return 0;
    // ... and whatever the origin's position (a hit point ON this OBB lies inside its inflated box): if the forward ray
    // passes through the deflated box at all (tnD <= tfD, tfD >= 0), the exact evaluation is inside the real box at
    // max(tnD, 0) with ~eta to spare, so it reports a hit, with a distance <= its exit parameter, which is inside the
    // inflated box for the cheap ray: <= tfI * (1 + 2e-6) < limit when tfI < limit * (1 - 1e-4).
    const float lim = limit * 0.9999f;
    const bool through = tnD <= tfD && tfD >= 0.0f;
    return through && ((tnI > 0.0f && tnD < lim) || tfI < lim) ? 1 : 2;
    // (Measured and dropped: a "certainly misses" rule for origins that lie outside the box along one local axis and move
    // further out -- a hit point on this OBB whose goal is on its own side of the face -- is bit-exact but its six compares per
    // undecided call cost more than the exact tests it saves: C3 queries 10.75 -> 11.17 ms.)
}

// ---- permeation variants (PM:265-328) -------------------------------------------------------------
__device__ __forceinline__ float slab_loss(float tEnter, float tExit, float dens)
{
    if (tEnter > tExit || tExit < 0.0f) return 0.0f;                            // PM:281
    const float enter = um_max(tEnter, 0.0f);                                   // PM:286
    return mulr(um_max(0.0f, subr(tExit, enter)), dens);                        // PM:287
}
// PM:303-328 (assumes a unit direction): b = dot(oc,d); disc = b*b - c; miss if disc < 0
__device__ __forceinline__ bool sphere_loss_fast_miss(f3 oc, float cc, f3 d, float& b)
{
    b = dot3(oc, d);
    return mulr(b, b) < cc;      // <=> b*b - c < 0 (the rounded difference keeps the sign of the exact one)
}
static __device__ __noinline__ float sphere_loss_exact(float b, float cc, float dens)
{
    const float disc = subr(mulr(b, b), cc);
    if (disc < 0.0f) return 0.0f;
    const float sqrtD = sqrtr(disc);
    const float tEnter = subr(-b, sqrtD);
    const float tExit = addr(-b, sqrtD);
    if (tExit < 0.0f) return 0.0f;
    const float enter = um_max(tEnter, 0.0f);
    return mulr(um_max(0.0f, subr(tExit, enter)), dens);
}
// PM:294-300
static __device__ __noinline__ float obb_loss_exact(float qx, float qy, float qz, float qw, float pcx, float pcy, float pcz,
                                             float hx, float hy, float hz, float dx, float dy, float dz, float dens)
{
    f4 q; q.x = qx; q.y = qy; q.z = qz; q.w = qw;
    const f3 lo = qmul3(q, mk3(pcx, pcy, pcz));
    const f3 ld = qmul3(q, mk3(dx, dy, dz));
    const float ix = rcpr(ld.x), iy = rcpr(ld.y), iz = rcpr(ld.z);
    float tEnter, tExit;
    slab<8>(subr(-hx, lo.x), subr(-hy, lo.y), subr(-hz, lo.z), subr(hx, lo.x), subr(hy, lo.y), subr(hz, lo.z),
            ix, iy, iz, tEnter, tExit);
    return slab_loss(tEnter, tExit, dens);
}
template <int CLS>
__device__ __forceinline__ float aabb_loss(float4 A, float2 B, f3 P, float ix, float iy, float iz, float dens)
{
    float tEnter, tExit;
    slab<CLS>(subr(A.x, P.x), subr(A.y, P.y), subr(A.z, P.z), subr(A.w, P.x), subr(B.x, P.y), subr(B.y, P.z),
              ix, iy, iz, tEnter, tExit);
    return slab_loss(tEnter, tExit, dens);
}

}  // namespace art
