// intersect.cuh -- ray / collider primitive tests shared by K1 (trace) and K2 (permeation).
// RT = Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs, PM = .../AudioPermeationJobBatched.cs.
#pragma once
#include "um_math.cuh"

namespace art {

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

// RT:323-355 with a = dot(d,d) hoisted; cc = dot(oc,oc) - R*R
__device__ __forceinline__ bool sphere_hit(f3 oc, float cc, f3 d, float fourA, float twoA, float& dist)
{
    float b = mulr(2.0f, dot3(oc, d));
    float disc = subr(mulr(b, b), mulr(fourA, cc));
    dist = 0.0f;
    if (disc < 0.0f) return false;
    float sq = sqrtr(disc);
    float t0 = divr(subr(-b, sq), twoA);
    if (t0 >= 0.0f) { dist = t0; return true; }
    float t1 = divr(addr(-b, sq), twoA);
    if (t1 >= 0.0f) { dist = t1; return true; }
    return false;
}

// RT:314-320 with the per-origin part (lo = q*(o-C)) supplied by the caller.
__device__ __forceinline__ bool obb_hit(f4 q, f3 lo, f3 h, f3 d, float& dist)
{
    f3 ld = qmul3(q, d);
    float ix = rcpr(ld.x), iy = rcpr(ld.y), iz = rcpr(ld.z);
    float tNear, tFar;
    slab<8>(subr(-h.x, lo.x), subr(-h.y, lo.y), subr(-h.z, lo.z), subr(h.x, lo.x), subr(h.y, lo.y), subr(h.z, lo.z),
            ix, iy, iz, tNear, tFar);
    return slab_hit(tNear, tFar, dist);
}

// Conservative rejection of an OBB by its bounding sphere: true only if the exact test is
// certain to report a miss. pc = o - C, cB = |pc|^2 - Rb^2 (Rb inflated), dd = dot(d,d).
// The ray's line misses the sphere when (pc.d)^2 - dd*cB < 0; the tolerance term covers the FP32
// cancellation error of that difference; a sphere behind an outside origin is missed as well.
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

}  // namespace art
