// grid_dev.cuh -- device side of the uniform collider grid (scene_dev.cuh: GridDesc, grid_host.h): the
// per-lane 3D-DDA walk and the exact per-collider distance helpers shared by the grid kernels.
#pragma once
#include "intersect.cuh"
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {

// -DART_DEBUG_BOUNDS: every index the grid kernels form is checked and violations are counted (ArtCounters.debugViolations);
// the GPU tests are run once against such a build (compute-sanitizer is not available on the test pool).
#ifdef ART_DEBUG_BOUNDS
#define ART_CHECK(counters, cond) do { if (!(cond)) atomicAdd(&(counters)[C_DEBUG_VIOLATIONS], 1ull); } while (0)
#else
#define ART_CHECK(counters, cond) do { } while (0)
#endif

// ---- per-lane 3D-DDA ------------------------------------------------------------------------------
struct Dda {
    int ix, iy, iz;
    float tmx, tmy, tmz;      // parameter at which the ray leaves the current cell along each axis
    float tdx, tdy, tdz;      // parameter step per cell
    float tEnd;               // stop once the next cell starts beyond this
    float tCur;               // parameter at which the ray entered the current cell (>= 0)
    int lastAxis;             // axis of the last step (0,1,2), -1 in the first cell
};

// Clip the ray o + t*d, t in [0, tLimit], to the grid and set up the walk. inv = 1/d (may be +-Inf).
// Returns false when the ray misses the grid altogether (then it misses every collider).
__device__ __forceinline__ bool dda_init(const GridDesc& g, f3 o, f3 d, f3 inv, float tLimit, Dda& w)
{
    // fminf/fmaxf drop NaNs (0 * Inf when the origin sits on a bound with a zero direction component):
    // that axis then imposes no constraint, which is the conservative reading.
    const float ax = (g.g0x - o.x) * inv.x, bx = (g.g1x - o.x) * inv.x;
    const float ay = (g.g0y - o.y) * inv.y, by = (g.g1y - o.y) * inv.y;
    const float az = (g.g0z - o.z) * inv.z, bz = (g.g1z - o.z) * inv.z;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    const float tStart = fmaxf(tn, 0.0f);
    w.tEnd = fminf(tf, tLimit);
    w.tCur = tStart;
    w.lastAxis = -1;
    if (!(tStart <= w.tEnd)) return false;
    const float sx = fmaf(d.x, tStart, o.x), sy = fmaf(d.y, tStart, o.y), sz = fmaf(d.z, tStart, o.z);
    w.ix = min(max((int)floorf((sx - g.g0x) * g.icx), 0), g.nx - 1);
    w.iy = min(max((int)floorf((sy - g.g0y) * g.icy), 0), g.ny - 1);
    w.iz = min(max((int)floorf((sz - g.g0z) * g.icz), 0), g.nz - 1);
    const float inf = pos_inf();
    if (d.x > 0.0f) { w.tmx = (g.g0x + (float)(w.ix + 1) * g.csx - o.x) * inv.x; w.tdx = g.csx * inv.x; }
    else if (d.x < 0.0f) { w.tmx = (g.g0x + (float)w.ix * g.csx - o.x) * inv.x; w.tdx = -g.csx * inv.x; }
    else { w.tmx = inf; w.tdx = 0.0f; }
    if (d.y > 0.0f) { w.tmy = (g.g0y + (float)(w.iy + 1) * g.csy - o.y) * inv.y; w.tdy = g.csy * inv.y; }
    else if (d.y < 0.0f) { w.tmy = (g.g0y + (float)w.iy * g.csy - o.y) * inv.y; w.tdy = -g.csy * inv.y; }
    else { w.tmy = inf; w.tdy = 0.0f; }
    if (d.z > 0.0f) { w.tmz = (g.g0z + (float)(w.iz + 1) * g.csz - o.z) * inv.z; w.tdz = g.csz * inv.z; }
    else if (d.z < 0.0f) { w.tmz = (g.g0z + (float)w.iz * g.csz - o.z) * inv.z; w.tdz = -g.csz * inv.z; }
    else { w.tmz = inf; w.tdz = 0.0f; }
    return true;
}
// Parameter at which the ray enters the next cell.
__device__ __forceinline__ float dda_next_t(const Dda& w) { return fminf(fminf(w.tmx, w.tmy), w.tmz); }
// Advance one cell; false when the walk leaves the grid.
__device__ __forceinline__ bool dda_step(const GridDesc& g, f3 d, Dda& w)
{
    if (w.tmx <= w.tmy && w.tmx <= w.tmz) {
        w.tCur = w.tmx; w.lastAxis = 0;
        w.ix += d.x > 0.0f ? 1 : -1; w.tmx += w.tdx;
        return (unsigned)w.ix < (unsigned)g.nx;
    }
    if (w.tmy <= w.tmz) {
        w.tCur = w.tmy; w.lastAxis = 1;
        w.iy += d.y > 0.0f ? 1 : -1; w.tmy += w.tdy;
        return (unsigned)w.iy < (unsigned)g.ny;
    }
    w.tCur = w.tmz; w.lastAxis = 2;
    w.iz += d.z > 0.0f ? 1 : -1; w.tmz += w.tdz;
    return (unsigned)w.iz < (unsigned)g.nz;
}
__device__ __forceinline__ uint2 dda_cell(const GridDesc& g, const Dda& w)
{
    return __ldg(&g.cells[(w.iz * g.ny + w.iy) * g.nx + w.ix]);   // < 2^23 cells: 32-bit index arithmetic
}

// ---- exact per-collider distances (NaN = miss), reference operation order ---------------------------
__device__ __forceinline__ float sphere_dist(const GeomView& gv, int id, f3 o, f3 d, float dd)
{
    const float4 s = gv.sph[id];
    const f3 oc = sub3(o, mk3(s.x, s.y, s.z));                       // RT:325
    const float cc = subr(dot3(oc, oc), s.w);                        // RT:328
    if (sphere_fast_miss(oc, cc, d, dd)) return quiet_nan();         // disc < 0 (RT:331)
    return sphere_dist_exact(oc.x, oc.y, oc.z, cc, d.x, d.y, d.z, dd);
}
__device__ __forceinline__ float aabb_dist(const GeomView& gv, int id, f3 o, f3 inv)
{
    const float4 A = gv.aabbA[id];
    const float2 B = gv.aabbB[id];
    float tNear, tFar, dist;
    slab<8>(subr(A.x, o.x), subr(A.y, o.y), subr(A.z, o.z), subr(A.w, o.x), subr(B.x, o.y), subr(B.y, o.z),
            inv.x, inv.y, inv.z, tNear, tFar);                       // RT:291-298
    return slab_hit(tNear, tFar, dist) ? dist : quiet_nan();         // RT:300-307
}
// any-hit form of aabb_dist: does the slab test report a distance < limit (RT:381-384 / RT:423-430)?
__device__ __forceinline__ bool aabb_blocks(const GeomView& gv, int id, f3 o, f3 inv, float limit)
{
    const float4 A = gv.aabbA[id];
    const float2 B = gv.aabbB[id];
    float tNear, tFar;
    slab<8>(subr(A.x, o.x), subr(A.y, o.y), subr(A.z, o.z), subr(A.w, o.x), subr(B.x, o.y), subr(B.y, o.z),
            inv.x, inv.y, inv.z, tNear, tFar);                       // RT:291-298
    const float dist = tNear > 0.0f ? tNear : tFar;                  // RT:306
    return !(tNear > tFar || tFar < 0.0f) && dist < limit;           // RT:300-304 + caller's compare
}
// nearest-hit OBB distance with an explicit rotation q4 (RT:314-320 passes the stored rotation, PM:172-179 its
// inverse): exact distance, or NaN when the collider misses or certainly lies beyond `best`
__device__ __forceinline__ float obb_dist_nearest_q(const GeomView& gv, float4 q4, int id, f3 o, f3 d, float dd, float errScale, float best)
{
    const float4 c4 = gv.obbC[id];
    const float2 h2 = gv.obbH[id];
    const f3 h = mk3(c4.w, h2.x, h2.y);
    const f3 pc = sub3(o, mk3(c4.x, c4.y, c4.z));                    // RT:316
    if (obb_sure_miss(pc, obb_cull_c(pc, h), d, dd)) return quiet_nan();
    if (!obb_maybe_nearer(q4, pc, h, d, errScale, best)) return quiet_nan();
    return obb_dist_exact(q4.x, q4.y, q4.z, q4.w, pc.x, pc.y, pc.z, h.x, h.y, h.z, d.x, d.y, d.z);
}

// any-hit: does the exact test report a distance < limit (RT:390 / RT:441)?
__device__ __forceinline__ bool obb_blocks(const GeomView& gv, int id, f3 o, f3 d, float dd, float errScale, float limit)
{
    const float4 c4 = gv.obbC[id];
    const float2 h2 = gv.obbH[id];
    const f3 h = mk3(c4.w, h2.x, h2.y);
    const f3 pc = sub3(o, mk3(c4.x, c4.y, c4.z));
    if (obb_sure_miss(pc, obb_cull_c(pc, h), d, dd)) return false;
    const float4 q4 = gv.obbQ[id];
    const int cls = obb_classify(q4, pc, h, d, errScale, limit);
    if (cls != 2) return cls == 1;
    return obb_dist_exact(q4.x, q4.y, q4.z, q4.w, pc.x, pc.y, pc.z, h.x, h.y, h.z, d.x, d.y, d.z) < limit;
}

}  // namespace art
