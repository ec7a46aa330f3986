// k1_bounce.cu -- the bounce loop of AudioRaytracerJobBatched.Execute (Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs:90-215)
// WITHOUT its echo / muffle queries: RT:124-173 only write EchoRayDistances / MuffleRayHits and feed nothing back into
// cRayOrigin / cRayDirection / cRayLife (RT:179-192), so this kernel only appends one record per hit point and
// query_fan_kernel (k1_query_fan.cu) evaluates the queries of all hit points afterwards. Used whenever the frame has target
// fans; without them trace_grid_kernel (k1_trace_grid.cu) runs bounce rays and queries together.
//
// One thread owns one ray; ShootRayCast (RT:225-280) walks the uniform grid (3D-DDA, grid_dev.cuh) with the exact per-collider
// tests in the reference's operation order, nearest hit = lexicographic (t, sphere < AABB < OBB, index) minimum. Unlike
// trace_grid_kernel the lanes of a warp are NOT kept on the same bounce: a warp step is "every walking lane processes one
// grid cell", and a lane whose segment has ended (hit or left the scene) waits only until a handful of lanes have gathered at
// that point; those then write their hit points, reflect (RT:456-532), start their next segment -- or fetch a new ray -- and
// walk on. Segments cross 1 .. 30 cells, so marching in bounce lock step left 5 of 32 lanes busy (ncu, round 2); what a lane
// computes for its ray is unchanged, hence so are all outputs.
#include <cstdlib>

#include "device_util.cuh"
#include "grid_dev.cuh"
#include "intersect.cuh"
#include "launchers.h"
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {

#ifndef ART_BOUNCE_WARPS
#define ART_BOUNCE_WARPS 32
#endif
#ifndef ART_BOUNCE_GATHER
#define ART_BOUNCE_GATHER 8
#endif
constexpr int kBounceWarps = ART_BOUNCE_WARPS;
constexpr int kBounceThreads = kBounceWarps * 32;
constexpr int kBounceGather = ART_BOUNCE_GATHER;     // lanes at a segment boundary before the warp handles them
constexpr uint32_t kNoHitKey = 0xFFFFFFFFu;

// nearest-hit OBB distance: exact, or NaN when the collider misses or certainly lies beyond `best`
__device__ __forceinline__ float bounce_obb_dist(const GeomView& gv, int id, f3 o, f3 d, float dd, float errScale, float best)
{
    const float4 c4 = gv.obbC[id];
    const float2 h2 = gv.obbH[id];
    const f3 h = mk3(c4.w, h2.x, h2.y);
    const f3 pc = sub3(o, mk3(c4.x, c4.y, c4.z));                    // RT:316
    if (obb_sure_miss(pc, obb_cull_c(pc, h), d, dd)) return quiet_nan();
    const float4 q4 = gv.obbQ[id];
    if (!obb_maybe_nearer(q4, pc, h, d, errScale, best)) return quiet_nan();
    return obb_dist_exact(q4.x, q4.y, q4.z, q4.w, pc.x, pc.y, pc.z, h.x, h.y, h.z, d.x, d.y, d.z);
}

template <bool SMEM, bool STATS>
__global__ void __launch_bounds__(kBounceThreads, 1) bounce_kernel(const TraceArgs a, const GridDesc g)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31;
    const unsigned char* geomBase = a.geom;
    if (SMEM) {
        stage_blob_to_smem(smem, a.geom, a.L.bytes, &bar);
        geomBase = smem;
    }
    const GeomView gv = make_view(geomBase, a.L);
    const f3 RayOrigin = mk3(a.ox, a.oy, a.oz);
    const uint32_t ltMask = (1u << lane) - 1u;

    // ---- per-lane state: the ray, and the segment it is walking
    bool hasRay = false, walking = false, queueEmpty = false;
    int j = 0, hits = 0;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 0), inv = mk3(0, 0, 0);
    float life = 0.0f, dd = 0.0f, best = kFloatMax;
    uint32_t bkey = kNoHitKey;             // (typeOrder << 28) | index ; typeOrder sphere 0, AABB 1, OBB 2
    Dda w;
    w.ix = w.iy = w.iz = 0; w.tmx = w.tmy = w.tmz = 0; w.tdx = w.tdy = w.tdz = 0; w.tEnd = 0; w.tCur = 0; w.lastAxis = -1;
    unsigned int nSegments = 0, nSegHits = 0;
    unsigned int st[4] = { 0, 0, 0, 0 };   // STATS: sphere / AABB / OBB tests, cells visited (this lane)

    for (;;) {
        // ================= segment boundary: lanes whose segment ended, lanes without a ray =================
        const bool canOwn = lane < a.raysPerWarp;
        const uint32_t atEnd = __ballot_sync(kFull, hasRay && !walking);
        const uint32_t idle = __ballot_sync(kFull, !hasRay && canOwn && !queueEmpty);
        const uint32_t walk = __ballot_sync(kFull, walking);
        if (!walk && !atEnd && !idle) break;                                       // every ray of this warp is finished
        if (!walk || __popc(atEnd | idle) >= kBounceGather) {
            // ---- finish the segments that ended (RT:108-193)
            const bool hit = hasRay && !walking && bkey != kNoHitKey;
            int hitType = 0, hitIdx = 0;
            float4 attr = make_float4(0, 0, 0, 0);
            f3 Pp = mk3(0, 0, 0);
            float echoL = 0.0f;
            int resultId = 0;
            if (hasRay && !walking && !hit) {                                      // RT:201-207 the ray left the scene
                if (a.hitCounts) a.hitCounts[j] = (uint8_t)hits;
                hasRay = false;
            }
            if (hit) {
                nSegHits++;
                hitType = (int)(bkey >> 28);
                hitIdx = (int)(bkey & 0x0FFFFFFFu);
                o = add3(o, mul3s(d, best));                                       // RT:111
                life = subr(life, best);                                           // RT:112
                hits += 1;                                                         // RT:113
                const size_t rayResultId = (size_t)j * a.H + hits - 1;             // RT:115
                resultId = (int)rayResultId;
                ART_CHECK(a.counters, j < a.map.nLocal && hits <= a.H && hitIdx < (hitType == 0 ? a.L.ns : (hitType == 1 ? a.L.na : a.L.no)));
                attr = hitType == 0 ? a.at.sphAttr[hitIdx] : (hitType == 1 ? a.at.aabbAttr[hitIdx] : a.at.obbAttr[hitIdx]);
                if (a.hitPoints) {                                                 // RT:118, 197
                    a.hitPoints[3 * rayResultId] = um_f32tof16(o.x);
                    a.hitPoints[3 * rayResultId + 1] = um_f32tof16(o.y);
                    a.hitPoints[3 * rayResultId + 2] = um_f32tof16(o.z);
                }
                if (a.hitIds) {
                    const uint32_t refType = hitType == 0 ? 3u : (hitType == 1 ? 1u : 2u);   // Enums/ColliderType.cs
                    a.hitIds[rayResultId] = (refType << 30) | (uint32_t)hitIdx;
                }
                Pp = sub3(o, mul3s(d, kEpsilon));                                  // RT:124 == RT:158
                const f3 wv = sub3(o, RayOrigin);                                  // RT:130
                echoL = sqrtr(dot3(wv, wv));
            }
            // one record per hit point, appended in whatever order the warps get here (the queries' results -- echo halves
            // indexed by rayResultId, integer muffle counts -- do not depend on it)
            const uint32_t hitMask = __ballot_sync(kFull, hit);
            if (hitMask) {
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(a.recCount, (unsigned)__popc(hitMask));
                base = __shfl_sync(kFull, base, 0);
                if (hit) {
                    const unsigned int idx = base + (unsigned)__popc(hitMask & ltMask);
                    a.recA[idx] = make_float4(Pp.x, Pp.y, Pp.z, echoL);
                    a.recB[idx] = make_float2(attr.y, __int_as_float(resultId));
                }
            }
            // ---- termination / reflection (RT:178-193)
            if (hit) {
                bool alive = true;
                if (hits >= a.H || life <= 0.0f) {
                    alive = false;
                } else {
                    f3 normal = mk3(0.0f, 0.0f, 0.0f);
                    if (hitType == 1) {
                        const float4 C = a.at.aabbCtr[hitIdx], Hx = a.at.aabbHalf[hitIdx];
                        const f3 lp = sub3(o, mk3(C.x, C.y, C.z));                                    // RT:465
                        const float ex = subr(Hx.x, fabsf(lp.x)), ey = subr(Hx.y, fabsf(lp.y)), ez = subr(Hx.z, fabsf(lp.z));
                        if (ex < ey && ex < ez) normal.x = um_sign(lp.x);                             // RT:471-482
                        else if (ey < ex && ey < ez) normal.y = um_sign(lp.y);
                        else normal.z = um_sign(lp.z);
                    } else if (hitType == 2) {
                        const float4 qi = a.at.obbQinv[hitIdx], q4 = gv.obbQ[hitIdx], c4 = gv.obbC[hitIdx];
                        const float4 Hx = a.at.obbHalf[hitIdx];                                       // raw OBB Size
                        f4 qinv; qinv.x = qi.x; qinv.y = qi.y; qinv.z = qi.z; qinv.w = qi.w;
                        f4 q; q.x = q4.x; q.y = q4.y; q.z = q4.z; q.w = q4.w;
                        const f3 lh = qmul3(qinv, sub3(o, mk3(c4.x, c4.y, c4.z)));                    // RT:489 (quirk Q3)
                        const float ex = subr(Hx.x, fabsf(lh.x)), ey = subr(Hx.y, fabsf(lh.y)), ez = subr(Hx.z, fabsf(lh.z));
                        f3 ln = mk3(0.0f, 0.0f, 0.0f);
                        if (ex < ey && ex < ez) ln.x = um_sign(lh.x);                                 // RT:497-508
                        else if (ey < ex && ey < ez) ln.y = um_sign(lh.y);
                        else ln.z = um_sign(lh.z);
                        normal = qmul3(q, ln);                                                        // RT:510
                    } else {
                        const float4 s = gv.sph[hitIdx];
                        normal = normalize3(sub3(o, mk3(s.x, s.y, s.z)));                             // RT:516
                    }
                    d = reflect3(d, normal);                                                          // RT:525
                    o = add3(o, mul3s(d, kEpsilon));                                                  // RT:528
                    life = subr(life, mulr(a.maxRayLife, attr.x));                                    // RT:531
                    if (life < 0.0f) alive = false;                                                   // RT:189
                }
                if (!alive) {
                    if (a.hitCounts) a.hitCounts[j] = (uint8_t)hits;                                  // RT:204, 212
                    hasRay = false;
                }
            }
            // ---- lanes without a ray take the next ones from the queue
            bool fresh = false;
            const uint32_t dead = __ballot_sync(kFull, !hasRay && canOwn);
            if (dead && !queueEmpty) {
                int base = 0;
                if (lane == 0) base = (int)atomicAdd(a.nextRay, (unsigned)__popc(dead));
                base = __shfl_sync(kFull, base, 0);
                if ((dead >> lane) & 1u) {
                    const int jj = base + __popc(dead & ltMask);
                    if (jj < a.map.nLocal) {
                        j = jj;
                        const int rayIndex = a.map.to_global(j);
                        d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                                um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));            // RT:94
                        o = RayOrigin;                                                                // RT:95
                        hits = 0;                                                                     // RT:97
                        life = a.maxRayLife;                                                          // RT:99
                        hasRay = true;
                        fresh = true;
                    }
                }
                if (base + __popc(dead) >= a.map.nLocal) queueEmpty = true;
            }
            // ---- start the next segment of every lane that stands at a boundary with a live ray (ShootRayCast, RT:225-280)
            if (hasRay && (hit || fresh)) {
                nSegments++;
                dd = dot3(d, d);                                                   // RT:326
                inv = mk3(rcpr(d.x), rcpr(d.y), rcpr(d.z));                        // RT:289
                best = kFloatMax;
                bkey = kNoHitKey;
                walking = dda_init(g, o, d, inv, pos_inf(), w);                    // (false: the ray misses every collider)
            }
        }

        // ================= one grid cell for every walking lane =================
        if (walking) {
            const uint2 hdr = dda_cell(g, w);
            const uint16_t* e = g.entries + hdr.x;
            const int nS = hdr.y & 1023, nA = (hdr.y >> 10) & 2047, nO = hdr.y >> 21;
            if (STATS) { st[0] += nS; st[1] += nA; st[2] += nO; st[3]++; }
            ART_CHECK(a.counters, (unsigned)w.ix < (unsigned)g.nx && (unsigned)w.iy < (unsigned)g.ny && (unsigned)w.iz < (unsigned)g.nz);
            ART_CHECK(a.counters, hdr.x + nS + nA + nO <= (unsigned)g.nEntries);
            for (int k = 0; k < nS; k++) {
                const int id = __ldg(e + k);
                ART_CHECK(a.counters, id < a.L.ns);
                const float dist = sphere_dist(gv, id, o, d, dd);
                const uint32_t key = (uint32_t)id;
                if (dist < best || (dist == best && key < bkey)) { best = dist; bkey = key; }
            }
            e += nS;
            for (int k = 0; k < nA; k++) {
                const int id = __ldg(e + k);
                ART_CHECK(a.counters, id < a.L.na);
                const float dist = aabb_dist(gv, id, o, inv);
                const uint32_t key = (1u << 28) | (uint32_t)id;
                if (dist < best || (dist == best && key < bkey)) { best = dist; bkey = key; }
            }
            e += nA;
            for (int k = 0; k < nO; k++) {
                const int id = __ldg(e + k);
                ART_CHECK(a.counters, id < a.L.no);
                const float dist = bounce_obb_dist(gv, id, o, d, dd, g.errScale, best);
                const uint32_t key = (2u << 28) | (uint32_t)id;
                if (dist < best || (dist == best && key < bkey)) { best = dist; bkey = key; }
            }
            // colliders listed only in later cells lie beyond the next cell boundary
            const float tNext = dda_next_t(w);
            if (tNext > w.tEnd || tNext > best) walking = false;
            else walking = dda_step(g, d, w);
        }
    }

    // segment counters: warp sum -> one atomic per warp
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        nSegments += __shfl_xor_sync(kFull, nSegments, s);
        nSegHits += __shfl_xor_sync(kFull, nSegHits, s);
    }
    if (lane == 0) {
        atomicAdd(&a.counters[C_SEGMENTS], (unsigned long long)nSegments);
        atomicAdd(&a.counters[C_SEGMENT_HITS], (unsigned long long)nSegHits);
    }
    if (STATS) {
        atomicAdd(&a.counters[C_GRID_RT_S], (unsigned long long)st[0]);
        atomicAdd(&a.counters[C_GRID_RT_A], (unsigned long long)st[1]);
        atomicAdd(&a.counters[C_GRID_RT_O], (unsigned long long)st[2]);
        atomicAdd(&a.counters[C_GRID_RT_CELLS], (unsigned long long)st[3]);
    }
}

size_t bounce_smem_bytes(const GeomLayout& L, bool geomInSmem) { return geomInSmem ? L.bytes : 0; }

// Lanes of a warp that own rays: all 32 unless the batch is so small that fewer fill the GPU evenly (as trace_grid_plan).
cudaError_t launch_bounce(const TraceArgs& a0, const GridDesc& g, int numCtas, bool geomInSmem, bool stats, cudaStream_t stream)
{
    TraceArgs a = a0;
    if (!a.recA || !a.recB || !a.recCount) return cudaErrorInvalidValue;
    const long long warps = (long long)numCtas * kBounceWarps;
    const long long k = (a.map.nLocal + warps * 32 - 1) / (warps * 32);
    long long r = k > 0 ? (a.map.nLocal + warps * k - 1) / (warps * k) : 32;
    a.raysPerWarp = (int)(r < 1 ? 1 : (r > 32 ? 32 : r));
    const size_t smem = bounce_smem_bytes(a.L, geomInSmem);
    void (*kern)(const TraceArgs, const GridDesc) = nullptr;
    if (stats) kern = geomInSmem ? bounce_kernel<true, true> : bounce_kernel<false, true>;
    else kern = geomInSmem ? bounce_kernel<true, false> : bounce_kernel<false, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<numCtas, kBounceThreads, smem, stream>>>(a, g);
    return cudaGetLastError();
}

}  // namespace art
