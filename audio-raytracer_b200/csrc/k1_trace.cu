// k1_trace.cu -- K1: the per-ray bounce loop of AudioRaytracerJobBatched.Execute
// (Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs:61-215) for sm_100a.
//
// Mapping: ONE WARP OWNS ONE RAY; the 32 lanes own 32 different colliders
// per step. All three query kinds are linear scans over the collider list, so a warp sweeps the
// list in "super-chunks" (32 lanes x R colliders held in registers):
//   * nearest hit (ShootRayCast, RT:225-280): every lane keeps its own (t, index) minimum, then one
//     lexicographic warp reduction (shuffle-min on t, redux-min on the canonical index) reproduces
//     the reference's "strict <, first in sphere->AABB->OBB order wins" rule exactly;
//   * echo return ray (CanRaySeePoint, RT:365-397) and the Na muffle rays (CanRaySeeAudioTarget,
//     RT:405-449) all start at the same offset hit point, so they are evaluated together: the
//     per-collider, origin-dependent part of each test (min-o, max-o, |oc|^2-R^2, q*(o-C)) is
//     computed once per super-chunk and reused for every still-unblocked query; a query leaves the
//     loop at the first super-chunk in which any lane reports a blocker (ballot), which keeps the
//     executed tests within one super-chunk of the reference's own early exit.
// The collider planes are staged into shared memory once per CTA with cp.async.bulk (TMA, UBLKCP)
// behind an mbarrier when they fit, otherwise read through L1/L2.
//
// Arithmetic: every value that can reach a reference-visible result is computed with the
// reference's operation order in separately rounded IEEE FP32 ops (um_math.cuh). The only FMAs are
// in the conservative bounding-sphere rejection of OBBs, which can only skip tests whose exact
// outcome is a miss.
#include "device_util.cuh"
#include "intersect.cuh"
#include "scene_dev.cuh"
#include "um_math.cuh"

#include <type_traits>

namespace art {

// ---- per-warp query records (shared memory) ------------------------------------------------------
// rec[0*32+q] = (inv.x, inv.y, inv.z, limit)   rec[1*32+q] = (d.x, d.y, d.z, dot(d,d))
// rec[2*32+q] = (bits: targetId, bits: class, 0, 0)
constexpr int kRecFloat4PerWarp = 96;
constexpr int kEchoOwnerId = -0x40000000;   // never equals an int16 AudioTargetId

struct WarpCounters {
    unsigned long long v[C_COUNT];
};

// Owner filter on the hit path (RT:413/426/439): drop blockers owned by the query's target.
template <int R>
__device__ __forceinline__ uint32_t owner_filter(uint32_t hm, const short* own, int base, int lane, int ownerId)
{
    if (hm) {
#pragma unroll
        for (int r = 0; r < R; r++)
            if ((hm >> r) & 1u)
                if ((int)own[base + r * 32 + lane] == ownerId) hm &= ~(1u << r);
    }
    return hm;
}

// Counter helper: tests the reference would have executed in one section for a query that was
// blocked at canonical index `firstIdx` of that section (owned colliders are skipped, not tested).
__device__ __forceinline__ int owned_upto(const short* own, int uptoInclusive, int ownerId, int lane)
{
    int c = 0;
    for (int i = lane; i <= uptoInclusive; i += 32) c += ((int)own[i] == ownerId) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
    return c;
}

template <bool SMEM, bool COUNT>
__global__ void __launch_bounds__(kThreads, 1) trace_kernel(const TraceArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Na = a.nTargets;
    const int G = (Na + 1 + 31) >> 5;          // query slot groups: slot 0 = echo, slot 1+t = target t
    const int slotsPad = G * 32;

    unsigned char* p = smem;
    const unsigned char* geomBase = a.geom;
    if (SMEM) {
        stage_blob_to_smem(p, a.geom, a.L.bytes, &bar);
        geomBase = p;
        p += a.L.bytes;
    }
    float4* rec = reinterpret_cast<float4*>(p) + warp * kRecFloat4PerWarp;
    p += kWarpsPerCta * kRecFloat4PerWarp * sizeof(float4);
    uint32_t* mcnt = reinterpret_cast<uint32_t*>(p) + warp * slotsPad;
    const GeomView gv = make_view(geomBase, a.L);
    const int ns = a.L.ns, na = a.L.na, no = a.L.no;
    const int nsPad = a.L.nsPad, naPad = a.L.naPad, noPad = a.L.noPad;

    if (a.muffleInSmem)
        for (int i = lane; i < slotsPad; i += 32) mcnt[i] = 0;
    __syncwarp();

    const f3 RayOrigin = mk3(a.ox, a.oy, a.oz);
    int curRow = -1;
    unsigned int nSegments = 0, nSegHits = 0;
    unsigned long long cnt[COUNT ? C_COUNT : 1];
    if (COUNT)
        for (int i = 0; i < C_COUNT; i++) cnt[i] = 0;

    auto flush_muffle = [&](int row) {
        if (a.muffleInSmem && row >= 0) {
            for (int i = lane; i < slotsPad; i += 32) {
                uint32_t v = mcnt[i];
                if (v && i >= 1 && i <= Na) atomicAdd(&a.muffleCounts[row * Na + (i - 1)], v);
                mcnt[i] = 0;
            }
        }
        __syncwarp();
    };

    for (;;) {
        int j = 0;
        if (lane == 0) j = (int)atomicAdd(a.nextRay, 1u);
        j = __shfl_sync(kFull, j, 0);
        if (j >= a.map.nLocal) break;
        const int rayIndex = a.map.to_global(j);
        const int row = rayIndex / a.batchSize;   // batch k of ART:161/191; slot row (RT:63-64) applied at finalisation
        if (row != curRow) { flush_muffle(curRow); curRow = row; }

        f3 d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                   um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));   // RT:94
        f3 o = RayOrigin;                                            // RT:95
        int hits = 0;                                                // RT:97 (byte; H <= 255)
        float life = a.maxRayLife;                                   // RT:99
        bool alive = true;

        while (alive) {                                              // RT:104
            // ================= ShootRayCast (RT:225-280) =================
            nSegments++;
            float best = kFloatMax;
            uint32_t bkey = 0xFFFFFFFFu;   // (typeOrder << 28) | index ; typeOrder sphere 0, AABB 1, OBB 2
            {
                const float dd = dot3(d, d);                 // RT:326
                for (int base = 0; base < nsPad; base += SC_S) {
                    f3 oc[RS]; float cc[RS];
                    uint32_t need = 0;
#pragma unroll
                    for (int r = 0; r < RS; r++) {
                        const float4 s = gv.sph[base + r * 32 + lane];
                        oc[r] = sub3(o, mk3(s.x, s.y, s.z));                   // RT:325
                        cc[r] = subr(dot3(oc[r], oc[r]), s.w);                 // RT:328
                        if (!sphere_fast_miss(oc[r], cc[r], d, dd)) need |= 1u << r;
                    }
                    if (need) {
#pragma unroll
                        for (int r = 0; r < RS; r++)
                            if ((need >> r) & 1u) {
                                const float dist = sphere_dist_exact(oc[r].x, oc[r].y, oc[r].z, cc[r], d.x, d.y, d.z, dd);
                                if (dist < best) { best = dist; bkey = (uint32_t)(base + r * 32 + lane); }
                            }
                    }
                }
                const float ix = rcpr(d.x), iy = rcpr(d.y), iz = rcpr(d.z);  // RT:289
                for (int base = 0; base < naPad; base += SC_A) {
#pragma unroll
                    for (int r = 0; r < RA; r++) {
                        const int idx = base + r * 32 + lane;
                        const float4 A = gv.aabbA[idx];
                        const float2 B = gv.aabbB[idx];
                        float tNear, tFar, dist;
                        slab<8>(subr(A.x, o.x), subr(A.y, o.y), subr(A.z, o.z), subr(A.w, o.x), subr(B.x, o.y), subr(B.y, o.z),
                                ix, iy, iz, tNear, tFar);                      // RT:291-298
                        if (slab_hit(tNear, tFar, dist) && dist < best) { best = dist; bkey = (1u << 28) | (uint32_t)idx; }
                    }
                }
                for (int base = 0; base < noPad; base += SC_O) {
#pragma unroll
                    for (int r = 0; r < RO; r++) {
                        const int idx = base + r * 32 + lane;
                        const float4 q4 = gv.obbQ[idx];
                        const float4 c4 = gv.obbC[idx];
                        const float2 h2 = gv.obbH[idx];
                        const f3 h = mk3(c4.w, h2.x, h2.y);
                        const f3 pc = sub3(o, mk3(c4.x, c4.y, c4.z));        // RT:316 rayOrigin - Center
                        if (!obb_sure_miss(pc, obb_cull_c(pc, h), d, dd)) {
                            const float dist = obb_dist_exact(q4.x, q4.y, q4.z, q4.w, pc.x, pc.y, pc.z, h.x, h.y, h.z, d.x, d.y, d.z);
                            if (dist < best) { best = dist; bkey = (2u << 28) | (uint32_t)idx; }
                        }
                    }
                }
                if (COUNT) { cnt[C_TRACE_S] += ns; cnt[C_TRACE_A] += na; cnt[C_TRACE_O] += no; }
            }
            // lexicographic (t, canonical index) minimum over the warp
            const float tmin = warp_min_f(best);
            uint32_t key = (best == tmin) ? bkey : 0xFFFFFFFFu;
            const uint32_t wkey = __reduce_min_sync(kFull, key);
            if (wkey == 0xFFFFFFFFu) break;                                   // RT:201-207 ray left the scene
            const int src = __ffs(__ballot_sync(kFull, key == wkey)) - 1;
            const float rayHitDist = __shfl_sync(kFull, best, src);
            const int hitType = (int)(wkey >> 28);                            // 0 sphere, 1 AABB, 2 OBB
            const int hitIdx = (int)(wkey & 0x0FFFFFFFu);
            nSegHits++;

            o = add3(o, mul3s(d, rayHitDist));                                // RT:111
            life = subr(life, rayHitDist);                                    // RT:112
            hits += 1;                                                        // RT:113
            const size_t rayResultId = (size_t)j * a.H + hits - 1;            // RT:115 (local indexing)

            const float4 attr = hitType == 0 ? a.at.sphAttr[hitIdx] : (hitType == 1 ? a.at.aabbAttr[hitIdx] : a.at.obbAttr[hitIdx]);
            if (lane == 0) {
                if (a.hitPoints) {                                            // RT:118, 197
                    a.hitPoints[3 * rayResultId] = um_f32tof16(o.x);
                    a.hitPoints[3 * rayResultId + 1] = um_f32tof16(o.y);
                    a.hitPoints[3 * rayResultId + 2] = um_f32tof16(o.z);
                }
                if (a.hitIds) {
                    const uint32_t refType = hitType == 0 ? 3u : (hitType == 1 ? 1u : 2u);   // Enums/ColliderType.cs
                    a.hitIds[rayResultId] = (refType << 30) | (uint32_t)hitIdx;
                }
            }

            // ================= echo + muffle queries (RT:121-175) =================
            const f3 Pp = sub3(o, mul3s(d, kEpsilon));                        // RT:124 == RT:158
            for (int g = 0; g < G; g++) {
                const int qslot = g * 32 + lane;
                const bool valid = qslot <= Na;
                float L = 0.0f;
                bool gate = false;
                int ownerId = kEchoOwnerId;
                if (valid) {
                    f3 T = RayOrigin;
                    if (qslot > 0) {
                        T = mk3(a.targets[3 * (qslot - 1)], a.targets[3 * (qslot - 1) + 1], a.targets[3 * (qslot - 1) + 2]);
                        ownerId = qslot - 1;
                    }
                    const f3 v = sub3(T, Pp);                                 // RT:127 / RT:162
                    const float len = sqrtr(dot3(v, v));
                    const f3 dir = smul3(rcpr(len), v);                       // normalize = rsqrt(dot) * v
                    if (qslot == 0) {
                        const f3 w = sub3(o, RayOrigin);                      // RT:130 distance(RayOrigin, cRayOrigin)
                        L = sqrtr(dot3(w, w));
                        gate = true;
                    } else {
                        L = len;                                              // RT:165
                        gate = L < a.maxMuffle;                               // RT:168
                    }
                    const float ix = rcpr(dir.x), iy = rcpr(dir.y), iz = rcpr(dir.z);
                    rec[lane] = make_float4(ix, iy, iz, L);
                    rec[32 + lane] = make_float4(dir.x, dir.y, dir.z, dot3(dir, dir));
                    rec[64 + lane] = make_float4(__int_as_float(ownerId), __int_as_float(slab_class(ix, iy, iz)), 0.0f, 0.0f);
                }
                __syncwarp();
                const uint32_t active = __ballot_sync(kFull, valid && gate);
                uint32_t open = active;          // queries not yet blocked
                if (COUNT) {
                    if (g == 0) cnt[C_ECHO_Q] += 1;
                    cnt[C_MUFFLE_Q] += __popc(g == 0 ? (active & ~1u) : active);
                }

                // tests the reference executes for a query blocked at `firstIdx` of a section / not blocked
                auto count_blocked = [&](int q, int section, int firstIdx) {
                    if (!COUNT) return;
                    const int oid = __float_as_int(rec[64 + q].x);
                    const bool isEcho = (g == 0 && q == 0);
                    const int cbase = isEcho ? C_ECHO_S : C_MUFFLE_S;
                    const int nSec[3] = { ns, na, no };
                    const short* own[3] = { a.at.ownS, a.at.ownA, a.at.ownO };
                    for (int s = 0; s < section; s++) {
                        int owned = (!isEcho && oid >= 0 && oid < Na) ? a.at.ownedCount[s * Na + oid] : 0;
                        cnt[cbase + s] += nSec[s] - owned;
                    }
                    int ownedBefore = 0;
                    if (!isEcho && oid >= 0 && oid < Na && a.at.ownedCount[section * Na + oid] > 0)
                        ownedBefore = owned_upto(own[section], firstIdx, oid, lane);
                    cnt[cbase + section] += firstIdx + 1 - ownedBefore;
                };

                // ---- spheres (RT:370-377 / 408-419)
                for (int base = 0; base < nsPad && open; base += SC_S) {
                    f3 oc[RS]; float ccm[RS];
#pragma unroll
                    for (int r = 0; r < RS; r++) {
                        const float4 s = gv.sph[base + r * 32 + lane];
                        oc[r] = sub3(Pp, mk3(s.x, s.y, s.z));
                        ccm[r] = sphere_cull_c(subr(dot3(oc[r], oc[r]), s.w), s.w);
                    }
                    uint32_t m = open;
                    while (m) {
                        const int q = __ffs(m) - 1;
                        m &= m - 1;
                        const float4 r1 = rec[32 + q];
                        const f3 qd = mk3(r1.x, r1.y, r1.z);
                        uint32_t need = 0, hm = 0;
#pragma unroll
                        for (int r = 0; r < RS; r++)
                            if (!sphere_sure_miss(oc[r], ccm[r], qd, r1.w)) need |= 1u << r;
                        if (need) {
                            const float limit = rec[q].w;
#pragma unroll
                            for (int r = 0; r < RS; r++)
                                if ((need >> r) & 1u) {
                                    const float cc = subr(dot3(oc[r], oc[r]), gv.sph[base + r * 32 + lane].w);   // RT:328
                                    if (sphere_dist_exact(oc[r].x, oc[r].y, oc[r].z, cc, qd.x, qd.y, qd.z, r1.w) < limit) hm |= 1u << r;
                                }
                        }
                        if (__any_sync(kFull, hm != 0)) {
                            if (a.anyOwned[0]) hm = owner_filter<RS>(hm, a.at.ownS, base, lane, __float_as_int(rec[64 + q].x));
                            const uint32_t bal = __ballot_sync(kFull, hm != 0);
                            if (bal) {
                                open &= ~(1u << q);
                                if (COUNT) {
                                    int first = hm ? base + (__ffs(hm) - 1) * 32 + lane : 0x7FFFFFFF;
                                    first = __reduce_min_sync(kFull, first);
                                    count_blocked(q, 0, first < ns ? first : ns - 1);
                                }
                            }
                        }
                    }
                }
                // ---- AABBs (RT:379-386 / 421-432)
                for (int base = 0; base < naPad && open; base += SC_A) {
                    float lox[RA], loy[RA], loz[RA], hix[RA], hiy[RA], hiz[RA];
#pragma unroll
                    for (int r = 0; r < RA; r++) {
                        const float4 A = gv.aabbA[base + r * 32 + lane];
                        const float2 B = gv.aabbB[base + r * 32 + lane];
                        lox[r] = subr(A.x, Pp.x); loy[r] = subr(A.y, Pp.y); loz[r] = subr(A.z, Pp.z);
                        hix[r] = subr(A.w, Pp.x); hiy[r] = subr(B.x, Pp.y); hiz[r] = subr(B.y, Pp.z);
                    }
                    uint32_t m = open;
                    while (m) {
                        const int q = __ffs(m) - 1;
                        m &= m - 1;
                        const float4 r0 = rec[q];
                        const int cls = __float_as_int(rec[64 + q].y);
                        bool hb[RA];
                        auto body = [&](auto clsTag) {
                            constexpr int CLS = decltype(clsTag)::value;
#pragma unroll
                            for (int r = 0; r < RA; r++) {
                                float tNear, tFar, dist;
                                slab<CLS>(lox[r], loy[r], loz[r], hix[r], hiy[r], hiz[r], r0.x, r0.y, r0.z, tNear, tFar);
                                hb[r] = slab_hit(tNear, tFar, dist) && dist < r0.w;
                            }
                        };
                        switch (cls) {
                        case 0: body(std::integral_constant<int, 0>{}); break;
                        case 1: body(std::integral_constant<int, 1>{}); break;
                        case 2: body(std::integral_constant<int, 2>{}); break;
                        case 3: body(std::integral_constant<int, 3>{}); break;
                        case 4: body(std::integral_constant<int, 4>{}); break;
                        case 5: body(std::integral_constant<int, 5>{}); break;
                        case 6: body(std::integral_constant<int, 6>{}); break;
                        case 7: body(std::integral_constant<int, 7>{}); break;
                        default: body(std::integral_constant<int, 8>{}); break;
                        }
                        bool anyb = false;
#pragma unroll
                        for (int r = 0; r < RA; r++) anyb |= hb[r];
                        if (__any_sync(kFull, anyb)) {
                            uint32_t hm = 0;
#pragma unroll
                            for (int r = 0; r < RA; r++) hm |= hb[r] ? (1u << r) : 0u;
                            if (a.anyOwned[1]) hm = owner_filter<RA>(hm, a.at.ownA, base, lane, __float_as_int(rec[64 + q].x));
                            const uint32_t bal = __ballot_sync(kFull, hm != 0);
                            if (bal) {
                                open &= ~(1u << q);
                                if (COUNT) {
                                    int first = hm ? base + (__ffs(hm) - 1) * 32 + lane : 0x7FFFFFFF;
                                    first = __reduce_min_sync(kFull, first);
                                    count_blocked(q, 1, first < na ? first : na - 1);
                                }
                            }
                        }
                    }
                }
                // ---- OBBs (RT:388-395 / 434-445)
                for (int base = 0; base < noPad && open; base += SC_O) {
                    float4 oq[RO]; f3 pc[RO], hh[RO]; float cB[RO];
#pragma unroll
                    for (int r = 0; r < RO; r++) {
                        oq[r] = gv.obbQ[base + r * 32 + lane];
                        const float4 c4 = gv.obbC[base + r * 32 + lane];
                        const float2 h2 = gv.obbH[base + r * 32 + lane];
                        hh[r] = mk3(c4.w, h2.x, h2.y);
                        pc[r] = sub3(Pp, mk3(c4.x, c4.y, c4.z));
                        cB[r] = obb_cull_c(pc[r], hh[r]);
                    }
                    uint32_t m = open;
                    while (m) {
                        const int q = __ffs(m) - 1;
                        m &= m - 1;
                        const float4 r1 = rec[32 + q];
                        const f3 qd = mk3(r1.x, r1.y, r1.z);
                        uint32_t need = 0, hm = 0;
#pragma unroll
                        for (int r = 0; r < RO; r++)
                            if (!obb_sure_miss(pc[r], cB[r], qd, r1.w)) need |= 1u << r;
                        if (need) {
                            const float limit = rec[q].w;
#pragma unroll
                            for (int r = 0; r < RO; r++)
                                if ((need >> r) & 1u)
                                    if (obb_dist_exact(oq[r].x, oq[r].y, oq[r].z, oq[r].w, pc[r].x, pc[r].y, pc[r].z,
                                                       hh[r].x, hh[r].y, hh[r].z, qd.x, qd.y, qd.z) < limit) hm |= 1u << r;
                        }
                        if (__any_sync(kFull, hm != 0)) {
                            if (a.anyOwned[2]) hm = owner_filter<RO>(hm, a.at.ownO, base, lane, __float_as_int(rec[64 + q].x));
                            const uint32_t bal = __ballot_sync(kFull, hm != 0);
                            if (bal) {
                                open &= ~(1u << q);
                                if (COUNT) {
                                    int first = hm ? base + (__ffs(hm) - 1) * 32 + lane : 0x7FFFFFFF;
                                    first = __reduce_min_sync(kFull, first);
                                    count_blocked(q, 2, first < no ? first : no - 1);
                                }
                            }
                        }
                    }
                }
                if (COUNT) {   // queries that stayed open ran all three sections to the end
                    uint32_t m = open;
                    while (m) {
                        const int q = __ffs(m) - 1;
                        m &= m - 1;
                        const int oid = __float_as_int(rec[64 + q].x);
                        const bool isEcho = (g == 0 && q == 0);
                        const int cbase = isEcho ? C_ECHO_S : C_MUFFLE_S;
                        const int nSec[3] = { ns, na, no };
                        for (int s = 0; s < 3; s++) {
                            int owned = (!isEcho && oid >= 0 && oid < Na) ? a.at.ownedCount[s * Na + oid] : 0;
                            cnt[cbase + s] += nSec[s] - owned;
                        }
                    }
                }

                // results of this group
                if (g == 0 && (open & 1u)) {                                  // RT:133-145 echo ray returned
                    if (lane == 0) {
                        const float echoMul = attr.y;                          // RT:135-141 material Echo
                        a.echo[rayResultId] = um_f32tof16(mulr(L, echoMul));   // RT:142-144 (lane 0 holds slot 0's L)
                    }
                }
                const uint32_t vis = (g == 0) ? (open & ~1u) : open;          // RT:168-172
                if ((vis >> lane) & 1u) {
                    if (a.muffleInSmem) mcnt[qslot] += 1;
                    else atomicAdd(&a.muffleCounts[row * Na + (qslot - 1)], 1u);
                }
                __syncwarp();
            }

            // ================= termination / reflection (RT:178-193) =================
            if (hits >= a.H || life <= 0.0f) {
                alive = false;
            } else {
                // ReflectRay RT:456-532
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
        }
        if (lane == 0 && a.hitCounts) a.hitCounts[j] = (uint8_t)hits;                            // RT:204, 212
    }

    flush_muffle(curRow);
    if (lane == 0) {
        atomicAdd(&a.counters[C_SEGMENTS], (unsigned long long)nSegments);
        atomicAdd(&a.counters[C_SEGMENT_HITS], (unsigned long long)nSegHits);
        if (COUNT)
            for (int i = C_TRACE_S; i <= C_MUFFLE_O; i++)
                if (cnt[i]) atomicAdd(&a.counters[i], cnt[i]);
    }
}

// ---- launcher -----------------------------------------------------------------------------------
size_t trace_smem_bytes(const GeomLayout& L, int nTargets, bool geomInSmem, bool muffleInSmem)
{
    size_t b = geomInSmem ? L.bytes : 0;
    b += (size_t)kWarpsPerCta * kRecFloat4PerWarp * sizeof(float4);
    if (muffleInSmem) b += (size_t)kWarpsPerCta * (((nTargets + 1 + 31) / 32) * 32) * sizeof(uint32_t);
    return b;
}

cudaError_t launch_trace(const TraceArgs& a, int numCtas, bool geomInSmem, bool count, cudaStream_t stream)
{
    const size_t smem = trace_smem_bytes(a.L, a.nTargets, geomInSmem, a.muffleInSmem != 0);
    void (*k)(const TraceArgs) = nullptr;
    if (geomInSmem) k = count ? trace_kernel<true, true> : trace_kernel<true, false>;
    else k = count ? trace_kernel<false, true> : trace_kernel<false, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<numCtas, kThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace art
