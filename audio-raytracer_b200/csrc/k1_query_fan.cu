// k1_query_fan.cu -- the echo-return ray (AudioRaytracerJobBatched.cs:124-145) and the Na muffle rays (RT:153-173) of
// every hit point, evaluated AFTER the bounce loop instead of inside it.
//
// RT:124-173 only WRITE EchoRayDistances / MuffleRayHits: nothing in that block feeds cRayOrigin / cRayDirection / cRayLife
// (RT:179-192). So the bounce tracer (trace_grid_kernel, bounce-only mode) just appends one record per hit point --
// (hit - eps*d, distance(RayOrigin, hit), material Echo, rayResultId) -- and this kernel evaluates the (Na + 1) any-hit
// queries of all records against the target fans (fan_dev.cuh). A query tests the goal's near list and the direction bin of
// (hit point - goal): AABBs first (nearest to the goal first), then spheres, then OBBs. The per-collider tests are the exact
// FP32 functions of the brute-force kernels (grid_dev.cuh / intersect.cuh), and an occlusion query is an "any" over all
// colliders (RT:365-449), so which non-blocking colliders are skipped and in which order the others are tested cannot change
// a result: outputs are bit-identical with k1_trace.cu and the oracle.
//
// Mapping. A warp takes 32 consecutive records; LANE = HIT POINT (its record stays in registers), the goals are walked in
// lock step, so the goal position and its near-list header are warp-uniform shared-memory broadcasts and the bin header +
// first AABB ids come with ONE 16-byte load (FanDesc::cells4):
//   pass 0   every (hit point, goal) query is tested against the first two AABBs of its lists -- with nearest-first lists
//            that blocks most queries -- by a CONSERVATIVE classification that needs no exact set-up: the ray is left
//            un-normalised (parameter s in [0, 1] along goal - P, three rcp.approx), and a test only answers "certainly
//            blocked" / "certainly not" when the reference's own FP32 evaluation cannot differ (q_classify_aabb: relative
//            error bounds on every compared quantity); everything else counts as undecided;
//   prepare  the queries pass 0 did not block (and the undecided ones, to be re-tested from the start) are queued with 16 B
//            each and PREPARED exactly, 32 at a time at full width (RT:127/162 normalize: exact sqrt and divide, RT:289:
//            three exact reciprocals) -- one third of the queries instead of all of them;
//   AABBs    the prepared queries wait WITH their state (64 B, self-contained) in a per-warp list in global scratch
//            (L2 resident). Whenever enough have gathered, every lane takes one into registers and tests one AABB per step;
//            a lane whose query is blocked takes the next one at once, so the lanes stay full whatever the list lengths.
//            When the queue runs dry the few queries still in flight are written back (with their cursor) and wait for the
//            next run -- no warp ever idles through the long lists of a few unobstructed queries;
//   S/O      a query no AABB blocks is queued the same way for the sphere and OBB lists; if nothing blocks, it sees its goal.
#include <algorithm>
#include <cstdlib>

#include "device_util.cuh"
#include "fan_dev.cuh"
#include "grid_dev.cuh"
#include "intersect.cuh"
#include "launchers.h"
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {

#ifndef ART_Q_WARPS
#define ART_Q_WARPS 32
#endif
#ifndef ART_Q_RUN
#define ART_Q_RUN 96
#endif
#ifndef ART_Q_MIN_LANES
#define ART_Q_MIN_LANES 20
#endif
constexpr int kQWarps = ART_Q_WARPS;
constexpr int kQThreads = kQWarps * 32;
constexpr int kQFirstTests = 2;                      // AABBs of pass 0 at most (their ids come with the headers, FanDesc::cells4)
constexpr int kQRun = ART_Q_RUN;                     // queued queries from which a refill loop runs
constexpr int kQMinLanes = ART_Q_MIN_LANES;          // a loop whose queue is dry stops (and writes its queries back) below this many busy lanes
constexpr int kQCap0 = kQRun + 32;                   // unprepared queries: < kQRun before a goal step, <= 32 more per step
constexpr int kQCapA = 2 * kQRun + 64;               // AABB queue: < kQRun + what one prepare run adds + <= 32 written back
constexpr int kQCapSO = 4 * kQRun + 128;             // sphere / OBB queue: additionally everything a prepare and an AABB run pass on
constexpr int kQEntry = 4;                           // float4 per prepared query
constexpr float kQRelErr = 1e-6f;                    // bound used for |reference value - conservative value| / |value| (derivation: q_classify_aabb)
constexpr int kQMuffleSmemMax = 4096;                // per-CTA muffle counters [T * Na] kept in shared memory up to this size

// A queued query, 64 B:  e0 = (1/dir.xyz, limit L)   e1 = (P.xyz, |goal - P|)   e2 = (bin header x, y, slot | cursor << 16, -)
//                        e3 = (material Echo, rayResultId, -, -)
struct QEnv {
    const QueryArgs& a;
    const FanDesc& f;
    const GeomView& gv;
    const float4* goalTab;       // shared memory (or null): goal position of slot s
    const uint4* nearTab;        // shared memory (or null): near-list header + first ids of slot s
    uint32_t* sMuffle;           // shared memory (or null): per-CTA muffle counters
    int lane;
    uint32_t ltMask;
    unsigned int* st;            // STATS: sphere / AABB / OBB tests, lists opened (this lane)
};

__device__ __forceinline__ int q_fan_of(const QueryArgs& a, int slot) { return slot == 0 ? a.nTargets : slot - 1; }
__device__ __forceinline__ f3 q_goal(const QEnv& E, int slot)
{
    if (E.goalTab) { const float4 g = E.goalTab[slot]; return mk3(g.x, g.y, g.z); }
    if (slot == 0) return mk3(E.a.ox, E.a.oy, E.a.oz);
    return mk3(E.a.targets[3 * (slot - 1)], E.a.targets[3 * (slot - 1) + 1], E.a.targets[3 * (slot - 1) + 2]);
}
__device__ __forceinline__ uint4 q_near(const QEnv& E, int slot)
{
    if (E.nearTab) return E.nearTab[slot];
    return __ldg(&E.f.cells4[(size_t)q_fan_of(E.a, slot) * kFanCells + 6 * kFanCellsPerFace]);
}
// the query sees its goal: RT:133-145 (echo ray, slot 0) / RT:168-172 (muffle ray of target slot - 1)
__device__ __forceinline__ void q_visible(const QEnv& E, int slot, float L, float echoMul, int resultId)
{
    if (slot == 0) E.a.echo[resultId] = um_f32tof16(mulr(L, echoMul));
    else {
        const int row = E.a.map.to_global(resultId / E.a.H) / E.a.batchSize;     // ART:161/191 batch of the ray
        const int idx = row * E.a.nTargets + (slot - 1);
        if (E.sMuffle) atomicAdd(&E.sMuffle[idx], 1u);
        else atomicAdd(&E.a.muffleCounts[idx], 1u);
    }
}

// Conservative AABB any-hit classification in un-normalised ray space: 1 = the reference's test (RT:284-308 + the caller's
// `dist < distToTarget`, RT:384 / RT:430) certainly reports a blocker, 0 = it certainly does not, 2 = undecided.
//
// The reference evaluates, per axis k, t0 = fl(A*inv), A = fl(min - P) (the same float here), inv = fl(1/nd), nd = fl(rl*v),
// rl = fl(1/len): t0 = (A/v) * (1/rl) * (1 + th), |th| <= 3 * 2^-24. Here a = fl(A * r), r = rcp.approx(v) (relative error
// <= 2^-23): a = (A/v) * (1 + th'), |th'| <= 3 * 2^-24. So rl * t0 and a differ by at most 6 * 2^-24 = 3.6e-7 of |a|; rl > 0
// is the same factor for every axis and collider, and min / max are monotone, hence rl * tNear lies within kQRelErr * |sn| of
// sn and rl * tFar within kQRelErr * |sf| of sf (kQRelErr = 1e-6). All values are finite (the caller guarantees 1e-12 <= |v_k|,
// 1e-12 <= |v|^2 <= 1e12), so the reference meets no NaN either. A comparison is decided only if it holds with the margin
// e = kQRelErr * max(|sn|, |sf|) + 1e-30 on both sides. The limit: rl * distToTarget lies in [limLo, limHi] (caller).
__device__ __forceinline__ int q_classify_aabb(const GeomView& gv, int id, f3 P, f3 r, float limLo, float limHi)
{
    const float4 A = gv.aabbA[id];
    const float2 B = gv.aabbB[id];
    const float ax = mulr(subr(A.x, P.x), r.x), bx = mulr(subr(A.w, P.x), r.x);
    const float ay = mulr(subr(A.y, P.y), r.y), by = mulr(subr(B.x, P.y), r.y);
    const float az = mulr(subr(A.z, P.z), r.z), bz = mulr(subr(B.y, P.z), r.z);
    const float sn = max3f(fminf(ax, bx), fminf(ay, by), fminf(az, bz));
    const float sf = min3f(fmaxf(ax, bx), fmaxf(ay, by), fmaxf(az, bz));
    const float e = fmaf(kQRelErr, fmaxf(fabsf(sn), fabsf(sf)), 1e-30f);
    const bool miss = (sn - sf > 2.0f * e) || (sf < -e);                  // tNear > tFar  or  tFar < 0, certainly
    const bool hit = (sf - sn > 2.0f * e) && (sf > e);                    // neither, certainly
    const bool pos = sn > e, neg = sn < -e;                               // the sign of tNear decides which distance is reported (RT:306)
    const float dist = pos ? sn : sf;
    if (hit && (pos || neg) && dist + e < limLo) return 1;
    if (miss || ((pos || neg) && dist - e >= limHi)) return 0;
    return 2;
}

// Exact set-up of the queries pass 0 queued (16 B each: record index, slot | first AABB to test << 16 | flags, bin header):
// 32 at a time, one lane per query -- RT:127 / RT:162 normalize (exact sqrt, exact divide), RT:289 the three reciprocals,
// RT:130 / RT:165 the limit, RT:168 the gate -- then on to the AABB queue, the sphere / OBB queue, or the goal is visible.
constexpr uint32_t kQFlagNoBin = 1u << 30;           // the direction goal -> hit point has no bin (hit point == goal, non-finite)
template <bool STATS>
__device__ __forceinline__ void q_prepare(const QEnv& E, const uint4* list0, int n0, float4* listA, int& nA, float4* listSO, int& nSO)
{
    const QueryArgs& a = E.a;
    for (int i0 = 0; i0 < n0; i0 += 32) {
        const int i = i0 + E.lane;
        int push = 0;
        float4 e0 = make_float4(0, 0, 0, 0), e1 = e0, e2 = e0, e3 = e0;
        if (i < n0) {
            const uint4 q = list0[i];
            const int slot = (int)(q.y & 0xFFFFu), cursor = (int)((q.y >> 16) & 0xFFFu);
            ART_CHECK(a.counters, slot <= a.nTargets && q.x < *a.recCount);
            const float4 ra = __ldg(&a.recA[q.x]);
            const float2 rb = __ldg(&a.recB[q.x]);
            const f3 P = mk3(ra.x, ra.y, ra.z);
            const f3 v = sub3(q_goal(E, slot), P);                             // RT:127 / RT:162
            const float len = sqrtr(dot3(v, v));
            float L = ra.w;                                                    // RT:130
            bool gate = true;
            if (slot > 0) { L = len; gate = L < a.maxMuffle; }                 // RT:165, 168
            if (gate) {
                if ((q.y & kQFlagNoBin) || len != len) {
                    q_visible(E, slot, L, rb.x, __float_as_int(rb.y));         // degenerate (hit point == goal): no test can block
                } else {
                    const f3 nd = smul3(rcpr(len), v);                         // normalize = rsqrt(dot) * v
                    const f3 inv = mk3(rcpr(nd.x), rcpr(nd.y), rcpr(nd.z));
                    const uint4 n4 = q_near(E, slot);
                    const int nAll = (int)((n4.y >> 10) & 2047u) + (int)((q.w >> 10) & 2047u);
                    if (cursor < nAll) push = 1;
                    else if (((n4.y | q.w) & 1023u) | ((n4.y | q.w) >> 21)) push = 2;
                    else q_visible(E, slot, L, rb.x, __float_as_int(rb.y));
                    e0 = make_float4(inv.x, inv.y, inv.z, L);
                    e1 = make_float4(P.x, P.y, P.z, len);
                    e2 = make_float4(__uint_as_float(q.z), __uint_as_float(q.w), __uint_as_float((uint32_t)slot | (push == 1 ? (uint32_t)cursor << 16 : 0u)), 0.0f);
                    e3 = make_float4(rb.x, rb.y, 0.0f, 0.0f);
                }
            }
        }
        const uint32_t am = __ballot_sync(kFull, push == 1), sm = __ballot_sync(kFull, push == 2);
        if (push) {
            const int pos = push == 1 ? nA + __popc(am & E.ltMask) : nSO + __popc(sm & E.ltMask);
            ART_CHECK(a.counters, pos < (push == 1 ? kQCapA : kQCapSO));
            float4* dst = (push == 1 ? listA : listSO) + (size_t)kQEntry * pos;
            dst[0] = e0; dst[1] = e1; dst[2] = e2; dst[3] = e3;
        }
        nA += __popc(am);
        nSO += __popc(sm);
    }
    __syncwarp();
}

// The AABB lists of the queued queries, entries [cursor, nA0 + nA1) of the run "near list, then bin". Every lane holds ONE
// query in registers and tests one AABB per step; a lane whose query is blocked (or has run through its lists) takes the
// next one from the queue at once. A query no AABB blocks goes to the sphere / OBB queue or sees its goal. With `drain`
// the loop runs until every query is resolved; otherwise it stops once the queue is empty and fewer than kQMinLanes lanes
// are busy, writes the queries in flight back to the front of the queue and returns their number.
template <bool STATS>
__device__ __forceinline__ int q_loop_aabb(const QEnv& E, float4* listA, int nIn, bool drain, float4* listSO, int& nSO)
{
    const QueryArgs& a = E.a; (void)a;
    int next = 0;
    bool have = false, anySO = false;
    float4 e0 = make_float4(0, 0, 0, 0), e1 = e0, e2 = e0, e3 = e0;
    uint32_t oN = 0, oB = 0;
    int nA0 = 0, nAll = 0, k = 0, idNext = 0;           // idNext: entry k, fetched one step ahead (the lists live in L2)
    for (;;) {
        const uint32_t need = __ballot_sync(kFull, !have);
        if (need && next < nIn) {
            const int idx = next + __popc(need & E.ltMask);
            if (!have && idx < nIn) {
                const float4* src = listA + (size_t)kQEntry * idx;
                e0 = src[0]; e1 = src[1]; e2 = src[2]; e3 = src[3];
                const uint32_t packed = __float_as_uint(e2.z);
                const int slot = (int)(packed & 0xFFFFu);
                ART_CHECK(a.counters, slot <= a.nTargets);
                const uint4 n4 = q_near(E, slot);
                const uint32_t hBx = __float_as_uint(e2.x), hBy = __float_as_uint(e2.y);
                const int nS0 = n4.y & 1023, nS1 = hBy & 1023, nA1 = (hBy >> 10) & 2047;
                nA0 = (n4.y >> 10) & 2047; nAll = nA0 + nA1;
                oN = n4.x + nS0; oB = hBx + nS1 - nA0;               // entry k: k < nA0 ? oN + k : oB + k
                k = (int)(packed >> 16);
                ART_CHECK(a.counters, k < nAll && n4.x + nS0 + nA0 <= (unsigned)E.f.nEntries && hBx + nS1 + nA1 <= (unsigned)E.f.nEntries);
                anySO = (((n4.y | hBy) & 1023u) | ((n4.y | hBy) >> 21)) != 0;
                idNext = (int)__ldg(E.f.entries + (k < nA0 ? oN : oB) + k);
                have = true;
            }
            next = min(nIn, next + __popc(need));
        }
        const uint32_t busy = __ballot_sync(kFull, have);
        if (!busy) break;                                // the queue is exhausted and every query has been resolved
        if (!drain && next >= nIn && __popc(busy) < kQMinLanes) break;
        bool toSO = false;
        if (have) {
            const int id = idNext;
            k++;
            if (k < nAll) idNext = (int)__ldg(E.f.entries + (k < nA0 ? oN : oB) + k);
            ART_CHECK(a.counters, id < a.L.na);
            if (STATS) E.st[1]++;
            if (aabb_blocks(E.gv, id, mk3(e1.x, e1.y, e1.z), mk3(e0.x, e0.y, e0.z), e0.w)) have = false;
            else if (k >= nAll) {
                have = false;
                if (anySO) toSO = true;
                else q_visible(E, (int)(__float_as_uint(e2.z) & 0xFFFFu), e0.w, e3.x, __float_as_int(e3.y));
            }
        }
        const uint32_t sm = __ballot_sync(kFull, toSO);
        if (sm) {
            if (toSO) {
                const int pos = nSO + __popc(sm & E.ltMask);
                ART_CHECK(a.counters, pos < kQCapSO);
                float4* dst = listSO + (size_t)kQEntry * pos;
                dst[0] = e0; dst[1] = e1;
                dst[2] = make_float4(e2.x, e2.y, __uint_as_float(__float_as_uint(e2.z) & 0xFFFFu), e2.w);   // cursor 0
                dst[3] = e3;
            }
            nSO += __popc(sm);
        }
    }
    // (only reached with queries in flight when !drain: the queue is empty, so they go to its front)
    const uint32_t busy = __ballot_sync(kFull, have);
    if (have) {
        float4* dst = listA + (size_t)kQEntry * __popc(busy & E.ltMask);
        dst[0] = e0; dst[1] = e1;
        dst[2] = make_float4(e2.x, e2.y, __uint_as_float((__float_as_uint(e2.z) & 0xFFFFu) | ((uint32_t)k << 16)), e2.w);
        dst[3] = e3;
    }
    __syncwarp();
    return __popc(busy);
}

// The queries no AABB blocks: sphere lists, then OBB lists (near list, then bin), one test per lane and step with the same
// refill scheme; the cursor runs over the spheres first, then the OBBs.
template <bool STATS>
__device__ __forceinline__ int q_loop_so(const QEnv& E, float4* listSO, int nIn, bool drain)
{
    const QueryArgs& a = E.a;
    int next = 0;
    bool have = false;
    float4 e1 = make_float4(0, 0, 0, 0), e2 = e1, e3 = e1;
    f3 d = mk3(0, 0, 0);
    float L = 0.0f, dd = 0.0f, i0x = 0.0f, i0y = 0.0f, i0z = 0.0f;
    uint32_t oSN = 0, oSB = 0, oON = 0, oOB = 0;
    int nS0 = 0, nS = 0, nO0 = 0, nTot = 0, k = 0, idNext = 0;
    for (;;) {
        const uint32_t need = __ballot_sync(kFull, !have);
        if (need && next < nIn) {
            const int idx = next + __popc(need & E.ltMask);
            if (!have && idx < nIn) {
                const float4* src = listSO + (size_t)kQEntry * idx;
                const float4 e0 = src[0];
                e1 = src[1]; e2 = src[2]; e3 = src[3];
                i0x = e0.x; i0y = e0.y; i0z = e0.z; L = e0.w;
                const uint32_t packed = __float_as_uint(e2.z);
                const int slot = (int)(packed & 0xFFFFu);
                ART_CHECK(a.counters, slot <= a.nTargets);
                const f3 v = sub3(q_goal(E, slot), mk3(e1.x, e1.y, e1.z));   // RT:127 / RT:162
                d = smul3(rcpr(e1.w), v);                                    // normalize = rsqrt(dot) * v, |v| from pass 0
                dd = dot3(d, d);
                const uint4 n4 = q_near(E, slot);
                const uint32_t hBx = __float_as_uint(e2.x), hBy = __float_as_uint(e2.y);
                const int nA0 = (n4.y >> 10) & 2047, nS1 = hBy & 1023, nA1 = (hBy >> 10) & 2047, nO1 = hBy >> 21;
                nS0 = n4.y & 1023; nO0 = n4.y >> 21;
                ART_CHECK(a.counters, n4.x + nS0 + nA0 + nO0 <= (unsigned)E.f.nEntries && hBx + nS1 + nA1 + nO1 <= (unsigned)E.f.nEntries);
                nS = nS0 + nS1; nTot = nS + nO0 + nO1;
                oSN = n4.x; oSB = hBx - nS0;                             // sphere k: k < nS0 ? oSN + k : oSB + k
                oON = n4.x + nS0 + nA0; oOB = hBx + nS1 + nA1 - nO0;     // OBB j = k - nS: j < nO0 ? oON + j : oOB + j
                k = (int)(packed >> 16);
                ART_CHECK(a.counters, k < nTot);
                idNext = (int)__ldg(E.f.entries + (k < nS ? (k < nS0 ? oSN : oSB) + k : (k - nS < nO0 ? oON : oOB) + (k - nS)));
                have = true;
            }
            next = min(nIn, next + __popc(need));
        }
        const uint32_t busy = __ballot_sync(kFull, have);
        if (!busy) break;
        if (!drain && next >= nIn && __popc(busy) < kQMinLanes) break;
        if (have) {
            const f3 P = mk3(e1.x, e1.y, e1.z);
            const int id = idNext;
            const bool isSphere = k < nS;
            k++;
            if (k < nTot) idNext = (int)__ldg(E.f.entries + (k < nS ? (k < nS0 ? oSN : oSB) + k : (k - nS < nO0 ? oON : oOB) + (k - nS)));
            bool blocked;
            if (isSphere) {
                ART_CHECK(a.counters, id < a.L.ns);
                if (STATS) E.st[0]++;
                blocked = sphere_dist(E.gv, id, P, d, dd) < L;           // RT:370-377 / RT:410-419
            } else {
                ART_CHECK(a.counters, id < a.L.no);
                if (STATS) E.st[2]++;
                blocked = obb_blocks(E.gv, id, P, d, dd, a.errScale, L); // RT:388-394 / RT:436-445
            }
            if (blocked) have = false;
            else if (k >= nTot) {
                have = false;
                q_visible(E, (int)(__float_as_uint(e2.z) & 0xFFFFu), L, e3.x, __float_as_int(e3.y));
            }
        }
    }
    const uint32_t busy = __ballot_sync(kFull, have);
    if (have) {
        float4* dst = listSO + (size_t)kQEntry * __popc(busy & E.ltMask);
        dst[0] = make_float4(i0x, i0y, i0z, L); dst[1] = e1;
        dst[2] = make_float4(e2.x, e2.y, __uint_as_float((__float_as_uint(e2.z) & 0xFFFFu) | ((uint32_t)k << 16)), e2.w);
        dst[3] = e3;
    }
    __syncwarp();
    return __popc(busy);
}

template <bool SMEM, bool STATS>
__global__ void __launch_bounds__(kQThreads, 1) query_fan_kernel(const QueryArgs a, const FanDesc f)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Na = a.nTargets;
    const int slots = Na + 1;                        // slot 0 = echo ray (goal: RayOrigin), slot 1 + t = muffle ray to target t

    unsigned char* p = smem;
    const unsigned char* geomBase = a.geom;
    if (SMEM) {
        stage_blob_to_smem(p, a.geom, a.L.bytes, &bar);
        geomBase = p;
        p += a.L.bytes;
    }
    float4* goalTab = nullptr; uint4* nearTab = nullptr; uint32_t* sMuffle = nullptr;
    if (a.tablesInSmem) {
        goalTab = reinterpret_cast<float4*>(p); p += (size_t)slots * sizeof(float4);
        nearTab = reinterpret_cast<uint4*>(p); p += (size_t)slots * sizeof(uint4);
        for (int s = threadIdx.x; s < slots; s += blockDim.x) {
            goalTab[s] = s == 0 ? make_float4(a.ox, a.oy, a.oz, 0.0f)
                                : make_float4(a.targets[3 * (s - 1)], a.targets[3 * (s - 1) + 1], a.targets[3 * (s - 1) + 2], 0.0f);
            nearTab[s] = __ldg(&f.cells4[(size_t)q_fan_of(a, s) * kFanCells + 6 * kFanCellsPerFace]);
        }
    }
    if (a.muffleInSmem) {
        sMuffle = reinterpret_cast<uint32_t*>(p);
        for (int i = threadIdx.x; i < a.muffleRows * Na; i += blockDim.x) sMuffle[i] = 0u;
    }
    __syncthreads();
    const GeomView gv = make_view(geomBase, a.L);
    const uint32_t ltMask = (1u << lane) - 1u;
    unsigned int st[4] = { 0, 0, 0, 0 };
    const QEnv E = { a, f, gv, goalTab, nearTab, sMuffle, lane, ltMask, st };
    float4* listA = a.scratch + ((size_t)blockIdx.x * kQWarps + warp) * (size_t)(kQEntry * (kQCapA + kQCapSO) + kQCap0);
    float4* listSO = listA + kQEntry * kQCapA;
    uint4* list0 = reinterpret_cast<uint4*>(listSO + kQEntry * kQCapSO);
    int n0 = 0, nA = 0, nSO = 0;                     // queued queries (kept across goals and record blocks)
    const unsigned int nRec = *a.recCount;           // the bounce tracer has finished (stream order)

    // Work unit = (block of 32 records, group of goals). With few records (a small batch against many targets) the launcher
    // splits the goals of a block over several warps so that the whole GPU is busy (QueryArgs::goalGroups); otherwise a unit
    // covers all goals of its block.
    const unsigned int nBlocks = (nRec + 31u) / 32u;

    unsigned int ticket = 0;                         // the NEXT unit: its ticket travels while the current unit is processed
    if (lane == 0) ticket = atomicAdd(a.queue, 1u);
    for (;;) {
        const unsigned int unit = __shfl_sync(kFull, ticket, 0);
        if (lane == 0) ticket = atomicAdd(a.queue, 1u);
        unsigned int blk = unit;
        int sBeg = 0, sEnd = slots;
        if (a.goalGroups > 1) {
            blk = unit / (unsigned)a.goalGroups;
            sBeg = (int)(unit - blk * (unsigned)a.goalGroups) * a.goalsPerGroup;
            sEnd = min(slots, sBeg + a.goalsPerGroup);
        }
        if (blk >= nBlocks) break;
        const unsigned int ri = blk * 32u + (unsigned)lane;
        const bool valid = ri < nRec;
        // the lane keeps only the hit point in registers; the rest of its record is re-read (L1) where a query needs it
        f3 P = mk3(0, 0, 0);
        if (valid) {
            const float4 ra = a.recA[ri];
            P = mk3(ra.x, ra.y, ra.z);
            ART_CHECK(a.counters, __float_as_int(a.recB[ri].y) >= 0 && __float_as_int(a.recB[ri].y) / a.H < a.map.nLocal);
        }
        // ---- pass 0: lane = hit point, all lanes walk the goals together; conservative tests only (q_classify_aabb)
        for (int s = sBeg; s < sEnd; s++) {
            bool push = false;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            const uint4 n4 = q_near(E, s);
            if (valid) {
                const f3 v = sub3(q_goal(E, s), P);                                // RT:127 / RT:162 (the reference's own operand)
                // the bin header depends on the direction's bin only: its load (L2) is issued first
                const int bin = fan_bin(-v.x, -v.y, -v.z);                         // direction goal -> hit point
                uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
                if (bin >= 0) c4 = __ldg(&f.cells4[(size_t)q_fan_of(a, s) * kFanCells + bin]);
                const float d2 = fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x));        // |v|^2 up to 2 ulp
                // un-normalised ray: X(s) = P + s * v, s = t / len; everything below compares rl * t = s-values
                const f3 r = mk3(rcp_fast(v.x), rcp_fast(v.y), rcp_fast(v.z));
                const bool sure = fminf(fminf(fabsf(v.x), fabsf(v.y)), fabsf(v.z)) >= 1e-12f && d2 >= 1e-12f && d2 <= 1e12f && bin >= 0;
                // RT:168 gate (muffle rays): |v| against MaxMuffleHitDistance, decided only outside a 2e-6 band
                const bool gateFail = s > 0 && sure && d2 > a.gateHi2, gateSure = s == 0 || d2 < a.gateLo2;
                float limLo = 1.0f - 1e-6f, limHi = 1.0f + 1e-6f;                  // muffle: distToTarget = len, len * fl(1 / len) = 1 +- 2^-24
                if (s == 0) {                                                      // echo: distToStartOrigin (RT:130) against |v|
                    const float sl = __ldg(&a.recA[ri].w) * rsqrt_fast(d2);        // rl within 6e-7 of rsqrt.approx(|v|^2)
                    limLo = fmaf(-2e-6f, sl, sl) - 1e-30f; limHi = fmaf(2e-6f, sl, sl) + 1e-30f;
                }
                const int nA0 = (n4.y >> 10) & 2047, nA1 = (c4.y >> 10) & 2047, nAll = nA0 + nA1;
                if (STATS) st[3] += 2;
                // the run "near list, then bin" starts with these ids: no dependent load of the entry lists here
                const uint32_t ids = nA0 >= 2 ? n4.z : (nA0 == 1 ? (n4.z & 0xFFFFu) | (c4.z << 16) : c4.z);
                const int nFirst = min(nAll, a.firstTests);
                bool blocked = false, unsure = !sure || !gateSure;
                if (sure && !gateFail) {
#pragma unroll 1
                    for (int t = 0; t < nFirst && !blocked; t++) {
                        const int id = (int)((ids >> (16 * t)) & 0xFFFFu);
                        ART_CHECK(a.counters, id < a.L.na);
                        if (STATS) st[1]++;
                        const int c = q_classify_aabb(gv, id, P, r, limLo, limHi);
                        blocked = c == 1;
                        unsure = unsure || c == 2;
#ifdef ART_Q_VERIFY
                        {   // cross-check against the exact evaluation (debug builds): a decided test must agree with it
                            const float lenX = sqrtr(dot3(v, v));
                            const f3 ndX = smul3(rcpr(lenX), v);
                            const f3 invX = mk3(rcpr(ndX.x), rcpr(ndX.y), rcpr(ndX.z));
                            const float LX = s == 0 ? a.recA[ri].w : lenX;
                            const bool bx = aabb_blocks(gv, id, P, invX, LX);
                            const bool gx = s == 0 || LX < a.maxMuffle;
                            if ((c == 1 && gx && !bx) || (c == 0 && bx) || (gateSure && !gx)) atomicAdd(&a.counters[C_DEBUG_VIOLATIONS], 1ull);
                        }
#endif
                    }
                }
#ifdef ART_Q_VERIFY
                if (gateFail && sqrtr(dot3(v, v)) < a.maxMuffle) atomicAdd(&a.counters[C_DEBUG_VIOLATIONS], 1ull);
#endif
                if (!blocked && !gateFail) {
                    // a decided-clear query with nothing left to test sees its goal; everything else is prepared exactly
                    // (q_prepare) and continues at AABB `nFirst`, or -- if anything was undecided -- from the start
                    if (!unsure && nAll <= nFirst && !(((n4.y | c4.y) & 1023u) | ((n4.y | c4.y) >> 21))) {
                        const float2 rb = __ldg(&a.recB[ri]);
                        q_visible(E, s, s == 0 ? __ldg(&a.recA[ri].w) : 0.0f, rb.x, __float_as_int(rb.y));
                    } else {
                        push = true;
                        q = make_uint4(ri, (uint32_t)s | ((uint32_t)(unsure ? 0 : nFirst) << 16) | (bin < 0 ? kQFlagNoBin : 0u), c4.x, c4.y);
                    }
                }
            }
            const uint32_t pm = __ballot_sync(kFull, push);
            if (push) {
                const int pos = n0 + __popc(pm & ltMask);
                ART_CHECK(a.counters, pos < kQCap0);
                list0[pos] = q;
            }
            n0 += __popc(pm);
            // ---- exact set-up of the queued queries, their remaining AABBs, their sphere and OBB lists -- each whenever
            //      enough have gathered to fill the lanes
            if (n0 >= kQRun) { __syncwarp(); q_prepare<STATS>(E, list0, n0, listA, nA, listSO, nSO); n0 = 0; }
            if (nA >= kQRun) { __syncwarp(); nA = q_loop_aabb<STATS>(E, listA, nA, false, listSO, nSO); }
            if (nSO >= kQRun) { __syncwarp(); nSO = q_loop_so<STATS>(E, listSO, nSO, false); }
        }
    }
    __syncwarp();
    if (n0 > 0) q_prepare<STATS>(E, list0, n0, listA, nA, listSO, nSO);
    if (nA > 0) q_loop_aabb<STATS>(E, listA, nA, true, listSO, nSO);
    if (nSO > 0) q_loop_so<STATS>(E, listSO, nSO, true);

    if (sMuffle) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.muffleRows * Na; i += blockDim.x) {
            const uint32_t v = sMuffle[i];
            if (v) atomicAdd(&a.muffleCounts[i], v);
        }
    }
    if (STATS) {
        atomicAdd(&a.counters[C_GRID_Q_S], (unsigned long long)st[0]);
        atomicAdd(&a.counters[C_GRID_Q_A], (unsigned long long)st[1]);
        atomicAdd(&a.counters[C_GRID_Q_O], (unsigned long long)st[2]);
        atomicAdd(&a.counters[C_GRID_Q_LISTS], (unsigned long long)st[3]);
    }
}

// ---- launcher -----------------------------------------------------------------------------------
size_t query_fan_smem_bytes(const GeomLayout& L, bool geomInSmem) { return geomInSmem ? L.bytes : 0; }
size_t query_fan_scratch_bytes(int numCtas) { return (size_t)numCtas * kQWarps * (kQEntry * (kQCapA + kQCapSO) + kQCap0) * sizeof(float4); }

cudaError_t launch_query_fan(const QueryArgs& a0, const FanDesc& fans, int numCtas, bool geomInSmem, bool stats, int maxSmemOptin, cudaStream_t stream)
{
    QueryArgs a = a0;
    size_t smem = query_fan_smem_bytes(a.L, geomInSmem);
    const size_t tables = (size_t)(a.nTargets + 1) * (sizeof(float4) + sizeof(uint4));
    a.tablesInSmem = 0; a.muffleInSmem = 0;
    const char* noTab = getenv("ART_K1_NO_GOAL_TABLES");         // (read per launch: test knob for the global-memory fallbacks)
    const bool tabs = !(noTab && atoi(noTab) != 0);
    // few records (upper bound: local rays x MaxHitsPerRay): split every block's goals over several warps
    {
        const long long maxBlocks = ((long long)a.map.nLocal * a.H + 31) / 32, warps = (long long)numCtas * kQWarps;
        const int slots = a.nTargets + 1;
        long long g = 1;
        if (maxBlocks > 0 && maxBlocks < 2 * warps) g = std::min<long long>(slots, (2 * warps + maxBlocks - 1) / maxBlocks);
        a.goalsPerGroup = (int)((slots + g - 1) / g);
        a.goalGroups = (slots + a.goalsPerGroup - 1) / a.goalsPerGroup;
    }
    // RT:168 gate, decided conservatively in pass 0: |v|^2 against MaxMuffleHitDistance^2 outside a relative band of 4e-6
    {
        const double m = (double)a.maxMuffle;
        a.gateLo2 = (float)(m * m * (1.0 - 4e-6));
        a.gateHi2 = (float)(m * m * (1.0 + 4e-6));
        if (!(m > 0.0)) { a.gateLo2 = 0.0f; a.gateHi2 = 0.0f; }      // (0, negative or NaN: nothing is certain, the exact compare decides)
        if (m != m) { a.gateLo2 = -1.0f; a.gateHi2 = 3.0e38f; }
    }
    // experiment knob (read per launch): AABBs tested in pass 0 (1 or 2)
    a.firstTests = kQFirstTests;
    if (const char* v = getenv("ART_Q_FIRST_TESTS")) { const int n = atoi(v); if (n >= 1 && n <= kQFirstTests) a.firstTests = n; }
    if (tabs && smem + tables <= (size_t)maxSmemOptin) { a.tablesInSmem = 1; smem += tables; }
    const size_t cnt = (size_t)a.muffleRows * a.nTargets;
    if (tabs && cnt <= (size_t)kQMuffleSmemMax && smem + cnt * sizeof(uint32_t) <= (size_t)maxSmemOptin) { a.muffleInSmem = 1; smem += cnt * sizeof(uint32_t); }
    if (smem > (size_t)maxSmemOptin) return cudaErrorInvalidValue;
    void (*k)(const QueryArgs, const FanDesc) = nullptr;
    if (stats) k = geomInSmem ? query_fan_kernel<true, true> : query_fan_kernel<false, true>;
    else k = geomInSmem ? query_fan_kernel<true, false> : query_fan_kernel<false, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<numCtas, kQThreads, smem, stream>>>(a, fans);
    return cudaGetLastError();
}

}  // namespace art
