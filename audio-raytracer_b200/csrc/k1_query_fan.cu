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
// Mapping. A warp takes 32 consecutive records; LANE = HIT POINT (its position stays in registers), the goals are walked in
// lock step, so the goal position and its near-list header are warp-uniform shared-memory broadcasts:
//   cull     every (hit point, goal) query first asks the fan for the COVERING DEPTH of its direction bin (k4_fan_build.cu):
//            beyond that depth some AABB fills the whole bin as seen from the goal, so the segment hit point -> goal runs
//            through it and the reference's test reports it, by margins a hundred times its FP32 rounding -- the query is
//            blocked (any-hit: nothing is written for it) at the price of one subtraction, the bin index, one 4-byte load and
//            two compares. About 70 % of C3's queries end here;
//   first    the others are queued with 8 B each (record, goal, bin) and, 32 at a time at full width, set up exactly as the
//            reference does (RT:127/162 normalize: exact sqrt and divide, RT:289: three exact reciprocals) and tested against
//            the first two AABBs of their lists (ids beside the bin header, FanDesc::cells4; nearest to the goal first);
//   AABBs    the queries still unblocked wait WITH their state (64 B, self-contained) in a per-warp list in global scratch
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
#define ART_Q_RUN 48     // 96 is 0.7 % faster on C3 but its queues no longer stay in L2: 5.6 GB of DRAM traffic per launch instead of 1.7
#endif
#ifndef ART_Q_MIN_LANES
#define ART_Q_MIN_LANES 20
#endif
constexpr int kQWarps = ART_Q_WARPS;
constexpr int kQThreads = kQWarps * 32;
constexpr int kQFirstTests = 2;                      // AABBs of the "first" step at most (their ids come with the headers, FanDesc::cells4)
#ifndef ART_Q_GOALS
#define ART_Q_GOALS 4
#endif
constexpr int kQGoals = ART_Q_GOALS;                 // goals per cull step
constexpr int kQCap0 = 32 + 32 * kQGoals;            // queries waiting for their set-up (uint2 each): < 32 before a step, <= 32 more per goal
constexpr int kQRun = ART_Q_RUN;                     // queued queries from which a refill loop runs
constexpr int kQMinLanes = ART_Q_MIN_LANES;          // a loop whose queue is dry stops (and writes its queries back) below this many busy lanes
constexpr int kQCapA = kQRun + 32 + 32 * kQGoals;    // AABB queue: < kQRun before a step, <= 32 more per goal of the step (q_first runs), <= 32 written back
constexpr int kQCapSO = 2 * kQRun + 64 + 64 * kQGoals;             // sphere / OBB queue: additionally everything one AABB run passes on
constexpr int kQEntry = 4;                           // float4 per queued query
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
constexpr int kQVerifyCulled = 0x40000000;          // -DART_Q_VERIFY builds: the cull would have dropped this query (travels in the result id)
__device__ __forceinline__ void q_visible(const QEnv& E, int slot, float L, float echoMul, int resultId)
{
#ifdef ART_Q_VERIFY
    if (resultId & kQVerifyCulled) { atomicAdd(&E.a.counters[C_DEBUG_VIOLATIONS], 1ull); resultId &= ~kQVerifyCulled; }
#endif
    if (slot == 0) E.a.echo[resultId] = um_f32tof16(mulr(L, echoMul));
    else {
        const int row = E.a.map.to_global(resultId / E.a.H) / E.a.batchSize;     // ART:161/191 batch of the ray
        const int idx = row * E.a.nTargets + (slot - 1);
        if (E.sMuffle) atomicAdd(&E.sMuffle[idx], 1u);
        else atomicAdd(&E.a.muffleCounts[idx], 1u);
    }
}

// Exact set-up and first tests of n <= 32 queries the cull did not decide (8 B each: record index, slot | bin << 16 | flags),
// one lane per query: RT:127 / RT:162 normalize (exact sqrt, exact divide), RT:289 the three reciprocals, RT:130 / RT:165 the
// limit, RT:168 the gate, then the first two AABBs of the run "near list, then bin" -- and on to the AABB queue, the sphere /
// OBB queue, or the goal is visible.
constexpr uint32_t kQFlagNoBin = 1u << 30;           // the direction goal -> hit point has no bin (hit point == goal, non-finite)
constexpr uint32_t kQFlagCulled = 1u << 29;          // -DART_Q_VERIFY builds only
template <bool STATS>
__device__ __forceinline__ void q_first(const QEnv& E, const uint2* src, int n, float4* listA, int& nA, float4* listSO, int& nSO)
{
    const QueryArgs& a = E.a;
    int push = 0;                                                        // 1: AABB queue, 2: sphere / OBB queue
    float4 e0 = make_float4(0, 0, 0, 0), e1 = e0, e2 = e0, e3 = e0;
    if (E.lane < n) {
        const uint2 q = src[E.lane];
        const int s = (int)(q.y & 0xFFFFu);
        ART_CHECK(a.counters, s <= a.nTargets && q.x < *a.recCount);
        // the bin header: issued first, it arrives while the exact square root and reciprocals are computed
        uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
        if (!(q.y & kQFlagNoBin)) c4 = __ldg(&E.f.cells4[(size_t)q_fan_of(a, s) * kFanCells + ((q.y >> 16) & 0x1FFFu)]);
        const float4 ra = __ldg(&a.recA[q.x]);
        float2 rb = __ldg(&a.recB[q.x]);
#ifdef ART_Q_VERIFY
        if (q.y & kQFlagCulled) rb.y = __int_as_float(__float_as_int(rb.y) | kQVerifyCulled);
#endif
        const uint4 n4 = q_near(E, s);
        const f3 P = mk3(ra.x, ra.y, ra.z);
        const f3 v = sub3(q_goal(E, s), P);                                // RT:127 / RT:162
        const float len = sqrtr(dot3(v, v));
        float L = ra.w;                                                    // RT:130
        bool gate = true;
        if (s > 0) { L = len; gate = L < a.maxMuffle; }                    // RT:165, 168
        if (gate) {
            if ((q.y & kQFlagNoBin) || len != len) {
                q_visible(E, s, L, rb.x, __float_as_int(rb.y));            // degenerate (hit point == goal): no test can block
            } else {
                const f3 nd = smul3(rcpr(len), v);                         // normalize = rsqrt(dot) * v
                const f3 inv = mk3(rcpr(nd.x), rcpr(nd.y), rcpr(nd.z));
                const int nA0 = (n4.y >> 10) & 2047, nA1 = (c4.y >> 10) & 2047, nAll = nA0 + nA1;
                if (STATS) E.st[3] += 2;
                // the run "near list, then bin" starts with these ids: no dependent load of the entry lists here
                const uint32_t ids = nA0 >= 2 ? n4.z : (nA0 == 1 ? (n4.z & 0xFFFFu) | (c4.z << 16) : c4.z);
                const int nFirst = min(nAll, a.firstTests);
                bool blocked = false;
#pragma unroll 1
                for (int t = 0; t < nFirst && !blocked; t++) {
                    const int id = (int)((ids >> (16 * t)) & 0xFFFFu);
                    ART_CHECK(a.counters, id < a.L.na);
                    if (STATS) E.st[1]++;
                    blocked = aabb_blocks(E.gv, id, P, inv, L);
                }
                if (!blocked) {
                    if (nAll > nFirst) push = 1;
                    else if (((n4.y | c4.y) & 1023u) | ((n4.y | c4.y) >> 21)) push = 2;
                    else q_visible(E, s, L, rb.x, __float_as_int(rb.y));
                    e0 = make_float4(inv.x, inv.y, inv.z, L);
                    e1 = make_float4(P.x, P.y, P.z, len);
                    e2 = make_float4(__uint_as_float(c4.x), __uint_as_float(c4.y),
                                     __uint_as_float((uint32_t)s | (push == 1 ? (uint32_t)nFirst << 16 : 0u)), 0.0f);
                    e3 = make_float4(rb.x, rb.y, 0.0f, 0.0f);
                }
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
                d = smul3(rcpr(e1.w), v);                                    // normalize = rsqrt(dot) * v, |v| from the set-up (q_first)
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
    float4* listA = a.scratch + ((size_t)blockIdx.x * kQWarps + warp) * (size_t)(kQEntry * (kQCapA + kQCapSO) + kQCap0 / 2);
    float4* listSO = listA + kQEntry * kQCapA;
    uint2* list0 = reinterpret_cast<uint2*>(listSO + kQEntry * kQCapSO);
    int n0 = 0, nA = 0, nSO = 0;                     // queued queries (kept across goals and record blocks)
    const unsigned int nRec = *a.recCount;           // the bounce tracer has finished (stream order)
    const uint32_t* coverCodes = reinterpret_cast<const uint32_t*>(f.cells4) + 3;  // cells4[i].w

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
        // ---- cull: lane = hit point, all lanes walk the goals together -- kQGoals goals per step, so that as many
        //      covering-depth loads (L2) are in flight per warp
        for (int s = sBeg; s < sEnd; s += kQGoals) {
            bool push[kQGoals];
            uint32_t qy[kQGoals];
#pragma unroll
            for (int u = 0; u < kQGoals; u++) {
                push[u] = false; qy[u] = 0u;
                if (valid && s + u < sEnd) {
                    const f3 g = q_goal(E, s + u);
                    float w;
                    int sub;
                    const int bin = fan_bin_w(P.x - g.x, P.y - g.y, P.z - g.z, w, sub);   // direction goal -> hit point: bin, depth on the face, sub-bin
                    push[u] = true;
                    qy[u] = (uint32_t)(s + u) | kQFlagNoBin;
                    if (bin >= 0) {
                        // beyond the covering depth of its sub-bin an AABB certainly blocks the query (k4_fan_build.cu; compared in
                        // the log domain, one code byte per sub-bin; w <= errScale is the range the margins were derived for)
                        const uint32_t codes = __ldg(coverCodes + 4 * ((size_t)q_fan_of(a, s + u) * kFanCells + bin));
                        const uint32_t code = (codes >> (8 * sub)) & 255u;
                        const bool culled = fmaf(__log2f(w), f.coverLogS, f.coverLogK) > (float)code + kFanCoverLogEps && code < 255u && w <= a.errScale;
                        qy[u] = (uint32_t)(s + u) | ((uint32_t)bin << 16);
#ifdef ART_Q_VERIFY
                        if (culled) qy[u] |= kQFlagCulled;
#else
                        push[u] = !culled;
#endif
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kQGoals; u++) {
                const uint32_t pm = __ballot_sync(kFull, push[u]);
                if (push[u]) list0[n0 + __popc(pm & ltMask)] = make_uint2(ri, qy[u]);
                n0 += __popc(pm);
            }
            // ---- exact set-up + first tests of the survivors, their remaining AABBs, their sphere and OBB lists -- each
            //      whenever enough queries have gathered to fill the lanes
#pragma unroll 1
            while (n0 >= 32) { __syncwarp(); n0 -= 32; q_first<STATS>(E, list0 + n0, 32, listA, nA, listSO, nSO); }
            if (nA >= kQRun) { __syncwarp(); nA = q_loop_aabb<STATS>(E, listA, nA, false, listSO, nSO); }
            if (nSO >= kQRun) { __syncwarp(); nSO = q_loop_so<STATS>(E, listSO, nSO, false); }
        }
    }
    __syncwarp();
    if (n0 > 0) q_first<STATS>(E, list0, n0, listA, nA, listSO, nSO);
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
size_t query_fan_scratch_bytes(int numCtas) { return (size_t)numCtas * kQWarps * (kQEntry * (kQCapA + kQCapSO) + kQCap0 / 2) * sizeof(float4); }

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
    // experiment knob (read per launch): AABBs of the first tests, q_first (1 or 2)
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
