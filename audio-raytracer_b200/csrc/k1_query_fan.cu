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
//   pass 0   every (hit point, goal) query is prepared (exact sqrt, four exact reciprocals) and tested against the first two
//            AABBs of its lists -- with nearest-first lists that blocks most queries;
//   rounds   the others go, WITH their prepared state (1/dir, limit, bin header: 32 B), to a per-warp survivor list in
//            global scratch (L2 resident) and are taken up again 32 at a time, one lane per survivor, for the next
//            4, 8, 16, ... AABBs; the list is compacted in place after every round, so the lanes stay full while the
//            population decays geometrically;
//   S/O      a query no AABB blocks is tested against the sphere and OBB lists; if nothing blocks it sees its goal.
// Goals are processed in chunks of at most 32 (<= 1,024 queries per warp in flight) so that the lists stay small.
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
#ifndef ART_Q_FIRST_SPAN
#define ART_Q_FIRST_SPAN 4
#endif
constexpr int kQWarps = ART_Q_WARPS;
constexpr int kQThreads = kQWarps * 32;
constexpr int kQGoalChunk = 32;                      // goals per pass
constexpr int kQListCap = kQGoalChunk * 32;          // survivors a warp can hold per list
constexpr int kQFirstTests = 2;                      // AABBs of pass 0 (their ids come with the headers, FanDesc::cells4)
constexpr int kQFirstSpan = ART_Q_FIRST_SPAN;        // AABBs of the first survivor round (doubles every round)
constexpr int kQMuffleSmemMax = 4096;                // per-CTA muffle counters [T * Na] kept in shared memory up to this size

struct QEnv {
    const QueryArgs& a;
    const FanDesc& f;
    const GeomView& gv;
    const float4* recP;          // per-warp records, shared memory: (px, py, pz, echoL)
    const float4* recQ;          //   (echoMul, rayResultId bits, batch row bits, -)
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
__device__ __forceinline__ void q_visible(const QEnv& E, int slot, float L, float echoMul, int resultId, int row)
{
    if (slot == 0) E.a.echo[resultId] = um_f32tof16(mulr(L, echoMul));
    else {
        const int idx = row * E.a.nTargets + (slot - 1);
        if (E.sMuffle) atomicAdd(&E.sMuffle[idx], 1u);
        else atomicAdd(&E.a.muffleCounts[idx], 1u);
    }
}

// One round over the survivor list: AABB entries [kBeg, kEnd) of (near list, then bin). Survivors with more AABBs left are
// written back in place (a survivor is only ever written below the entries already read), the ones whose AABB lists are
// exhausted go to the sphere / OBB list or see their goal.
template <bool STATS>
__device__ __forceinline__ int q_round_aabb(const QEnv& E, float4* listA, int nIn, int kBeg, int kEnd, float4* listSO, int& nSO)
{
    const QueryArgs& a = E.a; (void)a;
    int w = 0;
    for (int i0 = 0; i0 < nIn; i0 += 32) {
        const int i = i0 + E.lane;
        const bool on = i < nIn;
        float4 q0 = make_float4(0, 0, 0, 0), q1 = make_float4(0, 0, 0, 0);
        if (on) { q0 = listA[2 * i]; q1 = listA[2 * i + 1]; }
        __syncwarp();                                    // every entry of this step is read before any is overwritten
        bool keep = false, toSO = false;
        if (on) {
            const uint32_t packed = __float_as_uint(q1.z);
            const int slot = (int)(packed & 0xFFFFu), rc = (int)(packed >> 16);
            ART_CHECK(a.counters, rc < 32 && slot <= a.nTargets);
            const float4 rp = E.recP[rc];
            const f3 P = mk3(rp.x, rp.y, rp.z), inv = mk3(q0.x, q0.y, q0.z);
            const float L = q0.w;
            const uint4 n4 = q_near(E, slot);
            const uint32_t hBx = __float_as_uint(q1.x), hBy = __float_as_uint(q1.y);
            const int nS0 = n4.y & 1023, nA0 = (n4.y >> 10) & 2047, nS1 = hBy & 1023, nA1 = (hBy >> 10) & 2047;
            const int nAll = nA0 + nA1, kLast = min(kEnd, nAll);
            const uint16_t* eN = E.f.entries + n4.x + nS0;
            const uint16_t* eB = E.f.entries + hBx + nS1 - nA0;
            ART_CHECK(a.counters, kBeg < nAll && n4.x + nS0 + nA0 <= (unsigned)E.f.nEntries && hBx + nS1 + nA1 <= (unsigned)E.f.nEntries);
            bool blocked = false;
            int nxt = (int)__ldg((kBeg < nA0 ? eN : eB) + kBeg);
            for (int k = kBeg; k < kLast && !blocked; k++) {
                const int id = nxt;
                if (k + 1 < kLast) nxt = (int)__ldg((k + 1 < nA0 ? eN : eB) + k + 1);
                ART_CHECK(a.counters, id < a.L.na);
                if (STATS) E.st[1]++;
                blocked = aabb_blocks(E.gv, id, P, inv, L);
            }
            if (!blocked) {
                if (nAll > kEnd) keep = true;
                else if (((n4.y | hBy) & 1023u) | ((n4.y | hBy) >> 21)) toSO = true;
                else {
                    const float4 rq = E.recQ[rc];
                    q_visible(E, slot, L, rq.x, __float_as_int(rq.y), __float_as_int(rq.z));
                }
            }
        }
        const uint32_t km = __ballot_sync(kFull, keep), sm = __ballot_sync(kFull, toSO);
        if (keep) { const int pos = w + __popc(km & E.ltMask); listA[2 * pos] = q0; listA[2 * pos + 1] = q1; }
        if (toSO) {
            const int pos = nSO + __popc(sm & E.ltMask);
            ART_CHECK(a.counters, pos < kQListCap);
            listSO[2 * pos] = q0; listSO[2 * pos + 1] = q1;
        }
        w += __popc(km);
        nSO += __popc(sm);
    }
    __syncwarp();
    return w;
}

// The queries no AABB blocks: sphere lists, then OBB lists (near list, then bin). One lane per query.
template <bool STATS>
__device__ __forceinline__ void q_pass_so(const QEnv& E, const float4* listSO, int nIn)
{
    const QueryArgs& a = E.a;
    for (int i0 = 0; i0 < nIn; i0 += 32) {
        const int i = i0 + E.lane;
        if (i >= nIn) continue;
        const float4 q0 = listSO[2 * i], q1 = listSO[2 * i + 1];
        const uint32_t packed = __float_as_uint(q1.z);
        const int slot = (int)(packed & 0xFFFFu), rc = (int)(packed >> 16);
        ART_CHECK(a.counters, rc < 32 && slot <= a.nTargets);
        const float4 rp = E.recP[rc];
        const f3 P = mk3(rp.x, rp.y, rp.z);
        const float L = q0.w, len = q1.w;
        const f3 v = sub3(q_goal(E, slot), P);                           // RT:127 / RT:162
        const f3 d = smul3(rcpr(len), v);                                // normalize = rsqrt(dot) * v, len from pass 0
        const float dd = dot3(d, d);
        const uint4 n4 = q_near(E, slot);
        const uint32_t hBx = __float_as_uint(q1.x), hBy = __float_as_uint(q1.y);
        const int nS0 = n4.y & 1023, nA0 = (n4.y >> 10) & 2047, nO0 = n4.y >> 21;
        const int nS1 = hBy & 1023, nA1 = (hBy >> 10) & 2047, nO1 = hBy >> 21;
        ART_CHECK(a.counters, n4.x + nS0 + nA0 + nO0 <= (unsigned)E.f.nEntries && hBx + nS1 + nA1 + nO1 <= (unsigned)E.f.nEntries);
        bool blocked = false;
        {
            const uint16_t* eN = E.f.entries + n4.x;
            const uint16_t* eB = E.f.entries + hBx - nS0;
            const int n = nS0 + nS1;
            for (int k = 0; k < n && !blocked; k++) {
                const int id = (int)__ldg((k < nS0 ? eN : eB) + k);
                ART_CHECK(a.counters, id < a.L.ns);
                if (STATS) E.st[0]++;
                blocked = sphere_dist(E.gv, id, P, d, dd) < L;           // RT:370-377 / RT:410-419
            }
        }
        if (!blocked) {
            const uint16_t* eN = E.f.entries + n4.x + nS0 + nA0;
            const uint16_t* eB = E.f.entries + hBx + nS1 + nA1 - nO0;
            const int n = nO0 + nO1;
            for (int k = 0; k < n && !blocked; k++) {
                const int id = (int)__ldg((k < nO0 ? eN : eB) + k);
                ART_CHECK(a.counters, id < a.L.no);
                if (STATS) E.st[2]++;
                blocked = obb_blocks(E.gv, id, P, d, dd, a.errScale, L); // RT:388-394 / RT:436-445
            }
        }
        if (!blocked) {
            const float4 rq = E.recQ[rc];
            q_visible(E, slot, L, rq.x, __float_as_int(rq.y), __float_as_int(rq.z));
        }
    }
    __syncwarp();
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
    float4* recP = reinterpret_cast<float4*>(p) + warp * 32;
    float4* recQ = reinterpret_cast<float4*>(p) + (kQWarps + warp) * 32;
    p += (size_t)2 * kQWarps * 32 * sizeof(float4);
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
    const QEnv E = { a, f, gv, recP, recQ, goalTab, nearTab, sMuffle, lane, ltMask, st };
    float4* listA = a.scratch + ((size_t)blockIdx.x * kQWarps + warp) * (size_t)(4 * kQListCap);
    float4* listSO = listA + 2 * kQListCap;
    const unsigned int nRec = *a.recCount;           // the bounce tracer has finished (stream order)
    const int nChunks = (slots + kQGoalChunk - 1) / kQGoalChunk;
    const int chunk = (slots + nChunks - 1) / nChunks;

    for (;;) {
        unsigned int blk = 0;
        if (lane == 0) blk = atomicAdd(a.queue, 1u);
        blk = __shfl_sync(kFull, blk, 0);
        if ((unsigned long long)blk * 32ull >= (unsigned long long)nRec) break;
        const unsigned int ri = blk * 32u + (unsigned)lane;
        const bool valid = ri < nRec;
        f3 P = mk3(0, 0, 0);
        float echoL = 0.0f, echoMul = 0.0f;
        int resultId = 0, row = 0;
        if (valid) {
            const float4 ra = a.recA[ri];
            const float2 rb = a.recB[ri];
            P = mk3(ra.x, ra.y, ra.z); echoL = ra.w; echoMul = rb.x;
            resultId = __float_as_int(rb.y);
            ART_CHECK(a.counters, resultId >= 0 && resultId / a.H < a.map.nLocal);
            row = a.map.to_global(resultId / a.H) / a.batchSize;         // ART:161/191 batch of the ray
        }
        __syncwarp();                                                    // (the previous block's rounds have read their records)
        recP[lane] = make_float4(P.x, P.y, P.z, echoL);
        recQ[lane] = make_float4(echoMul, __int_as_float(resultId), __int_as_float(row), 0.0f);
        __syncwarp();

        for (int s0 = 0; s0 < slots; s0 += chunk) {
            const int s1 = min(slots, s0 + chunk);
            int nA = 0, nSO = 0;
            // ---- pass 0: lane = hit point, all lanes walk the goals together
            for (int s = s0; s < s1; s++) {
                bool pushA = false, pushSO = false;
                float4 q0 = make_float4(0, 0, 0, 0), q1 = make_float4(0, 0, 0, 0);
                const uint4 n4 = q_near(E, s);
                if (valid) {
                    const f3 v = sub3(q_goal(E, s), P);                            // RT:127 / RT:162
                    // the bin header depends on the direction's bin only: its load (L2) is issued first and completes
                    // while the exact square root and reciprocals below are computed
                    const int bin = fan_bin(-v.x, -v.y, -v.z);                     // direction goal -> hit point
                    uint4 c4 = make_uint4(0u, 0u, 0u, 0u);
                    if (bin >= 0) c4 = __ldg(&f.cells4[(size_t)q_fan_of(a, s) * kFanCells + bin]);
                    const float len = sqrtr(dot3(v, v));
                    float L = echoL;                                               // RT:130
                    bool gate = true;
                    if (s > 0) { L = len; gate = L < a.maxMuffle; }                // RT:165, 168
                    if (gate) {
                        if (bin < 0 || len != len) {
                            q_visible(E, s, L, echoMul, resultId, row);            // degenerate (hit point == goal): no test can block
                        } else {
                            const f3 nd = smul3(rcpr(len), v);                     // normalize = rsqrt(dot) * v
                            const f3 inv = mk3(rcpr(nd.x), rcpr(nd.y), rcpr(nd.z));
                            const int nA0 = (n4.y >> 10) & 2047, nA1 = (c4.y >> 10) & 2047, nAll = nA0 + nA1;
                            if (STATS) st[3] += 2;
                            // the run "near list, then bin" starts with these ids: no dependent load of the entry lists here
                            const uint32_t ids = nA0 >= 2 ? n4.z : (nA0 == 1 ? (n4.z & 0xFFFFu) | (c4.z << 16) : c4.z);
                            const int nFirst = min(nAll, kQFirstTests);
                            bool blocked = false;
#pragma unroll 1
                            for (int t = 0; t < nFirst && !blocked; t++) {
                                const int id = (int)((ids >> (16 * t)) & 0xFFFFu);
                                ART_CHECK(a.counters, id < a.L.na);
                                if (STATS) st[1]++;
                                blocked = aabb_blocks(gv, id, P, inv, L);
                            }
                            if (!blocked) {
                                if (nAll > kQFirstTests) pushA = true;
                                else if (((n4.y | c4.y) & 1023u) | ((n4.y | c4.y) >> 21)) pushSO = true;
                                else q_visible(E, s, L, echoMul, resultId, row);
                                q0 = make_float4(inv.x, inv.y, inv.z, L);
                                q1 = make_float4(__uint_as_float(c4.x), __uint_as_float(c4.y), __uint_as_float((uint32_t)s | ((uint32_t)lane << 16)), len);
                            }
                        }
                    }
                }
                const uint32_t am = __ballot_sync(kFull, pushA), sm = __ballot_sync(kFull, pushSO);
                if (pushA) { const int pos = nA + __popc(am & ltMask); listA[2 * pos] = q0; listA[2 * pos + 1] = q1; }
                if (pushSO) { const int pos = nSO + __popc(sm & ltMask); listSO[2 * pos] = q0; listSO[2 * pos + 1] = q1; }
                nA += __popc(am);
                nSO += __popc(sm);
            }
            __syncwarp();
            // ---- survivor rounds: the next 4, 8, 16, ... AABBs
            int kBeg = kQFirstTests, span = kQFirstSpan;
            while (nA > 0) {
                nA = q_round_aabb<STATS>(E, listA, nA, kBeg, kBeg + span, listSO, nSO);
                kBeg += span;
                if (span < 1024) span *= 2;
            }
            if (nSO > 0) q_pass_so<STATS>(E, listSO, nSO);
        }
    }

    if (sMuffle) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.muffleRows * Na; i += blockDim.x) {
            const uint32_t v = sMuffle[i];
            if (v) atomicAdd(&a.muffleCounts[i], v);
        }
    }
    if (STATS) {
        atomicAdd(&a.counters[C_GRID_RT_S], (unsigned long long)st[0]);
        atomicAdd(&a.counters[C_GRID_RT_A], (unsigned long long)st[1]);
        atomicAdd(&a.counters[C_GRID_RT_O], (unsigned long long)st[2]);
        atomicAdd(&a.counters[C_GRID_RT_CELLS], (unsigned long long)st[3]);
    }
}

// ---- launcher -----------------------------------------------------------------------------------
size_t query_fan_smem_bytes(const GeomLayout& L, bool geomInSmem) { return (geomInSmem ? L.bytes : 0) + (size_t)2 * kQWarps * 32 * sizeof(float4); }
size_t query_fan_scratch_bytes(int numCtas) { return (size_t)numCtas * kQWarps * 4 * kQListCap * sizeof(float4); }

cudaError_t launch_query_fan(const QueryArgs& a0, const FanDesc& fans, int numCtas, bool geomInSmem, bool stats, int maxSmemOptin, cudaStream_t stream)
{
    QueryArgs a = a0;
    size_t smem = query_fan_smem_bytes(a.L, geomInSmem);
    const size_t tables = (size_t)(a.nTargets + 1) * (sizeof(float4) + sizeof(uint4));
    a.tablesInSmem = 0; a.muffleInSmem = 0;
    const char* noTab = getenv("ART_K1_NO_GOAL_TABLES");         // (read per launch: test knob for the global-memory fallbacks)
    const bool tabs = !(noTab && atoi(noTab) != 0);
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
