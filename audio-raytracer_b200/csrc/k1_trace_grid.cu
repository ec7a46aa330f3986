// k1_trace_grid.cu -- K1 with the uniform-grid acceleration structure: the per-ray
// bounce loop of AudioRaytracerJobBatched.Execute (Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs:61-215).
//
// Mapping: ONE THREAD OWNS ONE RAY, one warp owns 32 rays (refilled from a global queue as rays die).
// Each round every live lane walks its segment through the grid (3D-DDA) to the nearest hit
// (ShootRayCast, RT:225-280); the warp then evaluates the echo-return ray (RT:124-145) and the Na
// muffle rays (RT:153-173) of all 32 hit points as ONE POOL of (ray, target) queries that lanes take
// from dynamically, one grid cell per step, so a lane whose query is blocked early immediately starts
// the next one. Finally every lane reflects its own ray (RT:456-532).
//
// Results are bit-identical with the brute-force kernel (k1_trace.cu) and the oracle: the per-collider
// tests are the same un-fused IEEE FP32 functions in the reference's operation order (intersect.cuh),
// the nearest hit is the lexicographic (t, sphere<AABB<OBB, index) minimum -- "strict <, first in
// canonical order wins" (RT:244, 257, 270) -- and an occlusion query is an "any"; the grid only skips
// colliders whose conservative bounds the ray does not come near (grid_host.h).
#include <cstdlib>

#include "device_util.cuh"
#include "grid_dev.cuh"
#include "intersect.cuh"
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {

#ifndef ART_GRID_WARPS
#define ART_GRID_WARPS 24
#endif
constexpr int kGridWarps = ART_GRID_WARPS;
constexpr int kGridThreads = kGridWarps * 32;
constexpr int kQueryWords = 16;          // per prepared query in the per-warp ring (3 x float4 + uint2, padded)
constexpr uint32_t kNoHit = 0xFFFFFFFFu;
constexpr int kRotateSlotBytes = 64 * sizeof(float4);   // one parked group: 32 lanes x 2 float4
#ifndef ART_CAP_A
#define ART_CAP_A 4
#endif
#ifndef ART_CAP_S
#define ART_CAP_S 1
#endif
#ifndef ART_CAP_O
#define ART_CAP_O 2
#endif
constexpr int kCapA = ART_CAP_A, kCapS = ART_CAP_S, kCapO = ART_CAP_O;   // tests per type per step of an occlusion query

// nearest-hit: exact distance, or NaN when the collider misses or certainly lies beyond `best`
__device__ __forceinline__ float obb_dist_nearest(const GeomView& gv, int id, f3 o, f3 d, float dd, float errScale, float best)
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
struct HitRec {            // per hit point, shared memory (one per lane)
    float px, py, pz;      // Pp = hit - eps*d   (RT:124 == RT:158)
    float echoL;           // distance(RayOrigin, hit)  (RT:130)
    float echoMul;         // material Echo of the hit collider (RT:135-141)
    int resultId;          // rayResultId (RT:115), local indexing
    int row;               // batch index of the ray
    int pad;
};


// ---- pooled occlusion queries -----------------------------------------------------------------------
constexpr int kChunkQ = 4096;            // queries per pool pass (bounds the survivor list: 16 KB of scratch per warp)
constexpr int kSkipCells = 4;
constexpr int kTwoStageSlots = 16;       // queries per hit point from which the AABB-first two-stage pool pays off            // empty cells a lane may step over in one pool step

struct PoolEnv {
    const TraceArgs& a;
    const GridDesc& g;
    const GeomView& gv;
    const HitRec* rec;                   // per-warp hit records (shared memory)
    float4* qbuf0; float4* qbuf1; float4* qbuf2; int* qbuf3;   // per-warp ring of prepared queries (shared memory)
    uint32_t* surv;                      // per-warp survivor list (global scratch): slot | rec << 16
    f3 RayOrigin;
    uint32_t hitMask;
    int slots, lane;
    uint32_t ltMask;
    unsigned int* st;                    // STATS: per-lane counters {sphere, AABB, OBB tests, cells}
    const float4* goalByPos;             // shared memory (or null): goal (xyz) and slot (w) of pool position p, [slots]
    const float4* goalBySlot;            //   ... and goal (xyz) of slot s, [slots]; slot 0 = the listener (echo ray)
};

__device__ __forceinline__ void query_visible(const PoolEnv& E, int slot, int recIdx, float L)
{
    const HitRec r = E.rec[recIdx];
    if (slot == 0) E.a.echo[r.resultId] = um_f32tof16(mulr(L, r.echoMul));             // RT:133-145
    else atomicAdd(&E.a.muffleCounts[r.row * E.a.nTargets + (slot - 1)], 1u);          // RT:168-172
}

// One pass over `count` queries. Queries are PREPARED 32 at a time by the whole warp (direction, exact
// reciprocals, DDA start) into the per-warp ring and CONSUMED by whichever lanes are idle, a bounded slice of
// one grid cell per step, so the lanes of a warp stay in step while early exits are taken at once.
//   STAGE 0: every (hit point, slot) query of the round, tested against the AABB lists only (the cheapest and
//            most numerous blockers). A query the AABBs do not block is appended to the survivor list.
//   STAGE 1: the survivors, against the sphere and OBB lists. A query that survives this too sees its goal.
//   STAGE 2: (few queries per hit point) all three lists in one pass.
// An occlusion query is an "any" over all colliders (RT:365-449), so the order of the tests is free.
template <int STAGE, bool STATS>
__device__ __forceinline__ int run_pool(const PoolEnv& E, int qFirst, int count, bool noSO)
{
    const TraceArgs& a = E.a;
    const GridDesc& g = E.g;
    const GeomView& gv = E.gv;
    const int lane = E.lane;
    const uint32_t ltMask = E.ltMask;
    const uint16_t* const ebase = g.entries;
    int nextQ = 0, bufNext = 0, bufCount = 0, survCount = 0;
    bool have = false;
    f3 qo = mk3(0, 0, 0), qd = mk3(0, 0, 0), qinv = mk3(0, 0, 0);
    float qL = 0.0f, qdd = 0.0f;
    int qslot = 0, qrec = 0;
    bool fresh = true;                // no cell fetched yet for this query
    uint2 hdr = make_uint2(0, 0);
    int kA = 0, kB = 0, kC = 0;       // cursors inside the current cell: AABB, sphere, OBB lists
    Dda w;
    w.ix = w.iy = w.iz = 0; w.tmx = w.tmy = w.tmz = 0; w.tdx = w.tdy = w.tdz = 0; w.tEnd = 0; w.tCur = 0; w.lastAxis = -1;
    for (;;) {
        const uint32_t idle = __ballot_sync(kFull, !have);
        if (idle) {
            if (bufNext == bufCount && nextQ < count) {
                // ---- prepare the next 32 queries (all lanes)
                const int qi = nextQ + lane;
                nextQ += 32;
                bool active = false;
                f3 nd = mk3(0, 0, 0), ninv = mk3(0, 0, 0);
                float nL = 0.0f;
                int nslot = 0, nrec = 0;
                Dda nw;
                nw.ix = nw.iy = nw.iz = 0; nw.tmx = nw.tmy = nw.tmz = 0; nw.tdx = nw.tdy = nw.tdz = 0; nw.tEnd = 0; nw.tCur = 0; nw.lastAxis = -1;
                if (qi < count) {
                    if (STAGE != 1) {
                        const int q = qFirst + qi;
                        const int ord = q / E.slots;
                        nslot = q - ord * E.slots;
                        if (E.goalByPos) nslot = __float_as_int(E.goalByPos[nslot].w);
                        else if (nslot > 0) nslot = a.targetOrder[nslot - 1] + 1;  // spatially sorted pool order
                        nrec = __fns(E.hitMask, 0, ord + 1);
                    } else {
                        const uint32_t packed = E.surv[qi];   // STAGE 1
                        nslot = (int)(packed & 0xFFFFu);
                        nrec = (int)(packed >> 16);
                    }
                    ART_CHECK(a.counters, (unsigned)nrec < 32u && nslot >= 0 && nslot <= a.nTargets);
                    const HitRec r = E.rec[nrec];
                    const f3 no = mk3(r.px, r.py, r.pz);
                    f3 T = E.RayOrigin;
                    if (E.goalBySlot) { const float4 gT = E.goalBySlot[nslot]; T = mk3(gT.x, gT.y, gT.z); }
                    else if (nslot > 0) T = mk3(a.targets[3 * (nslot - 1)], a.targets[3 * (nslot - 1) + 1], a.targets[3 * (nslot - 1) + 2]);
                    const f3 v = sub3(T, no);                                      // RT:127 / RT:162
                    const float len = sqrtr(dot3(v, v));
                    nd = smul3(rcpr(len), v);                                      // normalize = rsqrt(dot) * v
                    bool gate = true;
                    if (nslot == 0) nL = r.echoL;                                  // RT:130
                    else { nL = len; gate = nL < a.maxMuffle; }                    // RT:165, 168
                    if (gate) {
                        ninv = mk3(rcpr(nd.x), rcpr(nd.y), rcpr(nd.z));
                        active = dda_init(g, no, nd, ninv, nL, nw);
                        if (!active) query_visible(E, nslot, nrec, nL);            // nothing near the segment
                    }
                }
                const uint32_t act = __ballot_sync(kFull, active);
                if (active) {
                    const int pos = __popc(act & ltMask);
                    E.qbuf0[pos] = make_float4(nd.x, nd.y, nd.z, nL);
                    E.qbuf1[pos] = make_float4(ninv.x, ninv.y, ninv.z, nw.tEnd);
                    E.qbuf2[pos] = make_float4(nw.tmx, nw.tmy, nw.tmz, __int_as_float(nw.ix | (nw.iy << 8) | (nw.iz << 16) | (nrec << 24)));
                    E.qbuf3[pos] = nslot;
                }
                bufNext = 0;
                bufCount = __popc(act);
                __syncwarp();
            }
            if (bufNext < bufCount) {
                // ---- idle lanes take prepared queries
                const int pos = bufNext + __popc(idle & ltMask);
                if (!have && pos < bufCount) {
                    const float4 v0 = E.qbuf0[pos], v1 = E.qbuf1[pos], v2 = E.qbuf2[pos];
                    qslot = E.qbuf3[pos];
                    qd = mk3(v0.x, v0.y, v0.z); qL = v0.w;
                    qinv = mk3(v1.x, v1.y, v1.z); w.tEnd = v1.w;
                    const int packed = __float_as_int(v2.w);
                    qrec = (packed >> 24) & 31;
                    w.tmx = v2.x; w.tmy = v2.y; w.tmz = v2.z;
                    w.ix = packed & 255; w.iy = (packed >> 8) & 255; w.iz = (packed >> 16) & 255;
                    w.tdx = fabsf(g.csx * qinv.x); w.tdy = fabsf(g.csy * qinv.y); w.tdz = fabsf(g.csz * qinv.z);
                    qdd = dot3(qd, qd);
                    const HitRec r = E.rec[qrec];
                    qo = mk3(r.px, r.py, r.pz);
                    fresh = true; kA = kB = kC = 0; hdr = make_uint2(0, 0);
                    have = true;
                }
                bufNext = min(bufCount, bufNext + __popc(idle));
                __syncwarp();
            }
        }
        if (!__any_sync(kFull, have)) {
            if (bufNext == bufCount && nextQ >= count) break;
            continue;
        }
        bool survived = false;
        if (have && STAGE == 0) {   // (more than one slice per step was measured slower: 43 -> 45..70 ms on C3)
            // ---- a slice of up to kCapA AABB entries of the current cell (stepping over at most kSkipCells empty cells first)
            bool walkDone = false;
            for (int s = 0; s < kSkipCells; s++) {
                if (kA < (int)((hdr.y >> 10) & 2047)) break;
                if (!fresh) {
                    const float tNext = dda_next_t(w);
                    if (tNext > w.tEnd || !dda_step(g, qd, w)) { walkDone = true; break; }
                }
                fresh = false;
                hdr = dda_cell(g, w);
                if (STATS) E.st[3]++;
                kA = 0;
            }
            bool blocked = false;
            if (!walkDone) {
                const int nS = hdr.y & 1023, nA = (hdr.y >> 10) & 2047;
                const uint16_t* e = ebase + hdr.x + nS;
                const int ownerId = qslot - 1;         // -1 for the echo ray: never equals a valid owner below
                for (int c = 0; c < kCapA && kA < nA && !blocked; c++, kA++) {
                    const int id = __ldg(e + kA);
                    ART_CHECK(a.counters, id < a.L.na && hdr.x + nS + nA <= (unsigned)g.nEntries);
                    if (STATS) E.st[1]++;
                    if (aabb_blocks(gv, id, qo, qinv, qL))
                        blocked = !(qslot > 0 && a.anyOwned[1] && (int)a.at.ownA[id] == ownerId);   // RT:426
                }
            }
            if (blocked) {
                have = false;
            } else if (walkDone) {
                have = false;
                if (noSO) query_visible(E, qslot, qrec, qL);
                else survived = true;
            }
        }
        if (have && STAGE != 0) {
            // ---- a bounded slice of the current cell's lists (STAGE 1: spheres + OBBs, STAGE 2: all three types)
            bool walkDone = false;
            for (int s = 0; s < kSkipCells; s++) {
                const int nS = hdr.y & 1023, nA = (hdr.y >> 10) & 2047, nO = hdr.y >> 21;
                const bool pending = kB < nS || kC < nO || (STAGE == 2 && kA < nA);
                if (pending) break;
                if (!fresh) {
                    const float tNext = dda_next_t(w);
                    if (tNext > w.tEnd || !dda_step(g, qd, w)) { walkDone = true; break; }
                }
                fresh = false;
                hdr = dda_cell(g, w);
                if (STATS) E.st[3]++;
                kA = kB = kC = 0;
            }
            bool blocked = false;
            if (!walkDone) {
                const uint16_t* e = ebase + hdr.x;
                const int nS = hdr.y & 1023, nA = (hdr.y >> 10) & 2047, nO = hdr.y >> 21;
                const int ownerId = qslot - 1;
                if (STAGE == 2) {
                    for (int c = 0; c < kCapA && kA < nA && !blocked; c++, kA++) {
                        const int id = __ldg(e + nS + kA);
                        ART_CHECK(a.counters, id < a.L.na);
                        if (STATS) E.st[1]++;
                        if (aabb_blocks(gv, id, qo, qinv, qL))
                            blocked = !(qslot > 0 && a.anyOwned[1] && (int)a.at.ownA[id] == ownerId);   // RT:426
                    }
                }
                for (int c = 0; c < kCapS && kB < nS && !blocked; c++, kB++) {
                    const int id = __ldg(e + kB);
                    ART_CHECK(a.counters, id < a.L.ns);
                    if (STATS) E.st[0]++;
                    if (sphere_dist(gv, id, qo, qd, qdd) < qL)
                        blocked = !(qslot > 0 && a.anyOwned[0] && (int)a.at.ownS[id] == ownerId);       // RT:413
                }
                for (int c = 0; c < kCapO && kC < nO && !blocked; c++, kC++) {
                    const int id = __ldg(e + nS + nA + kC);
                    ART_CHECK(a.counters, id < a.L.no && hdr.x + nS + nA + nO <= (unsigned)g.nEntries);
                    if (STATS) E.st[2]++;
                    if (obb_blocks(gv, id, qo, qd, qdd, g.errScale, qL))
                        blocked = !(qslot > 0 && a.anyOwned[2] && (int)a.at.ownO[id] == ownerId);       // RT:439
                }
            }
            if (blocked) {
                have = false;
            } else if (walkDone) {
                have = false;
                query_visible(E, qslot, qrec, qL);     // walked the whole segment: the ray sees its goal
            }
        }
        if (STAGE == 0) {
            const uint32_t push = __ballot_sync(kFull, survived);
            if (push) {
                ART_CHECK(a.counters, survCount + __popc(push) <= kChunkQ);
                if (survived) E.surv[survCount + __popc(push & ltMask)] = (uint32_t)qslot | ((uint32_t)qrec << 16);
                survCount += __popc(push);
            }
        }
    }
    __syncwarp();
    return survCount;
}

// This kernel runs bounce rays AND occlusion queries, all on the grid walk; it is the path of frames without target fans
// (ART_FRAME_NO_FANS, fan overflow re-runs). With fans the frame runs bounce_kernel (k1_bounce.cu) + query_fan_kernel instead.
// ROT: group rotation (TraceArgs::migGroups, trace_grid_rotation) instead of the per-lane ray queue
template <bool SMEM, bool STATS, bool ROT>
__global__ void __launch_bounds__(kGridThreads, 1) trace_grid_kernel(const TraceArgs a, const GridDesc g)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Na = a.nTargets;
    const int slots = Na + 1;                    // slot 0 = echo ray, slot 1+t = muffle ray to target t

    unsigned char* p = smem;
    const unsigned char* geomBase = a.geom;
    if (SMEM) {
        stage_blob_to_smem(p, a.geom, a.L.bytes, &bar);
        geomBase = p;
        p += a.L.bytes;
    }
    HitRec* rec = reinterpret_cast<HitRec*>(p) + warp * 32;
    p += (size_t)kGridWarps * 32 * sizeof(HitRec);
    float4* qbase = reinterpret_cast<float4*>(p);
    float4* qbuf0 = qbase + warp * 32;                                // (dir.xyz, limit)
    float4* qbuf1 = qbase + (kGridWarps + warp) * 32;                 // (1/dir.xyz, tEnd)
    float4* qbuf2 = qbase + (2 * kGridWarps + warp) * 32;             // (tMax.xyz, cell | rec << 24)
    int* qbuf3 = reinterpret_cast<int*>(qbase + 3 * kGridWarps * 32) + warp * 32;   // slot
    // the goals of the pool (listener + targets, in pool order and by slot) in shared memory, so that a query starts with
    // one LDS instead of the dependent global loads targetOrder -> targets
    const float4* goalByPos = nullptr; const float4* goalBySlot = nullptr;
    if (a.goalsInSmem) {
        float4* gp = reinterpret_cast<float4*>(smem + a.goalsSmemOffset);
        for (int i = threadIdx.x; i < slots; i += blockDim.x) {
            const int s = i == 0 ? 0 : a.targetOrder[i - 1] + 1;
            gp[i] = s == 0 ? make_float4(a.ox, a.oy, a.oz, __int_as_float(0))
                           : make_float4(a.targets[3 * (s - 1)], a.targets[3 * (s - 1) + 1], a.targets[3 * (s - 1) + 2], __int_as_float(s));
            gp[slots + i] = i == 0 ? make_float4(a.ox, a.oy, a.oz, 0.0f)
                                   : make_float4(a.targets[3 * (i - 1)], a.targets[3 * (i - 1) + 1], a.targets[3 * (i - 1) + 2], 0.0f);
        }
        __syncthreads();
        goalByPos = gp; goalBySlot = gp + slots;
    }
    const GeomView gv = make_view(geomBase, a.L);
    const f3 RayOrigin = mk3(a.ox, a.oy, a.oz);
    const uint32_t ltMask = (1u << lane) - 1u;
    uint32_t* surv = a.scratch + ((size_t)blockIdx.x * kGridWarps + warp) * kChunkQ;

    // ---- per-lane ray state
    bool hasRay = false, queueEmpty = false;
    int j = 0, hits = 0, row = 0;
    f3 o = mk3(0, 0, 0), d = mk3(0, 0, 0);
    float life = 0.0f;
    unsigned int nSegments = 0, nSegHits = 0;
    unsigned int st[4] = { 0, 0, 0, 0 };          // STATS: sphere / AABB / OBB tests, cells visited (this lane)

    constexpr bool rotate = ROT;
    for (;;) {
        if (rotate) {
            // ================= group rotation: take the oldest waiting group (fresh groups first) =================
            // nextRay[0] = pop tickets, [1] = push tickets, [2] = finished groups. Tickets below migGroups are the fresh
            // groups; ticket migGroups + s is the s-th parked state of the log, valid once migFlags[s] is set.
            unsigned int h = 0;
            if (lane == 0) h = atomicAdd(&a.nextRay[0], 1u);
            h = __shfl_sync(kFull, h, 0);
            hasRay = false;
            if (h < (unsigned)a.migGroups) {
                j = (int)h * 32 + lane;
                if (j < a.map.nLocal) {
                    const int rayIndex = a.map.to_global(j);
                    row = rayIndex / a.batchSize;                              // ART:161/191 batch k
                    d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                            um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));    // RT:94
                    o = RayOrigin;                                             // RT:95
                    hits = 0;                                                  // RT:97
                    life = a.maxRayLife;                                       // RT:99
                    hasRay = true;
                }
            } else {
                const unsigned int slot = h - (unsigned)a.migGroups;
                int ok = slot < a.migSlots ? 1 : 0;
                if (ok && lane == 0) {
                    const volatile unsigned int* fl = a.migFlags + slot;
                    const volatile unsigned int* fin = a.nextRay + 2;
                    unsigned int spins = 0;
                    while (*fl == 0u) {
                        if (*fin >= (unsigned)a.migGroups) { ok = 0; break; }  // every group has finished: nothing will be parked any more
                        __nanosleep(256);
                        if (++spins > (1u << 22)) {                            // (seconds: cannot happen; never hang the device)
                            atomicAdd(&a.counters[C_DEBUG_VIOLATIONS], 1ull);
                            ok = 0;
                            break;
                        }
                    }
                    __threadfence();
                }
                ok = __shfl_sync(kFull, ok, 0);
                if (!ok) break;
                const float4* sp = a.migState + (size_t)slot * 64;
                const float4 s0 = __ldcg(sp + lane), s1 = __ldcg(sp + 32 + lane);
                const unsigned int packed = __float_as_uint(s1.w);
                o = mk3(s0.x, s0.y, s0.z); life = s0.w;
                d = mk3(s1.x, s1.y, s1.z);
                hits = (int)(packed & 255u);
                hasRay = ((packed >> 8) & 1u) != 0;
                j = (int)(packed >> 9) * 32 + lane;
                row = hasRay ? a.map.to_global(j) / a.batchSize : 0;
            }
        } else {
        // ================= refill dead lanes from the ray queue =================
        const uint32_t dead = __ballot_sync(kFull, !hasRay && lane < a.raysPerWarp);
        if (dead && !queueEmpty) {
            int base = 0;
            if (lane == 0) base = (int)atomicAdd(a.nextRay, (unsigned)__popc(dead));
            base = __shfl_sync(kFull, base, 0);
            if ((dead >> lane) & 1u) {
                const int jj = base + __popc(dead & ltMask);
                if (jj < a.map.nLocal) {
                    j = jj;
                    const int rayIndex = a.map.to_global(j);
                    row = rayIndex / a.batchSize;                              // ART:161/191 batch k
                    d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                            um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));    // RT:94
                    o = RayOrigin;                                             // RT:95
                    hits = 0;                                                  // RT:97
                    life = a.maxRayLife;                                       // RT:99
                    hasRay = true;
                }
            }
            if (base + __popc(dead) >= a.map.nLocal) queueEmpty = true;
        }
        }
        if (!rotate && !__any_sync(kFull, hasRay)) break;

        // ================= ShootRayCast (RT:225-280), one lane = one ray =================
        float best = kFloatMax;
        uint32_t bkey = kNoHit;            // (typeOrder << 28) | index ; typeOrder sphere 0, AABB 1, OBB 2
        if (hasRay) {
            nSegments++;
            const float dd = dot3(d, d);                                       // RT:326
            const f3 inv = mk3(rcpr(d.x), rcpr(d.y), rcpr(d.z));               // RT:289
            Dda w;
            bool walking = dda_init(g, o, d, inv, pos_inf(), w);
            while (walking) {
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
                    const float dist = obb_dist_nearest(gv, id, o, d, dd, g.errScale, best);
                    const uint32_t key = (2u << 28) | (uint32_t)id;
                    if (dist < best || (dist == best && key < bkey)) { best = dist; bkey = key; }
                }
                // colliders listed only in later cells lie beyond the next cell boundary
                const float tNext = dda_next_t(w);
                if (tNext > w.tEnd || tNext > best) break;
                walking = dda_step(g, d, w);
            }
        }
        const bool hit = hasRay && bkey != kNoHit;
        int hitType = 0, hitIdx = 0;
        float4 attr = make_float4(0, 0, 0, 0);
        if (hasRay && !hit) {                                                  // RT:201-207 ray left the scene
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
            const f3 Pp = sub3(o, mul3s(d, kEpsilon));                         // RT:124 == RT:158
            const f3 wv = sub3(o, RayOrigin);                                  // RT:130
            HitRec r;
            r.px = Pp.x; r.py = Pp.y; r.pz = Pp.z;
            r.echoL = sqrtr(dot3(wv, wv));
            r.echoMul = attr.y;
            r.resultId = (int)rayResultId;
            r.row = row;
            r.pad = 0;
            rec[lane] = r;
        }
        __syncwarp();

        // ================= echo + muffle queries of all hit points (RT:121-175) =================
        {
            const uint32_t hitMask = __ballot_sync(kFull, hit);
            const int total = __popc(hitMask) * slots;
            const PoolEnv E = { a, g, gv, rec, qbuf0, qbuf1, qbuf2, qbuf3, surv, RayOrigin, hitMask, slots, lane, ltMask, st, goalByPos, goalBySlot };
            const bool noSO = a.L.ns + a.L.no == 0;
            for (int q0 = 0; q0 < total; q0 += kChunkQ) {
                if (slots >= kTwoStageSlots) {
                    const int nSurv = run_pool<0, STATS>(E, q0, min(kChunkQ, total - q0), noSO);
                    if (nSurv > 0) run_pool<1, STATS>(E, 0, nSurv, noSO);
                } else {
                    run_pool<2, STATS>(E, q0, min(kChunkQ, total - q0), noSO);   // few queries per hit point: one pass over all types
                }
            }
            __syncwarp();
        }

        // ================= termination / reflection (RT:178-193) =================
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
        if (rotate) {
            // ================= park the group for whichever warp comes next, or retire it =================
            if (__any_sync(kFull, hasRay)) {
                unsigned int t = 0;
                if (lane == 0) t = atomicAdd(&a.nextRay[1], 1u);
                t = __shfl_sync(kFull, t, 0);
                ART_CHECK(a.counters, t < a.migSlots);
                float4* sp = a.migState + (size_t)t * 64;
                const unsigned int packed = (unsigned)hits | (hasRay ? 256u : 0u) | ((unsigned)(j >> 5) << 9);
                __stcg(sp + lane, make_float4(o.x, o.y, o.z, life));
                __stcg(sp + 32 + lane, make_float4(d.x, d.y, d.z, __uint_as_float(packed)));
                __threadfence();
                __syncwarp();
                if (lane == 0) { __threadfence(); atomicExch(&a.migFlags[t], 1u); }
            } else if (lane == 0) {
                __threadfence();
                atomicAdd(&a.nextRay[2], 1u);
            }
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

// ---- launcher -----------------------------------------------------------------------------------
// Launch shape of a batch of n local rays: warps per CTA and lanes per warp that own a ray. With R = 32 a batch takes
// k = ceil(n / (32 * warps)) rounds per lane and the last round is mostly empty when n is small (one shard of a
// ray-sharded frame); R = ceil(n / (k * warps)) fills all k rounds evenly instead. At least ~256 queries are kept in a
// round's pool. The warp count is always the compiled maximum: the kernel is latency bound, and trading warps for fuller
// lanes loses at every batch size (B200, C3 shard of 131,072 rays: 24 warps x 19 lanes 5.53 ms, 14 warps x 32 lanes
// 6.27 ms; 1M rays: 24 warps 35.5 ms, 20 warps 38.4 ms, 16 warps 42.7 ms -- tools/exp_k1_warps.py). ART_K1_WARPS (read
// per frame) forces a smaller count for that experiment.
void trace_grid_plan(int nLocal, int nTargets, int numCtas, int* warpsOut, int* raysPerWarpOut)
{
    const char* fv = getenv("ART_K1_WARPS");
    const int forced = fv ? atoi(fv) : 0;
    const int w = forced >= 1 && forced <= kGridWarps ? forced : kGridWarps;
    const long long warps = (long long)numCtas * w;
    const long long k = (nLocal + warps * 32 - 1) / (warps * 32);
    long long r = k > 0 ? (nLocal + warps * k - 1) / (warps * k) : 32;
    const long long minPool = (256 + nTargets) / (nTargets + 1);
    if (r < minPool) r = minPool;
    *warpsOut = w;
    *raysPerWarpOut = (int)(r < 1 ? 1 : (r > 32 ? 32 : r));
}

// Group rotation (TraceArgs::migGroups). Without it a warp keeps its rays for all H bounces, so 4,096 groups of 32 rays on
// 3,552 warps take 2 x H rounds on the critical path (however the lanes are filled); rotating the groups through the
// warps after every round takes 4,096 / 3,552 x H. That is what a shard of a ray-sharded frame looks like (B200, C3 / 8:
// 5.55 -> 4.86 ms), and it still pays at full size (C3: 35.5 -> 35.1 ms) -- but a rotating group is never refilled, so
// where many rays die early its lanes thin out while the ray queue would keep them busy. Hence: always for batches of at
// most 4 groups per warp (where the queue's last rounds are badly filled anyway), and for larger ones only when the
// caller saw rays living nearly all H bounces in the previous frame of the context (`fullLivedRays`).
// Returns the number of groups (0 = off) and the capacity of the log of parked states (a group is parked at most H - 1
// times; 1 KB each, capped at 1 GiB). ART_K1_ROTATE=0 disables it, ART_K1_ROTATE_MAX_K overrides the groups-per-warp
// limit (read per frame: experiment knobs).
int trace_grid_rotation(int nLocal, int H, int numCtas, int warpsPerCta, bool fullLivedRays, unsigned int* slotsOut)
{
    *slotsOut = 0;
    const char* on = getenv("ART_K1_ROTATE");
    if (on && atoi(on) == 0) return 0;
    const char* mk = getenv("ART_K1_ROTATE_MAX_K");
    const long long maxK = mk ? atoll(mk) : (fullLivedRays ? 1024 : 4);
    const long long warps = (long long)numCtas * warpsPerCta;
    const long long groups = ((long long)nLocal + 31) / 32;
    if (H < 2 || H > 255 || groups <= warps || groups > maxK * warps) return 0;
    const long long slots = groups * (H - 1);
    if (slots * (long long)kRotateSlotBytes > (1ll << 30) || groups >= (1ll << 22)) return 0;
    *slotsOut = (unsigned int)slots;
    return (int)groups;
}

size_t trace_grid_scratch_bytes(int numCtas) { return (size_t)numCtas * kGridWarps * kChunkQ * sizeof(uint32_t); }

size_t trace_grid_smem_bytes(const GeomLayout& L, bool geomInSmem)
{
    return (geomInSmem ? L.bytes : 0) + (size_t)kGridWarps * 32 * (sizeof(HitRec) + kQueryWords * 4);
}

// The occlusion queries walk the grid cells along their segments inside the bounce loop.
cudaError_t launch_trace_grid(const TraceArgs& a0, const GridDesc& g, int numCtas, bool geomInSmem, bool stats, cudaStream_t stream)
{
    TraceArgs a = a0;
    size_t smem = trace_grid_smem_bytes(a.L, geomInSmem);
    a.goalsInSmem = 0; a.goalsSmemOffset = 0;
    const char* noTab = getenv("ART_K1_NO_GOAL_TABLES");         // (read per launch: test knob for the global-memory fallback)
    if (!(noTab && atoi(noTab) != 0)) {                          // goal tables behind everything else, when they fit
        static int maxOptin = -1;
        if (maxOptin < 0) {
            int dev = 0;
            if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) maxOptin = 0;
        }
        const size_t off = (smem + 15) & ~(size_t)15;
        const size_t need = off + 2 * (size_t)(a.nTargets + 1) * sizeof(float4);
        if (need <= (size_t)maxOptin) { a.goalsInSmem = 1; a.goalsSmemOffset = (unsigned int)off; smem = need; }
    }
    void (*k)(const TraceArgs, const GridDesc) = nullptr;
    const bool rot = a.migGroups > 0;
    if (stats && rot) return cudaErrorInvalidValue;          // (stats frames are planned without rotation)
    if (stats) k = geomInSmem ? trace_grid_kernel<true, true, false> : trace_grid_kernel<false, true, false>;
    else if (rot) k = geomInSmem ? trace_grid_kernel<true, false, true> : trace_grid_kernel<false, false, true>;
    else k = geomInSmem ? trace_grid_kernel<true, false, false> : trace_grid_kernel<false, false, false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int warps = a.gridWarps >= 1 && a.gridWarps <= kGridWarps ? a.gridWarps : kGridWarps;
    k<<<numCtas, warps * 32, smem, stream>>>(a, g);
    return cudaGetLastError();
}

}  // namespace art
