// k2_permeation_grid.cu -- K2 with the uniform-grid acceleration structure:
// AudioPermeationJobBatched.Execute (Assets/C# Scripts/Jobs/AudioPermeationJobBatched.cs:34-91).
//
// Mapping: a warp takes 32 rays at a time from the queue.
//   Phase 1, one thread = one ray: first-hit DISTANCE by a 3D-DDA walk (PM:101-141; the OBB test uses the
//            inverse of the stored rotation, PM:174, quirk Q4). Exact tests, so firstHitDist -- the input of
//            perm_last_kernel, which produces the canonical PermeationPowerRemains bit-exactly -- is the same
//            float the brute-force kernel writes.
//   Phase 2, one thread = one (ray, target) pair: the through-material loss of PM:225-261 (no early exit, no
//            distance clip -- quirk Q7) accumulated cell by cell along the whole line, each collider's interval
//            clipped to the cell so that a collider listed in several cells is counted once. A lane keeps the
//            same target for all 32 rays, so the per-target sums live in registers and are flushed with two
//            integer atomics per target block (deterministic: integer part + 36-bit fixed-point fraction).
// The per-pair values feed only the permeationSum extension (tolerance 1e-5 relative to N*S, include/audiort.h);
// they are evaluated in cheap arithmetic (FMA, approximate reciprocals). What the reference keeps of this job --
// the last hitting ray of each batch (PM:85, quirk Q5/Q6) -- is recomputed by perm_last_kernel in the
// reference's own operation and summation order.
#include "device_util.cuh"
#include "fan_dev.cuh"
#include "grid_dev.cuh"
#include "intersect.cuh"
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {

#ifndef ART_PGRID_WARPS
#define ART_PGRID_WARPS 24
#endif
constexpr int kPGridWarps = ART_PGRID_WARPS;
constexpr int kPGridThreads = kPGridWarps * 32;

// length of [tEnter, tExit] inside the cell interval [tIn, tOut]
__device__ __forceinline__ float clip_len(float tEnter, float tExit, float tIn, float tOut)
{
    return fmaxf(0.0f, fminf(tExit, tOut) - fmaxf(tEnter, tIn));
}

// FAN: phase 2 does not walk the grid; a (ray, target) line tests the three lists the target's fan holds for it
// (fan_dev.cuh): the near list, the direction bin of (hit point - target) for the part of the line before the
// target and the opposite bin for the part behind it (PM:225 ignores the target distance, quirk Q7). Every
// collider is met once with its whole chord; the two bins are clipped at the target's parameter so that a collider
// listed in both (possible only for conservatively inflated bounds) is still counted once.
template <bool SMEM, bool STATS, bool FAN>
__global__ void __launch_bounds__(kPGridThreads, 1) permeation_grid_kernel(const PermArgs a, const GridDesc g, const FanDesc f)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Na = a.nTargets;

    unsigned char* p = smem;
    const unsigned char* geomBase = a.geom;
    if (SMEM) {
        stage_blob_to_smem(p, a.geom, a.L.bytes, &bar);
        geomBase = p;
        p += a.L.bytes;
    }
    float4* rec = reinterpret_cast<float4*>(p) + warp * 32;      // (Pp.xyz, hit flag) per ray of the warp
    p += (size_t)kPGridWarps * 32 * sizeof(float4);
    // FAN: true densities (PM:235/245/255 skip owned colliders per TARGET; the fans leave those out of the lists, so unlike
    // a.dens* these are not zeroed for owned colliders) -- in shared memory beside the geometry when that fits
    const float* densS = nullptr; const float* densA = nullptr; const float* densO = nullptr;
    if (FAN) {
        if (SMEM) {
            float* d = reinterpret_cast<float*>(p);
            for (int i = threadIdx.x; i < a.L.nsPad; i += kPGridThreads) d[i] = a.at.sphAttr[i].z;
            for (int i = threadIdx.x; i < a.L.naPad; i += kPGridThreads) d[a.L.nsPad + i] = a.at.aabbAttr[i].z;
            for (int i = threadIdx.x; i < a.L.noPad; i += kPGridThreads) d[a.L.nsPad + a.L.naPad + i] = a.at.obbAttr[i].z;
            __syncthreads();
            densS = d; densA = d + a.L.nsPad; densO = d + a.L.nsPad + a.L.naPad;
        } else {
            densS = a.trueDens; densA = a.trueDens + a.L.nsPad; densO = a.trueDens + a.L.nsPad + a.L.naPad;
        }
    }
    const GeomView gv = make_view(geomBase, a.L);
    const f3 RayOrigin = mk3(a.ox, a.oy, a.oz);
    unsigned int nRays = 0, nHitRays = 0;
    unsigned long long st[7] = { 0, 0, 0, 0, 0, 0, 0 };   // STATS: first-hit S/A/O tests, loss S/A/O tests, cells

    for (;;) {
        int base = 0;
        if (lane == 0) base = (int)atomicAdd(a.nextRay, (unsigned)a.raysPerWarp);
        base = __shfl_sync(kFull, base, 0);
        if (base >= a.map.nLocal) break;
        const int j = base + lane;
        const bool hasRay = lane < a.raysPerWarp && j < a.map.nLocal;

        // ================= Phase 1: ShootRayCast, distance only (PM:101-141) =================
        float best = pos_inf();                                                // math.INFINITY
        if (hasRay) {
            nRays++;
            const int rayIndex = a.map.to_global(j);
            const f3 d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                             um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));   // PM:53
            const f3 o = RayOrigin;
            const float dd = dot3(d, d);
            const f3 inv = mk3(rcpr(d.x), rcpr(d.y), rcpr(d.z));
            Dda w;
            bool walking = dda_init(g, o, d, inv, pos_inf(), w);
            while (walking) {
                const uint2 hdr = dda_cell(g, w);
                const uint16_t* e = g.entries + hdr.x;
                const int nS = hdr.y & 1023, nA = (hdr.y >> 10) & 2047, nO = hdr.y >> 21;
                if (STATS) { st[0] += nS; st[1] += nA; st[2] += nO; st[6]++; }
                ART_CHECK(a.counters, (unsigned)w.ix < (unsigned)g.nx && (unsigned)w.iy < (unsigned)g.ny && (unsigned)w.iz < (unsigned)g.nz);
                ART_CHECK(a.counters, hdr.x + nS + nA + nO <= (unsigned)g.nEntries);
                for (int k = 0; k < nS; k++) {
                    const float dist = sphere_dist(gv, __ldg(e + k), o, d, dd);
                    if (dist < best) best = dist;
                }
                e += nS;
                for (int k = 0; k < nA; k++) {
                    const float dist = aabb_dist(gv, __ldg(e + k), o, inv);
                    if (dist < best) best = dist;
                }
                e += nA;
                for (int k = 0; k < nO; k++) {
                    const int id = __ldg(e + k);
                    const float dist = obb_dist_nearest_q(gv, a.at.obbQinv[id], id, o, d, dd, g.errScale, best);   // PM:174 (quirk Q4)
                    if (dist < best) best = dist;
                }
                const float tNext = dda_next_t(w);
                if (tNext > w.tEnd || tNext > best) break;
                walking = dda_step(g, d, w);
            }
            a.firstHitDist[j] = best;
            const bool hit = best != pos_inf();                                // PM:140 / PM:58
            if (hit) {
                nHitRays++;
                atomicMax(&a.lastHitRay[rayIndex / a.batchSize], rayIndex);
                const f3 P = add3(o, mul3s(d, best));                          // PM:61
                const f3 Pp = sub3(P, mul3s(d, kEpsilon));                     // PM:72
                rec[lane] = make_float4(Pp.x, Pp.y, Pp.z, 1.0f);
            } else {
                rec[lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        } else {
            rec[lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        __syncwarp();
        if (a.hitPts) {                 // binned mode: the loss lines are evaluated by perm_loss_binned_kernel
            if (lane < a.raysPerWarp && j < a.map.nLocal) a.hitPts[j] = rec[lane];
            __syncwarp();
            continue;
        }

        // ================= Phase 2: per-target loss rays (PM:67-86), one lane = one (ray, target) pair =================
        for (int tb = 0; tb * 32 < Na; tb++) {
            const int tpb = min(32, Na - tb * 32);          // targets in this block
            const int rp = 32 / tpb;                        // rays handled side by side
            const int rsub = lane / tpb;
            const bool laneOn = rsub < rp;
            const int tgt = laneOn ? a.targetOrder[tb * 32 + (lane - rsub * tpb)] : 0;   // spatially sorted lane order
            const f3 T = laneOn ? mk3(a.targets[3 * tgt], a.targets[3 * tgt + 1], a.targets[3 * tgt + 2]) : mk3(0, 0, 0);
            long long accInt = 0, accFrac = 0;
            for (int r0 = 0; r0 < a.raysPerWarp; r0 += rp) {
                const int r = r0 + rsub;
                float4 rr = make_float4(0, 0, 0, 0);
                if (laneOn && r < 32) rr = rec[r];
                if (rr.w != 0.0f) {
                    const f3 Pp = mk3(rr.x, rr.y, rr.z);
                    const f3 toT = sub3(T, Pp);
                    const f3 dir = normalize3(toT);                            // PM:76
                    const f3 inv = mk3(rcpr(dir.x), rcpr(dir.y), rcpr(dir.z));             // PM:270
                    float loss = 0.0f;
                    if (FAN) {
                        const int bin = fan_bin(-toT.x, -toT.y, -toT.z);       // direction target -> hit point
                        if (bin >= 0) {
                            const float tT = sqrt_fast(fmaf(toT.z, toT.z, fmaf(toT.y, toT.y, toT.x * toT.x)));   // line parameter of the target
                            // cell 0: near list, whole line; 1: bin towards the hit point, t in [0, tT]; 2: opposite bin, t > tT
                            const int fanBase = tgt * kFanCells;
                            const int face = bin / kFanCellsPerFace, rb = bin - face * kFanCellsPerFace;
                            const uint2 h0 = __ldg(&f.cells[fanBase + 6 * kFanCellsPerFace]);
                            const uint2 h1 = __ldg(&f.cells[fanBase + bin]);
                            const uint2 h2 = __ldg(&f.cells[fanBase + (face ^ 1) * kFanCellsPerFace + (kFanCellsPerFace - 1 - rb)]);
                            const int nS0 = h0.y & 1023, nA0 = (h0.y >> 10) & 2047, nO0 = h0.y >> 21;
                            const int nS1 = h1.y & 1023, nA1 = (h1.y >> 10) & 2047, nO1 = h1.y >> 21;
                            const int nS2 = h2.y & 1023, nA2 = (h2.y >> 10) & 2047, nO2 = h2.y >> 21;
                            if (STATS) { st[3] += nS0 + nS1 + nS2; st[4] += nA0 + nA1 + nA2; st[5] += nO0 + nO1 + nO2; st[6] += 3; }
                            ART_CHECK(a.counters, h0.x + nS0 + nA0 + nO0 <= (unsigned)f.nEntries && h1.x + nS1 + nA1 + nO1 <= (unsigned)f.nEntries &&
                                                  h2.x + nS2 + nA2 + nO2 <= (unsigned)f.nEntries && tgt >= 0 && tgt < Na);
                            const float inf = pos_inf();
                            // each type's lists of the three cells run as ONE loop (the lanes of a warp have lists of different
                            // lengths: three loops instead of nine), the next index is fetched while the current collider is tested
                            {   // ---- AABBs, PM:265-288 in the reference's operation order
                                const int n01 = nA0 + nA1, n = n01 + nA2;
                                // 32-bit entry offsets of the three lists, biased so that offset + k addresses entry k of the run
                                const uint32_t o0 = h0.x + nS0, o1 = h1.x + nS1 - nA0, o2 = h2.x + nS2 - n01;
                                int nxt = n > 0 ? (int)__ldg(f.entries + (0 < nA0 ? o0 : (0 < n01 ? o1 : o2))) : 0;
                                for (int k = 0; k < n; k++) {
                                    const int id = nxt;
                                    const int k1 = k + 1;
                                    if (k1 < n) nxt = (int)__ldg(f.entries + ((k1 < nA0 ? o0 : (k1 < n01 ? o1 : o2)) + (uint32_t)k1));
                                    ART_CHECK(a.counters, id < a.L.na);
                                    const float tIn = k >= n01 ? tT : 0.0f, tOut = (k >= nA0 && k < n01) ? tT : inf;
                                    const float4 A = gv.aabbA[id];
                                    const float2 B = gv.aabbB[id];
                                    float tEnter, tExit;
                                    slab<8>(subr(A.x, Pp.x), subr(A.y, Pp.y), subr(A.z, Pp.z), subr(A.w, Pp.x), subr(B.x, Pp.y), subr(B.y, Pp.z),
                                            inv.x, inv.y, inv.z, tEnter, tExit);
                                    const float len = clip_len(tEnter, tExit, tIn, tOut);
                                    if (len > 0.0f) loss = fmaf(len, densA[id], loss);
                                }
                            }
                            {   // ---- spheres, PM:303-328 (unit direction)
                                const int n01 = nS0 + nS1, n = n01 + nS2;
                                const uint32_t o0 = h0.x, o1 = h1.x - nS0, o2 = h2.x - n01;
                                for (int k = 0; k < n; k++) {
                                    const int id = (int)__ldg(f.entries + ((k < nS0 ? o0 : (k < n01 ? o1 : o2)) + (uint32_t)k));
                                    ART_CHECK(a.counters, id < a.L.ns);
                                    const float tIn = k >= n01 ? tT : 0.0f, tOut = (k >= nS0 && k < n01) ? tT : inf;
                                    const float4 sp = gv.sph[id];
                                    const f3 oc = sub3(Pp, mk3(sp.x, sp.y, sp.z));
                                    const float cc = subr(dot3(oc, oc), sp.w);
                                    float b;
                                    if (sphere_loss_fast_miss(oc, cc, dir, b)) continue;   // disc < 0 (PM:311)
                                    const float sq = sqrtr(subr(mulr(b, b), cc));
                                    const float len = clip_len(subr(-b, sq), addr(-b, sq), tIn, tOut);
                                    if (len > 0.0f) loss = fmaf(len, densS[id], loss);
                                }
                            }
                            {   // ---- OBBs, PM:294-300 (stored rotation as is), cheap arithmetic about the point of closest approach
                                const int n01 = nO0 + nO1, n = n01 + nO2;
                                const uint32_t o0 = h0.x + nS0 + nA0, o1 = h1.x + nS1 + nA1 - nO0, o2 = h2.x + nS2 + nA2 - n01;
                                int nxt = n > 0 ? (int)__ldg(f.entries + (0 < nO0 ? o0 : (0 < n01 ? o1 : o2))) : 0;
                                for (int k = 0; k < n; k++) {
                                    const int id = nxt;
                                    const int k1 = k + 1;
                                    if (k1 < n) nxt = (int)__ldg(f.entries + ((k1 < nO0 ? o0 : (k1 < n01 ? o1 : o2)) + (uint32_t)k1));
                                    ART_CHECK(a.counters, id < a.L.no);
                                    const float tIn = k >= n01 ? tT : 0.0f, tOut = (k >= nO0 && k < n01) ? tT : inf;
                                    const float4 c4 = gv.obbC[id];
                                    const float2 h2o = gv.obbH[id];
                                    const f3 pc = mk3(Pp.x - c4.x, Pp.y - c4.y, Pp.z - c4.z);
                                    const float bq = fmaf(pc.z, dir.z, fmaf(pc.y, dir.y, pc.x * dir.x));
                                    const float pp = fmaf(pc.z, pc.z, fmaf(pc.y, pc.y, pc.x * pc.x));
                                    const float r2 = fmaf(h2o.y, h2o.y, fmaf(h2o.x, h2o.x, c4.w * c4.w));
                                    if (pp - bq * bq > r2 * 1.001f + 1e-4f) continue;   // the line passes the bounding sphere
                                    const float4 q4 = gv.obbQ[id];
                                    const f3 pn = mk3(fmaf(dir.x, -bq, pc.x), fmaf(dir.y, -bq, pc.y), fmaf(dir.z, -bq, pc.z));
                                    const f3 lo = qrot_fast(q4, pn), ld = qrot_fast(q4, dir);
                                    const float rx = rcp_fast(ld.x), ry = rcp_fast(ld.y), rz = rcp_fast(ld.z);
                                    const float ax = (-c4.w - lo.x) * rx, bx = (c4.w - lo.x) * rx;
                                    const float ay = (-h2o.x - lo.y) * ry, by = (h2o.x - lo.y) * ry;
                                    const float az = (-h2o.y - lo.z) * rz, bz = (h2o.y - lo.z) * rz;
                                    const float tEnter = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)) - bq;
                                    const float tExit = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) - bq;
                                    const float len = clip_len(tEnter, tExit, tIn, tOut);
                                    if (len > 0.0f) loss = fmaf(len, densO[id], loss);
                                }
                            }
                        }
                    } else {
                        Dda w;
                        bool walking = dda_init(g, Pp, dir, inv, pos_inf(), w);
                        const float tRay0 = w.tCur;                 // where the line enters the grid (0 inside it)
                        while (walking) {
                            const uint2 hdr = dda_cell(g, w);
                            const uint16_t* e = g.entries + hdr.x;
                            const int nS = hdr.y & 1023, nA = (hdr.y >> 10) & 2047, nO = hdr.y >> 21;
                            // the cell the walk came from differs along one axis: an OBB whose cell range contains that
                            // coordinate was already met there (the cells of a convex range along a line are contiguous)
                            const int axShift = 8 * w.lastAxis;
                            const int prevCoord = w.lastAxis == 0 ? w.ix - (dir.x > 0.0f ? 1 : -1)
                                                : (w.lastAxis == 1 ? w.iy - (dir.y > 0.0f ? 1 : -1) : w.iz - (dir.z > 0.0f ? 1 : -1));
                            const float tIn = w.tCur;
                            const float tOut = fminf(dda_next_t(w), w.tEnd);
                            if (STATS) { st[3] += nS; st[4] += nA; st[5] += nO; st[6]++; }
                            ART_CHECK(a.counters, (unsigned)w.ix < (unsigned)g.nx && (unsigned)w.iy < (unsigned)g.ny && (unsigned)w.iz < (unsigned)g.nz);
                            ART_CHECK(a.counters, hdr.x + nS + nA + nO <= (unsigned)g.nEntries && tgt >= 0 && tgt < Na);
                            // AABB and sphere intervals are evaluated in the reference's own operation order, so tEnter/tExit are
                            // the reference's floats (a near-tangent sphere crossing, 2*sqrt(disc) with disc ~ 0, would otherwise
                            // amplify harmless rounding into a visible difference); only the per-cell clipping is new.
                            for (int k = 0; k < nA; k++) {                          // PM:265-288
                                const int id = __ldg(e + nS + k);
                                ART_CHECK(a.counters, id < a.L.na);
                                const float4 A = gv.aabbA[id];
                                const float2 B = gv.aabbB[id];
                                float tEnter, tExit;
                                slab<8>(subr(A.x, Pp.x), subr(A.y, Pp.y), subr(A.z, Pp.z), subr(A.w, Pp.x), subr(B.x, Pp.y), subr(B.y, Pp.z),
                                        inv.x, inv.y, inv.z, tEnter, tExit);
                                const float len = clip_len(tEnter, tExit, tIn, tOut);
                                if (len > 0.0f) {
                                    const float4 at = a.at.aabbAttr[id];
                                    if (__float_as_int(at.w) != tgt) loss = fmaf(len, at.z, loss);     // PM:245 owner skip
                                }
                            }
                            for (int k = 0; k < nS; k++) {                          // PM:303-328 (unit direction)
                                const int id = __ldg(e + k);
                                const float4 s = gv.sph[id];
                                const f3 oc = sub3(Pp, mk3(s.x, s.y, s.z));
                                const float cc = subr(dot3(oc, oc), s.w);
                                float b;
                                if (sphere_loss_fast_miss(oc, cc, dir, b)) continue;   // disc < 0 (PM:311)
                                const float sq = sqrtr(subr(mulr(b, b), cc));
                                const float len = clip_len(subr(-b, sq), addr(-b, sq), tIn, tOut);
                                if (len > 0.0f) {
                                    const float4 at = a.at.sphAttr[id];
                                    if (__float_as_int(at.w) != tgt) loss = fmaf(len, at.z, loss);     // PM:235
                                }
                            }
                            for (int k = 0; k < nO; k++) {                          // PM:294-300 (stored rotation as is)
                                const int id = __ldg(e + nS + nA + k);
                                ART_CHECK(a.counters, id < a.L.no);
                                const float4 c4 = gv.obbC[id];
                                const float2 h2 = gv.obbH[id];
                                const f3 pc = mk3(Pp.x - c4.x, Pp.y - c4.y, Pp.z - c4.z);
                                // bounding sphere first: |pc x dir|^2 > r^2 means the line passes the box
                                const float bq = fmaf(pc.z, dir.z, fmaf(pc.y, dir.y, pc.x * dir.x));
                                const float pp = fmaf(pc.z, pc.z, fmaf(pc.y, pc.y, pc.x * pc.x));
                                const float r2 = fmaf(h2.y, h2.y, fmaf(h2.x, h2.x, c4.w * c4.w));
                                if (pp - bq * bq > r2 * 1.001f + 1e-4f) continue;
                                if (w.lastAxis >= 0) {              // count every OBB once: where the walk first meets it
                                    const uint2 rg = __ldg(&g.rangeO[id]);
                                    const int lo = (int)((rg.x >> axShift) & 255u), hi = (int)((rg.y >> axShift) & 255u);
                                    if (prevCoord >= lo && prevCoord <= hi) continue;
                                }
                                const float4 q4 = gv.obbQ[id];
                                // rotate the point of closest approach (|pn| <= r) instead of pc (|pc| can be the whole room):
                                // the rounding of the cheap rotation then scales with the box, not with the distance to it
                                const f3 pn = mk3(fmaf(dir.x, -bq, pc.x), fmaf(dir.y, -bq, pc.y), fmaf(dir.z, -bq, pc.z));
                                const f3 lo = qrot_fast(q4, pn), ld = qrot_fast(q4, dir);
                                const float rx = rcp_fast(ld.x), ry = rcp_fast(ld.y), rz = rcp_fast(ld.z);
                                const float ax = (-c4.w - lo.x) * rx, bx = (c4.w - lo.x) * rx;
                                const float ay = (-h2.x - lo.y) * ry, by = (h2.x - lo.y) * ry;
                                const float az = (-h2.y - lo.z) * rz, bz = (h2.y - lo.z) * rz;
                                const float tEnter = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)) - bq;
                                const float tExit = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) - bq;
                                const float len = fmaxf(0.0f, tExit - fmaxf(tEnter, tRay0));   // the whole chord, once (PM:286-287)
                                if (len > 0.0f) {
                                    const float4 at = a.at.obbAttr[id];
                                    if (__float_as_int(at.w) != tgt) loss = fmaf(len, at.z, loss);     // PM:255
                                }
                            }
                            if (dda_next_t(w) > w.tEnd) break;
                            walking = dda_step(g, dir, w);
                        }
                    }
                    const float v = subr(a.nTimesS, loss);                     // PM:260
                    const float ip = truncf(v);
                    accInt += (long long)ip;
                    accFrac += (long long)(((double)v - (double)ip) * 68719476736.0);
                }
            }
            if (laneOn && (accInt != 0 || accFrac != 0)) {
                atomicAdd(reinterpret_cast<unsigned long long*>(&a.permSumInt[tgt]), (unsigned long long)accInt);
                atomicAdd(reinterpret_cast<unsigned long long*>(&a.permSumFrac[tgt]), (unsigned long long)accFrac);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        nRays += __shfl_xor_sync(kFull, nRays, s);
        nHitRays += __shfl_xor_sync(kFull, nHitRays, s);
    }
    if (lane == 0) {
        atomicAdd(&a.counters[C_PERM_RAYS], (unsigned long long)nRays);
        atomicAdd(&a.counters[C_PERM_HIT_RAYS], (unsigned long long)nHitRays);
    }
    if (STATS) {
        for (int k = 0; k < 3; k++) {
            atomicAdd(&a.counters[C_GRID_PF_S + k], st[k]);
            atomicAdd(&a.counters[C_GRID_PL_S + k], st[3 + k]);
        }
        atomicAdd(&a.counters[C_GRID_PM_CELLS], st[6]);
    }
}

// Rays a warp takes at a time: chosen like trace_grid_rays_per_warp so that small batches are spread evenly over the
// warps (a multiple of the rays handled side by side when there are fewer than 32 targets).
int perm_grid_rays_per_warp(int nLocal, int nTargets, int numCtas)
{
    const long long warps = (long long)numCtas * kPGridWarps;
    const long long k = (nLocal + warps * 32 - 1) / (warps * 32);
    long long r = k > 0 ? (nLocal + warps * k - 1) / (warps * k) : 32;
    const int side = nTargets >= 32 ? 1 : 32 / nTargets;
    r = (r + side - 1) / side * side;
    return (int)(r < side ? side : (r > 32 ? 32 : r));
}

size_t perm_grid_smem_bytes(const GeomLayout& L, bool geomInSmem, bool fans)
{
    return (geomInSmem ? L.bytes + (fans ? ((size_t)L.nsPad + L.naPad + L.noPad) * sizeof(float) : 0) : 0) + (size_t)kPGridWarps * 32 * sizeof(float4);
}

// fans == nullptr: the loss lines walk the grid cells instead of using the target fans
cudaError_t launch_permeation_grid(const PermArgs& a, const GridDesc& g, const FanDesc* fans, int numCtas, bool geomInSmem, bool stats, cudaStream_t stream)
{
    const size_t smem = perm_grid_smem_bytes(a.L, geomInSmem, fans != nullptr);
    void (*k)(const PermArgs, const GridDesc, const FanDesc) = nullptr;
    if (fans) {
        if (geomInSmem) k = stats ? permeation_grid_kernel<true, true, true> : permeation_grid_kernel<true, false, true>;
        else k = stats ? permeation_grid_kernel<false, true, true> : permeation_grid_kernel<false, false, true>;
    } else {
        if (geomInSmem) k = stats ? permeation_grid_kernel<true, true, false> : permeation_grid_kernel<true, false, false>;
        else k = stats ? permeation_grid_kernel<false, true, false> : permeation_grid_kernel<false, false, false>;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    FanDesc fd{};
    if (fans) fd = *fans;
    k<<<numCtas, kPGridThreads, smem, stream>>>(a, g, fd);
    return cudaGetLastError();
}

}  // namespace art
