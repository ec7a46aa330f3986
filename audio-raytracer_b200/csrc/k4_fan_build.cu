// k4_fan_build.cu -- per-frame construction of the target fans (fan_dev.cuh) on the device.
//
// Two steps (after the per-goal ordering, fan_order_kernel). fan_project_kernel, one CTA per (goal, cube face): the colliders
// are swept in the goal's order (spheres | AABBs | OBBs, each nearest first), every thread projects one collider's
// conservative box onto the face (a bin rectangle, or nothing) and the non-empty rectangles are compacted IN ORDER into the
// face's rectangle list. fan_match_kernel, one CTA per (goal, face, strip of bin rows), one thread per direction bin: every
// warp walks the face's list for the rectangles that contain its bins. Pass 0 counts, a block scan + one atomicAdd reserves
// the CTA's span of the entry array, pass 1 walks the list again and writes the indices -- so every list is grouped by type
// and ordered like the sweep, and its content does not depend on scheduling (only its position in the entry array does).
#include <cstdlib>

#include "device_util.cuh"
#include "fan_dev.cuh"
#include "launchers.h"

namespace art {

constexpr uint32_t kRectEmpty = 0x000000FFu;   // a0 = 255 > a1 = 0

// Rectangle of the box [lo, hi] (already conservative, grid_host.h) seen from T on cube face (k, sgn), in SUB-BIN units
// (kFanSub x kFanSub sub-bins per bin; bin index = sub-bin index / kFanSub -- the lists are per bin, the covering depths per
// sub-bin), packed a0 | a1 << 8 | b0 << 16 | b1 << 24, or kRectEmpty. near: T lies within nearDist of the box (per axis).
__device__ __forceinline__ uint32_t fan_rect(const float lo[3], const float hi[3], const float T[3], int k, bool neg, float nearDist, bool& near, float& nearDepth)
{
    nearDepth = 0.0f;
    float rl[3], rh[3];
    near = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float e = 4e-6f * (fabsf(T[c]) + fabsf(lo[c]) + fabsf(hi[c])) + 1e-6f;   // rounding of the two subtractions
        rl[c] = lo[c] - T[c] - e;
        rh[c] = hi[c] - T[c] + e;
        near = near && rl[c] <= nearDist && rh[c] >= -nearDist;
    }
    if (near) return kRectEmpty;
    const int i = k == 2 ? 0 : k + 1, j = k == 0 ? 2 : k - 1;   // (k+1)%3, (k+2)%3
    const float w0 = neg ? -rh[k] : rl[k], w1 = neg ? -rl[k] : rh[k];
    nearDepth = w0;                                             // depth of the (inflated) box's near face: no AABB covers a bin from nearer
    if (!(w1 > 0.0f)) return kRectEmpty;
    const float x0 = rl[i], x1 = rh[i], y0 = rl[j], y1 = rh[j];
    const float xmin = (x0 <= 0.0f && x1 >= 0.0f) ? 0.0f : fminf(fabsf(x0), fabsf(x1));
    const float ymin = (y0 <= 0.0f && y1 >= 0.0f) ? 0.0f : fminf(fabsf(y0), fabsf(y1));
    // directions on this face have w >= |x|, |y| (1 % slack: a query direction within 2e-3 of the face edge)
    const float wlo = fmaxf(w0, 0.99f * fmaxf(xmin, ymin));
    if (wlo > w1) return kRectEmpty;
    const float inf = __int_as_float(0x7F800000);
    const float r1 = 1.0f / w1;
    const bool pos = wlo > 0.0f;
    const float rlo = pos ? 1.0f / wlo : 0.0f;
    const float amin = (x0 >= 0.0f ? x0 * r1 : (pos ? x0 * rlo : -inf)) - kFanTanMargin;
    const float amax = (x1 <= 0.0f ? x1 * r1 : (pos ? x1 * rlo : inf)) + kFanTanMargin;
    const float bmin = (y0 >= 0.0f ? y0 * r1 : (pos ? y0 * rlo : -inf)) - kFanTanMargin;
    const float bmax = (y1 <= 0.0f ? y1 * r1 : (pos ? y1 * rlo : inf)) + kFanTanMargin;
    if (amin > 1.0f || amax < -1.0f || bmin > 1.0f || bmax < -1.0f) return kRectEmpty;
    const float sc = 0.5f * kFanFine;
    const uint32_t a0 = (uint32_t)min(kFanFine - 1, max(0, (int)floorf((fmaxf(amin, -1.0f) + 1.0f) * sc)));
    const uint32_t a1 = (uint32_t)min(kFanFine - 1, max(0, (int)floorf((fminf(amax, 1.0f) + 1.0f) * sc)));
    const uint32_t b0 = (uint32_t)min(kFanFine - 1, max(0, (int)floorf((fmaxf(bmin, -1.0f) + 1.0f) * sc)));
    const uint32_t b1 = (uint32_t)min(kFanFine - 1, max(0, (int)floorf((fminf(bmax, 1.0f) + 1.0f) * sc)));
    return a0 | (a1 << 8) | (b0 << 16) | (b1 << 24);
}

// ---- covering depth of a bin --------------------------------------------------------------------
// An AABB COVERS bin (face, ia, ib) of goal G over the depth interval [w1, w2] (depth = coordinate along the face's major
// axis, measured from G) when every point G + z * (+-1, a, b), z in [w1, w2], (a, b) in the bin's tangent rectangle widened by
// kFanCoverTan, lies inside the box shrunk by the rounding of its subtraction from G. The segment from a hit point P in that
// bin at depth wP > w1 to G then runs INSIDE the box between the depths w1 and min(w2, wP), so the reference's slab test
// (RT:284-308) reports the box with a distance below the limit (RT:384 / RT:430) -- by margins that exceed its FP32 rounding
// a hundredfold (query_fan_kernel, "cull"):
//   * its t values are s * Lr * (1 + th), s the true segment parameter of a plane crossing, Lr = 1 / fl(1 / len) common to all of
//     them, |th| < 6e-7; the true entry and exit parameters differ by >= min(2^-10, (w2 - w1) / wP) >= 1e-4 (w2 - w1 >= 1e-4 * D
//     is required below, wP <= D is checked by the query), the exit parameter is >= 2^-10 (cover threshold w1 * (1 + 2^-9)), so
//     "tNear > tFar" and "tFar < 0" are both false;
//   * the reported distance is <= tFar <= len - zl + 6e-7 * len, zl >= 1e-3 * D >= 0.033 the depth of the box's near face (G is
//     outside the box); the limit is len (muffle rays) or distance(RayOrigin, hit point) >= len - 1.01e-4 (echo rays: P is the
//     hit point moved back by kEpsilon = 1e-4): dist < limit;
//   * no 0 * Inf: a zero direction component k means P_k == G_k, and then lo_k - P_k < 0 < hi_k - P_k by the tangent margin.
// Covering depths are kept per SUB-BIN (kFanSub x kFanSub per bin: the finer the cell, the more often one box fills it -- C3:
// 75 % of the queries lie beyond the covering depth of their bin, 86 % beyond that of their sub-bin), as the smallest threshold
// over the covering AABBs, one byte each in cells4.w: code c stands for the depth nearDist * 2^(c / coverLogS), rounded UP when
// encoded (fan_cover_code; a larger threshold only culls less), 255 = none. The query compares in the log domain,
// coverLogS * log2(w) + coverLogK > c + kFanCoverLogEps, with lg2.approx on both sides (abs. error < 3e-4 code units).
constexpr float kFanCoverTan = 1e-3f;
__device__ __forceinline__ uint32_t fan_cover_code(float depth, float logS, float logK)
{
    if (!(depth < 3.0e38f)) return 255u;
    const float c = ceilf(fmaf(__log2f(depth), logS, logK) + kFanCoverLogEps);
    return c <= 254.0f ? (uint32_t)fmaxf(c, 0.0f) : 255u;
}

// cover parameters of AABB [lo, hi] (the reference's own min / max, GeomView::aabbA/B) on face (k, neg) of goal T:
// depth range [c[0], c[1]], tangent ranges [c[2], c[3]] (axis i) and [c[4], c[5]] (axis j), all shrunk; c[0] > c[1] if unusable
// (degenerate, or nearer to the goal than minDepth)
__device__ __forceinline__ void fan_cover_params(const float lo[3], const float hi[3], const float T[3], int k, bool neg, float minDepth, float c[6])
{
    float rl[3], rh[3];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float e = 4e-6f * (fabsf(T[q]) + fabsf(lo[q]) + fabsf(hi[q])) + 1e-6f;
        rl[q] = lo[q] - T[q] + e;
        rh[q] = hi[q] - T[q] - e;
    }
    const int i = k == 2 ? 0 : k + 1, j = k == 0 ? 2 : k - 1;
    c[0] = neg ? -rh[k] : rl[k];
    c[1] = neg ? -rl[k] : rh[k];
    c[2] = rl[i]; c[3] = rh[i]; c[4] = rl[j]; c[5] = rh[j];
    // the near face must lie at least minDepth in front of the goal (the goal is outside the box, with room for kEpsilon)
    if (!(c[0] >= minDepth) || !(c[0] <= c[1]) || !(c[2] <= c[3]) || !(c[4] <= c[5])) { c[0] = 1.0f; c[1] = 0.0f; }
}
// the depth interval over which `coef * z <= bound` holds, intersected into [w1, w2]
__device__ __forceinline__ void fan_cover_clip(float coef, float bound, float& w1, float& w2)
{
    if (coef > 0.0f) w2 = fminf(w2, __fdividef(bound, coef));
    else if (coef < 0.0f) w1 = fmaxf(w1, __fdividef(bound, coef));
    else if (bound < 0.0f) w2 = -1.0f;
}

// Per goal: the colliders sorted by (type, distance of their box from the goal, index). The build sweeps them in this
// order, so every list comes out grouped by type and nearest-first: a collider close to the goal subtends more of the
// bin and lies between the goal and more hit points, hence an any-hit query (K1) finds its blocker in the first entries.
// One CTA per goal, bitonic sort of 64-bit (key, index) pairs in shared memory.
__global__ void __launch_bounds__(1024, 1) fan_order_kernel(const FanBuildArgs a, int nPow2)
{
    extern __shared__ unsigned long long sKeys[];
    const int fan = blockIdx.x, tid = threadIdx.x;
    const int nc = a.ns + a.na + a.no;
    float T[3];
    if (fan < a.nTargets) { T[0] = a.targets[3 * fan]; T[1] = a.targets[3 * fan + 1]; T[2] = a.targets[3 * fan + 2]; }
    else { T[0] = a.lx; T[1] = a.ly; T[2] = a.lz; }
    for (int g = tid; g < nPow2; g += 1024) {
        unsigned long long key = ~0ull;
        if (g < nc) {
            const float4 l4 = a.boxLo[g], h4 = a.boxHi[g];
            const float dx = fmaxf(fmaxf(l4.x - T[0], T[0] - h4.x), 0.0f), dy = fmaxf(fmaxf(l4.y - T[1], T[1] - h4.y), 0.0f),
                        dz = fmaxf(fmaxf(l4.z - T[2], T[2] - h4.z), 0.0f);
            float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (!(d2 >= 0.0f)) d2 = 0.0f;
            const uint32_t type = g < a.ns ? 0u : (g < a.ns + a.na ? 1u : 2u);
            key = ((unsigned long long)((type << 30) | (__float_as_uint(d2) >> 2)) << 32) | (uint32_t)g;
        }
        sKeys[g] = key;
    }
    __syncthreads();
    for (int k = 2; k <= nPow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < nPow2; i += 1024) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long x = sKeys[i], y = sKeys[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { sKeys[i] = y; sKeys[p] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int g = tid; g < nc; g += 1024) a.order[(size_t)fan * nc + g] = (uint32_t)(sKeys[g] & 0xFFFFFFFFull);
}

// ---- step 1: projection --------------------------------------------------------------------------
// One CTA per (goal, face): the colliders are swept in the goal's order, 1,024 at a time; every thread projects one
// collider's conservative box onto the face and the non-empty rectangles are compacted IN ORDER into the face's rectangle
// list in global scratch (8 B each: rectangle, local collider index | type << 16). The CTA of face 0 also collects the
// goal's near list.
__global__ void __launch_bounds__(1024, 2) fan_project_kernel(const FanBuildArgs a)
{
    constexpr int kThreads = 1024;
    __shared__ int sWarpCnt[32], sWarpNear[32];
    __shared__ unsigned int sTypeCnt[2];     // sphere / AABB rectangles of the face (the list is grouped by type, like the sweep)
    const int face = blockIdx.x, fan = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ltMask = (1u << lane) - 1u;
    const int nc = a.ns + a.na + a.no;
    float T[3];
    if (fan < a.nTargets) { T[0] = a.targets[3 * fan]; T[1] = a.targets[3 * fan + 1]; T[2] = a.targets[3 * fan + 2]; }
    else { T[0] = a.lx; T[1] = a.ly; T[2] = a.lz; }
    const int k = face >> 1;
    const bool neg = face & 1;
    const bool nearCta = face == 0;
    uint2* out = a.rects + ((size_t)fan * 6 + face) * nc;
    float* depthOut = a.rectDepth + ((size_t)fan * 6 + face) * nc;
    uint32_t* nearOut = a.nearList + (size_t)fan * kFanMaxNear;
    int total = 0, nearTotal = 0;
    if (tid < 2) sTypeCnt[tid] = 0u;
    __syncthreads();
    for (int base = 0; base < nc; base += kThreads) {
        uint32_t rect = kRectEmpty, idT = 0;
        bool near = false;
        float depth = 0.0f;
        int type = 2;
        if (base + tid < nc) {
            const int g = a.order ? (int)a.order[(size_t)fan * nc + base + tid] : base + tid;
            int id; short owner;
            if (g < a.ns) { type = 0; id = g; owner = a.ownS[id]; }
            else if (g < a.ns + a.na) { type = 1; id = g - a.ns; owner = a.ownA[id]; }
            else { type = 2; id = g - a.ns - a.na; owner = a.ownO[id]; }
            if (!(fan < a.nTargets && (int)owner == fan)) {          // RT:413/426/439, PM:235/245/255
                const float4 l4 = a.boxLo[g], h4 = a.boxHi[g];
                const float lo[3] = { l4.x, l4.y, l4.z }, hi[3] = { h4.x, h4.y, h4.z };
                rect = fan_rect(lo, hi, T, k, neg, a.nearDist, near, depth);
                idT = (uint32_t)id | ((uint32_t)type << 16);
            }
        }
        const uint32_t bal = __ballot_sync(kFull, rect != kRectEmpty);
        const uint32_t nbal = __ballot_sync(kFull, near && nearCta);
        const uint32_t balS = __ballot_sync(kFull, rect != kRectEmpty && type == 0), balA = __ballot_sync(kFull, rect != kRectEmpty && type == 1);
        if (lane == 0) {
            sWarpCnt[warp] = __popc(bal); sWarpNear[warp] = __popc(nbal);
            if (balS) atomicAdd(&sTypeCnt[0], (unsigned)__popc(balS));
            if (balA) atomicAdd(&sTypeCnt[1], (unsigned)__popc(balA));
        }
        __syncthreads();
        int incl = sWarpCnt[lane], nincl = sWarpNear[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, s), nv = __shfl_up_sync(kFull, nincl, s);
            if (lane >= s) { incl += v; nincl += nv; }
        }
        const int M = __shfl_sync(kFull, incl, 31), nM = __shfl_sync(kFull, nincl, 31);
        const int prefix = warp == 0 ? 0 : __shfl_sync(kFull, incl, warp - 1);
        const int nprefix = warp == 0 ? 0 : __shfl_sync(kFull, nincl, warp - 1);
        if (rect != kRectEmpty) {
            const int pos = total + prefix + __popc(bal & ltMask);
            out[pos] = make_uint2(rect, idT);
            depthOut[pos] = depth;
        }
        if (near && nearCta) {
            const int pos = nearTotal + nprefix + __popc(nbal & ltMask);
            if (pos < kFanMaxNear) nearOut[pos] = idT;
        }
        total += M; nearTotal += nM;
        __syncthreads();
    }
    if (tid == 0) {
        uint32_t* rc = a.rectCount + (size_t)(fan * 6 + face) * 3;
        rc[0] = (uint32_t)total; rc[1] = sTypeCnt[0]; rc[2] = sTypeCnt[0] + sTypeCnt[1];   // all | spheres | spheres + AABBs
        if (nearCta) a.nearCount[fan] = (uint32_t)nearTotal;       // (may exceed kFanMaxNear: the match step flags it)
    }
}

// ---- step 2: matching ----------------------------------------------------------------------------
// One CTA per (goal, face, strip of ROWS bin rows), one warp per bin row, one thread per bin; the warps do not synchronise
// while they walk the face's rectangle list: 32 rectangles per coalesced load, the warp picks the ones that overlap its row
// and every lane checks its column against those only (in order). Pass 0 counts, a block scan + one atomicAdd reserves the
// CTA's span of the entry array, pass 1 walks the list again and writes the indices, the first AABB ids and the covering
// depth. (Round 2: the sweep used to be part of this kernel, with a __syncthreads per 1,024 colliders -- rows with little to
// match waited at every barrier for the longest one, and cutting a face into strips repeated the whole sweep per strip. With
// the projection done once per face, strips of 8 rows are finer work units, and with per-lane masks the per-entry work runs once
// per entry of a bin instead of once per (row, rectangle) pair: C3's build 0.50 -> 0.35 ms, C4's 257 goals 1.03 -> 0.78 ms, C5's
// 9 goals x 16,384 colliders -- formerly 54 CTAs on 148 SMs -- 1.49 -> 0.87 ms, of which the one-CTA-per-goal sort is 0.20 and
// the longest row's walk through its face's list most of the rest. Prefetching the next 32 rectangles changed nothing.)
#ifndef ART_FAN_MATCH_THREADS_PER_SM
#define ART_FAN_MATCH_THREADS_PER_SM 1024     // 60 registers, no spills: measured faster than 2048 (32 registers, spills) on C3, C4 and C5
#endif
template <int ROWS>
__global__ void __launch_bounds__(ROWS * 32, ART_FAN_MATCH_THREADS_PER_SM / (ROWS * 32)) fan_match_kernel(const FanBuildArgs a)
{
    constexpr int kFanRows = ROWS, kFanThreads = ROWS * 32, kFanParts = kFanBins / ROWS;
    static_assert(kFanCellsPerFace == 1024 && kFanBins == 32 && kFanBins % kFanRows == 0, "one thread per bin, one warp per bin row");
    __shared__ int sWarpTot[32];
    __shared__ unsigned int sBase;
    __shared__ uint2 sStep[ROWS][32];        // pass 1: the 32 rectangles of the warp's current step
    __shared__ float sDepth[ROWS][32];       //         and the depths of their near faces

    const int face = blockIdx.x / kFanParts, part = blockIdx.x % kFanParts, fan = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ia = (uint32_t)lane, ib = (uint32_t)(part * kFanRows + warp);
    const int cellOfThread = face * kFanCellsPerFace + (int)ib * kFanBins + lane;
    const int nc = a.ns + a.na + a.no;
    float T[3];
    if (fan < a.nTargets) { T[0] = a.targets[3 * fan]; T[1] = a.targets[3 * fan + 1]; T[2] = a.targets[3 * fan + 2]; }
    else { T[0] = a.lx; T[1] = a.ly; T[2] = a.lz; }
    const int k = face >> 1;
    const bool neg = face & 1;
    const bool nearCta = face == 0 && part == 0;
    const uint2* rects = a.rects + ((size_t)fan * 6 + face) * nc;
    const float* rectDepth = a.rectDepth + ((size_t)fan * 6 + face) * nc;
    const uint32_t* rc = a.rectCount + (size_t)(fan * 6 + face) * 3;
    const int M = (int)rc[0], mS = (int)rc[1], mSA = (int)rc[2];
    const int nearTotal = nearCta ? (int)a.nearCount[fan] : 0;
    const uint32_t* nearList = a.nearList + (size_t)fan * kFanMaxNear;

    int cS = 0, cA = 0, cO = 0;
    unsigned int wpos = 0;                   // pass 1: next entry of this bin
    unsigned int aPos = 0;                   // pass 1: where this bin's AABB entries start
    uint32_t firstIds = 0;                   // pass 1: first two AABB entries of this bin, id0 | id1 << 16
    uint2 myCell = make_uint2(0u, 0u);       // this bin's header (written after pass 0, repeated in cells4 after pass 1)
    float cover[kFanSub * kFanSub];          // pass 1: smallest threshold depth of an AABB that covers sub-bin (sa, sb) of this bin
#pragma unroll
    for (int q = 0; q < kFanSub * kFanSub; q++) cover[q] = __int_as_float(0x7F800000);
    unsigned int blockTotal = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (int c0 = 0; c0 < M; c0 += 32) {
            const uint2 re = c0 + lane < M ? __ldg(&rects[c0 + lane]) : make_uint2(kRectEmpty, 0u);
            const uint32_t r = re.x;
            uint32_t rows = __ballot_sync(kFull, ib >= ((r >> 16) & 255u) / kFanSub && ib <= (r >> 24) / kFanSub && (r & 255u) <= ((r >> 8) & 255u));
            if (!rows) continue;
            // every lane's own mask of the step's rectangles that contain its bin: a short uniform loop over the rectangles
            // that overlap the row, full width (the per-entry work below then runs once per ENTRY OF A BIN, not once per
            // (row, rectangle) pair with the few lanes of the rectangle's columns)
            uint32_t mine = 0;
            while (rows) {
                const int j = __ffs(rows) - 1;
                rows &= rows - 1;
                const uint32_t rj = __shfl_sync(kFull, r, j);
                if (ia >= (rj & 255u) / kFanSub && ia <= ((rj >> 8) & 255u) / kFanSub) mine |= 1u << j;
            }
            if (pass == 0) {
                // the list is grouped by type like the sweep: entries [0, mS) spheres, [mS, mSA) AABBs, then OBBs
                const uint32_t belowS = mS <= c0 ? 0u : (mS - c0 >= 32 ? 0xFFFFFFFFu : (1u << (mS - c0)) - 1u);
                const uint32_t belowSA = mSA <= c0 ? 0u : (mSA - c0 >= 32 ? 0xFFFFFFFFu : (1u << (mSA - c0)) - 1u);
                cS += __popc(mine & belowS); cA += __popc(mine & belowSA & ~belowS); cO += __popc(mine & ~belowSA);
                continue;
            }
            __syncwarp();
            sStep[warp][lane] = re;
            sDepth[warp][lane] = c0 + lane < M ? __ldg(&rectDepth[c0 + lane]) : 0.0f;
            __syncwarp();
            while (mine) {
                const int j = __ffs(mine) - 1;
                mine &= mine - 1;
                const uint32_t rj = sStep[warp][j].x, e = sStep[warp][j].y;
                if (wpos == aPos) firstIds = (e & 0xFFFFu) | (e << 16);          // (a single AABB is listed twice)
                else if (wpos == aPos + 1u) firstIds = (firstIds & 0xFFFFu) | (e << 16);
                a.entries[wpos++] = (uint16_t)(e & 0xFFFFu);
                // A sub-bin on the border of the rectangle holds the edge of the projection at the depth where it is widest,
                // so it is covered at no depth -- unless the rectangle was clipped at the edge of the face: only the other
                // sub-bins do the interval arithmetic (none at all for most small or distant colliders).
                // (... and a box whose near face lies beyond the covering depths the sub-bins already have cannot lower any of them:
                // the sweep is nearest first, so that is most boxes once a bin is covered)
                uint32_t open = 0;                                     // sub-bins this box could still lower
                if ((e >> 16) == 1u && a.aabbA) {
                    const float d = sDepth[warp][j] * 1.001953125f;
#pragma unroll
                    for (int q = 0; q < kFanSub * kFanSub; q++) open |= (d < cover[q] ? 1u : 0u) << q;
                }
                if (open) {
                    const uint32_t fa0 = rj & 255u, fa1 = (rj >> 8) & 255u, fb0 = (rj >> 16) & 255u, fb1 = rj >> 24;
                    uint32_t candA = 0, candB = 0;
#pragma unroll
                    for (uint32_t q = 0; q < (uint32_t)kFanSub; q++) {
                        const uint32_t fa = ia * kFanSub + q, fb = ib * kFanSub + q;
                        if ((fa > fa0 || (fa == fa0 && fa0 == 0u)) && (fa < fa1 || (fa == fa1 && fa1 == (uint32_t)kFanFine - 1u))) candA |= 1u << q;
                        if ((fb > fb0 || (fb == fb0 && fb0 == 0u)) && (fb < fb1 || (fb == fb1 && fb1 == (uint32_t)kFanFine - 1u))) candB |= 1u << q;
                    }
                    // (candidate sub-bins as a 2 x 2 mask, bit sb * kFanSub + sa)
                    static_assert(kFanSub == 2, "mask arithmetic below");
                    open &= ((candB & 1u) ? candA : 0u) | ((candB & 2u) ? candA << 2 : 0u);
                    if (!open) candA = 0;
                    if (candA && candB) {
                        const float4 A = __ldg(&a.aabbA[e & 0xFFFFu]);
                        const float2 B = __ldg(&a.aabbB[e & 0xFFFFu]);
                        const float elo[3] = { A.x, A.y, A.z }, ehi[3] = { A.w, B.x, B.y };
                        float c[6];
                        fan_cover_params(elo, ehi, T, k, neg, a.nearDist, c);
                        if (c[0] <= c[1]) {
                            // the depth interval per sub-bin column / row (the constraints separate), then their intersections
                            float a1[kFanSub], a2[kFanSub], b1[kFanSub], b2[kFanSub];
#pragma unroll
                            for (int q = 0; q < kFanSub; q++) {
                                const float aLo = -1.0f + (float)(ia * kFanSub + q) * (2.0f / kFanFine) - kFanCoverTan;
                                const float aHi = -1.0f + (float)(ia * kFanSub + q + 1u) * (2.0f / kFanFine) + kFanCoverTan;
                                const float bLo = -1.0f + (float)(ib * kFanSub + q) * (2.0f / kFanFine) - kFanCoverTan;
                                const float bHi = -1.0f + (float)(ib * kFanSub + q + 1u) * (2.0f / kFanFine) + kFanCoverTan;
                                a1[q] = c[0]; a2[q] = c[1]; b1[q] = c[0]; b2[q] = c[1];
                                fan_cover_clip(aHi, c[3], a1[q], a2[q]);      //  a * z <= xh for every a <= aHi
                                fan_cover_clip(-aLo, -c[2], a1[q], a2[q]);    //  a * z >= xl for every a >= aLo
                                fan_cover_clip(bHi, c[5], b1[q], b2[q]);
                                fan_cover_clip(-bLo, -c[4], b1[q], b2[q]);
                            }
                            // (the clipped bounds carry the 2-ulp error of the fast division: far inside the margins)
#pragma unroll
                            for (int sb = 0; sb < kFanSub; sb++)
#pragma unroll
                                for (int sa = 0; sa < kFanSub; sa++) {
                                    const float w1 = fmaxf(a1[sa], b1[sb]), w2 = fminf(a2[sa], b2[sb]);
                                    if ((open >> (sb * kFanSub + sa) & 1u) && w2 >= w1 * 1.00390625f && w2 - w1 >= a.coverMinThickness)
                                        cover[sb * kFanSub + sa] = fminf(cover[sb * kFanSub + sa], w1 * 1.001953125f);
                                }
                        }
                    }
                }
            }
        }
        if (pass == 1) {
            if (cA > 0) a.firstA[(size_t)fan * kFanCells + cellOfThread] = firstIds;
            uint32_t codes = 0;
#pragma unroll
            for (int q = 0; q < kFanSub * kFanSub; q++) codes |= fan_cover_code(cover[q], a.coverLogS, a.coverLogK) << (8 * q);
            a.cells4[(size_t)fan * kFanCells + cellOfThread] = make_uint4(myCell.x, myCell.y, firstIds, codes);
            break;
        }
        // ---- reserve the CTA's span: block scan of the per-bin totals
        const int tot = cS + cA + cO;
        bool bad = cS > kGridMaxS || cA > kGridMaxA || cO > kGridMaxO || nearTotal > kFanMaxNear;
        int incl = tot;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, incl, s); if (lane >= s) incl += v; }
        if (lane == 31) sWarpTot[warp] = incl;
        __syncthreads();
        int wincl = lane < kFanRows ? sWarpTot[lane] : 0;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) { const int v = __shfl_up_sync(kFull, wincl, s); if (lane >= s) wincl += v; }
        blockTotal = (unsigned)__shfl_sync(kFull, wincl, 31);
        const unsigned binOff = (unsigned)((warp == 0 ? 0 : __shfl_sync(kFull, wincl, warp - 1)) + incl - tot);
        bad = __syncthreads_or(bad);
        if (tid == 0) {
            unsigned int b = 0xFFFFFFFFu;
            if (!bad) {
                b = atomicAdd(&a.ctl[0], blockTotal + (unsigned)nearTotal);
                if ((unsigned long long)b + blockTotal + (unsigned)nearTotal > (unsigned long long)a.capacity) b = 0xFFFFFFFFu;
            }
            if (b == 0xFFFFFFFFu) atomicExch(&a.ctl[1], 1u);          // the host re-runs the frame without fans
            sBase = b;
        }
        __syncthreads();
        const unsigned int bs = sBase;
        uint2* cells = a.cells + (size_t)fan * kFanCells;
        if (bs == 0xFFFFFFFFu) {
            cells[cellOfThread] = make_uint2(0u, 0u);
            a.cells4[(size_t)fan * kFanCells + cellOfThread] = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
            if (nearCta && tid == 0) {
                cells[6 * kFanCellsPerFace] = make_uint2(0u, 0u);
                a.cells4[(size_t)fan * kFanCells + 6 * kFanCellsPerFace] = make_uint4(0u, 0u, 0u, 0xFFFFFFFFu);
            }
            return;
        }
        wpos = bs + binOff;
        aPos = wpos + (unsigned)cS;
        myCell = make_uint2(wpos, (uint32_t)cS | ((uint32_t)cA << 10) | ((uint32_t)cO << 21));
        cells[cellOfThread] = myCell;
        if (nearCta && tid == 0) {
            uint32_t nS = 0, nA = 0, nO = 0, ids = 0, ids2 = 0;
            for (int q = 0; q < nearTotal; q++) {
                const uint32_t nq = nearList[q];
                const uint32_t t = nq >> 16;
                if (t == 1) {
                    if (nA == 0) ids = (nq & 0xFFFFu) | (nq << 16);
                    else if (nA == 1) ids = (ids & 0xFFFFu) | (nq << 16);
                    else if (nA == 2) ids2 = (nq & 0xFFFFu) | (nq << 16);
                    else if (nA == 3) ids2 = (ids2 & 0xFFFFu) | (nq << 16);
                }
                nS += t == 0; nA += t == 1; nO += t == 2;
            }
            const uint2 nc2 = make_uint2(bs + blockTotal, nS | (nA << 10) | (nO << 21));
            cells[6 * kFanCellsPerFace] = nc2;
            a.firstA[(size_t)fan * kFanCells + 6 * kFanCellsPerFace] = ids;
            a.cells4[(size_t)fan * kFanCells + 6 * kFanCellsPerFace] = make_uint4(nc2.x, nc2.y, ids, ids2);
        }
        if (nearCta)
            for (int q = tid; q < nearTotal; q += kFanThreads) a.entries[bs + blockTotal + q] = (uint16_t)(nearList[q] & 0xFFFFu);
    }
}

// a.order (may be null: canonical order) needs (nTargets + 1) * (ns + na + no) words; a.rects etc.: fan_build_scratch_bytes
size_t fan_build_scratch_bytes(int nFans, int nc)
{
    return (size_t)nFans * 6 * nc * (sizeof(uint2) + sizeof(float)) + (size_t)nFans * (18 + kFanMaxNear + 1) * sizeof(uint32_t);
}
// carves a.rects / rectCount / nearList / nearCount out of `scratch` (fan_build_scratch_bytes)
void fan_build_set_scratch(FanBuildArgs& a, void* scratch)
{
    const size_t nFans = (size_t)a.nTargets + 1, nc = (size_t)a.ns + a.na + a.no;
    a.rects = reinterpret_cast<uint2*>(scratch);
    a.rectDepth = reinterpret_cast<float*>(a.rects + nFans * 6 * nc);
    a.rectCount = reinterpret_cast<uint32_t*>(a.rectDepth + nFans * 6 * nc);
    a.nearList = a.rectCount + nFans * 18;
    a.nearCount = a.nearList + nFans * kFanMaxNear;
}

cudaError_t launch_fan_build(const FanBuildArgs& a, cudaStream_t stream)
{
    const int nFans = a.nTargets + 1;
    const int nc = a.ns + a.na + a.no;
    if (a.order) {
        int nPow2 = 2;
        while (nPow2 < nc) nPow2 <<= 1;
        const size_t smem = (size_t)nPow2 * sizeof(unsigned long long);
        cudaError_t e = cudaFuncSetAttribute(fan_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        fan_order_kernel<<<nFans, 1024, smem, stream>>>(a, nPow2);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    fan_project_kernel<<<dim3(6, nFans), 1024, 0, stream>>>(a);
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return e2;
    // strips of 8 bin rows unless ART_FAN_ROWS (experiment knob, read per launch) says otherwise
    int rows = 8;
    if (const char* v = getenv("ART_FAN_ROWS")) { const int r = atoi(v); if (r == 4 || r == 8 || r == 16 || r == 32) rows = r; }
    switch (rows) {
    case 32: fan_match_kernel<32><<<dim3(6, nFans), 1024, 0, stream>>>(a); break;
    case 16: fan_match_kernel<16><<<dim3(12, nFans), 512, 0, stream>>>(a); break;
    case 8: fan_match_kernel<8><<<dim3(24, nFans), 256, 0, stream>>>(a); break;
    default: fan_match_kernel<4><<<dim3(48, nFans), 128, 0, stream>>>(a); break;
    }
    return cudaGetLastError();
}

}  // namespace art
