// fan_dev.cuh -- "target fans": per-goal, direction-binned collider lists for the queries that all end in
// one of a few fixed points -- the listener (echo-return rays, RT:124-145) and the Na audio targets (muffle
// rays RT:153-173, permeation lines PM:67-86).
//
// Every such query runs along a line through its goal G, so seen FROM G it is a single direction. For each
// goal the directions are binned on a cube map (6 faces x B x B bins); bin (face, ia, ib) lists every
// collider whose conservative bounds (the same inflated boxes the uniform grid registers, grid_host.h)
// subtend that bin as seen from G. Colliders whose bounds come within `nearDist` of G -- for which "direction
// from G" is ill-conditioned -- are kept in one extra "near" list per goal that every query of that goal
// tests. Colliders owned by target a are left out of fan a (RT:413/426/439, PM:235/245/255 skip them).
//
// A query from P towards G therefore tests  near(G) + bin(G, P - G)  (any-hit, K1) or
// near(G) + bin(G, P - G) + bin(G, G - P)  (whole line, K2) -- about a dozen colliders per list at C3, each
// exactly once, with no cell walk. The per-collider tests are the same exact FP32 functions as everywhere
// else (intersect.cuh); the lists only decide WHICH colliders are tested, and they are conservative:
//   * the exact FP32 test can report collider c at parameter t only if the real ray point X(t) lies in c's
//     inflated box (the guarantee the grid relies on, grid_host.h);
//   * the ray leaves P with a direction that is normalize(G - P) up to ~3e-7 rad, so X(t) deviates from the
//     exact line P -> G by < 1e-6 * errScale; seen from G, at |X - G| >= nearDist = 1e-3 * errScale, that is
//     < 1e-3 rad, i.e. < 2e-3 in the tangent-plane coordinates of a cube face;
//   * every collider is registered in all bins its box projects onto, inflated by kFanTanMargin = 4e-3, with
//     1 % slack on the face-membership constraint, and in the near list when G is within nearDist of the box.
// The lists are rebuilt on the device every frame (goals move): fan_order_kernel, fan_project_kernel, fan_match_kernel
// (k4_fan_build.cu).
#pragma once
#include "scene_dev.cuh"

namespace art {

constexpr int kFanBins = 32;                         // B: bins per cube-face edge (one CTA thread per bin when building)
constexpr int kFanCellsPerFace = kFanBins * kFanBins;
constexpr int kFanCells = 6 * kFanCellsPerFace + 1;  // per goal; the last cell is the near list
constexpr int kFanSub = 2;                           // sub-bins per bin edge (covering depths are kept per sub-bin, cells4.w)
constexpr int kFanFine = kFanBins * kFanSub;
constexpr float kFanCoverLogEps = 0.01f;             // slack (in code units) of the log-domain covering-depth compare, k4_fan_build.cu
constexpr float kFanTanMargin = 4e-3f;
constexpr int kFanMaxNear = 512;                     // near-list capacity per goal (else the frame falls back to the grid walk)

struct FanDesc {
    int nFans;                 // Na + 1: fan a < Na = audio target a, fan Na = the listener (RayOrigin)
    const uint2* cells;        // [nFans * kFanCells]: x = first entry, y = nS | nA << 10 | nO << 21 (as GridDesc::cells)
    const uint16_t* entries;   // collider indices per cell: spheres, AABBs, OBBs, each nearest to the goal first
    int nEntries;              // capacity (bounds checks of debug builds)
    const uint32_t* firstA;    // [nFans * kFanCells]: the cell's first two AABB entries, id0 | id1 << 16 (K1's first pass tests
                               // exactly those: it reads them beside the header instead of chasing the entry list afterwards)
    const uint4* cells4;       // [nFans * kFanCells]: (cells[i].x, cells[i].y, AABB ids 0 | 1 << 16, covering-depth codes of the bin's
                               // kFanSub x kFanSub sub-bins, one byte each) -- header, first ids and the cull thresholds of the bin
                               // in ONE 16-byte entry (k4_fan_build.cu "covering depth"; the near cell's w is unused)
    float coverLogS, coverLogK; // code of depth w = coverLogS * log2(w) + coverLogK (set by the API, the same for build and query)
};

// arguments of the fan build kernels (k4_fan_build.cu)
struct FanBuildArgs {
    const float4* boxLo;       // [ns + na + no] conservative bounds of every collider, canonical order S | A | O (grid_host.h)
    const float4* boxHi;
    int ns, na, no;
    const short* ownS; const short* ownA; const short* ownO;   // AudioTargetId per collider
    const float* targets;      // float3 [nTargets]
    int nTargets;
    float lx, ly, lz;          // listener = goal of fan nTargets
    float nearDist;
    const float4* aabbA; const float2* aabbB;   // the AABBs' own min / max (GeomView), for the covering depth; null: no covering depths
    float coverMinThickness;   // 1e-4 * errScale: a covering depth interval must be at least this long
    float coverLogS, coverLogK; // see FanDesc
    uint2* cells;              // [(nTargets + 1) * kFanCells]
    uint32_t* firstA;          // [(nTargets + 1) * kFanCells], see FanDesc
    uint4* cells4;             // [(nTargets + 1) * kFanCells], see FanDesc
    uint16_t* entries;
    unsigned int capacity;     // entries available
    unsigned int* ctl;         // [0] next free entry, [1] overflow flag (zeroed by the host before the launch)
    uint32_t* order;           // [(nTargets + 1) * (ns + na + no)] per-goal sweep order (fan_order_kernel), or null
    // scratch between the projection and the matching step (fan_build_set_scratch)
    uint2* rects;              // [(nTargets + 1) * 6][ns + na + no]: per (goal, face) the non-empty rectangles in sweep order,
                               // x = a0 | a1 << 8 | b0 << 16 | b1 << 24 (bins), y = local collider index | type << 16
    float* rectDepth;          // the same shape: depth of the inflated box's near face (no AABB covers a bin from nearer)
    uint32_t* rectCount;       // [(nTargets + 1) * 6][3]: rectangles of the face | of its spheres | of its spheres and AABBs
    uint32_t* nearList;        // [(nTargets + 1)][kFanMaxNear]: index | type << 16 of the colliders near the goal, in sweep order
    uint32_t* nearCount;       // [(nTargets + 1)] (may exceed kFanMaxNear: overflow)
};

// Cell index (within one fan) of the bin that direction v (from the goal, any length) falls in.
// Returns -1 when v has no usable direction (zero or non-finite).
// `w` receives the depth of v on its face (the largest |component|).
// `sub` receives the sub-bin (sb * kFanSub + sa) of v inside its bin: floor(x * kFanSub) / kFanSub == floor(x) for x >= 0 and the
// scaling by kFanSub = 2 is exact, so the bin is the same as without sub-bins.
__device__ __forceinline__ int fan_bin_w(float vx, float vy, float vz, float& w, int& sub)
{
    static_assert(kFanSub == 2, "exact only for powers of two");
    // branch-free: selects instead of an if-chain, the validity check folded into the result
    const float ax = fabsf(vx), ay = fabsf(vy), az = fabsf(vz);
    const bool fx = ax >= ay && ax >= az, fy = !fx && ay >= az;
    w = fx ? ax : (fy ? ay : az);
    const float s = fx ? vx : (fy ? vy : vz), p = fx ? vy : (fy ? vz : vx), q = fx ? vz : (fy ? vx : vy);
    const float r = __fdividef(1.0f, w);
    // tangent-plane coordinates a = p * r, b = q * r in [-1, 1]; (a + 1) * (kFanFine / 2) as one FMA (the same float: the
    // scaling by a power of two is exact)
    const int ia = min(kFanFine - 1, max(0, __float2int_rd(fmaf(p * r, 0.5f * kFanFine, 0.5f * kFanFine))));
    const int ib = min(kFanFine - 1, max(0, __float2int_rd(fmaf(q * r, 0.5f * kFanFine, 0.5f * kFanFine))));
    const int face = (fx ? 0 : (fy ? 2 : 4)) + (s < 0.0f ? 1 : 0);
    sub = (ib % kFanSub) * kFanSub + (ia % kFanSub);
    return (w > 0.0f && w < 3.0e38f) ? face * kFanCellsPerFace + (ib / kFanSub) * kFanBins + ia / kFanSub : -1;
}
__device__ __forceinline__ int fan_bin_w(float vx, float vy, float vz, float& w) { int sub; return fan_bin_w(vx, vy, vz, w, sub); }
__device__ __forceinline__ int fan_bin(float vx, float vy, float vz) { float w; return fan_bin_w(vx, vy, vz, w); }

}  // namespace art
