// scene_dev.cuh -- HBM / shared-memory layout of the packed collider scene and the kernel
// argument blocks shared by K0 (pack), K1 (trace), K2 (permeation), K3 (reduce).
//
// The reference keeps colliders as half-precision AoS structs (ColliderAABBStruct 20 B,
// ColliderOBBStruct 26 B, ColliderSphereStruct 16 B) and re-derives min/max, R^2 and the
// normalised quaternion on every test (AudioRaytracerJobBatched.cs:286-287, 328;
// ColliderOBBStruct.cs:14-16 -> halfQuaternion.cs:34-46). K0 evaluates those pure per-collider
// functions ONCE, with the reference's own operation order, into FP32 structure-of-arrays
// "planes" laid out so that lane l of a warp reads collider (base + r*32 + l) with conflict-free
// 128/64-bit shared loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace art {

#ifndef ART_WARPS
#define ART_WARPS 32
#endif
constexpr int kWarpsPerCta = ART_WARPS;
constexpr int kThreads = kWarpsPerCta * 32;
// colliders held in registers per lane per "super-chunk"
constexpr int RS = 4;   // spheres
constexpr int RA = 4;   // AABBs
constexpr int RO = 2;   // OBBs
constexpr int SC_S = 32 * RS, SC_A = 32 * RA, SC_O = 32 * RO;

constexpr float kEpsilon = 0.0001f;          // RT:57, PM:30
constexpr float kFloatMax = 3.402823466e+38f; // float.MaxValue (RT:230)

// Geometry blob: one contiguous, 16-byte aligned buffer (global memory; K1/K2 bulk-copy it to
// shared memory with cp.async.bulk when it fits). Counts are padded to a multiple of the
// super-chunk with duplicates of the last real collider of that type (a duplicate can never
// change a nearest-hit winner -- lower index wins ties -- nor an any-hit result; K2 zeroes the
// density of pads).
struct GeomLayout {
    int ns, na, no;                 // real counts
    int nsPad, naPad, noPad;        // padded counts
    uint32_t offSph;                // float4 (cx, cy, cz, R*R)                      [nsPad]
    uint32_t offAabbA;              // float4 (min.x, min.y, min.z, max.x)          [naPad]
    uint32_t offAabbB;              // float2 (max.y, max.z)                        [naPad]
    uint32_t offObbQ;               // float4 Rotation getter value (x,y,z,w)       [noPad]
    uint32_t offObbC;               // float4 (cx, cy, cz, |hx|)                    [noPad]
    uint32_t offObbH;               // float2 (|hy|, |hz|)                          [noPad]
    uint32_t bytes;                 // total, multiple of 16
};

// Per-collider attributes that are only touched once per segment (hit path): global/L2 only.
struct AttrArrays {
    const float4* sphAttr;   // (absorption, echo, density, owner-as-int-bits)   [nsPad]
    const float4* aabbAttr;  //                                                  [naPad]
    const float4* obbAttr;   //                                                  [noPad]
    const float4* aabbCtr;   // (cx, cy, cz, 0)  raw Center                      [naPad]
    const float4* aabbHalf;  // (hx, hy, hz, 0)  raw Size (half extents)         [naPad]
    const float4* obbHalf;   // (hx, hy, hz, 0)  raw Size                         [noPad]
    const float4* obbQinv;   // math.inverse(Rotation)  (RT:489, PM:174)         [noPad]
    const short*  ownS;      // AudioTargetId per collider                       [nsPad]
    const short*  ownA;      //                                                  [naPad]
    const short*  ownO;      //                                                  [noPad]
    const int*    ownedCount;// [3][nTargetsCap] colliders of each type owned by target a (counters only)
};

// One plane of the blob: base pointer + byte offset. Every plane of a view shares the base, and the offsets are kernel
// parameters (constant bank), so a view costs the kernels one address register pair instead of six.
template <class T>
struct PlaneRef {
    const unsigned char* base;
    uint32_t off;
    __host__ __device__ __forceinline__ const T& operator[](int i) const { return reinterpret_cast<const T*>(base + off)[i]; }
};

struct GeomView {
    PlaneRef<float4> sph; PlaneRef<float4> aabbA; PlaneRef<float2> aabbB;
    PlaneRef<float4> obbQ; PlaneRef<float4> obbC; PlaneRef<float2> obbH;
};

__host__ __device__ __forceinline__ GeomView make_view(const unsigned char* base, const GeomLayout& L)
{
    GeomView v;
    v.sph = { base, L.offSph };
    v.aabbA = { base, L.offAabbA };
    v.aabbB = { base, L.offAabbB };
    v.obbQ = { base, L.offObbQ };
    v.obbC = { base, L.offObbC };
    v.obbH = { base, L.offObbH };
    return v;
}

// Device-side work counters (u64 each), index constants.
enum CounterIdx {
    C_SEGMENTS = 0, C_SEGMENT_HITS,
    C_TRACE_S, C_TRACE_A, C_TRACE_O,
    C_ECHO_Q, C_ECHO_S, C_ECHO_A, C_ECHO_O,
    C_MUFFLE_Q, C_MUFFLE_S, C_MUFFLE_A, C_MUFFLE_O,
    C_PERM_RAYS, C_PERM_HIT_RAYS, C_PERM_FIRST_S, C_PERM_FIRST_A, C_PERM_FIRST_O,
    C_PERM_PAIRS, C_PERM_LOSS_S, C_PERM_LOSS_A, C_PERM_LOSS_O,
    // tests the grid kernels actually executed (ART_FRAME_GRID_STATS): trace job, permeation first hit, permeation loss
    C_GRID_RT_S, C_GRID_RT_A, C_GRID_RT_O,
    C_GRID_PF_S, C_GRID_PF_A, C_GRID_PF_O,
    C_GRID_PL_S, C_GRID_PL_A, C_GRID_PL_O,
    C_GRID_RT_CELLS, C_GRID_PM_CELLS,
    C_GRID_Q_S, C_GRID_Q_A, C_GRID_Q_O, C_GRID_Q_LISTS,   // ... by query_fan_kernel (echo / muffle queries against the target fans)
    C_DEBUG_VIOLATIONS,        // -DART_DEBUG_BOUNDS builds: index checks that failed in the grid kernels (must stay 0)
    C_COUNT
};

// How this context's local rays map to global ray indices (art_set_ray_shard).
struct ShardMap {
    int nGlobal;      // RayDirections.Length
    int nLocal;       // rays traced by this context
    int shardIndex, shardCount, chunk;
    int dirsLocal;    // the device direction array holds only this context's rays, in local order (else all nGlobal rays)
    __host__ __device__ inline size_t dir_index(int j, int rayIndex) const { return (size_t)(dirsLocal ? j : rayIndex); }
    __host__ __device__ inline int to_global(int j) const
    {
        if (shardCount <= 1) return j;
        return ((j / chunk) * shardCount + shardIndex) * chunk + (j % chunk);
    }
};

struct TraceArgs {
    const unsigned char* geom;     // geometry blob (global)
    GeomLayout L;
    AttrArrays at;
    const uint16_t* dirs;          // half3 [nGlobal]
    ShardMap map;
    float ox, oy, oz;              // RayOrigin
    const float* targets;          // float3 [nTargets]
    const int* targetOrder;        // grid kernels: spatially sorted permutation of the targets (coherent lanes), or null
    int nTargets;
    float maxRayLife;
    int H;                         // MaxHitsPerRay
    float maxMuffle;               // MaxMuffleHitDistance
    int batchSize;                 // ART:161
    // outputs, local ray indexing
    uint16_t* echo;                // half  [nLocal*H]
    uint16_t* hitPoints;           // half3 [nLocal*H] or null
    uint8_t*  hitCounts;           // [nLocal] or null
    uint32_t* hitIds;              // [nLocal*H] or null
    uint32_t* muffleCounts;        // u32 [T*Na], row = batch index k = rayIndex / batchSize
    unsigned long long* counters;  // [C_COUNT]
    unsigned int* nextRay;         // dynamic ray queue
    uint32_t* scratch;             // grid kernel: per-warp survivor lists (trace_grid_scratch_bytes)
    int raysPerWarp;               // grid kernel: lanes of a warp that own a ray (32 unless the batch is too small to fill the GPU)
    int gridWarps;                 // grid kernel: warps per CTA of this launch (<= ART_GRID_WARPS, trace_grid_plan)
    // grid kernel, group rotation (small batches, trace_grid_plan): rays are dealt in fixed groups of 32 and a group changes
    // warps after every bounce round through a log of parked group states, so that all warps advance all groups evenly
    int migGroups;                 // > 0: rotation on, number of groups = ceil(nLocal / 32)
    unsigned int migSlots;         // capacity of the log: migGroups * (H - 1) parked states
    unsigned int* migFlags;        // [migSlots] 0 = not yet published (zeroed per frame)
    float4* migState;              // [migSlots][64]: per lane (o.xyz, life), (d.xyz, hits | alive << 8 | group << 9)
    int goalsInSmem;               // grid kernel, FAN 2: listener + target positions staged in shared memory (launch_trace_grid)
    unsigned int goalsSmemOffset;  //   byte offset of those tables in the dynamic shared memory
    int muffleInSmem;              // per-warp shared counters fit
    int anyOwned[3];               // does any sphere / AABB / OBB belong to a target < nTargets (RT:413/426/439)
    // grid kernel, bounce-only mode (target fans in use): the echo / muffle queries of a hit point do not feed the bounce
    // loop (RT:124-173 only WRITE EchoRayDistances / MuffleRayHits), so the tracer just appends one record per hit point
    // and query_fan_kernel (k1_query_fan.cu) evaluates all of them afterwards
    float4* recA;                  // [nLocal*H] (hit - eps*d).xyz (RT:124 == RT:158), distance(RayOrigin, hit) (RT:130)
    float2* recB;                  // [nLocal*H] material Echo of the hit collider (RT:135-141), rayResultId (RT:115) as int bits
    unsigned int* recCount;        // records appended so far (zeroed per frame)
};

// k1_query_fan.cu: the echo-return ray (RT:124-145) and the Na muffle rays (RT:153-173) of every hit point the bounce
// tracer recorded, evaluated against the target fans (fan_dev.cuh)
struct QueryArgs {
    const unsigned char* geom;     // geometry blob (global)
    GeomLayout L;
    const float4* recA; const float2* recB; const unsigned int* recCount;   // hit records (TraceArgs)
    ShardMap map;
    int H, batchSize;
    float ox, oy, oz;              // RayOrigin = goal of slot 0 (echo ray)
    const float* targets;          // float3 [nTargets] = goals of slots 1..Na (muffle rays)
    int nTargets;
    float maxMuffle;               // MaxMuffleHitDistance (RT:168)
    float errScale;                // GridDesc::errScale (conservative OBB pre-tests)
    uint16_t* echo;                // half [nLocal*H]
    uint32_t* muffleCounts;        // u32 [T*Na], row = batch index of the ray
    int muffleRows;                // T
    unsigned long long* counters;  // [C_COUNT]
    unsigned int* queue;           // next block of 32 records (zeroed per frame)
    float4* scratch;               // per-warp survivor lists (query_fan_scratch_bytes)
    int tablesInSmem;              // goal positions + near-list headers of all slots staged in shared memory
    int muffleInSmem;              // per-CTA muffle counters [T*Na] in shared memory
    int firstTests;                // AABBs every query the cull leaves is tested against at once, q_first (1 or 2)
    int goalGroups, goalsPerGroup; // > 1 group: the goals of a record block are split over several warps (small batches)
};

// Uniform grid over the collider scene (acceleration structure). Built on the host at
// scene upload (grid_host.h); cell (ix,iy,iz) lists the canonical indices of every collider whose
// conservatively inflated bounds overlap it, grouped spheres | AABBs | OBBs. The kernels walk a ray
// through the cells (3D-DDA) and run the SAME exact per-collider tests as the brute-force kernels on
// the listed colliders only; since the nearest-hit rule is a (t, canonical index) minimum and the
// occlusion rule an "any", the results do not depend on which non-hitting colliders were skipped.
struct GridDesc {
    float g0x, g0y, g0z;          // min corner
    float g1x, g1y, g1z;          // max corner
    float csx, csy, csz;          // cell size
    float icx, icy, icz;          // 1 / cell size
    int nx, ny, nz;
    float errScale;               // >= any distance a ray travels inside the scene (conservative pre-tests)
    int nEntries;                 // length of entries (bounds checks of debug builds)
    const uint2* cells;           // [nz*ny*nx]  x = first entry, y = nS | nA << 10 | nO << 21
    const uint16_t* entries;      // collider indices, per cell: spheres, AABBs, OBBs
    const uint2* rangeO;          // per OBB: cell range it is listed in, x = ix0 | iy0 << 8 | iz0 << 16, y = ix1 | iy1 << 8 | iz1 << 16
};
constexpr int kGridMaxS = 1023, kGridMaxA = 2047, kGridMaxO = 2047;

// k5_grid_build.cu: the grid's cell lists, filled on the device from the colliders' conservative boxes
struct GridBuildArgs {
    const float4* boxLo;          // [ns + na + no] conservative bounds, canonical order S | A | O (grid_host.h: grid_params)
    const float4* boxHi;
    int ns, na, no;
    float g0x, g0y, g0z;          // grid min corner
    float csx, csy, csz;          // cell size
    int nx, ny, nz;
    unsigned int* cnt;            // [nx*ny*nz * 3] scratch: counts, then write cursors
    uint2* cells;                 // [nx*ny*nz]
    uint16_t* entries;
    unsigned int capacity;        // entries available
    unsigned int* ctl;            // [0] entries used, [1] overflow flag (cell list too long for the header format / buffer too small)
};

// K0 arguments
struct PackArgs {
    const uint16_t* rawS;   // ColliderSphereStruct[ns] as 8 x u16
    const uint16_t* rawA;   // ColliderAABBStruct[na]  as 10 x u16
    const uint16_t* rawO;   // ColliderOBBStruct[no]   as 13 x u16
    GeomLayout L;
    unsigned char* geom;
    float4* sphAttr; float4* aabbAttr; float4* obbAttr;
    float4* aabbCtr; float4* aabbHalf; float4* obbHalf; float4* obbQinv;
    short* ownS; short* ownA; short* ownO;
};

// K2 arguments
struct PermArgs {
    const unsigned char* geom;
    GeomLayout L;
    AttrArrays at;
    const float* densS; const float* densA; const float* densO;   // padded; 0 for pads and for colliders owned by a target < Na
    const float* trueDens;     // padded [S | A | O]: density of every collider, owned or not (target-fan path, geometry not in smem)
    const int* ownedList;      // (section << 28) | index of every collider owned by a target < Na, canonical order (host built)
    int nOwned;
    const uint16_t* dirs;
    ShardMap map;
    float ox, oy, oz;
    const float* targets;
    const int* targetOrder;    // grid kernel: spatially sorted permutation of the targets, or null
    int nTargets;
    float nTimesS;             // (float)RayDirections.Length * PermeationStrengthPerRay (PM:260)
    int batchSize;
    float* firstHitDist;       // [nLocal] first-hit distance, +Inf = no hit
    int* lastHitRay;           // [T] max global ray index with a hit, per batch (init -1)
    long long* permSumInt;     // [Na] sum over hitting rays of trunc(value)
    long long* permSumFrac;    // [Na] sum of frac(value) * 2^36
    float* permLast;           // [T*Na] values of the last hitting ray of batch k (perm_last_kernel)
    unsigned long long* counters;
    unsigned int* nextRay;
    int raysPerWarp;           // grid kernel: rays a warp takes from the queue at a time (<= 32)
    float4* hitPts;            // grid kernel, binned mode (k2_permeation_binned.cu): [nLocal] (hit point - eps*d, hit flag) is
                               // stored here and the loss lines are NOT evaluated by permeation_grid_kernel; null otherwise
};

// k2_permeation_binned.cu: the loss lines grouped by (target, direction bin of the target's fan)
struct PermBinArgs {
    const float4* hitPts;      // [nLocal] from permeation_grid_kernel
    int nLocal;
    const float* targets;      // float3 [nTargets]
    int nTargets;
    int slices;                // ray slices per target (one counting CTA each)
    uint32_t* cnt;             // [nTargets][kBinBuckets][slices]: counts, then exclusive offsets into the target's region of pairRay
    uint32_t* pairRay;         // [nTargets][nLocal]: local ray indices of the hitting rays, grouped by bin
    uint32_t* binStart;        // [nTargets][kBinBuckets]: first line of every bucket (the last, always empty one = number of lines)
    uint32_t* blockBin;        // [nTargets][ceil(nLocal / kBinBlock)]: bucket of the first line of every block of sorted lines
    unsigned int* jobQueue;    // zeroed per frame
};

// K3 output (also part of the partial-result blob)
struct EchoStats {                 // device + blob layout
    long long fixedLo;             // sum of (value*2^24) & 0xFFFFF   (signed by the half's sign)
    long long fixedHi;             // sum of (value*2^24) >> 20
    unsigned long long zeros;      // entries == 0 (PA:42)
    unsigned long long entries;    // entries visited
    unsigned long long posInf, negInf, nan;
    float seqTotal;                // reverb_seq_kernel: reverbTotal (PA:47)
    float seqZeros;                // reverb_seq_kernel: echoRayReturnedHits (PA:44)
    unsigned int seqValid;
    unsigned int pad;
};

}  // namespace art
