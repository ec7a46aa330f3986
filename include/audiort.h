/*
 * audiort.h -- C ABI of libaudiort_cuda, the B200 (sm_100a) native plugin that
 * replaces the CPU job path of FirePixel8422/Audio-Raytracer:
 *
 *   AudioRaytracerJobBatched   Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs   (RT)
 *   AudioPermeationJobBatched  Assets/C# Scripts/Jobs/AudioPermeationJobBatched.cs  (PM)
 *   ProcessAudioDataJob        Assets/C# Scripts/Jobs/ProcessAudioDataJob.cs        (PA)
 *
 * The reference has no FFI layer; the seam this header replaces is the block of
 * Assets/C# Scripts/Audio/AudioRayTracer.cs (ART) that fills the three job
 * structs and calls Schedule / IsCompleted / Complete (ART:95-97, 161-237,
 * 241-254). Every entry point cites the reference lines it stands in for. The
 * C# P/Invoke binding a maintainer adds is shown in INTEGRATION.md.
 *
 * Conventions: plain C, blittable PODs, caller-owned pointers, no exceptions,
 * no callbacks. Return value 0 = OK, negative = ArtStatus error. There is NO
 * CPU fallback: without a usable CUDA device art_create fails with
 * ART_E_NO_DEVICE. One context may have at most one frame in flight (the
 * reference never overlaps two frames, ART:95-97). A context is not thread
 * safe; different contexts are independent.
 */
#ifndef AUDIORT_H
#define AUDIORT_H

#include <stddef.h>
#include <stdint.h>

#if defined(_WIN32)
#  define ART_API __declspec(dllexport)
#else
#  define ART_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ART_ABI_VERSION 3

typedef enum ArtStatus {
    ART_OK          =  0,
    ART_E_ARG       = -1,   /* invalid argument (also: reference would throw, e.g. 0 targets -> RT:63 divide by zero) */
    ART_E_CUDA      = -2,   /* CUDA runtime error; sticky errors poison the context */
    ART_E_PENDING   = -3,   /* a frame is already in flight / handle not complete */
    ART_E_NO_DEVICE = -4,   /* no CUDA device (there is no CPU fallback by design) */
    ART_E_STATE     = -5,   /* scene or rays not set, stale handle */
    ART_E_NOMEM     = -6
} ArtStatus;

/* ---- wire layouts: byte-for-byte the C# sequential structs (all members 2 bytes) ---- */
#pragma pack(push, 2)
/* DataTypes/Collider Structs/ColliderAABBStruct.cs:8-14 -- 20 bytes */
typedef struct ArtAABB {
    uint16_t center[3];     /* half3 Center */
    uint16_t size[3];       /* half3 Size (half extents, RT:257) */
    uint16_t absorption;    /* AudioMaterialProperties (CS/AudioMaterialProperties.cs:7-16) */
    uint16_t density;
    uint16_t echo;
    int16_t  audioTargetId; /* -1 = not owned by a target */
} ArtAABB;
/* DataTypes/Collider Structs/ColliderOBBStruct.cs:8-24 -- 26 bytes */
typedef struct ArtOBB {
    uint16_t center[3];
    uint16_t size[3];
    uint16_t rot[3];        /* halfQuaternion x,y,z (DataTypes/halfQuaternion.cs:7-11), stored inverted */
    uint16_t absorption, density, echo;
    int16_t  audioTargetId;
} ArtOBB;
/* DataTypes/Collider Structs/ColliderSphereStruct.cs:8-14 -- 16 bytes */
typedef struct ArtSphere {
    uint16_t center[3];
    uint16_t radius;
    uint16_t absorption, density, echo;
    int16_t  audioTargetId;
} ArtSphere;
#pragma pack(pop)

/* DataTypes/AudioTargetRTSettings.cs:8-24 -- 24 bytes */
typedef struct ArtTargetSettings {
    float muffleStrength;
    float reverbStrength;
    float reverbVolume;
    float percievedAudioPosition[3];
} ArtTargetSettings;

typedef struct ArtCtx ArtCtx;
typedef int32_t ArtHandle;   /* frame ticket, stands in for Unity's JobHandle (ART:46) */

/* art_create flags */
#define ART_CREATE_DEFAULT        0u

#define ART_MAX_DEVICES 16

typedef struct ArtConfig {
    int32_t  abiVersion;     /* must be ART_ABI_VERSION */
    int32_t  device;         /* CUDA device ordinal (used when nDevices <= 1) */
    uint32_t flags;
    /* nDevices > 1: ONE context over several GPUs of this process. The library owns one device context per entry of
     * devices[], the ray shard map (interleaved chunks of shardChunkRays rays; 0 = N / (8 x devices) rays, at most 16,384), the exchange and exact merge of the
     * per-source partial results and the scatter of the per-ray outputs into the caller's arrays; every entry point keeps
     * its single-device meaning and art_trace_schedule stays one call (ART:161-237). Results are bit-identical with a
     * one-device context. */
    int32_t  nDevices;
    int32_t  devices[ART_MAX_DEVICES];
    int32_t  shardChunkRays;
    int32_t  reserved[3];
} ArtConfig;

/* Which jobs a frame runs (ART:191, 213, 237). */
#define ART_JOB_RAYTRACE    1u   /* AudioRaytracerJobBatched */
#define ART_JOB_PERMEATION  2u   /* AudioPermeationJobBatched */
#define ART_JOB_PROCESS     4u   /* ProcessAudioDataJob (needs both of the above in the same frame) */
#define ART_JOB_ALL         7u

/* Frame flags */
#define ART_FRAME_COUNTERS         1u  /* also produce oracle-equivalent work counters (slower kernel variant) */
#define ART_FRAME_REVERB_SEQ_FP32  2u  /* PA:38-48 as the reference rounds it: one sequential FP32 sum (bit-exact
                                          with the reference, serial tail). Default: exact integer accumulation. */
#define ART_FRAME_NO_HOST_OUTPUTS  4u  /* keep per-ray outputs in HBM, copy back only the per-target results */
#define ART_FRAME_PARTIALS_ONLY    8u  /* sharded run: skip the PA finalisation, caller combines partials */
#define ART_FRAME_BRUTE_FORCE     16u  /* scan every collider for every query exactly as the reference's loops do
                                          (no acceleration structure). Default: uniform-grid traversal, which runs
                                          the same exact tests on the colliders near each ray only; all outputs are
                                          bit-identical. ART_FRAME_COUNTERS implies brute force, and so do batches of fewer
                                          than 4,096..32,768 rays (by scene size; ART_GRID_MIN_RAYS) unless ART_FRAME_FORCE_GRID is set. */
#define ART_FRAME_FORCE_GRID      64u  /* use the grid kernels even for batches so small (a few thousand rays) that the library
                                          would pick the brute-force kernels, whose warp-per-ray mapping has the lower latency there */
#define ART_FRAME_GRID_STATS      32u  /* grid kernels also count the collider tests and cells they actually visit
                                          (ArtCounters.grid*); slightly slower kernel variant */
#define ART_FRAME_NO_FANS        128u  /* grid kernels: do not build the per-frame target fans (direction-binned collider lists
                                          around the listener and every audio target); the echo / muffle / permeation queries
                                          then walk the grid cells along their lines instead. Results are bit-identical
                                          (permeationSum within its tolerance). ART_DISABLE_FANS=1 does the same for a context. */

/* One field per job-struct field (RT:12-52, PM:10-27, PA:10-25). */
typedef struct ArtParams {
    float        rayOrigin[3];             /* RT:12, PM:10, PA:25 */
    const float* audioTargetPositions;     /* RT:22  float3[totalAudioTargets]; copied by the call */
    int32_t      totalAudioTargets;        /* RT:23 */
    float        maxRayLife;               /* RT:25 */
    uint8_t      maxHitsPerRay;            /* RT:26 (1..255) */
    float        maxMuffleHitDistance;     /* RT:52 */
    float        permeationStrengthPerRay; /* PM:23 */
    float        muffleEffectiveness;      /* PA:14 */
    float        permeationEffectiveness;  /* PA:18 */
    float        maxReverbDistance;        /* PA:21 */
    int32_t      batchCount;               /* T = MuffleRayHits.Length / targets (ATM:112); batch size per ART:161 */
    uint32_t     jobs;                     /* ART_JOB_* */
    uint32_t     flags;                    /* ART_FRAME_* */
} ArtParams;

/* Caller-owned output arrays (NativeArray.GetUnsafePtr()); any pointer may be NULL = not wanted.
 * Per-ray arrays are indexed by the context's LOCAL ray index (== the global index unless
 * art_set_ray_shard was used). They must stay alive and untouched until art_complete returns. */
typedef struct ArtOutputs {
    uint16_t* echoRayDistances;       /* half  [n*H]     RT:42 */
    uint16_t* rayHitResults;          /* half3 [n*H]     RT:35 AudioRayHitResult.HitPoint */
    uint8_t*  rayHitResultCounts;     /* byte  [n]       RT:38 */
    uint16_t* muffleRayHits;          /* ushort[T*Na]    RT:50 */
    float*    permeationPowerRemains; /* float [T*Na]    PM:27 */
    ArtTargetSettings* audioTargetSettings; /* [Na]      PA:28 */
    /* extensions (not produced by the reference) */
    uint32_t* hitColliderIds;         /* [n*H] ColliderType<<30 | index (None=0,AABB=1,OBB=2,Sphere=3) */
    uint32_t* muffleTotals;           /* [Na] muffle hits summed over all rays, no ushort wrap (Q8) */
    double*   permeationSum;          /* [Na] sum over hitting rays of the PM:260 value (the intended reduction, Q5) */
} ArtOutputs;

/* Work counters; "tests" are counted exactly as the reference's loops would execute them
 * (early exits honoured). [3] index order: sphere, AABB, OBB. Filled only for ART_FRAME_COUNTERS,
 * except the *_ms and segments/segmentHits fields which are always valid. */
typedef struct ArtCounters {
    uint64_t segments;          /* ShootRayCast calls in the bounce loop (RT:108) */
    uint64_t segmentHits;
    uint64_t traceTests[3];
    uint64_t echoQueries;
    uint64_t echoTests[3];
    uint64_t muffleQueries;
    uint64_t muffleTests[3];
    uint64_t permRays;
    uint64_t permHitRays;
    uint64_t permFirstTests[3];
    uint64_t permPairs;
    uint64_t permLossTests[3];
    float    traceMs;           /* device time of the trace kernel */
    float    permeationMs;      /* device time of the permeation kernel(s) */
    float    reduceMs;          /* device time of the reduction kernel(s) */
    float    deviceMs;          /* first kernel start -> last kernel end, this frame */
    float    h2dMs, d2hMs;      /* copy time on the stream (0 when nothing was copied) */
    uint32_t kernelLaunches;    /* kernels of this library launched for the frame */
    uint32_t gridUsed;          /* bit 0: trace job used the uniform grid, bit 1: permeation job did, bit 2: the frame's
                                   goal-directed queries used the target fans, bit 3: the fan build overflowed its entry
                                   buffer and the frame was re-run on the grid walk (the buffer grows for the next frame), bit 4: the
                                   trace job rotated its ray groups through the warps (small batches / shards), bit 5: the
                                   permeation job evaluated its loss lines sorted by (target, direction bin) (large frames with
                                   the target fans; ART_K2_BINNED=0/1 forces it off/on) */
    /* ART_FRAME_GRID_STATS: collider tests the grid kernels actually executed ([3] = sphere, AABB, OBB) and grid
     * cells they visited; compare with traceTests + echoTests + muffleTests / permFirstTests / permLossTests, the
     * counts of the reference's full scans */
    uint64_t gridTraceTests[3];
    uint64_t gridPermFirstTests[3];
    uint64_t gridPermLossTests[3];
    uint64_t gridTraceCells, gridPermCells;
    uint64_t debugViolations;   /* builds with -DART_DEBUG_BOUNDS: failed index checks in the grid kernels (0 otherwise);
                                   any build: a group-rotation wait that timed out (art_complete then fails) */
    /* parts of traceMs (default path): per-frame target-fan build, bounce tracer (trace_grid_kernel, bounce-only mode),
     * echo / muffle query kernel (query_fan_kernel); 0 where a part did not run */
    float    fanBuildMs, bounceMs, queryMs;
    float    exchangeMs;        /* multi-process frames (art_comm_init): device time of the all-gather of the partial blobs */
    uint32_t devicesUsed;       /* GPUs that worked on this frame (multi-device context: nDevices) */
    uint32_t reserved0;
    /* ART_FRAME_GRID_STATS, default path: the part of gridTraceTests / gridTraceCells executed by query_fan_kernel (echo and
     * muffle queries against the target fans; "cells" = lists opened); the rest belongs to the bounce tracer's grid walk */
    uint64_t gridQueryTests[3];
    uint64_t gridQueryLists;
} ArtCounters;

/* ≙ AudioRayTracer.Awake/InitializeAudioRaytraceSystem (ART:53-87): one context per AudioRayTracer. */
ART_API int32_t art_create(const ArtConfig* cfg, ArtCtx** out);
/* ≙ OnDestroy (ART:241-254): completes pending work, frees device memory. */
ART_API void    art_destroy(ArtCtx* ctx);

/* ≙ the AABBColliders/OBBColliders/SphereColliders + *ColliderCount fields (RT:15-20, ART:168-175).
 * Data is copied (staged) before the call returns. Counts may be 0, pointers then may be NULL. */
ART_API int32_t art_set_scene(ArtCtx* ctx, const ArtAABB* aabbs, int32_t nAABB,
                              const ArtOBB* obbs, int32_t nOBB,
                              const ArtSphere* spheres, int32_t nSphere);

/* ≙ RayDirections (RT:13, ART:166): half3[n] as 3 x uint16 per ray. Copied before return. */
ART_API int32_t art_set_rays(ArtCtx* ctx, const uint16_t* half3Directions, int32_t rayCount);

/* ≙ FibonacciDirectionsJobParallel (Jobs/FibonacciDirectionsJobParallel.cs:15-35, ART:72-77) run on the
 * device: RayDirections = Fibonacci sphere of rayCount rays, never crossing PCIe. */
ART_API int32_t art_generate_fibonacci_rays(ArtCtx* ctx, int32_t rayCount);
/* Copy the context's current ray directions back (half3[n]). */
ART_API int32_t art_get_rays(ArtCtx* ctx, uint16_t* half3Directions, int32_t capacityRays);

/* Multi-GPU sharding (one context per GPU): this context traces only the rays of chunks
 * c with c % shardCount == shardIndex, a chunk being chunkRays consecutive global ray indices
 * (chunkRays = 0: one contiguous slice per shard, i.e. exactly a reference "batch", ART:161).
 * Local ray j maps to global ray ((j / chunk) * shardCount + shardIndex) * chunk + j % chunk.
 * RayDirections still describes the whole batch (PM:260 uses RayDirections.Length). */
ART_API int32_t art_set_ray_shard(ArtCtx* ctx, int32_t shardIndex, int32_t shardCount, int32_t chunkRays);
ART_API int32_t art_local_ray_count(ArtCtx* ctx);   /* n = rays owned by this context (or negative status) */

/* ≙ the three Schedule() calls ART:191 + 213 + 237 in one call. Asynchronous: returns once the
 * work is enqueued. Inputs may be reused by the caller as soon as it returns. */
ART_API int32_t art_trace_schedule(ArtCtx* ctx, const ArtParams* params, const ArtOutputs* outputs, ArtHandle* out);
/* ≙ JobHandle.IsCompleted (ART:95): 1 = done, 0 = running, negative = error. Never blocks. */
ART_API int32_t art_is_completed(ArtCtx* ctx, ArtHandle h);
/* ≙ JobHandle.Complete() (ART:97): blocks; on return 0 every requested output is valid. */
ART_API int32_t art_complete(ArtCtx* ctx, ArtHandle h);

ART_API int32_t art_get_counters(ArtCtx* ctx, ArtHandle h, ArtCounters* out);
/* UTF-8, context-owned, valid until the next call on that context. ctx may be NULL (creation errors). */
ART_API const char* art_last_error(ArtCtx* ctx);

/* ---- sharded frames: combine per-context partial results --------------------------
 * After art_complete of a frame scheduled with ART_FRAME_PARTIALS_ONLY, each context exports a
 * fixed-size blob of its per-target partial sums; blobs are merged with art_partials_merge (exact,
 * order independent: integer sums + a max-by-ray-index select) and art_finalize applies the
 * ProcessAudioDataJob formulas (PA:49-74) to the merged blob. Blobs are plain bytes and may be
 * exchanged with NCCL/MPI all-gather. */
ART_API int64_t art_partials_size(int32_t totalAudioTargets, int32_t batchCount);
ART_API int32_t art_get_partials(ArtCtx* ctx, ArtHandle h, void* blob, int64_t blobBytes);
ART_API int32_t art_partials_merge(void* accumBlob, const void* otherBlob, int64_t blobBytes);
ART_API int32_t art_finalize(const void* blob, int64_t blobBytes, const ArtParams* params, int32_t rayCount,
                             const ArtOutputs* outputs /* only the per-target arrays are used */);

/* ---- multi-process sharding (one rank per GPU, e.g. under torchrun / mpirun) -----------------------------------------
 * art_comm_init joins the context to a communicator of `world` ranks (NCCL, loaded with dlopen("libnccl.so.2") on first
 * use -- single-GPU hosts need no NCCL) and makes it trace shard `rank` of the batch (interleaved chunks of chunkRays rays,
 * 0 = N / (8 x ranks) rays, at most 16,384). From then on a frame scheduled WITHOUT ART_FRAME_PARTIALS_ONLY all-gathers the ranks' partial blobs on the
 * device (one ncclAllGather of a few KB on the context's stream, no host round trip), merges them exactly and finalises on
 * every rank: art_complete returns the same per-source outputs everywhere. Per-ray outputs stay local (local ray indexing).
 * uniqueId: 128 bytes from art_comm_unique_id on one rank, distributed by the caller (MPI_Bcast, torch.distributed ...). */
ART_API int32_t art_comm_unique_id(void* uniqueId128);
ART_API int32_t art_comm_init(ArtCtx* ctx, const void* uniqueId128, int32_t rank, int32_t world, int32_t chunkRays);

/* Host-only inspection of the acceleration structure art_set_scene builds (a uniform grid over the colliders;
 * csrc/grid_host.h): cell (x,y,z) = cells[2*((z*ny + y)*nx + x) + {0,1}] = {first entry, nS | nA << 10 | nO << 21},
 * entries = canonical collider indices, per cell spheres | AABBs | OBBs. Needs no context and no GPU (tests, tooling).
 * cells / entries may be NULL to query the sizes first. Returns ART_OK, or ART_E_STATE when the scene cannot be gridded
 * (the library then uses the brute-force kernels). */
typedef struct ArtGridInfo {
    int32_t nx, ny, nz;
    float   g0[3], g1[3], cellSize[3];
    float   margin;             /* m of grid_host.h */
    int64_t nCells, nEntries;
} ArtGridInfo;
ART_API int32_t art_grid_build_host(const ArtAABB* aabbs, int32_t nAABB, const ArtOBB* obbs, int32_t nOBB,
                                    const ArtSphere* spheres, int32_t nSphere, float cellScale, ArtGridInfo* info,
                                    uint32_t* cells, int64_t cellsCapacity, uint16_t* entries, int64_t entriesCapacity);

/* Inspection of the target fans the last completed frame built on the device (csrc/fan_dev.cuh; tests, tooling):
 * fan a < nTargets belongs to audio target a, fan nTargets to the listener. Per fan there are cellsPerFan cells:
 * 6 cube faces x binsPerFace x binsPerFace direction bins (cell = face * B*B + ib * B + ia; face = 2*axis + (negative ? 1 : 0),
 * ia / ib = tangent-plane coordinates of the other two axes in cyclic order, mapped from [-1,1] to [0,B)) followed by the
 * near list. cells[2*i + {0,1}] = {first entry, nS | nA << 10 | nO << 21}, entries = collider indices per cell,
 * spheres | AABBs | OBBs, each nearest to the goal first. cells / entries may be NULL to query the sizes first.
 * Returns ART_E_STATE when the last frame did not use the fans. */
typedef struct ArtFanInfo {
    int32_t nFans, binsPerFace, cellsPerFan;
    float   nearDist;           /* colliders whose bounds come this close to the goal are in its near list */
    int64_t nCells, nEntries;   /* nEntries: entries in use (the lists are packed in no particular cell order) */
} ArtFanInfo;
ART_API int32_t art_debug_get_fans(ArtCtx* ctx, ArtFanInfo* info, uint32_t* cells, int64_t cellsCapacity,
                                   uint16_t* entries, int64_t entriesCapacity);

/* ... and their covering depths (csrc/k4_fan_build.cu "covering depth": beyond that depth -- measured from the goal along the
 * cube face's major axis -- one AABB fills the whole sub-bin, so every echo / muffle query from there is blocked and the
 * query kernel drops it untested). codes[i] holds the four sub-bins of cell i (2 x 2 per bin: sub-bin sb * 2 + sa in byte
 * sb * 2 + sa; sa / sb = which half of the bin in ia / ib direction); a code c < 255 stands for the depth
 * 2^((c - logK) / logS), 255 for "none". The near-list cell of a fan (the last one) carries no codes.
 * capacity in uint32 (>= ArtFanInfo.nCells). Returns ART_E_STATE when the last frame did not use the fans. */
ART_API int32_t art_debug_get_fan_cover(ArtCtx* ctx, uint32_t* codes, int64_t capacity, float* logS, float* logK);

/* Device-side FP32 issue-rate microbenchmarks used for the roofline denominator (bench.py):
 * kind 0 = un-fused FADD/FMUL, 1 = FMNMX, 2 = FFMA. Returns achieved Gop/s (lane-ops). */
ART_API int32_t art_microbench(ArtCtx* ctx, int32_t kind, double* gops);

#ifdef __cplusplus
}
#endif
#endif /* AUDIORT_H */
