/*
 * audiort_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the three reference jobs on the acoustic hot path of
 * FirePixel8422/Audio-Raytracer:
 *   RT = Assets/C# Scripts/Jobs/AudioRaytracerJobBatched.cs
 *   PM = Assets/C# Scripts/Jobs/AudioPermeationJobBatched.cs
 *   PA = Assets/C# Scripts/Jobs/ProcessAudioDataJob.cs
 *   FIB = Assets/C# Scripts/Jobs/FibonacciDirectionsJobParallel.cs
 * plus the Unity.Mathematics 1.3.2 primitives they call (third-party package,
 * pinned in Packages/packages-lock.json:57-58, source NOT vendored in the
 * reference tree -> restated from the published package).
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
 * for this path and cannot be executed in this environment (C#, no .NET/Mono
 * toolchain). This oracle is pinned only by hand-derived known-answer vectors
 * (tests/test_oracle_kat.py). See DESIGN.md section 3.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library. The product path
 * (libaudiort_cuda) never links or calls it.
 */
#ifndef AUDIORT_ORACLE_H
#define AUDIORT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- wire layouts: exact C# sequential layouts, all members 2 bytes -------- */
/* CS/ColliderAABBStruct.cs:8-14 (20 B) */
typedef struct {
    uint16_t center[3];      /* half3 Center */
    uint16_t size[3];        /* half3 Size == half extents (RT:257) */
    uint16_t absorption;     /* AudioMaterialProperties.Absorption (half) */
    uint16_t density;        /* .Density */
    uint16_t echo;           /* .Echo */
    int16_t  audioTargetId;  /* -1 = not owned */
} OrAABB;

/* CS/ColliderOBBStruct.cs:8-24 (26 B) */
typedef struct {
    uint16_t center[3];
    uint16_t size[3];
    uint16_t rot[3];         /* halfQuaternion x,y,z (DT/halfQuaternion.cs:7-11); w reconstructed */
    uint16_t absorption, density, echo;
    int16_t  audioTargetId;
} OrOBB;

/* CS/ColliderSphereStruct.cs:8-14 (16 B) */
typedef struct {
    uint16_t center[3];
    uint16_t radius;
    uint16_t absorption, density, echo;
    int16_t  audioTargetId;
} OrSphere;

/* DT/AudioTargetRTSettings.cs:8-24 (24 B) */
typedef struct {
    float muffleStrength;
    float reverbStrength;
    float reverbVolume;
    float percievedAudioPosition[3];
} OrTargetSettings;

/* collider id encoding for the instrumented hit-id output: type<<30 | index,
 * type per Enums/ColliderType.cs (None=0, AABB=1, OBB=2, Sphere=3). */
#define OR_TYPE_NONE   0u
#define OR_TYPE_AABB   1u
#define OR_TYPE_OBB    2u
#define OR_TYPE_SPHERE 3u

/* Work counters. Index order of the [3] arrays: 0 = sphere, 1 = AABB, 2 = OBB
 * (the reference's scan order). "tests" = Ray*Intersects* calls actually
 * executed, early exits honoured. */
typedef struct {
    uint64_t segments;        /* ShootRayCast calls inside the bounce loop (RT:108) */
    uint64_t segment_hits;    /* of which returned true */
    uint64_t trace_tests[3];  /* RT ShootRayCast */
    uint64_t echo_queries;    /* CanRaySeePoint calls (RT:133) */
    uint64_t echo_tests[3];
    uint64_t muffle_queries;  /* CanRaySeeAudioTarget calls (RT:168, after the distance gate) */
    uint64_t muffle_tests[3];
    uint64_t perm_rays;       /* PM rays */
    uint64_t perm_hit_rays;   /* PM rays whose first-hit test succeeded (PM:58) */
    uint64_t perm_first_tests[3];
    uint64_t perm_pairs;      /* ShootPermeationRayCast calls (PM:82) */
    uint64_t perm_loss_tests[3];
} OrCounters;

/* Inputs shared by RT / PM / PA (one field per job-struct field). */
typedef struct {
    float           rayOrigin[3];        /* RT:12 / PM:10 / PA:25 */
    const uint16_t* rayDirections;       /* RT:13  half3[N] */
    int32_t         rayCount;            /* RayDirections.Length */
    const OrAABB*   aabbs;   int32_t nAABB;     /* RT:15-16 */
    const OrOBB*    obbs;    int32_t nOBB;      /* RT:17-18 */
    const OrSphere* spheres; int32_t nSphere;   /* RT:19-20 */
    const float*    targetPositions;     /* RT:22 float3[Na] */
    int32_t         nTargets;            /* RT:23 */
    float           maxRayLife;          /* RT:25 */
    uint8_t         maxHitsPerRay;       /* RT:26 */
    float           maxMuffleHitDistance;/* RT:52 */
    float           permeationStrengthPerRay; /* PM:23 */
    float           muffleEffectiveness;      /* PA:14 */
    float           permeationEffectiveness;  /* PA:18 */
    float           maxReverbDistance;        /* PA:21 */
    int32_t         batchCount;          /* T: MuffleRayHits.Length / Na (ATM:112) */
} OrScene;

/* Outputs. Any pointer may be NULL (skipped). */
typedef struct {
    uint16_t* echoRayDistances;     /* half[N*H]     RT:42 */
    uint16_t* rayHitResults;        /* half3[N*H]    RT:35 */
    uint8_t*  rayHitResultCounts;   /* byte[N]       RT:38 */
    uint16_t* muffleRayHits;        /* ushort[T*Na]  RT:50 */
    float*    permeationPowerRemains; /* float[T*Na] PM:27 */
    OrTargetSettings* settings;     /* [Na]          PA:28 */
    /* instrumentation / extensions (not reference outputs) */
    uint32_t* hitColliderIds;       /* [N*H] type<<30|index, 0 where no hit */
    float*    hitDistances;         /* [N*H] rayHitDist of each segment */
    uint32_t* muffleTotals;         /* [Na] sum over rays, no u16 wrap */
    double*   permeationSum;        /* [Na] sum over hitting rays of PM:260 values */
    OrTargetSettings* settingsFp64; /* [Na] PA evaluated in double on the same arrays */
} OrOutputs;

/* ---- Unity.Mathematics half conversions (math.f32tof16 / f16tof32) ------- */
uint16_t or_f32tof16(float x);
float    or_f16tof32(uint16_t h);

/* FIB:15-35, directions[i] for i in [first, first+count) of an N-ray sphere. */
void or_fibonacci_directions(int32_t N, int32_t first, int32_t count, uint16_t* outHalf3);

/* ART:161 batch size rule. */
int32_t or_batch_size(int32_t rayCount, int32_t batchCount);

/* One reference Execute(rayStartIndex, totalRays) call, as written.
 * faithfulReset != 0 reproduces the RT:72-80 reset range literally (quirk Q1). */
void or_rt_execute(const OrScene* s, const OrOutputs* o, int32_t rayStartIndex, int32_t totalRays,
                   int faithfulReset, OrCounters* c);
void or_pm_execute(const OrScene* s, const OrOutputs* o, int32_t rayStartIndex, int32_t totalRays,
                   OrCounters* c);
void or_pa_execute(const OrScene* s, const OrOutputs* o);

/* Canonical whole-frame evaluation (SURVEY 8a pins): outputs zero-initialised,
 * batches k = 0,1,.. run serially in ascending order, echo/hit arrays per the
 * T=1 reading of Q1. jobs bitmask: 1 = RT, 2 = PM, 4 = PA.
 * nThreads > 1 runs the (independent) batches of RT and the rays of PM on
 * that many pthreads -- same results, used only for the CPU baseline.
 * Returns 0, or -1 on invalid arguments. */
#define OR_JOB_RT 1
#define OR_JOB_PM 2
#define OR_JOB_PA 4
int or_run_frame(const OrScene* s, const OrOutputs* o, int jobs, int nThreads, OrCounters* c);

/* Literal serial emulation with T batches including quirk Q1 (diagnostic). */
int or_run_frame_faithful_q1(const OrScene* s, const OrOutputs* o, OrCounters* c);

/* Trace only rays [first, first+count) (canonical semantics, single batch slot
 * row `slotRow` of the muffle table); used for bounded CPU-baseline samples. */
int or_trace_range(const OrScene* s, const OrOutputs* o, int32_t first, int32_t count,
                   int nThreads, OrCounters* c);

/* Permeation work only (first hit + per-target loss rays) for rays
 * [first, first+count): counters, no slot writes. CPU-baseline timing helper. With nThreads <= 1 and
 * o->permeationSum set (zeroed by the caller) the window's per-target sums over rays are accumulated there. */
int or_permeation_range(const OrScene* s, const OrOutputs* o, int32_t first, int32_t count,
                        int nThreads, OrCounters* c);

#ifdef __cplusplus
}
#endif
#endif
