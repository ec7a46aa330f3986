// audiort_jobs.hpp -- header-only C++17 mirror of the reference's job-scheduling seam over the C ABI (audiort.h).
//
// The reference (C#, Unity) fills three Burst job structs field by field and schedules them every frame
// (Assets/C# Scripts/Audio/AudioRayTracer.cs:161-237), then polls / completes the combined JobHandle one frame later
// (ART:95-107). This header keeps those structs' field names (RT = Jobs/AudioRaytracerJobBatched.cs:12-52,
// PM = Jobs/AudioPermeationJobBatched.cs:10-27, PA = Jobs/ProcessAudioDataJob.cs:10-28) so that native host code reads
// like the reference's own scheduling block; `AudioRayTracerPlugin::Schedule` does what ART:191 + 213 + 237 do, with the
// three Schedule() calls replaced by one art_trace_schedule. Nothing here computes: every result comes from
// libaudiort_cuda (there is no CPU fallback -- the constructor throws when art_create reports ART_E_NO_DEVICE).
// The C# twin is csharp/AudioRtNative.cs, the Python twin audio-raytracer_b200/jobs.py.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "audiort.h"

namespace audiort {

struct half3 { uint16_t x, y, z; };                       // Unity.Mathematics.half3 (raw half bits)
struct float3 { float x, y, z; };
static_assert(sizeof(half3) == 6 && sizeof(float3) == 12, "blittable like the C# structs");

// NativeArray<T> as the job structs see it: a caller-owned pointer and a length (NativeArray.GetUnsafePtr()).
template <class T>
struct NativeArray {
    T* ptr = nullptr;
    int Length = 0;
    NativeArray() = default;
    NativeArray(T* p, int n) : ptr(p), Length(n) {}
    explicit NativeArray(std::vector<T>& v) : ptr(v.data()), Length((int)v.size()) {}
};

class ArtError : public std::runtime_error {
public:
    ArtError(int status, const std::string& what) : std::runtime_error(what), status(status) {}
    int status;                                           // ArtStatus
};

// RT:12-52
struct AudioRaytracerJobBatched {
    float3 RayOrigin{};                                   // RT:12
    NativeArray<const half3> RayDirections;               // RT:13
    NativeArray<const ArtAABB> AABBColliders;  int AABBColliderCount = 0;      // RT:15-16
    NativeArray<const ArtOBB> OBBColliders;    int OBBColliderCount = 0;       // RT:17-18
    NativeArray<const ArtSphere> SphereColliders; int SphereColliderCount = 0; // RT:19-20
    NativeArray<const float3> AudioTargetPositions;       // RT:22
    int TotalAudioTargets = 0;                            // RT:23
    float MaxRayLife = 0.0f;                              // RT:25
    uint8_t MaxHitsPerRay = 0;                            // RT:26
    NativeArray<half3> RayHitResults;                     // RT:35 AudioRayHitResult.HitPoint [N*H]
    NativeArray<uint8_t> RayHitResultCounts;              // RT:38 [N]
    NativeArray<uint16_t> EchoRayDistances;               // RT:42 half [N*H]
    NativeArray<uint16_t> MuffleRayHits;                  // RT:50 ushort [T*Na]
    float MaxMuffleHitDistance = 0.0f;                    // RT:52
};

// PM:10-27 (shares its inputs with the ray tracer job, ART:196-211)
struct AudioPermeationJobBatched {
    float PermeationStrengthPerRay = 0.0f;                // PM:23
    NativeArray<float> PermeationPowerRemains;            // PM:27 [T*Na]
};

// PA:10-28
struct ProcessAudioDataJob {
    float MuffleEffectiveness = 0.0f;                     // PA:14
    float PermeationEffectiveness = 0.0f;                 // PA:18
    float MaxReverbDistance = 0.0f;                       // PA:21
    NativeArray<ArtTargetSettings> AudioTargetSettings;   // PA:28 [Na]
};

// One per AudioRayTracer MonoBehaviour (ART:53-87 Awake / InitializeAudioRaytraceSystem, ART:241-254 OnDestroy).
class AudioRayTracerPlugin {
public:
    explicit AudioRayTracerPlugin(int device = 0)
    {
        ArtConfig cfg{};
        cfg.abiVersion = ART_ABI_VERSION;
        cfg.device = device;
        const int rc = art_create(&cfg, &ctx_);
        if (rc != ART_OK) { ctx_ = nullptr; throw ArtError(rc, std::string("art_create: ") + art_last_error(nullptr)); }
    }
    ~AudioRayTracerPlugin()
    {
        if (ctx_) art_destroy(ctx_);                      // completes pending work first (ART:241-254)
    }
    AudioRayTracerPlugin(const AudioRayTracerPlugin&) = delete;
    AudioRayTracerPlugin& operator=(const AudioRayTracerPlugin&) = delete;

    // ≙ the three Schedule() calls ART:191 + 213 + 237. `batchCount` = AudioRaytracingManager.ToUseThreadCount (the T of
    // MuffleRayHits.Length = T * Na, ATM:112). Inputs may be reused as soon as this returns; the output arrays must stay
    // alive and untouched until Complete() -- the contract Unity imposes between Schedule and Complete.
    void Schedule(const AudioRaytracerJobBatched& rt, const AudioPermeationJobBatched& pm, const ProcessAudioDataJob& pa,
                  int batchCount, uint32_t flags = 0)
    {
        if (pending_) throw ArtError(ART_E_PENDING, "a frame is already in flight (ART:95-97 never overlaps two frames)");
        check(art_set_scene(ctx_, rt.AABBColliders.ptr, rt.AABBColliderCount, rt.OBBColliders.ptr, rt.OBBColliderCount,
                            rt.SphereColliders.ptr, rt.SphereColliderCount), "art_set_scene");
        if (rt.RayDirections.ptr != rays_ || rt.RayDirections.Length != nRays_) {          // ART:166: the same array every frame
            check(art_set_rays(ctx_, reinterpret_cast<const uint16_t*>(rt.RayDirections.ptr), rt.RayDirections.Length), "art_set_rays");
            rays_ = rt.RayDirections.ptr; nRays_ = rt.RayDirections.Length;
        }
        ArtParams p{};
        p.rayOrigin[0] = rt.RayOrigin.x; p.rayOrigin[1] = rt.RayOrigin.y; p.rayOrigin[2] = rt.RayOrigin.z;
        p.audioTargetPositions = reinterpret_cast<const float*>(rt.AudioTargetPositions.ptr);
        p.totalAudioTargets = rt.TotalAudioTargets;
        p.maxRayLife = rt.MaxRayLife;
        p.maxHitsPerRay = rt.MaxHitsPerRay;
        p.maxMuffleHitDistance = rt.MaxMuffleHitDistance;
        p.permeationStrengthPerRay = pm.PermeationStrengthPerRay;
        p.muffleEffectiveness = pa.MuffleEffectiveness;
        p.permeationEffectiveness = pa.PermeationEffectiveness;
        p.maxReverbDistance = pa.MaxReverbDistance;
        p.batchCount = batchCount;
        p.jobs = ART_JOB_ALL;
        p.flags = flags;
        ArtOutputs o{};
        o.echoRayDistances = rt.EchoRayDistances.ptr;
        o.rayHitResults = reinterpret_cast<uint16_t*>(rt.RayHitResults.ptr);
        o.rayHitResultCounts = rt.RayHitResultCounts.ptr;
        o.muffleRayHits = rt.MuffleRayHits.ptr;
        o.permeationPowerRemains = pm.PermeationPowerRemains.ptr;
        o.audioTargetSettings = pa.AudioTargetSettings.ptr;
        check(art_trace_schedule(ctx_, &p, &o, &handle_), "art_trace_schedule");
        pending_ = true;
    }
    // ≙ mainJobHandle.IsCompleted (ART:95): never blocks.
    bool IsCompleted()
    {
        if (!pending_) return true;
        const int rc = art_is_completed(ctx_, handle_);
        if (rc < 0) check(rc, "art_is_completed");
        return rc == 1;
    }
    // ≙ mainJobHandle.Complete() (ART:97): blocks; afterwards every output array of the scheduled structs is valid.
    void Complete()
    {
        if (!pending_) return;
        pending_ = false;
        check(art_complete(ctx_, handle_), "art_complete");
    }
    ArtCounters Counters()
    {
        ArtCounters c{};
        check(art_get_counters(ctx_, handle_, &c), "art_get_counters");
        return c;
    }
    ArtCtx* ctx() const { return ctx_; }

private:
    void check(int rc, const char* what)
    {
        if (rc != ART_OK) throw ArtError(rc, std::string(what) + ": " + art_last_error(ctx_));
    }
    ArtCtx* ctx_ = nullptr;
    ArtHandle handle_ = 0;
    bool pending_ = false;
    const half3* rays_ = nullptr;
    int nRays_ = 0;
};

}  // namespace audiort
