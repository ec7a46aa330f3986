// host_jobs_demo.cpp -- one frame of the hot path scheduled from native host code through the job-struct mirror
// (include/audiort_jobs.hpp), the C++ twin of the reference's AudioRayTracer.OnUpdate block
// (Assets/C# Scripts/Audio/AudioRayTracer.cs:161-237, 95-97). Reads a binary scene dump (ARTD, scene_io.py), schedules,
// polls like the reference's async path, completes, writes the outputs as an ARTO file that
// `python tools/scene_dump.py diff scene.artd out.arto` compares with the oracle bit for bit.
//
//   python tools/scene_dump.py write c1 /tmp/c1.artd
//   g++ -std=c++17 -O2 cpp/host_jobs_demo.cpp -Iinclude -Laudio-raytracer_b200 -laudiort_cuda -Wl,-rpath,$PWD/audio-raytracer_b200 -o /tmp/host_jobs_demo
//   /tmp/host_jobs_demo /tmp/c1.artd /tmp/c1.cpp.arto && python tools/scene_dump.py diff /tmp/c1.artd /tmp/c1.cpp.arto
//
// Exit codes: 0 frame done, 2 bad arguments / dump, 3 no usable CUDA device (there is no CPU fallback), 1 any other error.
#include <cstdio>
#include <cstring>
#include <vector>

#include "audiort_jobs.hpp"

using namespace audiort;

template <class T>
static bool read_n(FILE* f, std::vector<T>& v, size_t n)
{
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, f) == n;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: host_jobs_demo scene.artd out.arto\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    char magic[4]; int32_t h[8]; float p[9];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ARTD", 4) != 0 || fread(h, 4, 8, f) != 8 || h[0] != 1 || fread(p, 4, 9, f) != 9) {
        fprintf(stderr, "%s: not an ARTD v1 dump\n", argv[1]);
        return 2;
    }
    const int nA = h[1], nO = h[2], nS = h[3], Na = h[4], N = h[5], H = h[6], T = h[7];
    std::vector<ArtAABB> aabbs; std::vector<ArtOBB> obbs; std::vector<ArtSphere> spheres;
    std::vector<float3> targets; std::vector<half3> dirs;
    if (!read_n(f, aabbs, nA) || !read_n(f, obbs, nO) || !read_n(f, spheres, nS) || !read_n(f, targets, Na) || !read_n(f, dirs, N)) {
        fprintf(stderr, "%s: truncated dump\n", argv[1]);
        return 2;
    }
    fclose(f);

    // the NativeArrays AudioRayTracer / AudioTargetManager own (ART:66-87, ATM:112-121)
    std::vector<uint16_t> echoRayDistances((size_t)N * H), muffleRayHits((size_t)T * Na);
    std::vector<half3> rayHitResults((size_t)N * H);
    std::vector<uint8_t> rayHitResultCounts(N);
    std::vector<float> permeationPowerRemains((size_t)T * Na);
    std::vector<ArtTargetSettings> audioTargetSettings(Na);

    try {
        AudioRayTracerPlugin plugin(0);

        AudioRaytracerJobBatched rt;                       // ART:161-190
        rt.RayOrigin = { p[0], p[1], p[2] };
        rt.RayDirections = NativeArray<const half3>(dirs.data(), N);
        rt.AABBColliders = NativeArray<const ArtAABB>(aabbs.data(), nA);          rt.AABBColliderCount = nA;
        rt.OBBColliders = NativeArray<const ArtOBB>(obbs.data(), nO);             rt.OBBColliderCount = nO;
        rt.SphereColliders = NativeArray<const ArtSphere>(spheres.data(), nS);    rt.SphereColliderCount = nS;
        rt.AudioTargetPositions = NativeArray<const float3>(targets.data(), Na);  rt.TotalAudioTargets = Na;
        rt.MaxRayLife = p[3];
        rt.MaxHitsPerRay = (uint8_t)H;
        rt.MaxMuffleHitDistance = p[4];
        rt.RayHitResults = NativeArray<half3>(rayHitResults);
        rt.RayHitResultCounts = NativeArray<uint8_t>(rayHitResultCounts);
        rt.EchoRayDistances = NativeArray<uint16_t>(echoRayDistances);
        rt.MuffleRayHits = NativeArray<uint16_t>(muffleRayHits);

        AudioPermeationJobBatched pm;                      // ART:196-212
        pm.PermeationStrengthPerRay = p[5];
        pm.PermeationPowerRemains = NativeArray<float>(permeationPowerRemains);

        ProcessAudioDataJob pa;                            // ART:218-236
        pa.MuffleEffectiveness = p[6];
        pa.PermeationEffectiveness = p[7];
        pa.MaxReverbDistance = p[8];
        pa.AudioTargetSettings = NativeArray<ArtTargetSettings>(audioTargetSettings);

        plugin.Schedule(rt, pm, pa, T);                    // ART:191 + 213 + 237
        unsigned long polls = 0;
        while (!plugin.IsCompleted()) polls++;             // ART:95 (AudioRaytracingManager.ComputeAsync)
        plugin.Complete();                                 // ART:97
        const ArtCounters c = plugin.Counters();

        FILE* o = fopen(argv[2], "wb");
        if (!o) { perror(argv[2]); return 2; }
        const int32_t ver = 1;
        fwrite("ARTO", 1, 4, o); fwrite(&ver, 4, 1, o);
        fwrite(echoRayDistances.data(), 2, echoRayDistances.size(), o);
        fwrite(rayHitResults.data(), 6, rayHitResults.size(), o);
        fwrite(rayHitResultCounts.data(), 1, rayHitResultCounts.size(), o);
        fwrite(muffleRayHits.data(), 2, muffleRayHits.size(), o);
        fwrite(permeationPowerRemains.data(), 4, permeationPowerRemains.size(), o);
        fwrite(audioTargetSettings.data(), sizeof(ArtTargetSettings), audioTargetSettings.size(), o);
        fclose(o);
        printf("{\"rays\": %d, \"targets\": %d, \"colliders\": %d, \"segments\": %llu, \"device_ms\": %.4f, \"polls\": %lu, \"muffle0\": %.6f}\n",
               N, Na, nA + nO + nS, (unsigned long long)c.segments, c.deviceMs, polls, Na ? audioTargetSettings[0].muffleStrength : 0.0f);
    } catch (const ArtError& e) {
        fprintf(stderr, "%s\n", e.what());
        return e.status == ART_E_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
