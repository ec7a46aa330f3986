/* abi_latency.c -- frame latency of libaudiort_cuda through the plain C ABI (no Python in the loop): what a C# caller
 * going through P/Invoke pays per AudioRayTracer.OnUpdate (Assets/C# Scripts/Audio/AudioRayTracer.cs:92-238).
 *
 *   python tools/scene_dump.py write c1 /tmp/c1.artd
 *   gcc -O2 -std=c11 tools/abi_latency.c -Iinclude -Laudio-raytracer_b200 -laudiort_cuda -Wl,-rpath,$PWD/audio-raytracer_b200 -o /tmp/abi_latency
 *   /tmp/abi_latency /tmp/c1.artd 2000 [static]
 *
 * Per frame (like the reference, ART:154-237): art_set_scene (dynamic colliders re-baked every frame; "static" skips it),
 * art_trace_schedule with host output arrays, art_complete. Prints wall-clock microseconds per frame.
 */
#define _POSIX_C_SOURCE 199309L
#include "audiort.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_us(void)
{
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}
static int cmp_double(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }

#define CK(call) do { int rc_ = (call); if (rc_ != ART_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, art_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: abi_latency scene.artd [frames] [static]\n"); return 2; }
    const int frames = argc > 2 ? atoi(argv[2]) : 1000;
    const int staticScene = argc > 3 && strcmp(argv[3], "static") == 0;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    char magic[4]; int32_t h[8]; float p[9];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ARTD", 4) || fread(h, 4, 8, f) != 8 || h[0] != 1 || fread(p, 4, 9, f) != 9) { fprintf(stderr, "bad dump\n"); return 2; }
    const int nA = h[1], nO = h[2], nS = h[3], Na = h[4], N = h[5], H = h[6], T = h[7];
    ArtAABB* aabbs = malloc(sizeof(ArtAABB) * (nA + 1)); ArtOBB* obbs = malloc(sizeof(ArtOBB) * (nO + 1)); ArtSphere* sph = malloc(sizeof(ArtSphere) * (nS + 1));
    float* targets = malloc(12 * (size_t)Na); uint16_t* dirs = malloc(6 * (size_t)N);
    if (fread(aabbs, sizeof(ArtAABB), nA, f) != (size_t)nA || fread(obbs, sizeof(ArtOBB), nO, f) != (size_t)nO || fread(sph, sizeof(ArtSphere), nS, f) != (size_t)nS ||
        fread(targets, 12, Na, f) != (size_t)Na || fread(dirs, 6, N, f) != (size_t)N) { fprintf(stderr, "truncated dump\n"); return 2; }
    fclose(f);

    ArtCtx* ctx = NULL;
    ArtConfig cfg; memset(&cfg, 0, sizeof cfg); cfg.abiVersion = ART_ABI_VERSION; cfg.device = 0;
    if (art_create(&cfg, &ctx) != ART_OK) { fprintf(stderr, "art_create: %s\n", art_last_error(NULL)); return 1; }
    ArtParams prm; memset(&prm, 0, sizeof prm);
    prm.rayOrigin[0] = p[0]; prm.rayOrigin[1] = p[1]; prm.rayOrigin[2] = p[2];
    prm.audioTargetPositions = targets; prm.totalAudioTargets = Na; prm.maxRayLife = p[3]; prm.maxHitsPerRay = (uint8_t)H;
    prm.maxMuffleHitDistance = p[4]; prm.permeationStrengthPerRay = p[5]; prm.muffleEffectiveness = p[6];
    prm.permeationEffectiveness = p[7]; prm.maxReverbDistance = p[8]; prm.batchCount = T; prm.jobs = ART_JOB_ALL; prm.flags = 0;
    ArtOutputs out; memset(&out, 0, sizeof out);
    out.echoRayDistances = malloc(2 * (size_t)N * H); out.rayHitResults = malloc(6 * (size_t)N * H); out.rayHitResultCounts = malloc(N);
    out.muffleRayHits = malloc(2 * (size_t)T * Na); out.permeationPowerRemains = malloc(4 * (size_t)T * Na);
    out.audioTargetSettings = malloc(sizeof(ArtTargetSettings) * Na);

    CK(art_set_scene(ctx, aabbs, nA, obbs, nO, sph, nS));
    CK(art_set_rays(ctx, dirs, N));
    double* us = malloc(sizeof(double) * frames);
    double tSet = 0, tSched = 0, tDone = 0;
    ArtHandle hnd = 0;
    for (int i = -20; i < frames; i++) {
        const double t0 = now_us();
        if (!staticScene) CK(art_set_scene(ctx, aabbs, nA, obbs, nO, sph, nS));
        const double t1 = now_us();
        CK(art_trace_schedule(ctx, &prm, &out, &hnd));
        const double t2 = now_us();
        CK(art_complete(ctx, hnd));
        const double t3 = now_us();
        if (i >= 0) { us[i] = t3 - t0; tSet += t1 - t0; tSched += t2 - t1; tDone += t3 - t2; }
    }
    ArtCounters c; CK(art_get_counters(ctx, hnd, &c));
    qsort(us, frames, sizeof(double), cmp_double);
    double sum = 0; for (int i = 0; i < frames; i++) sum += us[i];
    printf("{\"rays\": %d, \"targets\": %d, \"colliders\": %d, \"frames\": %d, \"scene_upload_per_frame\": %s, \"us_per_frame_mean\": %.1f, "
           "\"us_per_frame_median\": %.1f, \"us_per_frame_p99\": %.1f, \"device_us\": %.1f, \"set_scene_us\": %.1f, \"schedule_us\": %.1f, "
           "\"complete_us\": %.1f, \"segments\": %llu, \"muffle0\": %.6f}\n",
           N, Na, nA + nO + nS, frames, staticScene ? "false" : "true", sum / frames, us[frames / 2], us[(int)(frames * 0.99)],
           c.deviceMs * 1e3, tSet / frames, tSched / frames, tDone / frames, (unsigned long long)c.segments,
           out.audioTargetSettings[0].muffleStrength);
    art_destroy(ctx);
    return 0;
}
