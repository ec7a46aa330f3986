// AudioRtNative.cs -- P/Invoke binding of include/audiort.h (libaudiort_cuda) for the Unity project.
// Drop into Assets/C# Scripts/Audio/ next to AudioRayTracer.cs; put libaudiort_cuda.so (Linux) into
// Assets/Plugins/x86_64/. NOT compiled in this repository's CI (no .NET toolchain in the build image);
// field order, sizes and packing mirror include/audiort.h exactly.
using System;
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using Unity.Mathematics;

public static unsafe class AudioRtNative
{
    const string Lib = "audiort_cuda";
    public const int ART_ABI_VERSION = 3;
    public const int ART_MAX_DEVICES = 16;

    public const uint JOB_RAYTRACE = 1, JOB_PERMEATION = 2, JOB_PROCESS = 4, JOB_ALL = 7;
    public const uint FRAME_COUNTERS = 1, FRAME_REVERB_SEQ_FP32 = 2, FRAME_NO_HOST_OUTPUTS = 4, FRAME_PARTIALS_ONLY = 8;
    public const uint FRAME_BRUTE_FORCE = 16, FRAME_GRID_STATS = 32, FRAME_FORCE_GRID = 64, FRAME_NO_FANS = 128;

    [StructLayout(LayoutKind.Sequential)]
    public struct ArtConfig
    {
        public int abiVersion; public int device; public uint flags;
        public int nDevices;                       // > 1: one context over devices[0..nDevices) (the library shards the rays and merges)
        public fixed int devices[ART_MAX_DEVICES];
        public int shardChunkRays;
        public fixed int reserved[3];
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct ArtParams
    {
        public float rayOriginX, rayOriginY, rayOriginZ;   // RT:12
        public float3* audioTargetPositions;               // RT:22
        public int totalAudioTargets;                      // RT:23
        public float maxRayLife;                           // RT:25
        public byte maxHitsPerRay;                         // RT:26
        public float maxMuffleHitDistance;                 // RT:52
        public float permeationStrengthPerRay;             // PM:23
        public float muffleEffectiveness;                  // PA:14
        public float permeationEffectiveness;              // PA:18
        public float maxReverbDistance;                    // PA:21
        public int batchCount;                             // AudioRaytracingManager.ToUseThreadCount
        public uint jobs, flags;
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct ArtOutputs
    {
        public half* echoRayDistances;                     // RT:42
        public AudioRayHitResult* rayHitResults;           // RT:35
        public byte* rayHitResultCounts;                   // RT:38
        public ushort* muffleRayHits;                      // RT:50
        public float* permeationPowerRemains;              // PM:27
        public AudioTargetRTSettings* audioTargetSettings; // PA:28
        public uint* hitColliderIds;                       // extension
        public uint* muffleTotals;                         // extension
        public double* permeationSum;                      // extension
    }

    [DllImport(Lib)] public static extern int art_create(ref ArtConfig cfg, out IntPtr ctx);
    [DllImport(Lib)] public static extern void art_destroy(IntPtr ctx);
    [DllImport(Lib)] public static extern int art_set_scene(IntPtr ctx, ColliderAABBStruct* aabbs, int nAABB,
                                                            ColliderOBBStruct* obbs, int nOBB, ColliderSphereStruct* spheres, int nSphere);
    [DllImport(Lib)] public static extern int art_set_rays(IntPtr ctx, half3* dirs, int rayCount);
    [DllImport(Lib)] public static extern int art_generate_fibonacci_rays(IntPtr ctx, int rayCount);
    [DllImport(Lib)] public static extern int art_trace_schedule(IntPtr ctx, ref ArtParams p, ref ArtOutputs o, out int handle);
    [DllImport(Lib)] public static extern int art_is_completed(IntPtr ctx, int handle);
    [DllImport(Lib)] public static extern int art_complete(IntPtr ctx, int handle);
    [DllImport(Lib)] public static extern IntPtr art_last_error(IntPtr ctx);

    public static void Check(IntPtr ctx, int rc)
    {
        if (rc < 0) throw new InvalidOperationException("libaudiort_cuda: " + Marshal.PtrToStringAnsi(art_last_error(ctx)));
    }
}
