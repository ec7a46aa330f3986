// Headless driver: one frame of the reference's three jobs, exactly as AudioRayTracer.OnUpdate schedules them
// (Assets/C# Scripts/Audio/AudioRayTracer.cs:161-237), over a binary scene dump (audio-raytracer_b200/scene_io.py).
// Canonical execution = batches in ascending order, outputs zero-initialised (SURVEY.md 3.2); with a thread count > 1
// the batches run on that many threads, which is what the CPU baseline of BASELINE.json asks for.
using System;
using System.Diagnostics;
using System.IO;
using System.Threading.Tasks;
using Unity.Collections;
using Unity.Mathematics;

public static class Program
{
    static T[] ReadStructs<T>(BinaryReader r, int n, Func<BinaryReader, T> one)
    {
        var a = new T[n];
        for (int i = 0; i < n; i++) a[i] = one(r);
        return a;
    }
    static half H(BinaryReader r) { half h; h.value = r.ReadUInt16(); return h; }
    static half3 H3(BinaryReader r) { half3 v; v.x = H(r); v.y = H(r); v.z = H(r); return v; }
    static T[] ReadBlittable<T>(BinaryReader r, int n, int size) where T : struct
    {
        if (System.Runtime.InteropServices.Marshal.SizeOf<T>() != size) throw new InvalidOperationException($"{typeof(T).Name} is not {size} bytes");
        var a = new T[n];
        for (int i = 0; i < n; i++) a[i] = System.Runtime.InteropServices.MemoryMarshal.Read<T>(r.ReadBytes(size));
        return a;
    }

    public static int Main(string[] args)
    {
        if (args.Length < 2) { Console.Error.WriteLine("usage: AudioRtHarness scene.artd out.arto [threads]"); return 2; }
        int threads = args.Length > 2 ? int.Parse(args[2]) : 1;
        using var r = new BinaryReader(File.OpenRead(args[0]));
        if (new string(r.ReadChars(4)) != "ARTD" || r.ReadInt32() != 1) throw new InvalidDataException("not an ARTD v1 dump");
        int nA = r.ReadInt32(), nO = r.ReadInt32(), nS = r.ReadInt32(), Na = r.ReadInt32(), N = r.ReadInt32(), Hh = r.ReadInt32(), T = r.ReadInt32();
        var origin = new float3(r.ReadSingle(), r.ReadSingle(), r.ReadSingle());
        float maxRayLife = r.ReadSingle(), maxMuffle = r.ReadSingle(), strength = r.ReadSingle(), muffleEff = r.ReadSingle(),
              permEff = r.ReadSingle(), maxReverb = r.ReadSingle();
        // the collider structs are blittable, sequential and made of 2-byte members only (20 / 26 / 16 bytes, no padding):
        // read them as raw bytes so that the private halfQuaternion of ColliderOBBStruct keeps its exact stored bits
        var aabbs = ReadBlittable<ColliderAABBStruct>(r, nA, 20);
        var obbs = ReadBlittable<ColliderOBBStruct>(r, nO, 26);
        var spheres = ReadBlittable<ColliderSphereStruct>(r, nS, 16);
        var targets = ReadStructs(r, Na, b => new float3(b.ReadSingle(), b.ReadSingle(), b.ReadSingle()));
        var dirs = ReadStructs(r, N, H3);

        var echo = new NativeArray<half>(N * Hh, Allocator.Persistent);
        var hitResults = new NativeArray<AudioRayHitResult>(N * Hh, Allocator.Persistent);
        var hitCounts = new NativeArray<byte>(N, Allocator.Persistent);
        var muffle = new NativeArray<ushort>(T * Na, Allocator.Persistent);
        var perm = new NativeArray<float>(T * Na, Allocator.Persistent);
        var settings = new NativeArray<AudioTargetRTSettings>(Na, Allocator.Persistent);

        var rt = new AudioRaytracerJobBatched
        {
            RayOrigin = origin, RayDirections = new NativeArray<half3>(dirs),
            AABBColliders = new NativeArray<ColliderAABBStruct>(aabbs), AABBColliderCount = nA,
            OBBColliders = new NativeArray<ColliderOBBStruct>(obbs), OBBColliderCount = nO,
            SphereColliders = new NativeArray<ColliderSphereStruct>(spheres), SphereColliderCount = nS,
            AudioTargetPositions = new NativeArray<float3>(targets), TotalAudioTargets = Na,
            MaxHitsPerRay = (byte)Hh, MaxRayLife = maxRayLife,
            RayHitResults = hitResults, RayHitResultCounts = hitCounts, EchoRayDistances = echo,
            MuffleRayHits = muffle, MaxMuffleHitDistance = maxMuffle,
        };
        var pm = new AudioPermeationJobBatched
        {
            RayOrigin = origin, RayDirections = rt.RayDirections,
            AABBColliders = rt.AABBColliders, AABBColliderCount = nA, OBBColliders = rt.OBBColliders, OBBColliderCount = nO,
            SphereColliders = rt.SphereColliders, SphereColliderCount = nS,
            AudioTargetPositions = rt.AudioTargetPositions, TotalAudioTargets = Na,
            PermeationStrengthPerRay = strength, PermeationPowerRemains = perm,
        };
        var pa = new ProcessAudioDataJob
        {
            EchoRayDistances = echo, MaxReverbDistance = maxReverb, TotalAudioTargets = Na,
            AudioTargetPositions = rt.AudioTargetPositions, AudioTargetSettings = settings,
            MuffleRayHits = muffle, MuffleEffectiveness = muffleEff,
            PermeationPowerRemains = perm, PermeationStrengthPerRay = strength, PermeationEffectiveness = permEff,
            MaxHitsPerRay = Hh, RayCount = N, RayOriginWorld = origin,
        };

        int batch = (int)math.max(1, math.ceil((float)N / T));                   // ART:161
        int batches = (N + batch - 1) / batch;
        var sw = Stopwatch.StartNew();
        Action<int> runRt = k => rt.Execute(k * batch, Math.Min(batch, N - k * batch));
        Action<int> runPm = k => pm.Execute(k * batch, Math.Min(batch, N - k * batch));
        if (threads <= 1) { for (int k = 0; k < batches; k++) runRt(k); for (int k = 0; k < batches; k++) runPm(k); }
        else
        {
            var opt = new ParallelOptions { MaxDegreeOfParallelism = threads };
            Parallel.For(0, batches, opt, runRt);
            Parallel.For(0, batches, opt, runPm);
        }
        pa.Execute();
        sw.Stop();
        long segments = 0;
        for (int i = 0; i < N; i++) segments += hitCounts[i];
        Console.WriteLine($"{{\"rays\": {N}, \"threads\": {threads}, \"batches\": {batches}, \"seconds\": {sw.Elapsed.TotalSeconds:F6}, " +
                          $"\"segment_hits\": {segments}, \"segment_hits_per_s\": {segments / sw.Elapsed.TotalSeconds:F1}}}");

        using var w = new BinaryWriter(File.Create(args[1]));
        w.Write(new[] { 'A', 'R', 'T', 'O' }); w.Write(1);
        for (int i = 0; i < echo.Length; i++) w.Write(echo[i].value);
        for (int i = 0; i < hitResults.Length; i++) { var p = hitResults[i].HitPoint; w.Write(p.x.value); w.Write(p.y.value); w.Write(p.z.value); }
        for (int i = 0; i < N; i++) w.Write(hitCounts[i]);
        for (int i = 0; i < muffle.Length; i++) w.Write(muffle[i]);
        for (int i = 0; i < perm.Length; i++) w.Write(perm[i]);
        for (int i = 0; i < Na; i++)
        {
            var s = settings[i];
            w.Write(s.MuffleStrength); w.Write(s.ReverbStrength); w.Write(s.ReverbVolume);
            w.Write(s.PercievedAudioPosition.x); w.Write(s.PercievedAudioPosition.y); w.Write(s.PercievedAudioPosition.z);
        }
        return 0;
    }
}
