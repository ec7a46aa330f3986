// Minimal stand-ins for the Unity assemblies the reference job files compile against: attributes that only steer
// Burst / the job safety system (no-ops here), NativeArray<T> over a managed array, and the job interfaces.
// Behaviour that matters for results lives in UnityMathematics.cs.
using System;

namespace Unity.Burst
{
    [AttributeUsage(AttributeTargets.All, AllowMultiple = true)] public sealed class BurstCompileAttribute : Attribute { }
    [AttributeUsage(AttributeTargets.All, AllowMultiple = true)] public sealed class NoAliasAttribute : Attribute { }
}

namespace Unity.Collections
{
    public enum Allocator { Invalid, None, Temp, TempJob, Persistent }
    public enum NativeArrayOptions { UninitializedMemory, ClearMemory }
    [AttributeUsage(AttributeTargets.All)] public sealed class ReadOnlyAttribute : Attribute { }
    [AttributeUsage(AttributeTargets.All)] public sealed class WriteOnlyAttribute : Attribute { }
    [AttributeUsage(AttributeTargets.All)] public sealed class NativeDisableParallelForRestrictionAttribute : Attribute { }

    /// <summary>NativeArray&lt;T&gt; with the members the jobs use: Length and a get/set indexer (value semantics).</summary>
    public struct NativeArray<T> : IDisposable where T : struct
    {
        private T[] data;
        public NativeArray(int length, Allocator allocator, NativeArrayOptions options = NativeArrayOptions.ClearMemory) { data = new T[length]; }
        public NativeArray(T[] wrap) { data = wrap; }
        public int Length => data == null ? 0 : data.Length;
        public bool IsCreated => data != null;
        public T this[int index] { get => data[index]; set => data[index] = value; }
        public T[] ToArray() => data;
        public void Dispose() { data = null; }
    }
}

namespace Unity.Jobs
{
    public interface IJob { void Execute(); }
    public interface IJobParallelFor { void Execute(int index); }
    /// <summary>Unity invokes Execute(startIndex, count) once per batch: start = k*b, count = min(b, N - k*b).</summary>
    public interface IJobParallelForBatch { void Execute(int startIndex, int count); }
}

namespace UnityEngine
{
    [AttributeUsage(AttributeTargets.All)] public sealed class SerializeFieldAttribute : Attribute { }
    [AttributeUsage(AttributeTargets.All)] public sealed class HideInInspectorAttribute : Attribute { }
    [AttributeUsage(AttributeTargets.All)] public sealed class TooltipAttribute : Attribute { public TooltipAttribute(string s) { } }
    [AttributeUsage(AttributeTargets.All)] public sealed class HeaderAttribute : Attribute { public HeaderAttribute(string s) { } }
    [AttributeUsage(AttributeTargets.All)] public sealed class RangeAttribute : Attribute { public RangeAttribute(float a, float b) { } }

    public struct Vector3
    {
        public float x, y, z;
        public Vector3(float x, float y, float z) { this.x = x; this.y = y; this.z = z; }
    }
}
