// The subset of com.unity.mathematics 1.3.2 (Packages/packages-lock.json:57-58) the reference jobs call, restated from the
// published package semantics -- the same definitions oracle/audiort_oracle.c and csrc/um_math.cuh pin (SURVEY.md
// Appendix A). Every operation is a separately rounded binary32 operation in the order the package writes it.
// If this harness is built against the REAL package instead (drop this file, reference Unity.Mathematics.dll), a
// difference in the diff tool pins an error in Appendix A.
using System;
using System.Runtime.CompilerServices;
using UnityEngine;

namespace Unity.Mathematics
{
    public struct float3
    {
        public float x, y, z;
        public float3(float x, float y, float z) { this.x = x; this.y = y; this.z = z; }
        public float3(float v) { x = y = z = v; }
        public static float3 zero => new float3(0f, 0f, 0f);
        public float3 yzx => new float3(y, z, x);
        public static float3 operator +(float3 a, float3 b) => new float3(a.x + b.x, a.y + b.y, a.z + b.z);
        public static float3 operator -(float3 a, float3 b) => new float3(a.x - b.x, a.y - b.y, a.z - b.z);
        public static float3 operator *(float3 a, float3 b) => new float3(a.x * b.x, a.y * b.y, a.z * b.z);
        public static float3 operator /(float3 a, float3 b) => new float3(a.x / b.x, a.y / b.y, a.z / b.z);
        public static float3 operator *(float3 a, float b) => new float3(a.x * b, a.y * b, a.z * b);
        public static float3 operator *(float a, float3 b) => new float3(a * b.x, a * b.y, a * b.z);
        public static float3 operator /(float3 a, float b) => new float3(a.x / b, a.y / b, a.z / b);
        public static float3 operator /(float a, float3 b) => new float3(a / b.x, a / b.y, a / b.z);
        public static float3 operator +(float3 a, float b) => new float3(a.x + b, a.y + b, a.z + b);
        public static float3 operator -(float3 a, float b) => new float3(a.x - b, a.y - b, a.z - b);
        public static float3 operator -(float3 a) => new float3(-a.x, -a.y, -a.z);
        public static implicit operator float3(float v) => new float3(v);
        public static implicit operator float3(half3 h) => new float3(h.x, h.y, h.z);
        public static implicit operator float3(Vector3 v) => new float3(v.x, v.y, v.z);
        public static implicit operator Vector3(float3 v) => new Vector3(v.x, v.y, v.z);
    }

    public struct float4
    {
        public float x, y, z, w;
        public float4(float x, float y, float z, float w) { this.x = x; this.y = y; this.z = z; this.w = w; }
        public float3 xyz => new float3(x, y, z);
        public static float4 operator *(float4 a, float4 b) => new float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        public static float4 operator *(float a, float4 b) => new float4(a * b.x, a * b.y, a * b.z, a * b.w);
    }

    public struct quaternion
    {
        public float4 value;
        public quaternion(float x, float y, float z, float w) { value = new float4(x, y, z, w); }
        public quaternion(float4 v) { value = v; }
        public static readonly quaternion identity = new quaternion(0f, 0f, 0f, 1f);
    }

    /// <summary>IEEE binary16 storage; float -> half is math.f32tof16 (round to nearest, ties AWAY from zero).</summary>
    public struct half
    {
        public ushort value;
        public half(half h) { value = h.value; }
        public half(float v) { value = (ushort)math.f32tof16(v); }
        public static explicit operator half(float v) => new half(v);
        public static explicit operator half(double v) => new half((float)v);
        public static implicit operator float(half h) => math.f16tof32(h.value);
        public static bool operator ==(half a, half b) => a.value == b.value;
        public static bool operator !=(half a, half b) => a.value != b.value;
        public override bool Equals(object o) => o is half h && h.value == value;
        public override int GetHashCode() => value;
    }

    public struct half3
    {
        public half x, y, z;
        public half3(half x, half y, half z) { this.x = x; this.y = y; this.z = z; }
        public half3(float v) { x = y = z = (half)v; }
        public half3(float3 v) { x = (half)v.x; y = (half)v.y; z = (half)v.z; }
        public static explicit operator half3(float3 v) => new half3(v);
    }

    public static class math
    {
        public const float PI = 3.14159265f;
        public const float INFINITY = float.PositiveInfinity;

        [MethodImpl(MethodImplOptions.AggressiveInlining)] public static uint asuint(float x) => BitConverter.SingleToUInt32Bits(x);
        [MethodImpl(MethodImplOptions.AggressiveInlining)] public static float asfloat(uint x) => BitConverter.UInt32BitsToSingle(x);

        public static float min(float x, float y) => float.IsNaN(y) || x < y ? x : y;
        public static float max(float x, float y) => float.IsNaN(y) || x > y ? x : y;
        public static int min(int x, int y) => x < y ? x : y;
        public static int max(int x, int y) => x > y ? x : y;
        public static float3 min(float3 a, float3 b) => new float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z));
        public static float3 max(float3 a, float3 b) => new float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z));
        public static float abs(float x) => asfloat(asuint(x) & 0x7FFFFFFFu);
        public static float3 abs(float3 v) => new float3(abs(v.x), abs(v.y), abs(v.z));
        public static float sign(float x) => (x > 0f ? 1f : 0f) - (x < 0f ? 1f : 0f);
        public static float3 sign(float3 v) => new float3(sign(v.x), sign(v.y), sign(v.z));
        public static float sqrt(float x) => (float)Math.Sqrt((double)x);
        public static float rcp(float x) => 1.0f / x;
        public static float rsqrt(float x) => 1.0f / sqrt(x);
        public static float cos(float x) => (float)Math.Cos((double)x);
        public static float sin(float x) => (float)Math.Sin((double)x);
        public static float ceil(float x) => (float)Math.Ceiling((double)x);
        public static float floor(float x) => (float)Math.Floor((double)x);
        public static float saturate(float x) => max(0f, min(1f, x));
        public static float clamp(float x, float a, float b) => max(a, min(b, x));
        public static float lerp(float a, float b, float t) => a + t * (b - a);
        public static float dot(float3 a, float3 b) => a.x * b.x + a.y * b.y + a.z * b.z;
        public static float dot(float4 a, float4 b) => a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        public static float3 cross(float3 a, float3 b) => (a * b.yzx - a.yzx * b).yzx;
        public static float length(float3 v) => sqrt(dot(v, v));
        public static float distance(float3 a, float3 b) => length(b - a);
        public static float3 normalize(float3 v) => rsqrt(dot(v, v)) * v;
        public static quaternion normalize(quaternion q) => new quaternion(rsqrt(dot(q.value, q.value)) * q.value);
        public static float3 reflect(float3 i, float3 n) => i - 2f * n * dot(i, n);
        public static quaternion inverse(quaternion q) => new quaternion(rcp(dot(q.value, q.value)) * q.value * new float4(-1f, -1f, -1f, 1f));
        public static float3 mul(quaternion q, float3 v)
        {
            float3 t = 2f * cross(q.value.xyz, v);
            return v + q.value.w * t + cross(q.value.xyz, t);
        }

        /// <summary>math.f16tof32: exact.</summary>
        public static float f16tof32(uint x)
        {
            const uint shifted_exp = 0x7c00u << 13;
            uint uf = (x & 0x7fffu) << 13;
            uint e = uf & shifted_exp;
            uf += (127u - 15u) << 23;
            if (e == shifted_exp) uf += (128u - 16u) << 23;
            else if (e == 0u) uf = asuint(asfloat(uf + (1u << 23)) - 6.10351563e-05f);
            return asfloat(uf | (x & 0x8000u) << 16);
        }

        /// <summary>math.f32tof16: truncate 12 mantissa bits, rescale by 2^-112, clamp, +0x1000, >>13.</summary>
        public static uint f32tof16(float x)
        {
            const uint infinity_32 = 255u << 23;
            const uint msk = 0x7FFFF000u;
            uint ux = asuint(x);
            uint uux = ux & msk;
            uint sb = asuint(asfloat(uux) * 1.92592994e-34f);
            if (sb > 0x0F7FF000u) sb = 0x0F7FF000u;
            uint h = (sb + 0x1000u) >> 13;
            if (uux >= infinity_32) h = uux > infinity_32 ? 0x7e00u : 0x7c00u;
            return h | (ux & ~msk) >> 16;
        }
    }
}
