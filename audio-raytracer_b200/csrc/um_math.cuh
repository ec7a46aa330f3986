// um_math.cuh -- device restatement of the Unity.Mathematics 1.3.2 primitives the reference jobs
// call (com.unity.mathematics pinned in Packages/packages-lock.json:57-58; call sites
// AudioRaytracerJobBatched.cs:127,130,142,162,165,289-298,316-317,326-337,466-525).
//
// Every operation that feeds a reference-visible result is a separately rounded IEEE binary32
// operation issued through an explicit *_rn intrinsic, so nvcc can never contract a*b+c into an
// FMA and the results are bit-identical with managed C# (and with oracle/audiort_oracle.c).
// FMAs are used only in code marked "conservative" whose outcome cannot change a result.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace art {

struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };

__device__ __forceinline__ float mulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float addr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float subr(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float divr(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float rcpr(float a) { return __frcp_rn(a); }      // == 1.0f / a, correctly rounded
__device__ __forceinline__ float sqrtr(float a) { return __fsqrt_rn(a); }    // == (float)Math.Sqrt((double)a)

__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 add3(f3 a, f3 b) { return mk3(addr(a.x, b.x), addr(a.y, b.y), addr(a.z, b.z)); }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(subr(a.x, b.x), subr(a.y, b.y), subr(a.z, b.z)); }
__device__ __forceinline__ f3 mul3s(f3 a, float s) { return mk3(mulr(a.x, s), mulr(a.y, s), mulr(a.z, s)); }
__device__ __forceinline__ f3 smul3(float s, f3 a) { return mk3(mulr(s, a.x), mulr(s, a.y), mulr(s, a.z)); }

// math.dot(float3): a.x*b.x + a.y*b.y + a.z*b.z, left to right
__device__ __forceinline__ float dot3(f3 a, f3 b)
{
    return addr(addr(mulr(a.x, b.x), mulr(a.y, b.y)), mulr(a.z, b.z));
}
__device__ __forceinline__ float dot4(f4 a, f4 b)
{
    return addr(addr(addr(mulr(a.x, b.x), mulr(a.y, b.y)), mulr(a.z, b.z)), mulr(a.w, b.w));
}
// math.min / math.max: return the non-NaN operand; fminf/fmaxf agree except for the sign of a zero
// result, which no reference comparison can observe.
__device__ __forceinline__ float um_min(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float um_max(float a, float b) { return fmaxf(a, b); }
// three-input forms (one FMNMX3 on sm_100a); same value as the nested two-input calls
__device__ __forceinline__ float max3f(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float min3f(float a, float b, float c) { return fminf(fminf(a, b), c); }
// math.sign
__device__ __forceinline__ float um_sign(float x) { return subr(x > 0.0f ? 1.0f : 0.0f, x < 0.0f ? 1.0f : 0.0f); }
// math.saturate(x) = max(0, min(1, x))
__device__ __forceinline__ float um_saturate(float x) { return um_max(0.0f, um_min(1.0f, x)); }
// math.rsqrt(x) = 1.0f / sqrt(x)
__device__ __forceinline__ float um_rsqrt(float x) { return rcpr(sqrtr(x)); }
// math.cross(a,b) = (a*b.yzx - a.yzx*b).yzx
__device__ __forceinline__ f3 cross3(f3 a, f3 b)
{
    return mk3(subr(mulr(a.y, b.z), mulr(a.z, b.y)),
               subr(mulr(a.z, b.x), mulr(a.x, b.z)),
               subr(mulr(a.x, b.y), mulr(a.y, b.x)));
}
// math.mul(quaternion q, float3 v): t = 2*cross(q.xyz, v); v + q.w*t + cross(q.xyz, t)
__device__ __forceinline__ f3 qmul3(f4 q, f3 v)
{
    f3 qv = mk3(q.x, q.y, q.z);
    f3 t = smul3(2.0f, cross3(qv, v));
    return add3(add3(v, smul3(q.w, t)), cross3(qv, t));
}
// math.inverse(quaternion q) = rcp(dot(q,q)) * q * float4(-1,-1,-1,1)
__device__ __forceinline__ f4 qinverse(f4 q)
{
    float r = rcpr(dot4(q, q));
    f4 o;
    o.x = mulr(mulr(r, q.x), -1.0f); o.y = mulr(mulr(r, q.y), -1.0f);
    o.z = mulr(mulr(r, q.z), -1.0f); o.w = mulr(mulr(r, q.w), 1.0f);
    return o;
}
// math.reflect(i,n) = i - 2f*n*dot(i,n)
__device__ __forceinline__ f3 reflect3(f3 i, f3 n) { return sub3(i, mul3s(smul3(2.0f, n), dot3(i, n))); }
// math.normalize(float3)
__device__ __forceinline__ f3 normalize3(f3 v) { return smul3(um_rsqrt(dot3(v, v)), v); }

// math.f16tof32: exact (identical to IEEE binary16 -> binary32, including subnormals).
__device__ __forceinline__ float um_f16tof32(uint16_t h)
{
    const uint32_t shifted_exp = 0x7c00u << 13;
    uint32_t uf = ((uint32_t)h & 0x7fffu) << 13;
    uint32_t e = uf & shifted_exp;
    uf += (127u - 15u) << 23;
    if (e == shifted_exp) uf += (128u - 16u) << 23;
    if (e == 0) uf = __float_as_uint(subr(__uint_as_float(uf + (1u << 23)), 6.10351563e-05f));
    return __uint_as_float(uf | (((uint32_t)h & 0x8000u) << 16));
}
// math.f32tof16: truncate 12 mantissa bits, rescale by 2^-112 (IEEE multiply, subnormals kept),
// clamp, +0x1000, >>13  => round to nearest, ties AWAY from zero. Not __float2half_rn.
__device__ __forceinline__ uint16_t um_f32tof16(float x)
{
    const uint32_t infinity_32 = 255u << 23;
    const uint32_t msk = 0x7FFFF000u;
    uint32_t ux = __float_as_uint(x);
    uint32_t uux = ux & msk;
    uint32_t sb = __float_as_uint(mulr(__uint_as_float(uux), 1.92592994e-34f));
    sb = min(sb, 0x0F7FF000u);
    uint32_t h = (sb + 0x1000u) >> 13;
    if (uux >= infinity_32) h = (uux > infinity_32) ? 0x7e00u : 0x7c00u;
    return (uint16_t)(h | ((ux & ~msk) >> 16));
}

}  // namespace art
