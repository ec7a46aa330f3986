// k0_pack.cu -- K0: half-precision AoS collider structs -> FP32 SoA planes (scene_dev.cuh).
//
// Evaluates, once per scene upload and with the reference's own operation order, the pure
// per-collider functions the reference re-evaluates on every test:
//   min = Center - halfExtents, max = Center + halfExtents   (AudioRaytracerJobBatched.cs:286-287)
//   R*R                                                      (RT:328)
//   Rotation getter: w = sqrt(max(0, 1-|xyz|^2)), normalize  (DataTypes/halfQuaternion.cs:34-46)
//   math.inverse(Rotation)                                   (RT:489, AudioPermeationJobBatched.cs:174)
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {


__global__ void pack_kernel(const PackArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    float4* sph = reinterpret_cast<float4*>(a.geom + a.L.offSph);
    float4* aabbA = reinterpret_cast<float4*>(a.geom + a.L.offAabbA);
    float2* aabbB = reinterpret_cast<float2*>(a.geom + a.L.offAabbB);
    float4* obbQ = reinterpret_cast<float4*>(a.geom + a.L.offObbQ);
    float4* obbC = reinterpret_cast<float4*>(a.geom + a.L.offObbC);
    float2* obbH = reinterpret_cast<float2*>(a.geom + a.L.offObbH);

    for (int i = tid; i < a.L.nsPad; i += stride) {
        const uint16_t* w = a.rawS + 8 * (size_t)min(i, a.L.ns - 1);
        const float cx = um_f16tof32(w[0]), cy = um_f16tof32(w[1]), cz = um_f16tof32(w[2]), R = um_f16tof32(w[3]);
        sph[i] = make_float4(cx, cy, cz, mulr(R, R));
        a.sphAttr[i] = make_float4(um_f16tof32(w[4]), um_f16tof32(w[6]), um_f16tof32(w[5]), __int_as_float((int)(short)w[7]));
        a.ownS[i] = (short)w[7];
    }
    for (int i = tid; i < a.L.naPad; i += stride) {
        const uint16_t* w = a.rawA + 10 * (size_t)min(i, a.L.na - 1);
        const float cx = um_f16tof32(w[0]), cy = um_f16tof32(w[1]), cz = um_f16tof32(w[2]);
        const float hx = um_f16tof32(w[3]), hy = um_f16tof32(w[4]), hz = um_f16tof32(w[5]);
        const float mnx = subr(cx, hx), mny = subr(cy, hy), mnz = subr(cz, hz);
        const float mxx = addr(cx, hx), mxy = addr(cy, hy), mxz = addr(cz, hz);
        // lo <= hi per component: {t0, t1} is the same set, so min/max of it is unchanged
        aabbA[i] = make_float4(fminf(mnx, mxx), fminf(mny, mxy), fminf(mnz, mxz), fmaxf(mnx, mxx));
        aabbB[i] = make_float2(fmaxf(mny, mxy), fmaxf(mnz, mxz));
        a.aabbCtr[i] = make_float4(cx, cy, cz, 0.0f);
        a.aabbHalf[i] = make_float4(hx, hy, hz, 0.0f);
        a.aabbAttr[i] = make_float4(um_f16tof32(w[6]), um_f16tof32(w[8]), um_f16tof32(w[7]), __int_as_float((int)(short)w[9]));
        a.ownA[i] = (short)w[9];
    }
    for (int i = tid; i < a.L.noPad; i += stride) {
        const uint16_t* w = a.rawO + 13 * (size_t)min(i, a.L.no - 1);
        const float cx = um_f16tof32(w[0]), cy = um_f16tof32(w[1]), cz = um_f16tof32(w[2]);
        const float hx = um_f16tof32(w[3]), hy = um_f16tof32(w[4]), hz = um_f16tof32(w[5]);
        // halfQuaternion.QuaternionValue getter
        const float xx = um_f16tof32(w[6]), yy = um_f16tof32(w[7]), zz = um_f16tof32(w[8]);
        const float wSquared = subr(1.0f, addr(addr(mulr(xx, xx), mulr(yy, yy)), mulr(zz, zz)));
        const float ww = wSquared > 0.0f ? sqrtr(wSquared) : 0.0f;
        f4 q; q.x = xx; q.y = yy; q.z = zz; q.w = ww;
        const float rs = um_rsqrt(dot4(q, q));                       // math.normalize(quaternion)
        f4 qn; qn.x = mulr(rs, q.x); qn.y = mulr(rs, q.y); qn.z = mulr(rs, q.z); qn.w = mulr(rs, q.w);
        const f4 qi = qinverse(qn);
        obbQ[i] = make_float4(qn.x, qn.y, qn.z, qn.w);
        obbC[i] = make_float4(cx, cy, cz, fabsf(hx));
        obbH[i] = make_float2(fabsf(hy), fabsf(hz));
        a.obbHalf[i] = make_float4(hx, hy, hz, 0.0f);
        a.obbQinv[i] = make_float4(qi.x, qi.y, qi.z, qi.w);
        a.obbAttr[i] = make_float4(um_f16tof32(w[9]), um_f16tof32(w[11]), um_f16tof32(w[10]), __int_as_float((int)(short)w[12]));
        a.ownO[i] = (short)w[12];
    }
}

cudaError_t launch_pack(const PackArgs& a, cudaStream_t stream)
{
    const int n = max(max(a.L.nsPad, a.L.naPad), a.L.noPad);
    if (n == 0) return cudaSuccess;
    pack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace art
