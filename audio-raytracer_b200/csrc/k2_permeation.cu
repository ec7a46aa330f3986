// k2_permeation.cu -- K2: AudioPermeationJobBatched.Execute
// (Assets/C# Scripts/Jobs/AudioPermeationJobBatched.cs:34-91) for sm_100a.
//
// Per ray: nearest first-hit DISTANCE over all colliders (PM:101-141; its OBB test inverts the
// already inverted stored rotation, PM:174 -- quirk Q4 -- so it reads obbQinv), then for every audio
// target the through-material loss  sum_colliders max(0, tExit - max(tEnter,0)) * density
// (PM:225-261, no early exit, no distance clip -- quirk Q7).
//
// Mapping: one warp owns one ray, lanes own colliders (same planes and super-chunks as K1).
// Targets are processed in blocks of TBK: the origin-dependent part of each collider test is
// computed once per super-chunk and reused for the TBK targets of the block; every lane keeps one
// partial sum per target in shared memory (acc[target][lane], conflict free), and each target's total
// is formed by one lane adding the 32 partials in lane order -- a fixed order, hence deterministic.
//
// What the reference finally KEEPS of all this work is, per batch, only the values of the last
// hitting ray (PM:85 overwrites, quirk Q5/Q6). perm_last_kernel recomputes exactly that ray with
// the reference's sequential FP32 summation order, so the canonical PermeationPowerRemains is
// bit-exact; the all-rays sum is exported as the permeationSum extension.
#include "device_util.cuh"
#include "intersect.cuh"
#include "scene_dev.cuh"
#include "um_math.cuh"

#include <type_traits>

namespace art {

constexpr int TBK = 16;          // targets per accumulator block
constexpr int kAccStride = 33;   // acc[t * 33 + lane]: conflict free per target and in the final per-lane sums
constexpr int kPermRecFloat4PerWarp = 64;   // rec[q] = (inv.xyz, class bits), rec[32+q] = (d.xyz, dot(d,d))
constexpr int kPermWarpBytes = kPermRecFloat4PerWarp * 16 + ((TBK * kAccStride * 4 + 15) / 16) * 16;

template <bool SMEM>
__global__ void __launch_bounds__(kThreads, 1) permeation_kernel(const PermArgs a)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int Na = a.nTargets;

    unsigned char* p = smem;
    const unsigned char* geomBase = a.geom;
    if (SMEM) {
        stage_blob_to_smem(p, a.geom, a.L.bytes, &bar);
        geomBase = p;
        p += a.L.bytes;
    }
    float4* rec = reinterpret_cast<float4*>(p + (size_t)warp * kPermWarpBytes);
    float* acc = reinterpret_cast<float*>(p + (size_t)warp * kPermWarpBytes + kPermRecFloat4PerWarp * 16);
    const GeomView gv = make_view(geomBase, a.L);
    const int nsPad = a.L.nsPad, naPad = a.L.naPad, noPad = a.L.noPad;
    const f3 RayOrigin = mk3(a.ox, a.oy, a.oz);
    unsigned int nRays = 0, nHitRays = 0;

    for (;;) {
        int j = 0;
        if (lane == 0) j = (int)atomicAdd(a.nextRay, 1u);
        j = __shfl_sync(kFull, j, 0);
        if (j >= a.map.nLocal) break;
        const int rayIndex = a.map.to_global(j);
        nRays++;
        const f3 d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                         um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));                  // PM:53
        const f3 o = RayOrigin;

        // ================= ShootRayCast, distance only (PM:101-141) =================
        float best = pos_inf();   // math.INFINITY
        {
            const float dd = dot3(d, d);
            for (int base = 0; base < nsPad; base += SC_S) {
                f3 oc[RS]; float cc[RS];
                uint32_t need = 0;
#pragma unroll
                for (int r = 0; r < RS; r++) {
                    const float4 s = gv.sph[base + r * 32 + lane];
                    oc[r] = sub3(o, mk3(s.x, s.y, s.z));
                    cc[r] = subr(dot3(oc[r], oc[r]), s.w);
                    if (!sphere_fast_miss(oc[r], cc[r], d, dd)) need |= 1u << r;
                }
                if (need) {
#pragma unroll
                    for (int r = 0; r < RS; r++)
                        if ((need >> r) & 1u) {
                            const float dist = sphere_dist_exact(oc[r].x, oc[r].y, oc[r].z, cc[r], d.x, d.y, d.z, dd);
                            if (dist < best) best = dist;
                        }
                }
            }
            const float ix = rcpr(d.x), iy = rcpr(d.y), iz = rcpr(d.z);
            for (int base = 0; base < naPad; base += SC_A) {
#pragma unroll
                for (int r = 0; r < RA; r++) {
                    const float4 A = gv.aabbA[base + r * 32 + lane];
                    const float2 B = gv.aabbB[base + r * 32 + lane];
                    float tNear, tFar, dist;
                    slab<8>(subr(A.x, o.x), subr(A.y, o.y), subr(A.z, o.z), subr(A.w, o.x), subr(B.x, o.y), subr(B.y, o.z),
                            ix, iy, iz, tNear, tFar);
                    if (slab_hit(tNear, tFar, dist) && dist < best) best = dist;
                }
            }
            for (int base = 0; base < noPad; base += SC_O) {
#pragma unroll
                for (int r = 0; r < RO; r++) {
                    const int idx = base + r * 32 + lane;
                    const float4 c4 = gv.obbC[idx];
                    const float2 h2 = gv.obbH[idx];
                    const f3 h = mk3(c4.w, h2.x, h2.y);
                    const f3 pc = sub3(o, mk3(c4.x, c4.y, c4.z));
                    if (!obb_sure_miss(pc, obb_cull_c(pc, h), d, dd)) {
                        const float4 qi = a.at.obbQinv[idx];                               // PM:174 (quirk Q4)
                        const float dist = obb_dist_exact(qi.x, qi.y, qi.z, qi.w, pc.x, pc.y, pc.z, h.x, h.y, h.z, d.x, d.y, d.z);
                        if (dist < best) best = dist;
                    }
                }
            }
        }
        best = warp_min_f(best);
        if (lane == 0) a.firstHitDist[j] = best;
        if (!(best != pos_inf())) continue;                                                // PM:140 / PM:58
        nHitRays++;
        if (lane == 0) atomicMax(&a.lastHitRay[rayIndex / a.batchSize], rayIndex);

        const f3 P = add3(o, mul3s(d, best));                                              // PM:61
        const f3 Pp = sub3(P, mul3s(d, kEpsilon));                                         // PM:72

        // ================= per-target loss rays (PM:67-86) =================
        for (int g = 0; g * 32 < Na; g++) {
            const int tslot = g * 32 + lane;
            if (tslot < Na) {
                const f3 T = mk3(a.targets[3 * tslot], a.targets[3 * tslot + 1], a.targets[3 * tslot + 2]);
                const f3 dir = normalize3(sub3(T, Pp));                                    // PM:76
                const float ix = rcpr(dir.x), iy = rcpr(dir.y), iz = rcpr(dir.z);
                rec[lane] = make_float4(ix, iy, iz, __int_as_float(slab_class(ix, iy, iz)));
                rec[32 + lane] = make_float4(dir.x, dir.y, dir.z, dot3(dir, dir));
            }
            __syncwarp();
            const int nIn = min(32, Na - g * 32);
            for (int t0 = 0; t0 < nIn; t0 += TBK) {
                const int nT = min(TBK, nIn - t0);
#pragma unroll
                for (int t = 0; t < TBK; t++) acc[t * kAccStride + lane] = 0.0f;
                __syncwarp();

                for (int base = 0; base < nsPad; base += SC_S) {
                    f3 oc[RS]; float cc[RS], dn[RS];
#pragma unroll
                    for (int r = 0; r < RS; r++) {
                        const float4 s = gv.sph[base + r * 32 + lane];
                        oc[r] = sub3(Pp, mk3(s.x, s.y, s.z));
                        cc[r] = subr(dot3(oc[r], oc[r]), s.w);
                        dn[r] = a.densS[base + r * 32 + lane];
                    }
                    for (int t = 0; t < nT; t++) {
                        const float4 r1 = rec[32 + t0 + t];
                        const f3 qd = mk3(r1.x, r1.y, r1.z);
                        float bb[RS];
                        uint32_t need = 0;
#pragma unroll
                        for (int r = 0; r < RS; r++)
                            if (!sphere_loss_fast_miss(oc[r], cc[r], qd, bb[r])) need |= 1u << r;
                        if (need) {
                            float s = 0.0f;
#pragma unroll
                            for (int r = 0; r < RS; r++)
                                if ((need >> r) & 1u) s = addr(s, sphere_loss_exact(bb[r], cc[r], dn[r]));
                            acc[t * kAccStride + lane] = addr(acc[t * kAccStride + lane], s);
                        }
                    }
                }
                for (int base = 0; base < naPad; base += SC_A) {
                    float lox[RA], loy[RA], loz[RA], hix[RA], hiy[RA], hiz[RA], dn[RA];
#pragma unroll
                    for (int r = 0; r < RA; r++) {
                        const float4 A = gv.aabbA[base + r * 32 + lane];
                        const float2 B = gv.aabbB[base + r * 32 + lane];
                        dn[r] = a.densA[base + r * 32 + lane];
                        // origin-relative planes, shared by all targets of the block
                        lox[r] = subr(A.x, Pp.x); loy[r] = subr(A.y, Pp.y); loz[r] = subr(A.z, Pp.z);
                        hix[r] = subr(A.w, Pp.x); hiy[r] = subr(B.x, Pp.y); hiz[r] = subr(B.y, Pp.z);
                    }
                    for (int t = 0; t < nT; t++) {
                        const float4 r0 = rec[t0 + t];
                        const int cls = __float_as_int(r0.w);
                        float s = 0.0f;
                        auto body = [&](auto clsTag) {
                            constexpr int CLS = decltype(clsTag)::value;
#pragma unroll
                            for (int r = 0; r < RA; r++) {
                                float tEnter, tExit;
                                slab<CLS>(lox[r], loy[r], loz[r], hix[r], hiy[r], hiz[r], r0.x, r0.y, r0.z, tEnter, tExit);
                                s = addr(s, slab_loss(tEnter, tExit, dn[r]));
                            }
                        };
                        switch (cls) {
                        case 0: body(std::integral_constant<int, 0>{}); break;
                        case 1: body(std::integral_constant<int, 1>{}); break;
                        case 2: body(std::integral_constant<int, 2>{}); break;
                        case 3: body(std::integral_constant<int, 3>{}); break;
                        case 4: body(std::integral_constant<int, 4>{}); break;
                        case 5: body(std::integral_constant<int, 5>{}); break;
                        case 6: body(std::integral_constant<int, 6>{}); break;
                        case 7: body(std::integral_constant<int, 7>{}); break;
                        default: body(std::integral_constant<int, 8>{}); break;
                        }
                        if (s != 0.0f) acc[t * kAccStride + lane] = addr(acc[t * kAccStride + lane], s);
                    }
                }
                for (int base = 0; base < noPad; base += SC_O) {
                    float4 oq[RO]; f3 pc[RO], hh[RO]; float cB[RO], dn[RO];
#pragma unroll
                    for (int r = 0; r < RO; r++) {
                        oq[r] = gv.obbQ[base + r * 32 + lane];                             // PM:296 stored rotation as is
                        const float4 c4 = gv.obbC[base + r * 32 + lane];
                        const float2 h2 = gv.obbH[base + r * 32 + lane];
                        hh[r] = mk3(c4.w, h2.x, h2.y);
                        pc[r] = sub3(Pp, mk3(c4.x, c4.y, c4.z));
                        cB[r] = obb_cull_c(pc[r], hh[r]);
                        dn[r] = a.densO[base + r * 32 + lane];
                    }
                    for (int t = 0; t < nT; t++) {
                        const float4 r1 = rec[32 + t0 + t];
                        const f3 qd = mk3(r1.x, r1.y, r1.z);
                        uint32_t need = 0;
#pragma unroll
                        for (int r = 0; r < RO; r++)
                            if (!obb_sure_miss(pc[r], cB[r], qd, r1.w)) need |= 1u << r;
                        if (need) {
                            float s = 0.0f;
#pragma unroll
                            for (int r = 0; r < RO; r++)
                                if ((need >> r) & 1u)
                                    s = addr(s, obb_loss_exact(oq[r].x, oq[r].y, oq[r].z, oq[r].w, pc[r].x, pc[r].y, pc[r].z,
                                                               hh[r].x, hh[r].y, hh[r].z, qd.x, qd.y, qd.z, dn[r]));
                            acc[t * kAccStride + lane] = addr(acc[t * kAccStride + lane], s);
                        }
                    }
                }
                // colliders owned by some target were given density 0 above; add them back for every
                // target except their owner (PM:235/245/255 skip only the owner's own colliders)
                for (int e = lane; e < a.nOwned; e += 32) {
                    const int code = a.ownedList[e];
                    const int sec = code >> 28, idx = code & 0x0FFFFFFF;
                    const float4 at = sec == 0 ? a.at.sphAttr[idx] : (sec == 1 ? a.at.aabbAttr[idx] : a.at.obbAttr[idx]);
                    const int owner = __float_as_int(at.w);
                    const float dens = at.z;
                    for (int t = 0; t < nT; t++) {
                        if (owner == g * 32 + t0 + t) continue;
                        const float4 r0 = rec[t0 + t], r1 = rec[32 + t0 + t];
                        const f3 qd = mk3(r1.x, r1.y, r1.z);
                        float c;
                        if (sec == 0) {
                            const float4 s = gv.sph[idx];
                            const f3 oc = sub3(Pp, mk3(s.x, s.y, s.z));
                            c = sphere_loss_exact(dot3(oc, qd), subr(dot3(oc, oc), s.w), dens);
                        } else if (sec == 1) {
                            c = aabb_loss<8>(gv.aabbA[idx], gv.aabbB[idx], Pp, r0.x, r0.y, r0.z, dens);
                        } else {
                            const float4 q4 = gv.obbQ[idx], c4 = gv.obbC[idx]; const float2 h2 = gv.obbH[idx];
                            const f3 pc = sub3(Pp, mk3(c4.x, c4.y, c4.z));
                            c = obb_loss_exact(q4.x, q4.y, q4.z, q4.w, pc.x, pc.y, pc.z, c4.w, h2.x, h2.y, qd.x, qd.y, qd.z, dens);
                        }
                        acc[t * kAccStride + lane] = addr(acc[t * kAccStride + lane], c);
                    }
                }
                __syncwarp();
                // lane t adds its target's 32 per-lane partials in lane order (fixed order); value = N*S - loss (PM:260)
                if (lane < nT) {
                    float loss = 0.0f;
#pragma unroll 8
                    for (int i = 0; i < 32; i++) loss = addr(loss, acc[lane * kAccStride + i]);
                    const float v = subr(a.nTimesS, loss);
                    const int tgt = g * 32 + t0 + lane;
                    // deterministic accumulation over rays: integer part + 36-bit fixed-point fraction
                    const float ip = truncf(v);
                    const long long ipart = (long long)ip;
                    const long long fpart = (long long)(((double)v - (double)ip) * 68719476736.0);
                    atomicAdd(reinterpret_cast<unsigned long long*>(&a.permSumInt[tgt]), (unsigned long long)ipart);
                    atomicAdd(reinterpret_cast<unsigned long long*>(&a.permSumFrac[tgt]), (unsigned long long)fpart);
                }
                __syncwarp();
            }
        }
    }
    if (lane == 0) {
        atomicAdd(&a.counters[C_PERM_RAYS], (unsigned long long)nRays);
        atomicAdd(&a.counters[C_PERM_HIT_RAYS], (unsigned long long)nHitRays);
    }
}

// One warp per (batch k, target a): the reference's value for the last hitting ray of batch k,
// summed in the reference's sequential order (spheres, AABBs, OBBs; PM:229-258). Lanes evaluate 32
// colliders at a time; the non-zero contributions are then added one by one in index order
// (adding +0 never changes the running FP32 sum, so skipping the zeros is exact).
__global__ void __launch_bounds__(128) perm_last_kernel(const PermArgs a, int T)
{
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int k = blockIdx.x;
    const int tgt = blockIdx.y * 4 + wib;
    if (k >= T || tgt >= a.nTargets) return;
    const int rayIndex = a.lastHitRay[k];
    if (rayIndex < 0) { if (lane == 0) a.permLast[k * a.nTargets + tgt] = 0.0f; return; }
    // global -> local index of this shard
    int j = rayIndex;
    if (a.map.shardCount > 1) j = ((rayIndex / a.map.chunk) / a.map.shardCount) * a.map.chunk + rayIndex % a.map.chunk;
    const GeomView gv = make_view(a.geom, a.L);
    const f3 d = mk3(um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex)]), um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 1]),
                     um_f16tof32(a.dirs[3 * a.map.dir_index(j, rayIndex) + 2]));
    const f3 o = mk3(a.ox, a.oy, a.oz);
    const float t = a.firstHitDist[j];
    const f3 P = add3(o, mul3s(d, t));
    const f3 Pp = sub3(P, mul3s(d, kEpsilon));
    const f3 T3 = mk3(a.targets[3 * tgt], a.targets[3 * tgt + 1], a.targets[3 * tgt + 2]);
    const f3 dir = normalize3(sub3(T3, Pp));
    const float ix = rcpr(dir.x), iy = rcpr(dir.y), iz = rcpr(dir.z);
    float loss = 0.0f;
    auto fold = [&](float c) {
        uint32_t m = __ballot_sync(kFull, __float_as_uint(c) != 0u && __float_as_uint(c) != 0x80000000u);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            loss = addr(loss, __shfl_sync(kFull, c, src));
        }
    };
    for (int base = 0; base < a.L.ns; base += 32) {
        const int i = base + lane;
        float c = 0.0f;
        if (i < a.L.ns && (int)a.at.ownS[i] != tgt) {
            const float4 s = gv.sph[i];
            const f3 oc = sub3(Pp, mk3(s.x, s.y, s.z));
            c = sphere_loss_exact(dot3(oc, dir), subr(dot3(oc, oc), s.w), a.at.sphAttr[i].z);
        }
        fold(c);
    }
    for (int base = 0; base < a.L.na; base += 32) {
        const int i = base + lane;
        float c = 0.0f;
        if (i < a.L.na && (int)a.at.ownA[i] != tgt) c = aabb_loss<8>(gv.aabbA[i], gv.aabbB[i], Pp, ix, iy, iz, a.at.aabbAttr[i].z);
        fold(c);
    }
    for (int base = 0; base < a.L.no; base += 32) {
        const int i = base + lane;
        float c = 0.0f;
        if (i < a.L.no && (int)a.at.ownO[i] != tgt) {
            const float4 q4 = gv.obbQ[i], c4 = gv.obbC[i]; const float2 h2 = gv.obbH[i];
            const f3 pc = sub3(Pp, mk3(c4.x, c4.y, c4.z));
            c = obb_loss_exact(q4.x, q4.y, q4.z, q4.w, pc.x, pc.y, pc.z, c4.w, h2.x, h2.y, dir.x, dir.y, dir.z, a.at.obbAttr[i].z);
        }
        fold(c);
    }
    if (lane == 0) a.permLast[k * a.nTargets + tgt] = subr(a.nTimesS, loss);   // PM:260
}

cudaError_t launch_perm_last(const PermArgs& a, int T, cudaStream_t stream);

size_t perm_smem_bytes(const GeomLayout& L, bool geomInSmem)
{
    return (geomInSmem ? L.bytes : 0) + (size_t)kWarpsPerCta * kPermWarpBytes;
}

cudaError_t launch_permeation(const PermArgs& a, int numCtas, bool geomInSmem, int T, cudaStream_t stream)
{
    const size_t smem = perm_smem_bytes(a.L, geomInSmem);
    void (*k)(const PermArgs) = geomInSmem ? permeation_kernel<true> : permeation_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<numCtas, kThreads, smem, stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_perm_last(a, T, stream);
}

cudaError_t launch_perm_last(const PermArgs& a, int T, cudaStream_t stream)
{
    dim3 grid((unsigned)T, (unsigned)((a.nTargets + 3) / 4));
    perm_last_kernel<<<grid, 128, 0, stream>>>(a, T);
    return cudaGetLastError();
}

}  // namespace art
