/*
 * audiort_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See audiort_oracle.h for scope, citations and the "PARITY UNPINNED" note.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off: managed C# never contracts a*b+c into an FMA; every
 * operation below is a separately rounded IEEE binary32 operation, written in
 * the order the reference writes it.
 *
 * Citation shorthands: RT/PM/PA/FIB as in the header; UM = Unity.Mathematics
 * 1.3.2 math.cs (restated, package source not in the reference tree).
 */
#include "audiort_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ========================================================================== */
/* Unity.Mathematics subset                                                    */
/* ========================================================================== */

typedef struct { float x, y, z; } f3;
typedef struct { float x, y, z, w; } f4; /* quaternion.value */

static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* UM math.f32tof16: mask off the low 12 mantissa bits, rescale by 2^-112, clamp,
 * add 0x1000, shift by 13 => round-to-nearest, ties AWAY from zero. */
uint16_t or_f32tof16(float x)
{
    const uint32_t infinity_32 = 255u << 23;
    const uint32_t msk = 0x7FFFF000u;
    uint32_t ux = as_uint(x);
    uint32_t uux = ux & msk;
    uint32_t sb = as_uint(as_float(uux) * 1.92592994e-34f); /* 2^-112 */
    if (sb > 0x0F7FF000u) sb = 0x0F7FF000u;                 /* clamp to (signed) infinity if overflowed */
    uint32_t h = (sb + 0x1000u) >> 13;
    if (uux >= infinity_32) h = (uux > infinity_32) ? 0x7e00u : 0x7c00u; /* NaN -> qNaN, Inf -> Inf */
    return (uint16_t)(h | ((ux & ~msk) >> 16));
}

/* UM math.f16tof32: exact. */
float or_f16tof32(uint16_t h)
{
    const uint32_t shifted_exp = 0x7c00u << 13;
    uint32_t uf = ((uint32_t)h & 0x7fffu) << 13;
    uint32_t e = uf & shifted_exp;
    uf += (127u - 15u) << 23;
    if (e == shifted_exp) uf += (128u - 16u) << 23;
    if (e == 0) uf = as_uint(as_float(uf + (1u << 23)) - 6.10351563e-05f);
    return as_float(uf | (((uint32_t)h & 0x8000u) << 16));
}

static inline f3 F3(float x, float y, float z) { f3 r = { x, y, z }; return r; }
static inline f3 h3(const uint16_t* h) { return F3(or_f16tof32(h[0]), or_f16tof32(h[1]), or_f16tof32(h[2])); }
static inline f3 add3(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline f3 sub3(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline f3 mul3(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline f3 mul3s(f3 a, float s) { return F3(a.x * s, a.y * s, a.z * s); }
static inline f3 smul3(float s, f3 a) { return F3(s * a.x, s * a.y, s * a.z); }

/* UM: min(x,y) = (isnan(y) || x < y) ? x : y ; max likewise with > */
static inline float um_min(float x, float y) { return (isnan(y) || x < y) ? x : y; }
static inline float um_max(float x, float y) { return (isnan(y) || x > y) ? x : y; }
static inline f3 min3(f3 a, f3 b) { return F3(um_min(a.x, b.x), um_min(a.y, b.y), um_min(a.z, b.z)); }
static inline f3 max3(f3 a, f3 b) { return F3(um_max(a.x, b.x), um_max(a.y, b.y), um_max(a.z, b.z)); }
static inline float um_abs(float x) { return as_float(as_uint(x) & 0x7FFFFFFFu); }
static inline f3 abs3(f3 a) { return F3(um_abs(a.x), um_abs(a.y), um_abs(a.z)); }
/* UM: sign(x) = (x > 0 ? 1 : 0) - (x < 0 ? 1 : 0) */
static inline float um_sign(float x) { return (x > 0.0f ? 1.0f : 0.0f) - (x < 0.0f ? 1.0f : 0.0f); }
/* UM: saturate(x) = clamp(x,0,1) = max(0, min(1, x)) */
static inline float um_saturate(float x) { return um_max(0.0f, um_min(1.0f, x)); }
/* UM: dot(float3) = a.x*b.x + a.y*b.y + a.z*b.z (left to right) */
static inline float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float dot4(f4 a, f4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
/* UM: sqrt(float) = (float)Math.Sqrt((double)x) == correctly rounded sqrtf */
static inline float um_sqrt(float x) { return sqrtf(x); }
/* UM: rsqrt(x) = 1.0f / sqrt(x) */
static inline float um_rsqrt(float x) { return 1.0f / um_sqrt(x); }
/* UM: normalize(float3 v) = rsqrt(dot(v,v)) * v */
static inline f3 normalize3(f3 v) { return smul3(um_rsqrt(dot3(v, v)), v); }
/* UM: length(v) = sqrt(dot(v,v)); distance(x,y) = length(y - x) */
static inline float distance3(f3 x, f3 y) { f3 d = sub3(y, x); return um_sqrt(dot3(d, d)); }
/* UM: cross(x,y) = (x * y.yzx - x.yzx * y).yzx */
static inline f3 cross3(f3 a, f3 b)
{
    return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* UM: reflect(i,n) = i - 2f * n * dot(i,n)   ((2f*n) * dot) */
static inline f3 reflect3(f3 i, f3 n) { return sub3(i, mul3s(smul3(2.0f, n), dot3(i, n))); }
/* UM: mul(quaternion q, float3 v): t = 2*cross(q.xyz, v); return v + q.w*t + cross(q.xyz, t) */
static inline f3 qmul3(f4 q, f3 v)
{
    f3 qv = F3(q.x, q.y, q.z);
    f3 t = smul3(2.0f, cross3(qv, v));
    return add3(add3(v, smul3(q.w, t)), cross3(qv, t));
}
/* UM: inverse(quaternion q) = rcp(dot(q,q)) * q * float4(-1,-1,-1,1) ; rcp(x) = 1.0f/x */
static inline f4 qinverse(f4 q)
{
    float r = 1.0f / dot4(q, q);
    f4 o = { (r * q.x) * -1.0f, (r * q.y) * -1.0f, (r * q.z) * -1.0f, (r * q.w) * 1.0f };
    return o;
}
/* DT/halfQuaternion.cs:34-46 QuaternionValue getter (runs on every .Rotation access,
 * CS/ColliderOBBStruct.cs:14-16): w = sqrt(max(0, 1-|xyz|^2)), then UM normalize(quaternion)
 * = rsqrt(dot(q,q)) * q. */
static inline f4 obb_rotation(const OrOBB* o)
{
    float xx = or_f16tof32(o->rot[0]);
    float yy = or_f16tof32(o->rot[1]);
    float zz = or_f16tof32(o->rot[2]);
    float wSquared = 1.0f - (xx * xx + yy * yy + zz * zz);
    float w = wSquared > 0.0f ? um_sqrt(wSquared) : 0.0f;
    f4 q = { xx, yy, zz, w };
    float rs = um_rsqrt(dot4(q, q));
    f4 n = { rs * q.x, rs * q.y, rs * q.z, rs * q.w };
    return n;
}

#define OR_EPSILON 0.0001f /* RT:57, PM:30 */

/* ========================================================================== */
/* Intersection tests                                                          */
/* ========================================================================== */

/* RT:284-308 == PM:145-169 */
static inline int ray_aabb(f3 rayOrigin, f3 rayDir, f3 Center, f3 halfExtents, float* distance)
{
    f3 mn = sub3(Center, halfExtents);
    f3 mx = add3(Center, halfExtents);
    f3 invDir = F3(1.0f / rayDir.x, 1.0f / rayDir.y, 1.0f / rayDir.z);
    f3 t0 = mul3(sub3(mn, rayOrigin), invDir);
    f3 t1 = mul3(sub3(mx, rayOrigin), invDir);
    f3 tmin = min3(t0, t1);
    f3 tmax = max3(t0, t1);
    float tNear = um_max(um_max(tmin.x, tmin.y), tmin.z);
    float tFar = um_min(um_min(tmax.x, tmax.y), tmax.z);
    if (tNear > tFar || tFar < 0) { *distance = 0; return 0; }
    *distance = tNear > 0 ? tNear : tFar;
    return 1;
}

/* RT:314-320: stored Rotation is used directly as "invRotation". */
static inline int ray_obb_rt(f3 rayOrigin, f3 rayDir, f3 Center, f3 halfExtents, f4 invRotation, float* distance)
{
    f3 localOrigin = qmul3(invRotation, sub3(rayOrigin, Center));
    f3 localDir = qmul3(invRotation, rayDir);
    return ray_aabb(localOrigin, localDir, F3(0, 0, 0), halfExtents, distance);
}

/* PM:172-179: applies math.inverse() to the (already inverted) stored rotation (quirk Q4). */
static inline int ray_obb_pm(f3 rayOrigin, f3 rayDir, f3 Center, f3 halfExtents, f4 rotation, float* distance)
{
    f4 invRotation = qinverse(rotation);
    f3 localOrigin = qmul3(invRotation, sub3(rayOrigin, Center));
    f3 localDir = qmul3(invRotation, rayDir);
    return ray_aabb(localOrigin, localDir, F3(0, 0, 0), halfExtents, distance);
}

/* RT:323-355 == PM:182-214 */
static inline int ray_sphere(f3 rayOrigin, f3 rayDir, f3 Center, float Radius, float* distance)
{
    f3 oc = sub3(rayOrigin, Center);
    float a = dot3(rayDir, rayDir);
    float b = 2.0f * dot3(oc, rayDir);
    float c = dot3(oc, oc) - Radius * Radius;
    float discriminant = b * b - 4 * a * c;
    if (discriminant < 0) { *distance = 0; return 0; }
    float sqrtDiscriminant = um_sqrt(discriminant);
    float t0 = (-b - sqrtDiscriminant) / (2.0f * a);
    float t1 = (-b + sqrtDiscriminant) / (2.0f * a);
    if (t0 >= 0) { *distance = t0; return 1; }
    else if (t1 >= 0) { *distance = t1; return 1; }
    *distance = 0;
    return 0;
}

/* ========================================================================== */
/* AudioRaytracerJobBatched                                                    */
/* ========================================================================== */

typedef struct { uint32_t type; int32_t index; } HitRef;

/* RT:225-280: spheres, then AABBs, then OBBs; strict '<' keeps the earliest on ties. */
static int rt_shoot_raycast(const OrScene* s, f3 o, f3 d, HitRef* hit, float* closestDist, OrCounters* c)
{
    float dist;
    *closestDist = 3.402823466e+38f; /* float.MaxValue */
    hit->type = OR_TYPE_NONE; hit->index = -1;
    for (int i = 0; i < s->nSphere; i++) {
        const OrSphere* t = &s->spheres[i];
        if (ray_sphere(o, d, h3(t->center), or_f16tof32(t->radius), &dist) && dist < *closestDist) {
            hit->type = OR_TYPE_SPHERE; hit->index = i; *closestDist = dist;
        }
    }
    for (int i = 0; i < s->nAABB; i++) {
        const OrAABB* t = &s->aabbs[i];
        if (ray_aabb(o, d, h3(t->center), h3(t->size), &dist) && dist < *closestDist) {
            hit->type = OR_TYPE_AABB; hit->index = i; *closestDist = dist;
        }
    }
    for (int i = 0; i < s->nOBB; i++) {
        const OrOBB* t = &s->obbs[i];
        if (ray_obb_rt(o, d, h3(t->center), h3(t->size), obb_rotation(t), &dist) && dist < *closestDist) {
            hit->type = OR_TYPE_OBB; hit->index = i; *closestDist = dist;
        }
    }
    if (c) { c->trace_tests[0] += s->nSphere; c->trace_tests[1] += s->nAABB; c->trace_tests[2] += s->nOBB; }
    return hit->type != OR_TYPE_NONE;
}

/* RT:365-397 (audioTargetId < 0) and RT:405-449 (audioTargetId >= 0: skip owned colliders). */
static int rt_can_ray_see(const OrScene* s, f3 o, f3 d, float distLimit, int skipOwned, int audioTargetId,
                          uint64_t tests[3])
{
    float dist;
    for (int i = 0; i < s->nSphere; i++) {
        const OrSphere* t = &s->spheres[i];
        if (skipOwned && t->audioTargetId == audioTargetId) continue;
        tests[0]++;
        if (ray_sphere(o, d, h3(t->center), or_f16tof32(t->radius), &dist) && dist < distLimit) return 0;
    }
    for (int i = 0; i < s->nAABB; i++) {
        const OrAABB* t = &s->aabbs[i];
        if (skipOwned && t->audioTargetId == audioTargetId) continue;
        tests[1]++;
        if (ray_aabb(o, d, h3(t->center), h3(t->size), &dist) && dist < distLimit) return 0;
    }
    for (int i = 0; i < s->nOBB; i++) {
        const OrOBB* t = &s->obbs[i];
        if (skipOwned && t->audioTargetId == audioTargetId) continue;
        tests[2]++;
        if (ray_obb_rt(o, d, h3(t->center), h3(t->size), obb_rotation(t), &dist) && dist < distLimit) return 0;
    }
    return 1;
}

/* RT:456-532 */
static void rt_reflect_ray(const OrScene* s, HitRef hit, f3* cRayOrigin, f3* cRayDir, float* cRayLife)
{
    f3 normal = F3(0, 0, 0);
    float absorption = 0;
    switch (hit.type) {
    case OR_TYPE_AABB: {
        const OrAABB* b = &s->aabbs[hit.index];
        f3 localPoint = sub3(*cRayOrigin, h3(b->center));
        f3 absPoint = abs3(localPoint);
        f3 halfExtents = h3(b->size);
        if (halfExtents.x - absPoint.x < halfExtents.y - absPoint.y && halfExtents.x - absPoint.x < halfExtents.z - absPoint.z)
            normal.x = um_sign(localPoint.x);
        else if (halfExtents.y - absPoint.y < halfExtents.x - absPoint.x && halfExtents.y - absPoint.y < halfExtents.z - absPoint.z)
            normal.y = um_sign(localPoint.y);
        else
            normal.z = um_sign(localPoint.z);
        absorption = or_f16tof32(b->absorption);
        break;
    }
    case OR_TYPE_OBB: {
        const OrOBB* b = &s->obbs[hit.index];
        /* quirk Q3: inverse(Rotation) here, Rotation for the normal (RT:489, 510) */
        f3 localHit = qmul3(qinverse(obb_rotation(b)), sub3(*cRayOrigin, h3(b->center)));
        f3 localHalfExtents = h3(b->size);
        f3 absPointOBB = abs3(localHit);
        f3 delta = sub3(localHalfExtents, absPointOBB);
        f3 localNormal = F3(0, 0, 0);
        if (delta.x < delta.y && delta.x < delta.z) localNormal.x = um_sign(localHit.x);
        else if (delta.y < delta.x && delta.y < delta.z) localNormal.y = um_sign(localHit.y);
        else localNormal.z = um_sign(localHit.z);
        normal = qmul3(obb_rotation(b), localNormal);
        absorption = or_f16tof32(b->absorption);
        break;
    }
    case OR_TYPE_SPHERE: {
        const OrSphere* b = &s->spheres[hit.index];
        normal = normalize3(sub3(*cRayOrigin, h3(b->center)));
        absorption = or_f16tof32(b->absorption);
        break;
    }
    default: break;
    }
    *cRayDir = reflect3(*cRayDir, normal);                       /* RT:525 */
    *cRayOrigin = add3(*cRayOrigin, mul3s(*cRayDir, OR_EPSILON)); /* RT:528 */
    *cRayLife -= s->maxRayLife * absorption;                     /* RT:531 */
}

static inline int32_t mul_i32_wrap(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

/* RT:61-215 */
void or_rt_execute(const OrScene* s, const OrOutputs* o, int32_t rayStartIndex, int32_t totalRays,
                   int faithfulReset, OrCounters* c)
{
    const int32_t Na = s->nTargets;
    const int32_t H = s->maxHitsPerRay;
    const int32_t muffleLen = s->batchCount * Na;                   /* MuffleRayHits.Length (ATM:112) */
    int32_t batchCount = muffleLen / Na;                            /* RT:63 */
    int32_t batchId = mul_i32_wrap(rayStartIndex, batchCount) / s->rayCount; /* RT:64 */
    const f3 RayOrigin = F3(s->rayOrigin[0], s->rayOrigin[1], s->rayOrigin[2]);
    OrCounters local; memset(&local, 0, sizeof local);

    if (faithfulReset) {                                            /* RT:72-80, quirk Q1 */
        for (int32_t i = 0; i < totalRays * H; i++) {
            int32_t rayIndex = rayStartIndex + i;
            if (o->echoRayDistances) o->echoRayDistances[rayIndex] = 0;
            if (o->rayHitResults) { o->rayHitResults[3 * rayIndex] = 0; o->rayHitResults[3 * rayIndex + 1] = 0; o->rayHitResults[3 * rayIndex + 2] = 0; }
        }
    }
    if (o->muffleRayHits)                                           /* RT:82-85 */
        for (int32_t i = 0; i < Na; i++) o->muffleRayHits[batchId * Na + i] = 0;

    for (int32_t localRayId = 0; localRayId < totalRays; localRayId++) {   /* RT:90 */
        int32_t rayIndex = rayStartIndex + localRayId;
        f3 cRayDir = h3(&s->rayDirections[3 * rayIndex]);           /* RT:94 */
        f3 cRayOrigin = RayOrigin;
        uint8_t cRayHits = 0;
        float cRayLife = s->maxRayLife;
        int isRayAlive = 1;

        while (isRayAlive) {                                        /* RT:104 */
            HitRef hit; float rayHitDist;
            local.segments++;
            if (rt_shoot_raycast(s, cRayOrigin, cRayDir, &hit, &rayHitDist, &local)) {
                local.segment_hits++;
                cRayOrigin = add3(cRayOrigin, mul3s(cRayDir, rayHitDist));     /* RT:111 */
                cRayLife -= rayHitDist;                                        /* RT:112 */
                cRayHits += 1;                                                 /* RT:113 */
                int32_t rayResultId = rayIndex * H + cRayHits - 1;             /* RT:115 */
                uint16_t hitPoint[3] = { or_f32tof16(cRayOrigin.x), or_f32tof16(cRayOrigin.y), or_f32tof16(cRayOrigin.z) }; /* RT:118 */
                if (o->hitColliderIds) o->hitColliderIds[rayResultId] = (hit.type << 30) | (uint32_t)hit.index;
                if (o->hitDistances) o->hitDistances[rayResultId] = rayHitDist;

                /* echo ray RT:124-145 */
                f3 offsetted = sub3(cRayOrigin, mul3s(cRayDir, OR_EPSILON));   /* RT:124 */
                f3 returnRayDir = normalize3(sub3(RayOrigin, offsetted));      /* RT:127 */
                float distToStartOrigin = distance3(RayOrigin, cRayOrigin);    /* RT:130 */
                local.echo_queries++;
                if (rt_can_ray_see(s, offsetted, returnRayDir, distToStartOrigin, 0, -1, local.echo_tests)) {
                    uint16_t echoHalf;
                    switch (hit.type) {                                        /* RT:135-141 */
                    case OR_TYPE_AABB: echoHalf = s->aabbs[hit.index].echo; break;
                    case OR_TYPE_OBB: echoHalf = s->obbs[hit.index].echo; break;
                    case OR_TYPE_SPHERE: echoHalf = s->spheres[hit.index].echo; break;
                    default: echoHalf = 0x3C00; break;
                    }
                    /* Half.Multiply(in float,in float,out half) UT/HalfDataTypesUtility.cs:86-90 */
                    uint16_t echoRayPower = or_f32tof16(distToStartOrigin * or_f16tof32(echoHalf));
                    if (o->echoRayDistances) o->echoRayDistances[rayResultId] = echoRayPower;  /* RT:144 */
                }

                /* muffle rays RT:153-173 */
                for (int32_t a = 0; a < Na; a++) {
                    int32_t muffleRayId = batchId * Na + a;
                    offsetted = sub3(cRayOrigin, mul3s(cRayDir, OR_EPSILON));  /* RT:158 */
                    f3 tp = F3(s->targetPositions[3 * a], s->targetPositions[3 * a + 1], s->targetPositions[3 * a + 2]);
                    f3 rayToTargetDir = normalize3(sub3(tp, offsetted));       /* RT:162 */
                    float distToTarget = distance3(offsetted, tp);             /* RT:165 */
                    if (distToTarget < s->maxMuffleHitDistance) {              /* RT:168 */
                        local.muffle_queries++;
                        if (rt_can_ray_see(s, offsetted, rayToTargetDir, distToTarget, 1, a, local.muffle_tests)) {
                            if (o->muffleRayHits) o->muffleRayHits[muffleRayId] = (uint16_t)(o->muffleRayHits[muffleRayId] + 1); /* RT:171 */
                            if (o->muffleTotals) __atomic_fetch_add(&o->muffleTotals[a], 1u, __ATOMIC_RELAXED);
                        }
                    }
                }

                if (cRayHits >= H || cRayLife <= 0) {                          /* RT:179 */
                    isRayAlive = 0;
                } else {
                    rt_reflect_ray(s, hit, &cRayOrigin, &cRayDir, &cRayLife);  /* RT:186 */
                    if (cRayLife < 0) isRayAlive = 0;                          /* RT:189 */
                }
                if (o->rayHitResults) {                                        /* RT:197 */
                    o->rayHitResults[3 * rayResultId] = hitPoint[0];
                    o->rayHitResults[3 * rayResultId + 1] = hitPoint[1];
                    o->rayHitResults[3 * rayResultId + 2] = hitPoint[2];
                }
            } else {
                if (o->rayHitResultCounts) o->rayHitResultCounts[rayIndex] = cRayHits; /* RT:204 */
                break;
            }
        }
        if (o->rayHitResultCounts) o->rayHitResultCounts[rayIndex] = cRayHits;         /* RT:212 */
    }

    if (c) {
        __atomic_fetch_add(&c->segments, local.segments, __ATOMIC_RELAXED);
        __atomic_fetch_add(&c->segment_hits, local.segment_hits, __ATOMIC_RELAXED);
        __atomic_fetch_add(&c->echo_queries, local.echo_queries, __ATOMIC_RELAXED);
        __atomic_fetch_add(&c->muffle_queries, local.muffle_queries, __ATOMIC_RELAXED);
        for (int k = 0; k < 3; k++) {
            __atomic_fetch_add(&c->trace_tests[k], local.trace_tests[k], __ATOMIC_RELAXED);
            __atomic_fetch_add(&c->echo_tests[k], local.echo_tests[k], __ATOMIC_RELAXED);
            __atomic_fetch_add(&c->muffle_tests[k], local.muffle_tests[k], __ATOMIC_RELAXED);
        }
    }
}

/* ========================================================================== */
/* AudioPermeationJobBatched                                                   */
/* ========================================================================== */

/* PM:101-141 */
static int pm_shoot_raycast(const OrScene* s, f3 o, f3 d, float* closestDist, OrCounters* c)
{
    float dist;
    *closestDist = INFINITY;
    for (int i = 0; i < s->nSphere; i++) {
        const OrSphere* t = &s->spheres[i];
        if (ray_sphere(o, d, h3(t->center), or_f16tof32(t->radius), &dist) && dist < *closestDist) *closestDist = dist;
    }
    for (int i = 0; i < s->nAABB; i++) {
        const OrAABB* t = &s->aabbs[i];
        if (ray_aabb(o, d, h3(t->center), h3(t->size), &dist) && dist < *closestDist) *closestDist = dist;
    }
    for (int i = 0; i < s->nOBB; i++) {
        const OrOBB* t = &s->obbs[i];
        if (ray_obb_pm(o, d, h3(t->center), h3(t->size), obb_rotation(t), &dist) && dist < *closestDist) *closestDist = dist;
    }
    c->perm_first_tests[0] += s->nSphere; c->perm_first_tests[1] += s->nAABB; c->perm_first_tests[2] += s->nOBB;
    return *closestDist != INFINITY;
}

/* PM:265-288 */
static inline void aabb_permeation(f3 rayOrigin, f3 rayDir, f3 Center, f3 halfExtents, float density, float* loss)
{
    f3 mn = sub3(Center, halfExtents);
    f3 mx = add3(Center, halfExtents);
    f3 invDir = F3(1.0f / rayDir.x, 1.0f / rayDir.y, 1.0f / rayDir.z);
    f3 t0 = mul3(sub3(mn, rayOrigin), invDir);
    f3 t1 = mul3(sub3(mx, rayOrigin), invDir);
    f3 tmin = min3(t0, t1);
    f3 tmax = max3(t0, t1);
    float tEnter = um_max(um_max(tmin.x, tmin.y), tmin.z);
    float tExit = um_min(um_min(tmax.x, tmax.y), tmax.z);
    if (tEnter > tExit || tExit < 0.0f) return;
    float enter = um_max(tEnter, 0.0f);
    *loss += um_max(0.0f, tExit - enter) * density;
}

/* PM:294-300: stored rotation used directly. */
static inline void obb_permeation(f3 rayOrigin, f3 rayDir, f3 Center, f3 halfExtents, f4 invRotation, float density, float* loss)
{
    f3 localOrigin = qmul3(invRotation, sub3(rayOrigin, Center));
    f3 localDir = qmul3(invRotation, rayDir);
    aabb_permeation(localOrigin, localDir, F3(0, 0, 0), halfExtents, density, loss);
}

/* PM:303-328: assumes a unit direction. */
static inline void sphere_permeation(f3 rayOrigin, f3 rayDir, f3 Center, float Radius, float density, float* loss)
{
    f3 oc = sub3(rayOrigin, Center);
    float b = dot3(oc, rayDir);
    float c = dot3(oc, oc) - Radius * Radius;
    float discriminant = b * b - c;
    if (discriminant < 0.0f) return;
    float sqrtD = um_sqrt(discriminant);
    float tEnter = -b - sqrtD;
    float tExit = -b + sqrtD;
    if (tExit < 0.0f) return;
    float enter = um_max(tEnter, 0.0f);
    *loss += um_max(0.0f, tExit - enter) * density;
}

/* PM:225-261 (distToStartOrigin parameter is unused by the reference: quirk Q7) */
static float pm_shoot_permeation(const OrScene* s, f3 o, f3 d, int audioTargetId, OrCounters* c)
{
    float loss = 0;
    for (int i = 0; i < s->nSphere; i++) {
        const OrSphere* t = &s->spheres[i];
        if (t->audioTargetId == audioTargetId) continue;
        c->perm_loss_tests[0]++;
        sphere_permeation(o, d, h3(t->center), or_f16tof32(t->radius), or_f16tof32(t->density), &loss);
    }
    for (int i = 0; i < s->nAABB; i++) {
        const OrAABB* t = &s->aabbs[i];
        if (t->audioTargetId == audioTargetId) continue;
        c->perm_loss_tests[1]++;
        aabb_permeation(o, d, h3(t->center), h3(t->size), or_f16tof32(t->density), &loss);
    }
    for (int i = 0; i < s->nOBB; i++) {
        const OrOBB* t = &s->obbs[i];
        if (t->audioTargetId == audioTargetId) continue;
        c->perm_loss_tests[2]++;
        obb_permeation(o, d, h3(t->center), h3(t->size), obb_rotation(t), or_f16tof32(t->density), &loss);
    }
    return (float)s->rayCount * s->permeationStrengthPerRay - loss;   /* PM:260 */
}

/* PM:34-91 */
void or_pm_execute(const OrScene* s, const OrOutputs* o, int32_t rayStartIndex, int32_t totalRays, OrCounters* c)
{
    const int32_t Na = s->nTargets;
    const int32_t permLen = s->batchCount * Na;                     /* PermeationPowerRemains.Length (ATM:121) */
    int32_t batchCount = permLen / totalRays / Na;                  /* PM:36 (quirk Q6: 0 in practice) */
    int32_t batchId = mul_i32_wrap(rayStartIndex, batchCount) / s->rayCount; /* PM:37 */
    const f3 RayOrigin = F3(s->rayOrigin[0], s->rayOrigin[1], s->rayOrigin[2]);
    OrCounters local; memset(&local, 0, sizeof local);

    if (o->permeationPowerRemains)
        for (int32_t i = 0; i < Na; i++) o->permeationPowerRemains[batchId * Na + i] = 0.0f; /* PM:43-46 */

    for (int32_t localRayId = 0; localRayId < totalRays; localRayId++) {
        int32_t rayIndex = rayStartIndex + localRayId;
        f3 cRayDir = h3(&s->rayDirections[3 * rayIndex]);
        f3 cRayOrigin = RayOrigin;
        float rayHitDist;
        local.perm_rays++;
        if (pm_shoot_raycast(s, cRayOrigin, cRayDir, &rayHitDist, &local)) {         /* PM:58 */
            local.perm_hit_rays++;
            cRayOrigin = add3(cRayOrigin, mul3s(cRayDir, rayHitDist));               /* PM:61 */
            for (int32_t a = 0; a < Na; a++) {
                int32_t permeationRayId = batchId * Na + a;
                f3 offsetted = sub3(cRayOrigin, mul3s(cRayDir, OR_EPSILON));         /* PM:72 */
                f3 tp = F3(s->targetPositions[3 * a], s->targetPositions[3 * a + 1], s->targetPositions[3 * a + 2]);
                f3 rayToTargetDir = normalize3(sub3(tp, offsetted));                 /* PM:76 */
                local.perm_pairs++;
                float remains = pm_shoot_permeation(s, offsetted, rayToTargetDir, a, &local); /* PM:82 */
                if (o->permeationPowerRemains) o->permeationPowerRemains[permeationRayId] = remains; /* PM:85 overwrite (Q5) */
                if (o->permeationSum) {
                    o->permeationSum[a] += (double)remains;   /* extension; PM runs serially in the drivers */
                }
            }
        }
    }
    if (c) {
        __atomic_fetch_add(&c->perm_rays, local.perm_rays, __ATOMIC_RELAXED);
        __atomic_fetch_add(&c->perm_hit_rays, local.perm_hit_rays, __ATOMIC_RELAXED);
        __atomic_fetch_add(&c->perm_pairs, local.perm_pairs, __ATOMIC_RELAXED);
        for (int k = 0; k < 3; k++) {
            __atomic_fetch_add(&c->perm_first_tests[k], local.perm_first_tests[k], __ATOMIC_RELAXED);
            __atomic_fetch_add(&c->perm_loss_tests[k], local.perm_loss_tests[k], __ATOMIC_RELAXED);
        }
    }
}

/* ========================================================================== */
/* ProcessAudioDataJob                                                         */
/* ========================================================================== */

/* DT/AudioTargetRTSettings.cs:18-24 constructor */
static OrTargetSettings make_settings(float muffle, float reverbStrength, float reverbVolume, const float* pos)
{
    OrTargetSettings r;
    r.muffleStrength = um_saturate(muffle);
    r.reverbStrength = um_saturate(reverbStrength);
    r.reverbVolume = um_saturate(reverbVolume);
    r.percievedAudioPosition[0] = pos[0]; r.percievedAudioPosition[1] = pos[1]; r.percievedAudioPosition[2] = pos[2];
    return r;
}

/* PA:32-76, sequential FP32 exactly as written (and an FP64 evaluation of the same formulas). */
void or_pa_execute(const OrScene* s, const OrOutputs* o)
{
    const int32_t Na = s->nTargets;
    int32_t maxBatchSize = (s->batchCount * Na) / Na;               /* PA:34 */
    int32_t maxRayHits = (int32_t)s->maxHitsPerRay * s->rayCount;   /* PA:35 */

    float reverbTotal = 0, echoRayReturnedHits = 0;
    double reverbTotal64 = 0, zeros64 = 0;
    for (int32_t i = 0; i < maxRayHits; i++) {                      /* PA:40-48 */
        float e = or_f16tof32(o->echoRayDistances[i]);
        if (e == 0) { echoRayReturnedHits += 1; zeros64 += 1; continue; }
        reverbTotal += e;
        reverbTotal64 += (double)e;
    }
    float avgReverbDist = reverbTotal / maxRayHits;                 /* PA:49 */
    float reverbStrength = avgReverbDist / s->maxReverbDistance;    /* PA:50 */
    float reverbVolume = echoRayReturnedHits / maxRayHits;          /* PA:51 */
    double reverbStrength64 = reverbTotal64 / maxRayHits / (double)s->maxReverbDistance;
    double reverbVolume64 = zeros64 / maxRayHits;

    for (int32_t a = 0; a < Na; a++) {                              /* PA:55 */
        int32_t totalMuffleRayhits = 0;
        float totalPermeationPower = 0;
        double totalPermeationPower64 = 0;
        for (int32_t i = 0; i < maxBatchSize; i++) {                /* PA:61-65 */
            totalMuffleRayhits += o->muffleRayHits[Na * i + a];
            totalPermeationPower += o->permeationPowerRemains[Na * i + a];
            totalPermeationPower64 += (double)o->permeationPowerRemains[Na * i + a];
        }
        /* PA:68: (float)hits / (RayCount*MaxHitsPerRay) [int product -> float] * MuffleEffectiveness */
        float muffle = 1 - (float)totalMuffleRayhits / (float)(s->rayCount * (int32_t)s->maxHitsPerRay) * s->muffleEffectiveness;
        /* PA:69 */
        float permeation = (float)totalPermeationPower / (float)s->rayCount / s->permeationStrengthPerRay * s->permeationEffectiveness;
        muffle = um_saturate(muffle - permeation);                  /* PA:71 */
        if (o->settings) o->settings[a] = make_settings(muffle, reverbStrength, reverbVolume, &s->targetPositions[3 * a]);
        if (o->settingsFp64) {
            double m64 = 1.0 - (double)totalMuffleRayhits / ((double)s->rayCount * s->maxHitsPerRay) * (double)s->muffleEffectiveness;
            double p64 = totalPermeationPower64 / s->rayCount / (double)s->permeationStrengthPerRay * (double)s->permeationEffectiveness;
            o->settingsFp64[a] = make_settings((float)(m64 - p64), (float)reverbStrength64, (float)reverbVolume64, &s->targetPositions[3 * a]);
        }
    }
}

/* ========================================================================== */
/* FibonacciDirectionsJobParallel                                              */
/* ========================================================================== */

/* FIB:25-34. UM cos/sin(float) = (float)Math.Cos((double)x). */
void or_fibonacci_directions(int32_t N, int32_t first, int32_t count, uint16_t* out)
{
    const float PI_F = 3.14159274f;                       /* math.PI as float */
    float phi = PI_F * (3.0f - um_sqrt(5.0f));
    for (int32_t k = 0; k < count; k++) {
        int32_t i = first + k;
        float y = 1.0f - ((float)i / (float)(N - 1)) * 2.0f;
        float radius = um_sqrt(1.0f - y * y);
        float theta = phi * (float)i;
        float x = (float)cos((double)theta) * radius;
        float z = (float)sin((double)theta) * radius;
        out[3 * k] = or_f32tof16(x);
        out[3 * k + 1] = or_f32tof16(y);
        out[3 * k + 2] = or_f32tof16(z);
    }
}

/* ========================================================================== */
/* Frame drivers                                                               */
/* ========================================================================== */

/* ART:161: (int)math.max(1, math.ceil((float)rayCount / ToUseThreadCount)) */
int32_t or_batch_size(int32_t rayCount, int32_t batchCount)
{
    float q = ceilf((float)rayCount / (float)batchCount);
    if (!(q > 1.0f)) q = 1.0f;
    return (int32_t)q;
}

static int check_scene(const OrScene* s)
{
    if (!s || s->rayCount <= 0 || s->nTargets <= 0 || s->maxHitsPerRay == 0 || s->batchCount <= 0) return -1;
    if (s->nAABB < 0 || s->nOBB < 0 || s->nSphere < 0) return -1;
    return 0;
}

static void zero_outputs(const OrScene* s, const OrOutputs* o, int jobs)
{
    size_t NH = (size_t)s->rayCount * s->maxHitsPerRay;
    size_t TNa = (size_t)s->batchCount * s->nTargets;
    if (jobs & OR_JOB_RT) {
        if (o->echoRayDistances) memset(o->echoRayDistances, 0, NH * 2);
        if (o->rayHitResults) memset(o->rayHitResults, 0, NH * 6);
        if (o->rayHitResultCounts) memset(o->rayHitResultCounts, 0, (size_t)s->rayCount);
        if (o->muffleRayHits) memset(o->muffleRayHits, 0, TNa * 2);
        if (o->hitColliderIds) memset(o->hitColliderIds, 0, NH * 4);
        if (o->hitDistances) memset(o->hitDistances, 0, NH * 4);
        if (o->muffleTotals) memset(o->muffleTotals, 0, (size_t)s->nTargets * 4);
    }
    if (jobs & OR_JOB_PM) {
        if (o->permeationPowerRemains) memset(o->permeationPowerRemains, 0, TNa * 4);
        if (o->permeationSum) memset(o->permeationSum, 0, (size_t)s->nTargets * 8);
    }
}

typedef struct {
    const OrScene* s; const OrOutputs* o; OrCounters* c;
    int32_t* next; int32_t nItems; int32_t itemSize; int32_t first; int32_t end; int kind;
} Work;

/* kind 0: RT batches (item = batch k of size itemSize starting at first + k*itemSize);
 * kind 1: RT arbitrary sub-ranges for bounded samples (muffle u16 table not touched);
 * kind 2: PM arbitrary sub-ranges, work/counters only (no slot writes: Q5 is order dependent);
 * kind 3: as 2 but the per-target sums over rays (permeationSum extension) are accumulated -- single thread only. */
static void* worker(void* p)
{
    Work* w = (Work*)p;
    for (;;) {
        int32_t k = __atomic_fetch_add(w->next, 1, __ATOMIC_RELAXED);
        if (k >= w->nItems) break;
        int32_t start = w->first + k * w->itemSize;
        int32_t count = w->end - start < w->itemSize ? w->end - start : w->itemSize;
        if (w->kind == 0) {
            or_rt_execute(w->s, w->o, start, count, 0, w->c);
        } else if (w->kind == 2 || w->kind == 3) {
            OrOutputs o2 = *w->o; o2.permeationPowerRemains = NULL;
            if (w->kind == 2) o2.permeationSum = NULL;
            or_pm_execute(w->s, &o2, start, count, w->c);
        } else {
            OrOutputs o2 = *w->o; o2.muffleRayHits = NULL;
            or_rt_execute(w->s, &o2, start, count, 0, w->c);
        }
    }
    return NULL;
}

static void run_items(const OrScene* s, const OrOutputs* o, OrCounters* c, int kind,
                      int32_t first, int32_t end, int32_t itemSize, int nThreads)
{
    int32_t nItems = (end - first + itemSize - 1) / itemSize;
    int32_t next = 0;
    Work w = { s, o, c, &next, nItems, itemSize, first, end, kind };
    if (nThreads <= 1) { worker(&w); return; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nThreads);
    for (int t = 0; t < nThreads; t++) pthread_create(&th[t], NULL, worker, &w);
    for (int t = 0; t < nThreads; t++) pthread_join(th[t], NULL);
    free(th);
}

int or_run_frame(const OrScene* s, const OrOutputs* o, int jobs, int nThreads, OrCounters* c)
{
    if (check_scene(s) || !o) return -1;
    OrCounters dummy; if (!c) { c = &dummy; } memset(c, 0, sizeof *c);
    zero_outputs(s, o, jobs);
    int32_t b = or_batch_size(s->rayCount, s->batchCount);
    if (jobs & OR_JOB_RT)
        run_items(s, o, c, 0, 0, s->rayCount, b, nThreads);
    if (jobs & OR_JOB_PM) {
        /* canonical serial order (Q5: last writer wins) */
        for (int32_t start = 0; start < s->rayCount; start += b) {
            int32_t count = s->rayCount - start < b ? s->rayCount - start : b;
            or_pm_execute(s, o, start, count, c);
        }
    }
    if (jobs & OR_JOB_PA) {
        if (!o->echoRayDistances || !o->muffleRayHits || !o->permeationPowerRemains) return -1;
        or_pa_execute(s, o);
    }
    return 0;
}

int or_run_frame_faithful_q1(const OrScene* s, const OrOutputs* o, OrCounters* c)
{
    if (check_scene(s) || !o) return -1;
    OrCounters dummy; if (!c) { c = &dummy; } memset(c, 0, sizeof *c);
    zero_outputs(s, o, OR_JOB_RT | OR_JOB_PM);
    int32_t b = or_batch_size(s->rayCount, s->batchCount);
    for (int32_t start = 0; start < s->rayCount; start += b) {
        int32_t count = s->rayCount - start < b ? s->rayCount - start : b;
        or_rt_execute(s, o, start, count, 1, c);
    }
    return 0;
}

int or_trace_range(const OrScene* s, const OrOutputs* o, int32_t first, int32_t count, int nThreads, OrCounters* c)
{
    if (check_scene(s) || !o || first < 0 || count < 0 || first + count > s->rayCount) return -1;
    OrCounters dummy; if (!c) { c = &dummy; } memset(c, 0, sizeof *c);
    int32_t item = 16;
    run_items(s, o, c, 1, first, first + count, item, nThreads);
    return 0;
}

int or_permeation_range(const OrScene* s, const OrOutputs* o, int32_t first, int32_t count, int nThreads, OrCounters* c)
{
    if (check_scene(s) || !o || first < 0 || count < 0 || first + count > s->rayCount) return -1;
    OrCounters dummy; if (!c) { c = &dummy; } memset(c, 0, sizeof *c);
    /* one thread: the caller's permeationSum (zeroed by the caller) receives the window's per-target sums */
    run_items(s, o, c, nThreads <= 1 && o->permeationSum ? 3 : 2, first, first + count, 16, nThreads);
    return 0;
}
