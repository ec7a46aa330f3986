// k2_permeation_binned.cu -- the loss lines of AudioPermeationJobBatched (PM:67-86, 225-261) grouped by
// (audio target, direction bin of the target's fan).
//
// permeation_grid_kernel (k2_permeation_grid.cu) gives every lane its own (ray, target) line, and the lanes of a warp then
// walk fan lists of very different lengths: 14.5 of 32 lanes active (ncu, round 1d/1e). But a line's lists are decided by
// two things only -- its target and the direction bin its hit point lies in as seen from that target (fan_dev.cuh). So the
// lines are first SORTED by (target, bin) -- a counting sort in shared memory, three small kernels -- and a warp then takes 32
// lines of ONE (target, bin) at a time: all 32 lanes walk the same three lists, every collider index is a warp-uniform load,
// every geometry read a shared-memory broadcast, and the loops have no divergence apart from the tests' own early outs.
//
//   perm_bin_count_kernel    one CTA per (ray slice, target): histogram of the slice's hit points over the target's bins
//   perm_bin_scan_kernel     one CTA per target: exclusive scan of the (bin, slice) counts -> where every piece is written
//   perm_bin_scatter_kernel  one CTA per (ray slice, target): local ray indices written bin by bin
//   perm_bin_blocks_kernel   the sorted lines of a target are cut into blocks of kBinBlock lines (the unit of work of the
//                            loss kernel: hit points cluster around the listener, so seen from a distant target a few bins
//                            hold thousands of lines each); per block, the bin its first line belongs to
//   perm_loss_binned_kernel  persistent; a warp pops (target, block) jobs and evaluates the block's lines bin by bin, 32 of
//                            one bin at a time
//
// The per-line arithmetic -- set-up, tests, the order AABBs | spheres | OBBs, near list | bin | opposite bin -- is that of the
// fan path of permeation_grid_kernel, statement for statement, so every line's value is the same float; the per-target sums
// are integer (trunc + 36-bit fixed-point fraction) and therefore do not depend on the order of the lines: the binned path
// returns bit-identical permeationSum values. First-hit distances (phase 1) still come from permeation_grid_kernel, which
// in this mode stores the hit points instead of evaluating the lines (PermArgs::hitPts).
#include "device_util.cuh"
#include "fan_dev.cuh"
#include "grid_dev.cuh"
#include "intersect.cuh"
#include "launchers.h"
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {

constexpr int kBinBuckets = 6 * kFanCellsPerFace + 2;   // direction bins | lines without a usable direction | sentinel (always empty)
constexpr int kBucketNoDir = 6 * kFanCellsPerFace;
constexpr int kBinThreads = 1024;
constexpr int kBinBlock = 256;                          // lines per job of the loss kernel
#ifndef ART_PBIN_WARPS
#define ART_PBIN_WARPS 24
#endif
constexpr int kPBinWarps = ART_PBIN_WARPS;
constexpr int kPBinThreads = kPBinWarps * 32;

__device__ __forceinline__ int line_bucket(const float4 p, const f3 T)
{
    const f3 toT = sub3(T, mk3(p.x, p.y, p.z));                      // PM:76 operand, as in permeation_grid_kernel
    const int bin = fan_bin(-toT.x, -toT.y, -toT.z);
    return bin < 0 ? kBucketNoDir : bin;
}

__global__ void __launch_bounds__(kBinThreads, 1) perm_bin_count_kernel(const PermBinArgs a)
{
    extern __shared__ uint32_t sh[];
    const int sl = blockIdx.x, t = blockIdx.y;
    for (int b = threadIdx.x; b < kBinBuckets; b += kBinThreads) sh[b] = 0u;
    __syncthreads();
    const f3 T = mk3(a.targets[3 * t], a.targets[3 * t + 1], a.targets[3 * t + 2]);
    const int per = (a.nLocal + a.slices - 1) / a.slices;
    const int j1 = min(a.nLocal, (sl + 1) * per);
    for (int j = sl * per + threadIdx.x; j < j1; j += kBinThreads) {
        const float4 p = a.hitPts[j];
        if (p.w != 0.0f) atomicAdd(&sh[line_bucket(p, T)], 1u);
    }
    __syncthreads();
    uint32_t* out = a.cnt + (size_t)t * kBinBuckets * a.slices;
    for (int b = threadIdx.x; b < kBinBuckets; b += kBinThreads) out[(size_t)b * a.slices + sl] = sh[b];
}

// in place: counts -> exclusive offsets inside the target's region of pairRay, order (bin, slice)
__global__ void __launch_bounds__(kBinThreads, 1) perm_bin_scan_kernel(const PermBinArgs a)
{
    __shared__ uint32_t sWarp[32];
    const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* c = a.cnt + (size_t)t * kBinBuckets * a.slices;
    const int n = kBinBuckets * a.slices;
    const int per = (n + kBinThreads - 1) / kBinThreads;
    const int i0 = min(n, (int)threadIdx.x * per), i1 = min(n, i0 + per);
    uint32_t sum = 0;
    for (int i = i0; i < i1; i++) sum += c[i];
    uint32_t incl = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const uint32_t v = __shfl_up_sync(kFull, incl, s); if (lane >= s) incl += v; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    uint32_t w = sWarp[lane];
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const uint32_t v = __shfl_up_sync(kFull, w, s); if (lane >= s) w += v; }
    uint32_t run = (warp == 0 ? 0u : __shfl_sync(kFull, w, warp - 1)) + incl - sum;
    for (int i = i0; i < i1; i++) { const uint32_t v = c[i]; c[i] = run; run += v; }
    __syncthreads();
    for (int b = threadIdx.x; b < kBinBuckets; b += kBinThreads) a.binStart[(size_t)t * kBinBuckets + b] = c[(size_t)b * a.slices];
}

// per (target, block of kBinBlock sorted lines): the bucket that holds the block's first line (binary search in binStart)
__global__ void perm_bin_blocks_kernel(const PermBinArgs a, int nBlocks)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.nTargets * nBlocks) return;
    const int t = i / nBlocks, blk = i - t * nBlocks;
    const uint32_t* st = a.binStart + (size_t)t * kBinBuckets;
    const uint32_t line = (uint32_t)blk * kBinBlock;
    int lo = 0, hi = kBinBuckets - 1;                  // largest b with st[b] <= line  (st[0] = 0; st[last] = number of lines)
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (st[mid] <= line) lo = mid; else hi = mid - 1;
    }
    a.blockBin[i] = (uint32_t)lo;
}

__global__ void __launch_bounds__(kBinThreads, 1) perm_bin_scatter_kernel(const PermBinArgs a)
{
    extern __shared__ uint32_t sh[];
    const int sl = blockIdx.x, t = blockIdx.y;
    const uint32_t* off = a.cnt + (size_t)t * kBinBuckets * a.slices;
    for (int b = threadIdx.x; b < kBinBuckets; b += kBinThreads) sh[b] = off[(size_t)b * a.slices + sl];
    __syncthreads();
    const f3 T = mk3(a.targets[3 * t], a.targets[3 * t + 1], a.targets[3 * t + 2]);
    const int per = (a.nLocal + a.slices - 1) / a.slices;
    const int j1 = min(a.nLocal, (sl + 1) * per);
    uint32_t* dst = a.pairRay + (size_t)t * a.nLocal;
    for (int j = sl * per + threadIdx.x; j < j1; j += kBinThreads) {
        const float4 p = a.hitPts[j];
        if (p.w != 0.0f) dst[atomicAdd(&sh[line_bucket(p, T)], 1u)] = (uint32_t)j;
    }
}

__device__ __forceinline__ float clip_len_b(float tEnter, float tExit, float tIn, float tOut)
{
    return fmaxf(0.0f, fminf(tExit, tOut) - fmaxf(tEnter, tIn));
}

// body(id) for every entry of one list, in list order. The list is the same for all lanes of the warp: 32 ids per coalesced
// load, handed round by shuffle. Every lane must call this (warp-uniform off / n).
template <class Body>
__device__ __forceinline__ void perm_each_entry(const uint16_t* entries, uint32_t off, int n, int lane, Body body)
{
    for (int base = 0; base < n; base += 32) {
        const int mine = base + lane < n ? (int)__ldg(entries + off + base + lane) : 0;
        const int m = min(32, n - base);
        for (int j = 0; j < m; j++) body(__shfl_sync(kFull, mine, j));
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(kPBinThreads, 1) perm_loss_binned_kernel(const PermArgs a, const PermBinArgs ba, const FanDesc f)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31;

    unsigned char* p = smem;
    const unsigned char* geomBase = a.geom;
    const float* densS; const float* densA; const float* densO;
    if (SMEM) {
        stage_blob_to_smem(p, a.geom, a.L.bytes, &bar);
        geomBase = p;
        p += a.L.bytes;
        float* d = reinterpret_cast<float*>(p);       // true densities (owned colliders are left out of the fans, not zeroed)
        for (int i = threadIdx.x; i < a.L.nsPad; i += kPBinThreads) d[i] = a.at.sphAttr[i].z;
        for (int i = threadIdx.x; i < a.L.naPad; i += kPBinThreads) d[a.L.nsPad + i] = a.at.aabbAttr[i].z;
        for (int i = threadIdx.x; i < a.L.noPad; i += kPBinThreads) d[a.L.nsPad + a.L.naPad + i] = a.at.obbAttr[i].z;
        __syncthreads();
        densS = d; densA = d + a.L.nsPad; densO = d + a.L.nsPad + a.L.naPad;
    } else {
        densS = a.trueDens; densA = a.trueDens + a.L.nsPad; densO = a.trueDens + a.L.nsPad + a.L.naPad;
    }
    const GeomView gv = make_view(geomBase, a.L);
    const int nBlocks = (ba.nLocal + kBinBlock - 1) / kBinBlock;
    const int nJobs = a.nTargets * nBlocks;
    const float inf = pos_inf();

    for (;;) {
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(ba.jobQueue, 1u);
        job = __shfl_sync(kFull, job, 0);
        if (job >= nJobs) break;
        const int tgt = job / nBlocks, blk = job - tgt * nBlocks;
        const uint32_t* st = ba.binStart + (size_t)tgt * kBinBuckets;
        const uint32_t nLines = st[kBinBuckets - 1];                     // hitting rays (the sentinel bucket starts behind the last line)
        uint32_t cur = (uint32_t)blk * kBinBlock;
        if (cur >= nLines) continue;
        const uint32_t blockEnd = min(nLines, cur + (uint32_t)kBinBlock);
        int bin = (int)ba.blockBin[job];
        long long accInt = 0, accFrac = 0;
        while (cur < blockEnd) {
            // the bucket that holds line `cur`: the first b >= bin whose successor starts behind cur (32 candidates per step)
            for (;;) {
                const int cb = bin + lane;
                const bool behind = cb + 1 < kBinBuckets && st[cb + 1] > cur;
                const uint32_t m = __ballot_sync(kFull, behind);
                if (m) { bin += __ffs(m) - 1; break; }
                bin += 32;
            }
            const uint32_t start = cur, end = min(blockEnd, st[bin + 1]);
            cur = end;
            if (bin == kBucketNoDir) {
                // no usable direction (hit point == target): no list to test, loss 0 (as permeation_grid_kernel)
                if (lane == 0) {
                    const float v = subr(a.nTimesS, 0.0f);                             // PM:260
                    const float ip = truncf(v);
                    accInt += (long long)(end - start) * (long long)ip;
                    accFrac += (long long)(end - start) * (long long)(((double)v - (double)ip) * 68719476736.0);
                }
            } else {
                const f3 T = mk3(a.targets[3 * tgt], a.targets[3 * tgt + 1], a.targets[3 * tgt + 2]);
                // cell 0: near list, whole line; 1: bin towards the hit point, t in [0, tT]; 2: opposite bin, t > tT
                const int fanBase = tgt * kFanCells;
                const int face = bin / kFanCellsPerFace, rb = bin - face * kFanCellsPerFace;
                const uint2 h0 = __ldg(&f.cells[fanBase + 6 * kFanCellsPerFace]);
                const uint2 h1 = __ldg(&f.cells[fanBase + bin]);
                const uint2 h2 = __ldg(&f.cells[fanBase + (face ^ 1) * kFanCellsPerFace + (kFanCellsPerFace - 1 - rb)]);
                const int nS0 = h0.y & 1023, nA0 = (h0.y >> 10) & 2047, nO0 = h0.y >> 21;
                const int nS1 = h1.y & 1023, nA1 = (h1.y >> 10) & 2047, nO1 = h1.y >> 21;
                const int nS2 = h2.y & 1023, nA2 = (h2.y >> 10) & 2047, nO2 = h2.y >> 21;
                ART_CHECK(a.counters, h0.x + nS0 + nA0 + nO0 <= (unsigned)f.nEntries && h1.x + nS1 + nA1 + nO1 <= (unsigned)f.nEntries &&
                                      h2.x + nS2 + nA2 + nO2 <= (unsigned)f.nEntries);
                const uint32_t* rays = ba.pairRay + (size_t)tgt * ba.nLocal;
                for (uint32_t i0 = start; i0 < end; i0 += 32) {
                    const bool on = i0 + lane < end;
                    float4 rr = make_float4(0, 0, 0, 0);
                    if (on) rr = ba.hitPts[rays[i0 + lane]];
                    const f3 Pp = mk3(rr.x, rr.y, rr.z);
                    const f3 toT = sub3(T, Pp);
                    // (an idle lane of the last chunk computes on zeros and is ignored)
                    const f3 dir = normalize3(toT);                                    // PM:76
                    const f3 inv = mk3(rcpr(dir.x), rcpr(dir.y), rcpr(dir.z));         // PM:270
                    const float tT = sqrt_fast(fmaf(toT.z, toT.z, fmaf(toT.y, toT.y, toT.x * toT.x)));   // line parameter of the target
                    float loss = 0.0f;
                    // The three lists of a type are walked one after the other -- near list (whole line), bin towards the hit
                    // point (t in [0, tT]), opposite bin (t > tT) -- with the clip interval a constant of the list. All 32
                    // lines of the segment walk the SAME lists, so the warp fetches 32 entry ids with one coalesced load and
                    // hands them round by shuffle (perm_each_entry) instead of every lane loading every id.
                    {   // ---- AABBs, PM:265-288 in the reference's operation order
                        auto test = [&](int id, float tIn, float tOut) {
                            ART_CHECK(a.counters, id < a.L.na);
                            const float4 A = gv.aabbA[id];
                            const float2 B = gv.aabbB[id];
                            float tEnter, tExit;
                            slab<8>(subr(A.x, Pp.x), subr(A.y, Pp.y), subr(A.z, Pp.z), subr(A.w, Pp.x), subr(B.x, Pp.y), subr(B.y, Pp.z),
                                    inv.x, inv.y, inv.z, tEnter, tExit);
                            const float len = clip_len_b(tEnter, tExit, tIn, tOut);
                            if (len > 0.0f) loss = fmaf(len, densA[id], loss);
                        };
                        perm_each_entry(f.entries, h0.x + nS0, nA0, lane, [&](int id) { test(id, 0.0f, inf); });
                        perm_each_entry(f.entries, h1.x + nS1, nA1, lane, [&](int id) { test(id, 0.0f, tT); });
                        perm_each_entry(f.entries, h2.x + nS2, nA2, lane, [&](int id) { test(id, tT, inf); });
                    }
                    {   // ---- spheres, PM:303-328 (unit direction)
                        auto test = [&](int id, float tIn, float tOut) {
                            ART_CHECK(a.counters, id < a.L.ns);
                            const float4 sp = gv.sph[id];
                            const f3 oc = sub3(Pp, mk3(sp.x, sp.y, sp.z));
                            const float cc = subr(dot3(oc, oc), sp.w);
                            float b;
                            if (sphere_loss_fast_miss(oc, cc, dir, b)) return;         // disc < 0 (PM:311)
                            const float sq = sqrtr(subr(mulr(b, b), cc));
                            const float len = clip_len_b(subr(-b, sq), addr(-b, sq), tIn, tOut);
                            if (len > 0.0f) loss = fmaf(len, densS[id], loss);
                        };
                        perm_each_entry(f.entries, h0.x, nS0, lane, [&](int id) { test(id, 0.0f, inf); });
                        perm_each_entry(f.entries, h1.x, nS1, lane, [&](int id) { test(id, 0.0f, tT); });
                        perm_each_entry(f.entries, h2.x, nS2, lane, [&](int id) { test(id, tT, inf); });
                    }
                    {   // ---- OBBs, PM:294-300 (stored rotation as is), cheap arithmetic about the point of closest approach
                        auto test = [&](int id, float tIn, float tOut) {
                            ART_CHECK(a.counters, id < a.L.no);
                            const float4 c4 = gv.obbC[id];
                            const float2 h2o = gv.obbH[id];
                            const f3 pc = mk3(Pp.x - c4.x, Pp.y - c4.y, Pp.z - c4.z);
                            const float bq = fmaf(pc.z, dir.z, fmaf(pc.y, dir.y, pc.x * dir.x));
                            const float pp = fmaf(pc.z, pc.z, fmaf(pc.y, pc.y, pc.x * pc.x));
                            const float r2 = fmaf(h2o.y, h2o.y, fmaf(h2o.x, h2o.x, c4.w * c4.w));
                            if (pp - bq * bq > r2 * 1.001f + 1e-4f) return;            // the line passes the bounding sphere
                            const float4 q4 = gv.obbQ[id];
                            const f3 pn = mk3(fmaf(dir.x, -bq, pc.x), fmaf(dir.y, -bq, pc.y), fmaf(dir.z, -bq, pc.z));
                            const f3 lo = qrot_fast(q4, pn), ld = qrot_fast(q4, dir);
                            const float rx = rcp_fast(ld.x), ry = rcp_fast(ld.y), rz = rcp_fast(ld.z);
                            const float ax = (-c4.w - lo.x) * rx, bx = (c4.w - lo.x) * rx;
                            const float ay = (-h2o.x - lo.y) * ry, by = (h2o.x - lo.y) * ry;
                            const float az = (-h2o.y - lo.z) * rz, bz = (h2o.y - lo.z) * rz;
                            const float tEnter = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz)) - bq;
                            const float tExit = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) - bq;
                            const float len = clip_len_b(tEnter, tExit, tIn, tOut);
                            if (len > 0.0f) loss = fmaf(len, densO[id], loss);
                        };
                        perm_each_entry(f.entries, h0.x + nS0 + nA0, nO0, lane, [&](int id) { test(id, 0.0f, inf); });
                        perm_each_entry(f.entries, h1.x + nS1 + nA1, nO1, lane, [&](int id) { test(id, 0.0f, tT); });
                        perm_each_entry(f.entries, h2.x + nS2 + nA2, nO2, lane, [&](int id) { test(id, tT, inf); });
                    }
                    if (on) {
                        const float v = subr(a.nTimesS, loss);                         // PM:260
                        const float ip = truncf(v);
                        accInt += (long long)ip;
                        accFrac += (long long)(((double)v - (double)ip) * 68719476736.0);
                    }
                }
            }
        }   // next bin segment of the block
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            accInt += __shfl_xor_sync(kFull, accInt, s);
            accFrac += __shfl_xor_sync(kFull, accFrac, s);
        }
        if (lane == 0 && (accInt != 0 || accFrac != 0)) {
            atomicAdd(reinterpret_cast<unsigned long long*>(&a.permSumInt[tgt]), (unsigned long long)accInt);
            atomicAdd(reinterpret_cast<unsigned long long*>(&a.permSumFrac[tgt]), (unsigned long long)accFrac);
        }
    }
}

int perm_binned_slices(int nLocal, int nTargets, int numSms)
{
    // enough (slice, target) CTAs to fill the GPU twice, at least 16,384 rays per slice
    int s = (2 * numSms + nTargets - 1) / nTargets;
    const int maxS = (nLocal + 16383) / 16384;
    if (s > maxS) s = maxS;
    if (s > 32) s = 32;
    return s < 1 ? 1 : s;
}
size_t perm_binned_cnt_bytes(int nTargets, int slices) { return (size_t)nTargets * kBinBuckets * slices * sizeof(uint32_t); }
size_t perm_binned_start_bytes(int nTargets) { return (size_t)nTargets * kBinBuckets * sizeof(uint32_t); }
size_t perm_binned_block_bytes(int nLocal, int nTargets) { return (size_t)nTargets * ((nLocal + kBinBlock - 1) / kBinBlock) * sizeof(uint32_t); }
size_t perm_binned_smem_bytes(const GeomLayout& L, bool geomInSmem)
{
    return geomInSmem ? L.bytes + ((size_t)L.nsPad + L.naPad + L.noPad) * sizeof(float) : 0;
}

cudaError_t launch_permeation_binned(const PermArgs& a, const PermBinArgs& ba, const FanDesc& fans, int numCtas, bool geomInSmem, cudaStream_t stream)
{
    const size_t shBins = (size_t)kBinBuckets * sizeof(uint32_t);
    perm_bin_count_kernel<<<dim3(ba.slices, a.nTargets), kBinThreads, shBins, stream>>>(ba);
    perm_bin_scan_kernel<<<a.nTargets, kBinThreads, 0, stream>>>(ba);
    perm_bin_scatter_kernel<<<dim3(ba.slices, a.nTargets), kBinThreads, shBins, stream>>>(ba);
    const int nBlocks = (ba.nLocal + kBinBlock - 1) / kBinBlock;
    perm_bin_blocks_kernel<<<(a.nTargets * nBlocks + 255) / 256, 256, 0, stream>>>(ba, nBlocks);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem = perm_binned_smem_bytes(a.L, geomInSmem);
    void (*k)(const PermArgs, const PermBinArgs, const FanDesc) = geomInSmem ? perm_loss_binned_kernel<true> : perm_loss_binned_kernel<false>;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<numCtas, kPBinThreads, smem, stream>>>(a, ba, fans);
    return cudaGetLastError();
}

}  // namespace art
