// k5_grid_build.cu -- the cell lists of the uniform collider grid (scene_dev.cuh: GridDesc), filled on the device.
//
// The host computes only what is O(colliders) and decides the launch shapes -- collider bounds, grid dimensions, margins
// (grid_host.h: grid_params) -- and uploads the conservative boxes; which collider is listed in which cell is worked out
// here, so a scene that changes every frame (AudioColliderManager.UpdateJobBatch re-bakes the dynamic colliders every frame,
// Audio/AudioColliderManager.cs:115-122) costs the host no cell walk and no stream synchronisation:
//   grid_count_kernel   one warp per collider, lanes over the cells its box overlaps: counts per (cell, type)
//   grid_scan_kernel    one CTA: exclusive scan of the counts -> cell headers (first entry, nS | nA << 10 | nO << 21)
//   grid_fill_kernel    one warp per collider: collider index appended to each of its cells' lists
//   grid_sort_kernel    one thread per cell: every list ascending (the fill order is scheduling dependent)
// Every collider is entered into all cells its conservative box overlaps -- the same rule as the host version
// (build_grid, kept for art_grid_build_host); a cell list longer than the header format allows or an entry buffer that is
// too small raise a flag, the frame is then re-run on the brute-force kernels and the buffer grown (audiort_api.cu).
#include "device_util.cuh"
#include "launchers.h"
#include "scene_dev.cuh"

namespace art {

__device__ __forceinline__ void grid_cell_range(const GridBuildArgs& a, int g, int i0[3], int i1[3])
{
    const float4 lo = a.boxLo[g], hi = a.boxHi[g];
    const float l[3] = { lo.x, lo.y, lo.z }, h[3] = { hi.x, hi.y, hi.z };
    const float g0[3] = { a.g0x, a.g0y, a.g0z }, cs[3] = { a.csx, a.csy, a.csz };
    const int dim[3] = { a.nx, a.ny, a.nz };
#pragma unroll
    for (int k = 0; k < 3; k++) {
        i0[k] = max(0, min(dim[k] - 1, (int)floorf(__fdiv_rn(__fsub_rn(l[k], g0[k]), cs[k]))));
        i1[k] = max(0, min(dim[k] - 1, (int)floorf(__fdiv_rn(__fsub_rn(h[k], g0[k]), cs[k]))));
    }
}

template <bool FILL>
__global__ void __launch_bounds__(256) grid_visit_kernel(const GridBuildArgs a)
{
    const int lane = threadIdx.x & 31;
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    const int nc = a.ns + a.na + a.no;
    if (g >= nc) return;
    if (FILL && a.ctl[1] != 0u) return;                  // overflow: the lists are not used
    const int type = g < a.ns ? 0 : (g < a.ns + a.na ? 1 : 2);
    const int id = type == 0 ? g : (type == 1 ? g - a.ns : g - a.ns - a.na);
    int i0[3], i1[3];
    grid_cell_range(a, g, i0, i1);
    const int wx = i1[0] - i0[0] + 1, wy = i1[1] - i0[1] + 1, wz = i1[2] - i0[2] + 1;
    const int n = wx * wy * wz;
    for (int c = lane; c < n; c += 32) {
        const int x = i0[0] + c % wx, y = i0[1] + (c / wx) % wy, z = i0[2] + c / (wx * wy);
        const size_t cell = ((size_t)z * a.ny + y) * a.nx + x;
        const unsigned int pos = atomicAdd(&a.cnt[cell * 3 + type], 1u);
        if (FILL) {
            if (pos < a.capacity) a.entries[pos] = (uint16_t)id;
            else atomicExch(&a.ctl[1], 1u);
        }
    }
}

// counts -> cell headers; cnt[cell * 3 + type] becomes the write cursor of that list
__global__ void __launch_bounds__(1024, 1) grid_scan_kernel(const GridBuildArgs a)
{
    __shared__ unsigned int sWarp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nCells = a.nx * a.ny * a.nz;
    const int per = (nCells + 1023) / 1024;
    const int c0 = min(nCells, (int)threadIdx.x * per), c1 = min(nCells, c0 + per);
    unsigned int sum = 0;
    bool bad = false;
    for (int c = c0; c < c1; c++) {
        const unsigned int s = a.cnt[(size_t)c * 3], aa = a.cnt[(size_t)c * 3 + 1], o = a.cnt[(size_t)c * 3 + 2];
        bad = bad || s > (unsigned)kGridMaxS || aa > (unsigned)kGridMaxA || o > (unsigned)kGridMaxO;
        sum += s + aa + o;
    }
    unsigned int incl = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const unsigned int v = __shfl_up_sync(kFull, incl, s); if (lane >= s) incl += v; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    unsigned int w = sWarp[lane];
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const unsigned int v = __shfl_up_sync(kFull, w, s); if (lane >= s) w += v; }
    unsigned int off = (warp == 0 ? 0u : __shfl_sync(kFull, w, warp - 1)) + incl - sum;
    const unsigned int total = __shfl_sync(kFull, w, 31);
    // a cell list too long for the header format, or more entries than the buffer holds: every cell is left EMPTY (the
    // kernels of this frame then walk an empty grid -- harmless, their results are discarded by the re-run)
    const bool overflow = __syncthreads_or(bad) || total > a.capacity;
    for (int c = c0; c < c1; c++) {
        const unsigned int s = a.cnt[(size_t)c * 3], aa = a.cnt[(size_t)c * 3 + 1], o = a.cnt[(size_t)c * 3 + 2];
        a.cells[c] = overflow ? make_uint2(0u, 0u) : make_uint2(off, s | (aa << 10) | (o << 21));
        a.cnt[(size_t)c * 3] = off; a.cnt[(size_t)c * 3 + 1] = off + s; a.cnt[(size_t)c * 3 + 2] = off + s + aa;
        off += s + aa + o;
    }
    if (threadIdx.x == 0) {
        if (overflow) atomicExch(&a.ctl[1], 1u);
        a.ctl[0] = total;
    }
}

__global__ void __launch_bounds__(256) grid_sort_kernel(const GridBuildArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.nx * a.ny * a.nz || a.ctl[1] != 0u) return;
    const uint2 h = a.cells[c];
    const int n[3] = { (int)(h.y & 1023u), (int)((h.y >> 10) & 2047u), (int)(h.y >> 21) };
    uint16_t* e = a.entries + h.x;
    for (int t = 0; t < 3; t++) {
        for (int i = 1; i < n[t]; i++) {                 // insertion sort: the lists hold a handful of entries
            const uint16_t v = e[i];
            int j = i - 1;
            while (j >= 0 && e[j] > v) { e[j + 1] = e[j]; j--; }
            e[j + 1] = v;
        }
        e += n[t];
    }
}

cudaError_t launch_grid_build(const GridBuildArgs& a, cudaStream_t stream)
{
    const int nc = a.ns + a.na + a.no;
    const int nCells = a.nx * a.ny * a.nz;
    cudaError_t e = cudaMemsetAsync(a.cnt, 0, (size_t)nCells * 3 * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(a.ctl, 0, 2 * sizeof(unsigned int), stream)) != cudaSuccess) return e;
    const int blocks = (nc * 32 + 255) / 256;
    grid_visit_kernel<false><<<blocks, 256, 0, stream>>>(a);
    grid_scan_kernel<<<1, 1024, 0, stream>>>(a);
    grid_visit_kernel<true><<<blocks, 256, 0, stream>>>(a);
    grid_sort_kernel<<<(nCells + 255) / 256, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace art
