// device_util.cuh -- mbarrier / bulk-copy (TMA) helpers and warp primitives shared by K1/K2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace art {

constexpr uint32_t kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dstSmem, const void* srcGmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dstSmem)), "l"(srcGmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Stage `bytes` (multiple of 16) of the geometry blob into shared memory; all threads return once
// the data has landed. One elected thread issues <=32 KiB bulk copies against one mbarrier.
__device__ __forceinline__ void stage_blob_to_smem(unsigned char* dst, const unsigned char* src, uint32_t bytes,
                                                   uint64_t* bar)
{
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, bytes);
        constexpr uint32_t kChunk = 32768;
        for (uint32_t off = 0; off < bytes; off += kChunk) {
            uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
            bulk_g2s(dst + off, src + off, n, bar);
        }
    }
    mbar_wait(bar, 0);
}

__device__ __forceinline__ float warp_min_f(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
    return v;
}

}  // namespace art
