// k3_reduce.cu -- K3: the O(N*H) part of ProcessAudioDataJob.Execute
// (Assets/C# Scripts/Jobs/ProcessAudioDataJob.cs:38-48): total of the echo distances and the
// count of zero entries.
//
// Two variants:
//  * echo_stats_kernel (default): EXACT. A half is m * 2^e, so value * 2^24 is an integer < 2^40;
//    the kernel sums those integers (split hi/lo so nothing can overflow) with integer atomics.
//    Integer addition is associative, so the result is independent of the order in which warps
//    finish: deterministic and free of the FP32 accumulation drift the reference's running sum
//    has. HBM-bound: 2 B per entry.
//  * reverb_seq_kernel (ART_FRAME_REVERB_SEQ_FP32): the reference's own rounding -- one FP32
//    accumulator, entries added in index order (PA:40-48), including the float zero counter that
//    saturates at 2^24. One thread runs the dependent FADD chain while the rest of the CTA
//    stages and converts tiles through shared memory.
#include "scene_dev.cuh"
#include "um_math.cuh"

namespace art {


__global__ void __launch_bounds__(256) echo_stats_kernel(const uint16_t* __restrict__ echo, size_t n, EchoStats* out)
{
    long long lo = 0, hi = 0;
    unsigned int zeros = 0, pinf = 0, ninf = 0, nnan = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    // 8 halves per 16-byte load when aligned
    const size_t n8 = n / 8;
    const uint4* e8 = reinterpret_cast<const uint4*>(echo);
    auto acc = [&](uint32_t h) {
        const uint32_t mag = h & 0x7FFFu;
        if (mag == 0) { zeros++; return; }
        const uint32_t e = mag >> 10, m = mag & 1023u;
        if (e == 31) { if (m) nnan++; else if (h & 0x8000u) ninf++; else pinf++; return; }
        const unsigned long long fx = e == 0 ? (unsigned long long)m : ((unsigned long long)(1024u + m) << (e - 1));
        const long long l = (long long)(fx & 0xFFFFFull), hh = (long long)(fx >> 20);
        if (h & 0x8000u) { lo -= l; hi -= hh; } else { lo += l; hi += hh; }
    };
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 v = e8[i];
        acc(v.x & 0xFFFFu); acc(v.x >> 16); acc(v.y & 0xFFFFu); acc(v.y >> 16);
        acc(v.z & 0xFFFFu); acc(v.z >> 16); acc(v.w & 0xFFFFu); acc(v.w >> 16);
    }
    for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc(echo[i]);

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo += __shfl_xor_sync(0xFFFFFFFFu, lo, o);
        hi += __shfl_xor_sync(0xFFFFFFFFu, hi, o);
        zeros += __shfl_xor_sync(0xFFFFFFFFu, zeros, o);
        pinf += __shfl_xor_sync(0xFFFFFFFFu, pinf, o);
        ninf += __shfl_xor_sync(0xFFFFFFFFu, ninf, o);
        nnan += __shfl_xor_sync(0xFFFFFFFFu, nnan, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(&out->fixedLo), (unsigned long long)lo);
        atomicAdd(reinterpret_cast<unsigned long long*>(&out->fixedHi), (unsigned long long)hi);
        atomicAdd(&out->zeros, (unsigned long long)zeros);
        if (pinf) atomicAdd(&out->posInf, (unsigned long long)pinf);
        if (ninf) atomicAdd(&out->negInf, (unsigned long long)ninf);
        if (nnan) atomicAdd(&out->nan, (unsigned long long)nnan);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out->entries = (unsigned long long)n;
}

constexpr int kSeqTile = 4096;

__global__ void __launch_bounds__(256, 1) reverb_seq_kernel(const uint16_t* __restrict__ echo, size_t n, EchoStats* out)
{
    __shared__ float buf[2][kSeqTile];
    const size_t nTiles = (n + kSeqTile - 1) / kSeqTile;
    float total = 0.0f, zeros = 0.0f;
    auto load_tile = [&](size_t t, int b) {
        const size_t base = t * kSeqTile;
        for (int i = threadIdx.x; i < kSeqTile; i += blockDim.x) {
            const size_t g = base + i;
            buf[b][i] = g < n ? um_f16tof32(echo[g]) : 0.0f;
        }
    };
    if (nTiles > 0) load_tile(0, 0);
    __syncthreads();
    for (size_t t = 0; t < nTiles; t++) {
        const int b = (int)(t & 1);
        if (threadIdx.x == 0) {
            const size_t cnt = (t + 1 == nTiles) ? n - t * kSeqTile : (size_t)kSeqTile;
            const float* s = buf[b];
#pragma unroll 8
            for (size_t i = 0; i < cnt; i++) {
                const float e = s[i];
                zeros = addr(zeros, e == 0.0f ? 1.0f : 0.0f);   // PA:42-45 (float counter)
                total = addr(total, e == 0.0f ? 0.0f : e);      // PA:47 (zero entries are skipped)
            }
        } else if (t + 1 < nTiles) {
            // 255 threads stage the next tile while thread 0 runs the dependent chain
            const size_t base = (t + 1) * kSeqTile;
            for (int i = threadIdx.x - 1; i < kSeqTile; i += blockDim.x - 1) {
                const size_t g = base + i;
                buf[b ^ 1][i] = g < n ? um_f16tof32(echo[g]) : 0.0f;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out->seqTotal = total; out->seqZeros = zeros; out->seqValid = 1u; }
}

cudaError_t launch_echo_stats(const uint16_t* echo, size_t n, EchoStats* out, bool sequential, int numSms, cudaStream_t stream)
{
    if (n > 0) {
        size_t blocks = (n / 8 + 255) / 256;
        if (blocks < 1) blocks = 1;
        if (blocks > (size_t)numSms * 8) blocks = (size_t)numSms * 8;
        echo_stats_kernel<<<(unsigned)blocks, 256, 0, stream>>>(echo, n, out);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (sequential) {
        reverb_seq_kernel<<<1, 256, 0, stream>>>(echo, n, out);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

}  // namespace art
