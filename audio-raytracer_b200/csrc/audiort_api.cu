// audiort_api.cu -- C ABI of libaudiort_cuda (include/audiort.h): context, scene/ray upload,
// frame scheduling on one CUDA stream, completion, partial-result merge and the
// ProcessAudioDataJob finalisation. Stands in for AudioRayTracer.OnUpdate's scheduling block
// (Assets/C# Scripts/Audio/AudioRayTracer.cs:95-97, 161-237) -- there is no CPU compute path here:
// every intersection test, sum and count is produced by the kernels in k0..k3.
#include "../../include/audiort.h"

#include "grid_host.h"
#include "launchers.h"
#include "scene_dev.cuh"
#include "um_math.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <new>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace art {

// ---- FIB: FibonacciDirectionsJobParallel.Execute (Jobs/FibonacciDirectionsJobParallel.cs:25-34) --
__global__ void fibonacci_kernel(uint16_t* dirs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float phi = mulr(3.14159274f, subr(3.0f, sqrtr(5.0f)));                  // FIB:26
    const float y = subr(1.0f, mulr(divr((float)i, (float)(n - 1)), 2.0f));        // FIB:27
    const float radius = sqrtr(subr(1.0f, mulr(y, y)));                            // FIB:28
    const float theta = mulr(phi, (float)i);                                       // FIB:29
    const float x = mulr((float)cos((double)theta), radius);                       // FIB:31 math.cos(float) = (float)Math.Cos(double)
    const float z = mulr((float)sin((double)theta), radius);                       // FIB:32
    dirs[3 * (size_t)i] = um_f32tof16(x);
    dirs[3 * (size_t)i + 1] = um_f32tof16(y);
    dirs[3 * (size_t)i + 2] = um_f32tof16(z);
}
cudaError_t launch_fibonacci(uint16_t* dirs, int n, cudaStream_t stream)
{
    fibonacci_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dirs, n);
    return cudaGetLastError();
}

// ---- FP32 issue-rate microbenchmarks (roofline denominators) -------------------------
template <int KIND>
__global__ void __launch_bounds__(1024, 2) microbench_kernel(float* sink, int iters, float seed)
{
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = seed + (float)(threadIdx.x + k);
    const float m = 1.0000001f, c = 0.9999999f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (KIND == 0) { v[k] = __fmul_rn(v[k], m); v[k] = __fadd_rn(v[k], c); }          // un-fused FMUL + FADD
                else if (KIND == 1) { v[k] = fminf(v[k], v[(k + 3) & 7]); v[k] = fmaxf(v[k], v[(k + 5) & 7]); } // FMNMX pairs
                else { v[k] = __fmaf_rn(v[k], m, c); v[k] = __fmaf_rn(v[k], c, m); }             // FFMA
            }
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += v[k];
    if (s == 123.456f) sink[0] = s;
}
cudaError_t launch_microbench(int kind, int numSms, float* sink, long long* laneOps, cudaStream_t stream)
{
    const int iters = 4096, blocks = numSms * 2, threads = 1024;
    if (kind == 0) microbench_kernel<0><<<blocks, threads, 0, stream>>>(sink, iters, 1.0f);
    else if (kind == 1) microbench_kernel<1><<<blocks, threads, 0, stream>>>(sink, iters, 1.0f);
    else microbench_kernel<2><<<blocks, threads, 0, stream>>>(sink, iters, 1.0f);
    *laneOps = (long long)blocks * threads * iters * 4 * 8 * 2;
    return cudaGetLastError();
}

}  // namespace art

using namespace art;

// =================================================================================================
// Host side
// =================================================================================================
namespace {

constexpr uint32_t kBlobMagic = 0x41525442u;   // "ARTB"

struct BlobHeader {
    uint32_t magic;
    int32_t nTargets;
    int32_t batchCount;
    int32_t shards;              // number of contexts merged into this blob
    EchoStats echo;
    unsigned long long counters[C_COUNT];
};
// blob = header | int32 lastHitRay[T] (+pad to 8) | uint32 muffleCounts[T*Na] (+pad) | float permLast[T*Na] (+pad)
//        | int64 permSumInt[Na] | int64 permSumFrac[Na]
struct BlobLayout {
    size_t offLastHit, offMuffle, offPermLast, offSumInt, offSumFrac, bytes;
};
inline size_t align8(size_t x) { return (x + 7) & ~(size_t)7; }
BlobLayout blob_layout(int Na, int T)
{
    BlobLayout b;
    size_t o = align8(sizeof(BlobHeader));
    b.offLastHit = o; o = align8(o + sizeof(int32_t) * (size_t)T);
    b.offMuffle = o; o = align8(o + sizeof(uint32_t) * (size_t)T * Na);
    b.offPermLast = o; o = align8(o + sizeof(float) * (size_t)T * Na);
    b.offSumInt = o; o += sizeof(long long) * (size_t)Na;
    b.offSumFrac = o; o += sizeof(long long) * (size_t)Na;
    b.bytes = o;
    return b;
}

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ---- NCCL, loaded on first use (art_comm_*): single-GPU hosts need no libnccl ------------------------------------------
struct NcclUniqueId { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool tried = false;
    bool ok() const { return lib && GetUniqueId && CommInitRank && AllGather && CommDestroy && GetErrorString; }
};
NcclApi& nccl_api()
{
    static NcclApi api;
    if (!api.tried) {
        api.tried = true;
        const char* names[] = { getenv("ART_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(dlsym(api.lib, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<int (*)(void**, int, NcclUniqueId, int)>(dlsym(api.lib, "ncclCommInitRank"));
            api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, void*, cudaStream_t)>(dlsym(api.lib, "ncclAllGather"));
            api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
            api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
        }
    }
    return api;
}
constexpr int kNcclChar = 0;   // ncclInt8 / ncclChar

inline float h2f(uint16_t h)   // IEEE binary16 -> binary32 (== math.f16tof32), host copy for scene prep
{
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, mag = h & 0x7FFFu, bits;
    if (mag >= 0x7C00u) bits = sign | 0x7F800000u | ((mag & 0x3FFu) << 13);
    else if (mag >= 0x0400u) bits = sign | ((mag << 13) + ((127u - 15u) << 23));
    else if (mag == 0) bits = sign;
    else { float f = (float)mag * 5.9604644775390625e-08f; memcpy(&bits, &f, 4); bits |= sign; }
    float out; memcpy(&out, &bits, 4); return out;
}

thread_local std::string g_createError;

}  // namespace

struct ArtCtx {
    int device = 0;
    int numSms = 0;
    int maxSmemOptin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;             // per-ray outputs go back to the host while the permeation job runs
    cudaStream_t stream2 = nullptr;                // small frames: the permeation job runs beside the trace job (ART:213 does the same)
    cudaEvent_t ev[7] = {};
    cudaEvent_t evTraceDone = nullptr, evCopyDone = nullptr, evReady = nullptr, evP0 = nullptr, evP1 = nullptr;
    bool frameOverlap = false;
    std::string err;
    bool poisoned = false;

    // scene (host staging = caller's structs, device raw + packed)
    std::vector<uint16_t> hostS, hostA, hostO;     // raw words, kept for host-side prep
    PinBuf pinScene;
    DevBuf rawScene, geom, attrs, owners, perm;    // perm: dens arrays + owned list
    HostGrid grid;                                 // uniform grid over the scene (grid_host.h)
    DevBuf gridCells, gridEntries, gridRangeO, gridScratch, rotateLog, gridCnt, gridCtl;
    PinBuf pinGrid, pinGridCtl;
    size_t gridEntriesPerCollider = 64;            // entry capacity of the device-built grid (grows after an overflow)
    bool frameGridBuilt = false;                   // the frame in flight uses cell lists built for it (overflow flag to check)
    bool rerunNoGrid = false;                      // second pass of a frame whose grid build overflowed: brute-force kernels
    int dirtyLo[3] = { 0, 0, 0 }, dirtyHi[3] = { 0, 0, 0 };   // per type (S, A, O): struct range changed since the last upload
    bool sceneLayoutChanged = true;                // counts changed: everything is uploaded
    DevBuf hitRecs, queryScratch;                  // bounce-only trace job: hit records + survivor lists of query_fan_kernel
    DevBuf permHitPts, permBinCnt, permPairs;       // binned loss lines (k2_permeation_binned.cu)
    DevBuf fanBoxes, fanCells, fanEntries, fanCtl, fanOrder, fanScratch;  // target fans (fan_dev.cuh): collider bounds (per scene), lists (per frame)
    PinBuf pinFanCtl;
    bool fansDisabled = false;                     // ART_DISABLE_FANS=1
    float frameCoverLogS = 0.0f, frameCoverLogK = 0.0f;   // covering-depth code scale of the last frame's fans (art_debug_get_fan_cover)
    size_t fanEntriesPerPair = 64;                 // entry capacity = fans * colliders * this (grows after an overflow); ART_FAN_ENTRIES_PER_PAIR
    bool frameFans = false;                        // the frame in flight uses the fans
    bool frameCopiedEarly = false;                 // per-ray outputs of the frame in flight travel on copyStream behind the trace job
    bool rerunning = false;                        // art_complete is re-running an overflowed frame without fans
    bool gridDisabled = false;                     // ART_DISABLE_GRID=1
    bool gridBuilt = false;                        // grid (or the decision that there is none) is current for the scene
    float gridCellScale = 1.1f;                    // ART_GRID_CELL_SCALE
    int gridMinRays = -1;                          // ART_GRID_MIN_RAYS: smaller batches use the brute-force kernels (-1: heuristic)
    uint32_t frameGridUsed = 0;
    float lastHitFill = 0.0f;        // previous trace frame: hits / (rays x MaxHitsPerRay); ~1 = rays live all their bounces (group rotation)
    GeomLayout L{};
    bool haveScene = false, sceneDirty = false;
    int permPreparedForTargets = -1;
    int nOwned = 0;

    // rays
    int nGlobal = 0;
    PinBuf pinRays; DevBuf dirs;
    bool haveRays = false, raysDirty = false;
    int shardIndex = 0, shardCount = 1, chunkRays = 0;
    bool chunkAuto = false;                        // chunk size derived from the batch size (library-owned shard maps)

    // frame
    DevBuf targets, ownedCount, outAll, firstHit, partials, queue;   // outAll: echo | hit ids | hit points | hit counts, one memset, one copy
    PinBuf pinTargets, pinOwnedCount, pinPartials, pinAll, pinPerm;
    size_t offEcho = 0, offHitIds = 0, offHitPts = 0, offHitCnt = 0, outBytes = 0;
    ArtParams params{};
    ArtOutputs userOut{};
    bool haveUserOut = false;
    bool inFlight = false;
    bool frameDone = false;
    ArtHandle handle = 0;
    ShardMap map{};
    int frameH = 0, frameNa = 0, frameT = 0;
    uint32_t frameFlags = 0, frameJobs = 0;
    uint32_t kernelLaunches = 0;
    ArtCounters counters{};
    std::vector<unsigned char> lastBlob;
    std::vector<int> frameOwnedCount;
    cudaEvent_t evFan = nullptr, evBounce = nullptr, evX0 = nullptr, evX1 = nullptr;   // sub-times of the trace job, blob exchange
    cudaEvent_t evF0 = nullptr;                    // start of a fan build that runs on stream2 beside the bounce tracer
    bool frameFanBeside = false;
    bool frameIsRerun = false;                     // the frame in flight is the second pass of a frame whose fan build overflowed
    bool frameSplit = false;                       // the frame in flight ran the bounce-only tracer + query kernel
    bool rayHostValid = false;                     // pinRays holds the whole batch (art_set_rays); false for device-generated rays
    bool rayHostLocal = false;                     // pinRays holds only this context's shard, packed in local order
    int rayStamp[3] = { 0, 1, 0 };                 //   ... for this (shardIndex, shardCount, chunk)
    bool dirsLocal = false;                        // the device direction array holds only this context's shard, in local order

    // multi-device context (ArtConfig.nDevices > 1): this object owns no CUDA state itself, only one child per device
    std::vector<ArtCtx*> children;
    bool scatterGlobal = false;                    // child: per-ray outputs go to the caller's arrays at GLOBAL ray positions
    std::vector<float> parentTargets;              // parent: library-owned copy of the frame's target positions

    // multi-process communicator (art_comm_init)
    void* comm = nullptr;
    int commRank = 0, commWorld = 1;
    DevBuf gathered; PinBuf pinGathered;
    bool frameComm = false;                        // the frame in flight all-gathers the ranks' partial blobs on the device
};

namespace {

// pinned staging -> caller arrays; large arrays are copied by a few threads (one memcpy stream saturates ~10 GB/s)
void big_memcpy(void* dst, const void* src, size_t bytes)
{
    constexpr size_t kMinPerThread = (size_t)4 << 20;
    unsigned n = (unsigned)(bytes / kMinPerThread);
    const unsigned hw = std::thread::hardware_concurrency();
    if (n > 6) n = 6;
    if (hw && n > hw) n = hw;
    if (n <= 1) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t per = ((bytes / n) + 63) & ~(size_t)63;
    for (unsigned i = 0; i < n; i++) {
        const size_t off = (size_t)i * per;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len); });
    }
    for (auto& t : th) t.join();
}

int32_t fail(ArtCtx* c, int32_t code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_createError = buf;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->poisoned = (cudaGetLastError(), cudaPeekAtLastError() != cudaSuccess) || ctx->poisoned; \
            return fail(ctx, ART_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                          \
    } while (0)

int local_ray_count(int nGlobal, int shardIndex, int shardCount, int chunk)
{
    if (shardCount <= 1) return nGlobal;
    long long n = 0;
    const long long nChunks = ((long long)nGlobal + chunk - 1) / chunk;
    for (long long c = shardIndex; c < nChunks; c += shardCount) {
        long long lo = c * chunk, hi = lo + chunk;
        if (hi > nGlobal) hi = nGlobal;
        n += hi - lo;
    }
    return (int)n;
}

int effective_chunk(const ArtCtx* c)
{
    if (c->shardCount <= 1) return c->nGlobal > 0 ? c->nGlobal : 1;
    if (c->chunkAuto) {
        // Library-owned shard maps: interleaved chunks of N / (8 * shards) rays, at most 16,384. Long chunks keep a shard's
        // first-hit points (one latitude band of the Fibonacci sphere per chunk) together, so the permeation job's (source,
        // direction bin) groups stay long enough to fill warps when a shard has few rays (B200, C3 / 8: chunks of 256 rays
        // 5.47 ms per rank, 16,384 rays 5.18 ms, one contiguous slice 4.7 .. 5.5 ms depending on the rank); with many rays per
        // shard the groups are long anyway and more, shorter chunks balance the bands' different geometry better (C3 / 2 with
        // 8 chunks of 65,536 rays: one rank waits 0.3 ms of 16.7 for the other).
        long long ch = ((long long)c->nGlobal + 8LL * c->shardCount - 1) / (8LL * c->shardCount);
        ch = (ch + 255) / 256 * 256;
        if (ch > 16384) ch = 16384;
        return (int)(ch < 256 ? 256 : ch);
    }
    if (c->chunkRays > 0) return c->chunkRays;
    return (c->nGlobal + c->shardCount - 1) / c->shardCount;   // contiguous slices
}

// ART:161: (int)math.max(1, math.ceil((float)rayCount / ToUseThreadCount))
int batch_size(int rayCount, int T)
{
    float q = std::ceil((float)rayCount / (float)T);
    if (!(q > 1.0f)) q = 1.0f;
    return (int)q;
}

// math.saturate(x) = max(0, min(1, x)) with Unity's NaN rule
inline float um_min_h(float x, float y) { return (std::isnan(y) || x < y) ? x : y; }
inline float um_max_h(float x, float y) { return (std::isnan(y) || x > y) ? x : y; }
inline float saturate_h(float x) { return um_max_h(0.0f, um_min_h(1.0f, x)); }

int32_t finalize_blob(const unsigned char* blob, size_t bytes, const ArtParams* p, int32_t rayCount, const ArtOutputs* out,
                      std::string* err)
{
    if (bytes < sizeof(BlobHeader)) { if (err) *err = "blob too small"; return ART_E_ARG; }
    BlobHeader h; memcpy(&h, blob, sizeof h);
    if (h.magic != kBlobMagic || h.nTargets != p->totalAudioTargets || h.batchCount != p->batchCount) {
        if (err) *err = "blob does not match params"; return ART_E_ARG;
    }
    const int Na = h.nTargets, T = h.batchCount, N = rayCount, H = p->maxHitsPerRay;
    const BlobLayout bl = blob_layout(Na, T);
    if (bytes < bl.bytes) { if (err) *err = "blob truncated"; return ART_E_ARG; }
    const int32_t* lastHit = reinterpret_cast<const int32_t*>(blob + bl.offLastHit);
    const uint32_t* mcnt = reinterpret_cast<const uint32_t*>(blob + bl.offMuffle);
    const float* permLast = reinterpret_cast<const float*>(blob + bl.offPermLast);
    const long long* sumInt = reinterpret_cast<const long long*>(blob + bl.offSumInt);
    const long long* sumFrac = reinterpret_cast<const long long*>(blob + bl.offSumFrac);

    const int b = batch_size(N, T);
    const int numBatches = (N + b - 1) / b;

    // ---- MuffleRayHits table: batch k writes slot row batchId = k*b*T / N (RT:63-64, int32 wrap),
    //      zeroing it first (RT:82-85); batches are applied in ascending order (canonical serial order).
    std::vector<uint16_t> table((size_t)T * Na, 0);
    std::vector<float> permTable((size_t)T * Na, 0.0f);
    for (int k = 0; k < numBatches; k++) {
        const int32_t start = k * b;
        const int32_t row = (int32_t)((uint32_t)start * (uint32_t)T) / N;
        if (row < 0 || row >= T) { if (err) *err = "batch slot out of range (reference would throw)"; return ART_E_ARG; }
        for (int a = 0; a < Na; a++) table[(size_t)row * Na + a] = (uint16_t)mcnt[(size_t)k * Na + a];   // ushort wrap (Q8)
        // ---- PermeationPowerRemains: PM:36-37 batchCount = Length / totalRays / Na (quirk Q6)
        const int32_t totalRays = (N - start < b) ? N - start : b;
        const int32_t pbc = (T * Na) / totalRays / Na;
        const int32_t prow = (int32_t)((uint32_t)start * (uint32_t)pbc) / N;
        if (prow < 0 || prow >= T) { if (err) *err = "permeation slot out of range"; return ART_E_ARG; }
        for (int a = 0; a < Na; a++)
            permTable[(size_t)prow * Na + a] = lastHit[k] >= 0 ? permLast[(size_t)k * Na + a] : 0.0f;   // PM:43-46, 85
    }
    if (out->muffleRayHits) memcpy(out->muffleRayHits, table.data(), table.size() * 2);
    if (out->permeationPowerRemains) memcpy(out->permeationPowerRemains, permTable.data(), permTable.size() * 4);
    if (out->muffleTotals)
        for (int a = 0; a < Na; a++) {
            uint64_t s = 0;
            for (int k = 0; k < numBatches; k++) s += mcnt[(size_t)k * Na + a];
            out->muffleTotals[a] = (uint32_t)s;
        }
    if (out->permeationSum)
        for (int a = 0; a < Na; a++) out->permeationSum[a] = (double)sumInt[a] + (double)sumFrac[a] / 68719476736.0;

    if (!out->audioTargetSettings || !(p->jobs & ART_JOB_PROCESS)) return ART_OK;

    // ---- ProcessAudioDataJob.Execute (PA:32-76)
    const int maxRayHits = H * N;                                                    // PA:35
    float reverbStrength, reverbVolume;
    if ((p->flags & ART_FRAME_REVERB_SEQ_FP32) && h.echo.seqValid && h.shards == 1) {
        const float avgReverbDist = h.echo.seqTotal / (float)maxRayHits;             // PA:49
        reverbStrength = avgReverbDist / p->maxReverbDistance;                       // PA:50
        reverbVolume = h.echo.seqZeros / (float)maxRayHits;                          // PA:51
    } else {
        // exact total of the half values: (hi * 2^20 + lo) / 2^24
        long double total = ((long double)h.echo.fixedHi * 1048576.0L + (long double)h.echo.fixedLo) / 16777216.0L;
        if (h.echo.nan || (h.echo.posInf && h.echo.negInf)) total = NAN;
        else if (h.echo.posInf) total = INFINITY;
        else if (h.echo.negInf) total = -INFINITY;
        reverbStrength = (float)((double)total / (double)maxRayHits / (double)p->maxReverbDistance);
        reverbVolume = (float)((double)h.echo.zeros / (double)maxRayHits);
    }
    const int maxBatchSize = (T * Na) / Na;                                          // PA:34
    for (int a = 0; a < Na; a++) {
        int totalMuffleRayhits = 0;
        float totalPermeationPower = 0.0f;
        for (int i = 0; i < maxBatchSize; i++) {                                     // PA:61-65
            totalMuffleRayhits += table[(size_t)Na * i + a];
            totalPermeationPower += permTable[(size_t)Na * i + a];
        }
        float muffle = 1.0f - (float)totalMuffleRayhits / (float)(N * H) * p->muffleEffectiveness;               // PA:68
        const float permeation = totalPermeationPower / (float)N / p->permeationStrengthPerRay * p->permeationEffectiveness; // PA:69
        muffle = saturate_h(muffle - permeation);                                    // PA:71
        ArtTargetSettings s;                                                         // ctor DT/AudioTargetRTSettings.cs:18-24
        s.muffleStrength = saturate_h(muffle);
        s.reverbStrength = saturate_h(reverbStrength);
        s.reverbVolume = saturate_h(reverbVolume);
        s.percievedAudioPosition[0] = p->audioTargetPositions ? p->audioTargetPositions[3 * a] : 0.0f;
        s.percievedAudioPosition[1] = p->audioTargetPositions ? p->audioTargetPositions[3 * a + 1] : 0.0f;
        s.percievedAudioPosition[2] = p->audioTargetPositions ? p->audioTargetPositions[3 * a + 2] : 0.0f;
        out->audioTargetSettings[a] = s;
    }
    return ART_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

// frees everything a (possibly half-constructed) device context owns; used by art_destroy and by every failure path of
// art_create
static void release_ctx(ArtCtx* ctx)
{
    if (!ctx) return;
    for (ArtCtx* c : ctx->children) release_ctx(c);
    ctx->children.clear();
    if (ctx->stream || ctx->copyStream || ctx->stream2) cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    if (ctx->copyStream) cudaStreamSynchronize(ctx->copyStream);
    if (ctx->comm && nccl_api().ok()) nccl_api().CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    for (DevBuf* b : { &ctx->rawScene, &ctx->geom, &ctx->attrs, &ctx->owners, &ctx->perm, &ctx->gridCells, &ctx->gridEntries, &ctx->gridRangeO, &ctx->gridScratch, &ctx->rotateLog, &ctx->gridCnt, &ctx->gridCtl, &ctx->hitRecs, &ctx->queryScratch, &ctx->permHitPts, &ctx->permBinCnt, &ctx->permPairs, &ctx->fanBoxes, &ctx->fanCells, &ctx->fanEntries, &ctx->fanCtl, &ctx->fanOrder, &ctx->fanScratch, &ctx->dirs, &ctx->targets, &ctx->ownedCount,
                       &ctx->outAll, &ctx->firstHit, &ctx->partials, &ctx->queue, &ctx->gathered })
        b->release();
    for (PinBuf* b : { &ctx->pinScene, &ctx->pinRays, &ctx->pinTargets, &ctx->pinOwnedCount, &ctx->pinPartials, &ctx->pinAll, &ctx->pinPerm, &ctx->pinFanCtl, &ctx->pinGathered, &ctx->pinGrid, &ctx->pinGridCtl })
        b->release();
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t evx : { ctx->evReady, ctx->evP0, ctx->evP1, ctx->evTraceDone, ctx->evCopyDone, ctx->evFan, ctx->evBounce, ctx->evX0, ctx->evX1, ctx->evF0 })
        if (evx) cudaEventDestroy(evx);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

ART_API int32_t art_create(const ArtConfig* cfg, ArtCtx** out)
{
    if (!cfg || !out) return fail(nullptr, ART_E_ARG, "art_create: null argument");
    *out = nullptr;
    if (cfg->abiVersion != ART_ABI_VERSION) return fail(nullptr, ART_E_ARG, "art_create: ABI version %d != %d", cfg->abiVersion, ART_ABI_VERSION);
    int nDev = 0;
    cudaError_t e = cudaGetDeviceCount(&nDev);
    if (e != cudaSuccess || nDev <= 0) {
        cudaGetLastError();
        return fail(nullptr, ART_E_NO_DEVICE, "art_create: no CUDA device (%s); libaudiort_cuda has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (cfg->nDevices > 1) {
        // ---- one context over several GPUs: a parent that owns one ordinary device context per entry of devices[]
        if (cfg->nDevices > ART_MAX_DEVICES) return fail(nullptr, ART_E_ARG, "art_create: nDevices %d > %d", cfg->nDevices, ART_MAX_DEVICES);
        if (cfg->shardChunkRays < 0) return fail(nullptr, ART_E_ARG, "art_create: shardChunkRays < 0");
        ArtCtx* parent = new (std::nothrow) ArtCtx();
        if (!parent) return fail(nullptr, ART_E_NOMEM, "art_create: out of memory");
        parent->device = cfg->devices[0];
        for (int i = 0; i < cfg->nDevices; i++) {
            ArtConfig c1 = *cfg;
            c1.nDevices = 0; c1.device = cfg->devices[i];
            ArtCtx* child = nullptr;
            const int32_t rc = art_create(&c1, &child);
            if (rc != ART_OK) { release_ctx(parent); return rc; }      // (g_createError holds the child's message)
            child->shardIndex = i; child->shardCount = cfg->nDevices;
            child->chunkRays = cfg->shardChunkRays; child->chunkAuto = cfg->shardChunkRays == 0;
            child->scatterGlobal = true;
            parent->children.push_back(child);
        }
        *out = parent;
        return ART_OK;
    }
    if (cfg->device < 0 || cfg->device >= nDev) return fail(nullptr, ART_E_ARG, "art_create: device %d out of range [0,%d)", cfg->device, nDev);
    ArtCtx* ctx = new (std::nothrow) ArtCtx();
    if (!ctx) return fail(nullptr, ART_E_NOMEM, "art_create: out of memory");
    ctx->device = cfg->device;
    auto bail = [&](cudaError_t ee, const char* what) {
        int32_t rc = fail(nullptr, ART_E_CUDA, "art_create: %s: %s", what, cudaGetErrorString(ee));
        release_ctx(ctx);
        return rc;
    };
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    // the library carries one arch-specific cubin (sm_100a, no PTX): it runs on compute capability 10.0 only
    cudaFuncAttributes fattr;
    if (prop.major != 10 || prop.minor != 0 || cudaFuncGetAttributes(&fattr, art::fibonacci_kernel) != cudaSuccess) {
        cudaGetLastError();
        release_ctx(ctx);
        return fail(nullptr, ART_E_NO_DEVICE, "art_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
    }
    ctx->numSms = prop.multiProcessorCount;
    ctx->maxSmemOptin = (int)prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->evReady, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->evP0)) != cudaSuccess || (e = cudaEventCreate(&ctx->evP1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->evFan)) != cudaSuccess || (e = cudaEventCreate(&ctx->evBounce)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->evX0)) != cudaSuccess || (e = cudaEventCreate(&ctx->evX1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreate(&ctx->evF0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->evTraceDone, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaEventCreateWithFlags(&ctx->evCopyDone, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if (const char* v = getenv("ART_DISABLE_GRID")) ctx->gridDisabled = atoi(v) != 0;
    if (const char* v = getenv("ART_GRID_MIN_RAYS")) ctx->gridMinRays = atoi(v);
    if (const char* v = getenv("ART_DISABLE_FANS")) ctx->fansDisabled = atoi(v) != 0;
    if (const char* v = getenv("ART_FAN_ENTRIES_PER_PAIR")) { const long n = atol(v); if (n >= 1 && n <= 65536) ctx->fanEntriesPerPair = (size_t)n; }
    if (const char* v = getenv("ART_GRID_ENTRIES_PER_COLLIDER")) { const long n = atol(v); if (n >= 0 && n <= 16384) ctx->gridEntriesPerCollider = (size_t)n; }
    if (const char* v = getenv("ART_GRID_CELL_SCALE")) { const float f = (float)atof(v); if (f > 0.05f && f < 50.0f) ctx->gridCellScale = f; }
    *out = ctx;
    return ART_OK;
}

ART_API void art_destroy(ArtCtx* ctx) { release_ctx(ctx); }

ART_API const char* art_last_error(ArtCtx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

ART_API int32_t art_set_scene(ArtCtx* ctx, const ArtAABB* aabbs, int32_t nAABB, const ArtOBB* obbs, int32_t nOBB,
                              const ArtSphere* spheres, int32_t nSphere)
{
    if (!ctx) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_set_scene: a frame is in flight");
    if (!ctx->children.empty()) {
        for (ArtCtx* c : ctx->children) {
            const int32_t rc = art_set_scene(c, aabbs, nAABB, obbs, nOBB, spheres, nSphere);
            if (rc != ART_OK) return fail(ctx, rc, "%s", c->err.c_str());
        }
        ctx->haveScene = true;
        return ART_OK;
    }
    if (nAABB < 0 || nOBB < 0 || nSphere < 0 || (nAABB && !aabbs) || (nOBB && !obbs) || (nSphere && !spheres))
        return fail(ctx, ART_E_ARG, "art_set_scene: bad counts/pointers");
    if (nAABB >= (1 << 24) || nOBB >= (1 << 24) || nSphere >= (1 << 24)) return fail(ctx, ART_E_ARG, "art_set_scene: too many colliders");
    static_assert(sizeof(ArtAABB) == 20 && sizeof(ArtOBB) == 26 && sizeof(ArtSphere) == 16, "wire layout");
    // The reference double-buffers its collider arrays (DataTypes/NativeJobBatch.cs:36-50) and re-bakes only the dynamic
    // colliders per frame (ACM:115-122): diff the new payload against the previous one and remember, per type, the range of
    // structs that changed -- only that range is uploaded by the next frame (nothing at all when the scene is unchanged).
    {
        const uint16_t* src[3] = { reinterpret_cast<const uint16_t*>(spheres), reinterpret_cast<const uint16_t*>(aabbs), reinterpret_cast<const uint16_t*>(obbs) };
        std::vector<uint16_t>* dst[3] = { &ctx->hostS, &ctx->hostA, &ctx->hostO };
        const int words[3] = { 8, 10, 13 }, cnt[3] = { nSphere, nAABB, nOBB };
        const bool sameLayout = ctx->haveScene && ctx->hostS.size() == 8 * (size_t)nSphere && ctx->hostA.size() == 10 * (size_t)nAABB &&
                                ctx->hostO.size() == 13 * (size_t)nOBB;
        bool any = !sameLayout;
        for (int t = 0; t < 3; t++) {
            int lo = cnt[t], hi = 0;
            if (sameLayout) {
                const uint16_t* old = dst[t]->data();
                const size_t rowBytes = 2 * (size_t)words[t];
                for (int i = 0; i < cnt[t]; i++)
                    if (memcmp(old + (size_t)i * words[t], src[t] + (size_t)i * words[t], rowBytes) != 0) { lo = i; break; }
                for (int i = cnt[t] - 1; i >= lo && lo < cnt[t]; i--)
                    if (memcmp(old + (size_t)i * words[t], src[t] + (size_t)i * words[t], rowBytes) != 0) { hi = i + 1; break; }
                if (lo < hi) {
                    any = true;
                    memcpy(dst[t]->data() + (size_t)lo * words[t], src[t] + (size_t)lo * words[t], rowBytes * (size_t)(hi - lo));
                }
            } else {
                lo = 0; hi = cnt[t];
                dst[t]->assign(src[t], src[t] + (size_t)words[t] * cnt[t]);
            }
            // (ranges of frames that were never scheduled accumulate)
            if (ctx->sceneDirty && sameLayout) { ctx->dirtyLo[t] = std::min(ctx->dirtyLo[t], lo); ctx->dirtyHi[t] = std::max(ctx->dirtyHi[t], hi); }
            else { ctx->dirtyLo[t] = lo; ctx->dirtyHi[t] = hi; }
        }
        if (!sameLayout) ctx->sceneLayoutChanged = true;
        if (!any) return ART_OK;                       // identical scene: packed geometry, grid and owner tables stay valid
    }
    GeomLayout& L = ctx->L;
    L.ns = nSphere; L.na = nAABB; L.no = nOBB;
    L.nsPad = (nSphere + SC_S - 1) / SC_S * SC_S;
    L.naPad = (nAABB + SC_A - 1) / SC_A * SC_A;
    L.noPad = (nOBB + SC_O - 1) / SC_O * SC_O;
    uint32_t o = 0;
    L.offSph = o; o += 16u * L.nsPad;
    L.offAabbA = o; o += 16u * L.naPad;
    L.offAabbB = o; o += 8u * L.naPad;
    L.offObbQ = o; o += 16u * L.noPad;
    L.offObbC = o; o += 16u * L.noPad;
    L.offObbH = o; o += 8u * L.noPad;
    L.bytes = (o + 15u) & ~15u;
    ctx->haveScene = true;
    ctx->sceneDirty = true;
    ctx->permPreparedForTargets = -1;
    return ART_OK;
}

ART_API int32_t art_set_rays(ArtCtx* ctx, const uint16_t* dirs, int32_t rayCount)
{
    if (!ctx) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_set_rays: a frame is in flight");
    if (!dirs || rayCount <= 0) return fail(ctx, ART_E_ARG, "art_set_rays: bad arguments");
    if (!ctx->children.empty()) {
        for (ArtCtx* c : ctx->children) {
            const int32_t rc = art_set_rays(c, dirs, rayCount);
            if (rc != ART_OK) return fail(ctx, rc, "%s", c->err.c_str());
        }
        ctx->nGlobal = rayCount; ctx->haveRays = true;
        return ART_OK;
    }
    cudaSetDevice(ctx->device);
    ctx->nGlobal = rayCount;
    if (ctx->shardCount > 1) {
        // a shard needs only its own directions: take the chunks it owns straight from the caller's array, in local order
        // (one rank of 8 stages 0.8 MB of C3's 6.3 MB). The context then holds only its shard: after a change of the shard
        // map the rays have to be set again.
        const int chunk = effective_chunk(ctx);
        const size_t nLoc = (size_t)local_ray_count(rayCount, ctx->shardIndex, ctx->shardCount, chunk);
        ShardMap m;
        m.nGlobal = rayCount; m.nLocal = (int)nLoc; m.shardIndex = ctx->shardIndex; m.shardCount = ctx->shardCount; m.chunk = chunk; m.dirsLocal = 1;
        CK(ctx->pinRays.ensure(6 * nLoc + 16));
        unsigned char* dst = ctx->pinRays.as<unsigned char>();
        const unsigned char* src = reinterpret_cast<const unsigned char*>(dirs);
        for (size_t j0 = 0; j0 < nLoc; j0 += (size_t)chunk) {
            const size_t cnt = std::min((size_t)chunk, nLoc - j0);
            memcpy(dst + 6 * j0, src + 6 * (size_t)m.to_global((int)j0), 6 * cnt);
        }
        ctx->rayHostLocal = true; ctx->rayHostValid = false;
        ctx->rayStamp[0] = ctx->shardIndex; ctx->rayStamp[1] = ctx->shardCount; ctx->rayStamp[2] = chunk;
    } else {
        CK(ctx->pinRays.ensure(6 * (size_t)rayCount));
        memcpy(ctx->pinRays.p, dirs, 6 * (size_t)rayCount);
        ctx->rayHostLocal = false; ctx->rayHostValid = true;
    }
    ctx->haveRays = true;
    ctx->raysDirty = true;
    return ART_OK;
}

ART_API int32_t art_generate_fibonacci_rays(ArtCtx* ctx, int32_t rayCount)
{
    if (!ctx) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_generate_fibonacci_rays: a frame is in flight");
    if (rayCount <= 0) return fail(ctx, ART_E_ARG, "art_generate_fibonacci_rays: bad ray count");
    if (!ctx->children.empty()) {
        for (ArtCtx* c : ctx->children) {
            const int32_t rc = art_generate_fibonacci_rays(c, rayCount);
            if (rc != ART_OK) return fail(ctx, rc, "%s", c->err.c_str());
        }
        ctx->nGlobal = rayCount; ctx->haveRays = true;
        return ART_OK;
    }
    cudaSetDevice(ctx->device);
    CK(ctx->dirs.ensure(6 * (size_t)rayCount));
    CK(launch_fibonacci(ctx->dirs.as<uint16_t>(), rayCount, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->nGlobal = rayCount;
    ctx->haveRays = true;
    ctx->raysDirty = false;
    ctx->rayHostValid = false; ctx->rayHostLocal = false;
    ctx->dirsLocal = false;                        // the whole batch lives on the device
    return ART_OK;
}

ART_API int32_t art_get_rays(ArtCtx* ctx, uint16_t* dirs, int32_t capacityRays)
{
    if (!ctx || !dirs) return ART_E_ARG;
    if (!ctx->haveRays) return fail(ctx, ART_E_STATE, "art_get_rays: no rays set");
    if (!ctx->children.empty()) return art_get_rays(ctx->children[0], dirs, capacityRays);
    if (capacityRays < ctx->nGlobal) return fail(ctx, ART_E_ARG, "art_get_rays: capacity %d < %d", capacityRays, ctx->nGlobal);
    cudaSetDevice(ctx->device);
    if (ctx->rayHostValid) { memcpy(dirs, ctx->pinRays.p, 6 * (size_t)ctx->nGlobal); return ART_OK; }
    if (ctx->rayHostLocal) return fail(ctx, ART_E_STATE, "art_get_rays: a sharded context holds only its own shard of the directions");
    CK(cudaMemcpyAsync(dirs, ctx->dirs.p, 6 * (size_t)ctx->nGlobal, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ART_OK;
}

ART_API int32_t art_set_ray_shard(ArtCtx* ctx, int32_t shardIndex, int32_t shardCount, int32_t chunkRays)
{
    if (!ctx) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_set_ray_shard: a frame is in flight");
    if (shardCount < 1 || shardIndex < 0 || shardIndex >= shardCount || chunkRays < 0)
        return fail(ctx, ART_E_ARG, "art_set_ray_shard: bad shard %d/%d chunk %d", shardIndex, shardCount, chunkRays);
    if (!ctx->children.empty() || ctx->scatterGlobal) return fail(ctx, ART_E_STATE, "art_set_ray_shard: a multi-device context owns its shard map");
    if (ctx->comm) return fail(ctx, ART_E_STATE, "art_set_ray_shard: the communicator owns the shard map (art_comm_init)");
    if (shardIndex != ctx->shardIndex || shardCount != ctx->shardCount || chunkRays != ctx->chunkRays) {
        if (ctx->rayHostValid) ctx->raysDirty = true;   // the device holds only the shard's directions: upload the new shard
    }
    ctx->shardIndex = shardIndex; ctx->shardCount = shardCount; ctx->chunkRays = chunkRays; ctx->chunkAuto = false;
    return ART_OK;
}

ART_API int32_t art_local_ray_count(ArtCtx* ctx)
{
    if (!ctx) return ART_E_ARG;
    if (!ctx->haveRays) return fail(ctx, ART_E_STATE, "art_local_ray_count: no rays set");
    if (!ctx->children.empty()) return ctx->nGlobal;    // every ray is local to a multi-device context
    return local_ray_count(ctx->nGlobal, ctx->shardIndex, ctx->shardCount, effective_chunk(ctx));
}

// ---- multi-device context: fan the frame out over the children, merge their partial results -------------------------
static int32_t multi_schedule(ArtCtx* ctx, const ArtParams* prm, const ArtOutputs* outputs, ArtHandle* outHandle)
{
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_trace_schedule: a frame is already in flight");
    if (!ctx->haveScene || !ctx->haveRays) return fail(ctx, ART_E_STATE, "art_trace_schedule: scene or rays not set");
    if (prm->totalAudioTargets <= 0 || !prm->audioTargetPositions) return fail(ctx, ART_E_ARG, "totalAudioTargets must be >= 1 and audioTargetPositions non-null");
    if (prm->flags & ART_FRAME_REVERB_SEQ_FP32) return fail(ctx, ART_E_ARG, "ART_FRAME_REVERB_SEQ_FP32 needs the whole echo array on one device");
    ArtParams p = *prm;
    p.flags |= ART_FRAME_PARTIALS_ONLY;               // the children export partial blobs; this context merges and finalises
    const size_t n = ctx->children.size();
    std::vector<int32_t> rcs(n, ART_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; i++)
        th.emplace_back([&, i] { ArtHandle h = 0; rcs[i] = art_trace_schedule(ctx->children[i], &p, outputs, &h); });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < n; i++)
        if (rcs[i] != ART_OK) {
            for (size_t k = 0; k < n; k++)
                if (rcs[k] == ART_OK) art_complete(ctx->children[k], ctx->children[k]->handle);   // do not leave frames in flight
            return fail(ctx, rcs[i], "device %d: %s", ctx->children[i]->device, ctx->children[i]->err.c_str());
        }
    ctx->parentTargets.assign(prm->audioTargetPositions, prm->audioTargetPositions + 3 * (size_t)prm->totalAudioTargets);
    ctx->params = *prm;
    ctx->params.audioTargetPositions = ctx->parentTargets.data();
    ctx->haveUserOut = outputs != nullptr;
    if (outputs) ctx->userOut = *outputs; else memset(&ctx->userOut, 0, sizeof ctx->userOut);
    ctx->frameFlags = prm->flags; ctx->frameJobs = prm->jobs;
    ctx->frameNa = prm->totalAudioTargets; ctx->frameT = prm->batchCount; ctx->frameH = prm->maxHitsPerRay;
    ctx->inFlight = true; ctx->frameDone = false;
    ctx->handle++;
    *outHandle = ctx->handle;
    return ART_OK;
}

static int32_t multi_is_completed(ArtCtx* ctx)
{
    for (ArtCtx* c : ctx->children) {
        const int32_t rc = art_is_completed(c, c->handle);
        if (rc < 0) return fail(ctx, rc, "device %d: %s", c->device, c->err.c_str());
        if (rc == 0) return 0;
    }
    return 1;
}

static int32_t multi_complete(ArtCtx* ctx)
{
    const size_t n = ctx->children.size();
    std::vector<int32_t> rcs(n, ART_OK);
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; i++)
        th.emplace_back([&, i] { rcs[i] = art_complete(ctx->children[i], ctx->children[i]->handle); });   // (scatters its per-ray outputs)
    for (auto& t : th) t.join();
    ctx->inFlight = false;
    for (size_t i = 0; i < n; i++)
        if (rcs[i] != ART_OK) return fail(ctx, rcs[i], "device %d: %s", ctx->children[i]->device, ctx->children[i]->err.c_str());
    // exact, order-independent merge of the per-source partials (integer sums + max-by-ray-index select)
    ctx->lastBlob = ctx->children[0]->lastBlob;
    for (size_t i = 1; i < n; i++) {
        const int32_t rc = art_partials_merge(ctx->lastBlob.data(), ctx->children[i]->lastBlob.data(), (int64_t)ctx->lastBlob.size());
        if (rc != ART_OK) return fail(ctx, rc, "merging the partial results of device %d failed", ctx->children[i]->device);
    }
    ArtCounters& c = ctx->counters;
    c = ArtCounters{};
    for (size_t i = 0; i < n; i++) {
        const ArtCounters& k = ctx->children[i]->counters;
        const uint64_t* src = &k.segments; uint64_t* dst = &c.segments;
        for (int w = 0; w < 22; w++) dst[w] += src[w];              // segments .. permLossTests: 22 consecutive u64 counters
        for (int w = 0; w < 3; w++) { c.gridTraceTests[w] += k.gridTraceTests[w]; c.gridPermFirstTests[w] += k.gridPermFirstTests[w]; c.gridPermLossTests[w] += k.gridPermLossTests[w]; }
        c.gridTraceCells += k.gridTraceCells; c.gridPermCells += k.gridPermCells; c.debugViolations += k.debugViolations;
        for (int w = 0; w < 3; w++) c.gridQueryTests[w] += k.gridQueryTests[w];
        c.gridQueryLists += k.gridQueryLists;
        c.traceMs = std::max(c.traceMs, k.traceMs); c.permeationMs = std::max(c.permeationMs, k.permeationMs);
        c.reduceMs = std::max(c.reduceMs, k.reduceMs); c.deviceMs = std::max(c.deviceMs, k.deviceMs);
        c.h2dMs = std::max(c.h2dMs, k.h2dMs); c.d2hMs = std::max(c.d2hMs, k.d2hMs);
        c.fanBuildMs = std::max(c.fanBuildMs, k.fanBuildMs); c.bounceMs = std::max(c.bounceMs, k.bounceMs); c.queryMs = std::max(c.queryMs, k.queryMs);
        c.kernelLaunches += k.kernelLaunches;
        c.gridUsed |= k.gridUsed;
    }
    c.devicesUsed = (uint32_t)n;
    ctx->frameDone = true;
    if (!(ctx->frameFlags & ART_FRAME_PARTIALS_ONLY) && ctx->haveUserOut) {
        std::string err;
        const int32_t rc = finalize_blob(ctx->lastBlob.data(), ctx->lastBlob.size(), &ctx->params, ctx->nGlobal, &ctx->userOut, &err);
        if (rc != ART_OK) return fail(ctx, rc, "finalize: %s", err.c_str());
    }
    return ART_OK;
}

ART_API int32_t art_trace_schedule(ArtCtx* ctx, const ArtParams* prm, const ArtOutputs* outputs, ArtHandle* outHandle)
{
    if (!ctx || !prm || !outHandle) return ART_E_ARG;
    if (!ctx->children.empty()) return multi_schedule(ctx, prm, outputs, outHandle);
    if (ctx->poisoned) return fail(ctx, ART_E_CUDA, "context poisoned by an earlier CUDA error");
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_trace_schedule: a frame is already in flight");
    if (!ctx->haveScene || !ctx->haveRays) return fail(ctx, ART_E_STATE, "art_trace_schedule: scene or rays not set");
    const int Na = prm->totalAudioTargets, T = prm->batchCount, H = prm->maxHitsPerRay, N = ctx->nGlobal;
    if (Na <= 0) return fail(ctx, ART_E_ARG, "totalAudioTargets must be >= 1 (RT:63 divides by it)");
    if (Na > 32767) return fail(ctx, ART_E_ARG, "totalAudioTargets must fit a C# short (RT:153)");
    if (!prm->audioTargetPositions) return fail(ctx, ART_E_ARG, "audioTargetPositions is null");
    if (H < 1) return fail(ctx, ART_E_ARG, "maxHitsPerRay must be >= 1");
    if (T < 1) return fail(ctx, ART_E_ARG, "batchCount must be >= 1");
    if ((prm->jobs & ART_JOB_ALL) == 0) return fail(ctx, ART_E_ARG, "no jobs requested");
    if ((prm->jobs & ART_JOB_PROCESS) && (prm->jobs & (ART_JOB_RAYTRACE | ART_JOB_PERMEATION)) != (ART_JOB_RAYTRACE | ART_JOB_PERMEATION))
        return fail(ctx, ART_E_ARG, "ART_JOB_PROCESS needs ART_JOB_RAYTRACE and ART_JOB_PERMEATION in the same frame (ART:237)");
    if ((long long)N * H > 0x7FFFFFFFLL) return fail(ctx, ART_E_ARG, "rayCount * maxHitsPerRay overflows int32 (RT:115)");
    if ((long long)T * Na > 0x7FFFFFFFLL) return fail(ctx, ART_E_ARG, "batchCount * targets overflows int32");
    const bool wantRT = prm->jobs & ART_JOB_RAYTRACE, wantPM = prm->jobs & ART_JOB_PERMEATION;
    const bool count = prm->flags & ART_FRAME_COUNTERS;
    const bool hostOut = !(prm->flags & ART_FRAME_NO_HOST_OUTPUTS);
    if ((prm->flags & ART_FRAME_REVERB_SEQ_FP32) && ctx->shardCount > 1)
        return fail(ctx, ART_E_ARG, "ART_FRAME_REVERB_SEQ_FP32 needs the whole echo array on one device (shardCount == 1)");
    const int b = batch_size(N, T);
    {   // the reference would index out of range if a batch slot fell outside the table
        const int numBatches = (N + b - 1) / b;
        for (int k = 0; k < numBatches; k++) {
            const int32_t row = (int32_t)((uint32_t)(k * b) * (uint32_t)T) / N;
            if (row < 0 || row >= T) return fail(ctx, ART_E_ARG, "batch %d maps to slot row %d outside [0,%d) (RT:64 int32 overflow)", k, row, T);
        }
    }
    cudaSetDevice(ctx->device);
    ctx->kernelLaunches = 0;
    ctx->frameGridUsed = 0;
    if (!ctx->rerunning) ctx->frameIsRerun = false;

    const int chunk = effective_chunk(ctx);
    ShardMap map;
    map.nGlobal = N; map.shardIndex = ctx->shardIndex; map.shardCount = ctx->shardCount; map.chunk = chunk;
    map.nLocal = local_ray_count(N, ctx->shardIndex, ctx->shardCount, chunk);
    const size_t nLoc = (size_t)map.nLocal, NH = nLoc * H;
    const GeomLayout& L = ctx->L;

    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    // ---------------- uploads ----------------
    if (ctx->sceneDirty) {
        const size_t bS = ctx->hostS.size() * 2, bA = ctx->hostA.size() * 2, bO = ctx->hostO.size() * 2;
        const size_t offA = (bS + 15) & ~(size_t)15, offO = (offA + bA + 15) & ~(size_t)15, tot = offO + bO + 16;
        const bool fullUpload = ctx->sceneLayoutChanged || ctx->pinScene.cap < tot || ctx->rawScene.cap < tot;
        CK(ctx->pinScene.ensure(tot));
        CK(ctx->rawScene.ensure(tot));
        unsigned char* hp = ctx->pinScene.as<unsigned char>();
        if (fullUpload) {
            if (bS) memcpy(hp, ctx->hostS.data(), bS);
            if (bA) memcpy(hp + offA, ctx->hostA.data(), bA);
            if (bO) memcpy(hp + offO, ctx->hostO.data(), bO);
            CK(cudaMemcpyAsync(ctx->rawScene.p, hp, tot, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            // only the structs that changed since the last upload (art_set_scene diffed the payloads)
            const size_t offs[3] = { 0, offA, offO }, row[3] = { 16, 20, 26 };
            const std::vector<uint16_t>* srcv[3] = { &ctx->hostS, &ctx->hostA, &ctx->hostO };
            for (int t = 0; t < 3; t++) {
                if (ctx->dirtyLo[t] >= ctx->dirtyHi[t]) continue;
                const size_t b0 = row[t] * (size_t)ctx->dirtyLo[t], nb = row[t] * (size_t)(ctx->dirtyHi[t] - ctx->dirtyLo[t]);
                memcpy(hp + offs[t] + b0, reinterpret_cast<const unsigned char*>(srcv[t]->data()) + b0, nb);
                CK(cudaMemcpyAsync(ctx->rawScene.as<unsigned char>() + offs[t] + b0, hp + offs[t] + b0, nb, cudaMemcpyHostToDevice, ctx->stream));
            }
        }
        ctx->sceneLayoutChanged = false;
        CK(ctx->geom.ensure(L.bytes + 16));
        const size_t nAttr4 = (size_t)L.nsPad + 3 * (size_t)L.naPad + 3 * (size_t)L.noPad;
        CK(ctx->attrs.ensure(nAttr4 * sizeof(float4) + 16));
        CK(ctx->owners.ensure(((size_t)L.nsPad + L.naPad + L.noPad) * sizeof(short) + 16));
        PackArgs pa;
        unsigned char* rp = ctx->rawScene.as<unsigned char>();
        pa.rawS = reinterpret_cast<const uint16_t*>(rp); pa.rawA = reinterpret_cast<const uint16_t*>(rp + offA);
        pa.rawO = reinterpret_cast<const uint16_t*>(rp + offO);
        pa.L = L; pa.geom = ctx->geom.as<unsigned char>();
        float4* at = ctx->attrs.as<float4>();
        pa.sphAttr = at; at += L.nsPad;
        pa.aabbAttr = at; at += L.naPad;
        pa.aabbCtr = at; at += L.naPad;
        pa.aabbHalf = at; at += L.naPad;
        pa.obbAttr = at; at += L.noPad;
        pa.obbHalf = at; at += L.noPad;
        pa.obbQinv = at;
        short* ow = ctx->owners.as<short>();
        pa.ownS = ow; pa.ownA = ow + L.nsPad; pa.ownO = ow + L.nsPad + L.naPad;
        CK(launch_pack(pa, ctx->stream));
        ctx->kernelLaunches++;
        ctx->grid.ok = false;
        ctx->gridBuilt = false;                        // built lazily by the first frame that wants it
        ctx->sceneDirty = false;
    }
    if (ctx->rayHostLocal && (ctx->rayStamp[0] != ctx->shardIndex || ctx->rayStamp[1] != ctx->shardCount || ctx->rayStamp[2] != chunk))
        return fail(ctx, ART_E_STATE, "the shard map changed after art_set_rays: set the rays again (a sharded context keeps only its own directions)");
    if (ctx->raysDirty) {
        if (ctx->rayHostLocal) {
            CK(ctx->dirs.ensure(6 * nLoc + 16));
            CK(cudaMemcpyAsync(ctx->dirs.p, ctx->pinRays.p, 6 * nLoc, cudaMemcpyHostToDevice, ctx->stream));
            ctx->dirsLocal = true;
        } else if (ctx->shardCount > 1) {
            // the batch was set before the shard map: pack the chunks this shard owns (local order) behind it in pinned memory
            const size_t all = (6 * (size_t)N + 63) & ~(size_t)63;
            if (ctx->pinRays.cap < all + 6 * nLoc) {
                PinBuf bigger;
                CK(bigger.ensure(all + 6 * nLoc));
                memcpy(bigger.p, ctx->pinRays.p, 6 * (size_t)N);
                ctx->pinRays.release();
                ctx->pinRays = bigger;
            }
            const unsigned char* src = ctx->pinRays.as<unsigned char>();
            unsigned char* dst = ctx->pinRays.as<unsigned char>() + all;
            for (size_t j0 = 0; j0 < nLoc; j0 += (size_t)chunk) {
                const size_t cnt = std::min((size_t)chunk, nLoc - j0);
                memcpy(dst + 6 * j0, src + 6 * (size_t)map.to_global((int)j0), 6 * cnt);
            }
            CK(ctx->dirs.ensure(6 * nLoc + 16));
            CK(cudaMemcpyAsync(ctx->dirs.p, dst, 6 * nLoc, cudaMemcpyHostToDevice, ctx->stream));
            ctx->dirsLocal = true;
        } else {
            CK(ctx->dirs.ensure(6 * (size_t)N));
            CK(cudaMemcpyAsync(ctx->dirs.p, ctx->pinRays.p, 6 * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
            ctx->dirsLocal = false;
        }
        ctx->raysDirty = false;
    }
    map.dirsLocal = ctx->dirsLocal ? 1 : 0;
    CK(ctx->pinTargets.ensure(16 * (size_t)Na));
    CK(ctx->targets.ensure(16 * (size_t)Na));
    memmove(ctx->pinTargets.p, prm->audioTargetPositions, 12 * (size_t)Na);   // (a re-run passes the library-owned copy back in)
    {   // Morton order of the targets: neighbouring lanes of the grid kernels then walk towards neighbouring targets
        // (same cells, same list lengths). Results do not depend on the order.
        int* order = reinterpret_cast<int*>(ctx->pinTargets.as<unsigned char>() + 12 * (size_t)Na);
        const float* tp = prm->audioTargetPositions;
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        for (int t = 0; t < Na; t++)
            for (int k = 0; k < 3; k++) { lo[k] = std::fmin(lo[k], tp[3 * t + k]); hi[k] = std::fmax(hi[k], tp[3 * t + k]); }
        std::vector<std::pair<uint32_t, int>> keys((size_t)Na);
        for (int t = 0; t < Na; t++) {
            uint32_t code = 0;
            uint32_t q[3];
            for (int k = 0; k < 3; k++) {
                const float ext = hi[k] - lo[k];
                float u = ext > 0.0f ? (tp[3 * t + k] - lo[k]) / ext : 0.0f;
                if (!(u >= 0.0f)) u = 0.0f;
                if (u > 1.0f) u = 1.0f;
                q[k] = (uint32_t)(u * 1023.0f);
            }
            for (int b = 9; b >= 0; b--)
                for (int k = 0; k < 3; k++) code = (code << 1) | ((q[k] >> b) & 1u);
            keys[(size_t)t] = { code, t };
        }
        std::sort(keys.begin(), keys.end());
        for (int t = 0; t < Na; t++) order[t] = keys[(size_t)t].second;
    }
    CK(cudaMemcpyAsync(ctx->targets.p, ctx->pinTargets.p, 16 * (size_t)Na, cudaMemcpyHostToDevice, ctx->stream));

    // owned-collider bookkeeping (depends on the target count): counts per (section, target) for the
    // counters, density planes + owned list for K2
    std::vector<int> ownedCount((size_t)3 * Na, 0);
    {
        auto tally = [&](const std::vector<uint16_t>& raw, int words, int sec) {
            const size_t n = raw.size() / words;
            for (size_t i = 0; i < n; i++) {
                const int t = (int)(short)raw[i * words + words - 1];
                if (t >= 0 && t < Na) ownedCount[(size_t)sec * Na + t]++;
            }
        };
        tally(ctx->hostS, 8, 0); tally(ctx->hostA, 10, 1); tally(ctx->hostO, 13, 2);
    }
    if (count && wantRT) {
        CK(ctx->pinOwnedCount.ensure(ownedCount.size() * 4));
        CK(ctx->ownedCount.ensure(ownedCount.size() * 4));
        memcpy(ctx->pinOwnedCount.p, ownedCount.data(), ownedCount.size() * 4);
        CK(cudaMemcpyAsync(ctx->ownedCount.p, ctx->pinOwnedCount.p, ownedCount.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    const size_t nPad = (size_t)L.nsPad + L.naPad + L.noPad;
    if (wantPM && ctx->permPreparedForTargets != Na) {
        // [dens S | dens A | dens O | ownedList] staged in pinned memory and copied asynchronously (a dynamic scene comes
        // through here every frame)
        const size_t bytes = nPad * 12 + 16;      // + true densities (target-fan path)
        CK(ctx->pinPerm.ensure(bytes));
        CK(ctx->perm.ensure(bytes));
        float* dens = ctx->pinPerm.as<float>();
        int* owned = reinterpret_cast<int*>(ctx->pinPerm.as<unsigned char>() + nPad * 4);
        float* trueDens = reinterpret_cast<float*>(ctx->pinPerm.as<unsigned char>() + nPad * 8);
        memset(dens, 0, nPad * 4);
        memset(trueDens, 0, nPad * 4);
        int nOwned = 0;
        auto fill = [&](const std::vector<uint16_t>& raw, int words, int densWord, int sec, size_t off) {
            const size_t n = raw.size() / words;
            for (size_t i = 0; i < n; i++) {
                const int t = (int)(short)raw[i * words + words - 1];
                if (t >= 0 && t < Na) owned[nOwned++] = (sec << 28) | (int)i;
                else dens[off + i] = h2f(raw[i * words + densWord]);
                trueDens[off + i] = h2f(raw[i * words + densWord]);
            }
        };
        fill(ctx->hostS, 8, 5, 0, 0); fill(ctx->hostA, 10, 7, 1, L.nsPad); fill(ctx->hostO, 13, 10, 2, (size_t)L.nsPad + L.naPad);
        ctx->nOwned = nOwned;
        CK(cudaMemcpyAsync(ctx->perm.p, ctx->pinPerm.p, nPad * 4 + (size_t)nOwned * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->perm.as<unsigned char>() + nPad * 8, trueDens, nPad * 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->permPreparedForTargets = Na;
    }

    // ---------------- device outputs ----------------
    const BlobLayout bl = blob_layout(Na, T);
    const size_t queueOff = (bl.bytes + 63) & ~(size_t)63;         // the two ray-queue counters live behind the blob: one memset
    CK(ctx->partials.ensure(queueOff + 64));
    CK(ctx->pinPartials.ensure(bl.bytes));
    CK(cudaMemsetAsync(ctx->partials.p, 0, queueOff + 64, ctx->stream));
    CK(cudaMemsetAsync(ctx->partials.as<unsigned char>() + bl.offLastHit, 0xFF, sizeof(int32_t) * (size_t)T, ctx->stream));
    unsigned char* pb = ctx->partials.as<unsigned char>();
    BlobHeader* dh = reinterpret_cast<BlobHeader*>(pb);
    const bool wantHitPts = outputs && outputs->rayHitResults, wantHitCnt = outputs && outputs->rayHitResultCounts,
               wantHitIds = outputs && outputs->hitColliderIds;
    if (wantRT) {
        // echo | hit ids | hit points | hit counts in one allocation: one memset now, one copy back later
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        size_t o = 0;
        ctx->offEcho = o; o = up(o + NH * 2);
        ctx->offHitIds = o; if (wantHitIds) o = up(o + NH * 4);
        ctx->offHitPts = o; if (wantHitPts) o = up(o + NH * 6);
        ctx->offHitCnt = o; if (wantHitCnt) o = up(o + nLoc);
        ctx->outBytes = o;
        CK(ctx->outAll.ensure(o + 256));
        CK(cudaMemsetAsync(ctx->outAll.p, 0, o, ctx->stream));
    }
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));

    AttrArrays at;
    {
        const float4* a4 = ctx->attrs.as<float4>();
        at.sphAttr = a4; a4 += L.nsPad;
        at.aabbAttr = a4; a4 += L.naPad;
        at.aabbCtr = a4; a4 += L.naPad;
        at.aabbHalf = a4; a4 += L.naPad;
        at.obbAttr = a4; a4 += L.noPad;
        at.obbHalf = a4; a4 += L.noPad;
        at.obbQinv = a4;
        const short* ow = ctx->owners.as<short>();
        at.ownS = ow; at.ownA = ow + L.nsPad; at.ownO = ow + L.nsPad + L.naPad;
        at.ownedCount = ctx->ownedCount.as<int>();
    }

    // acceleration structure: the same exact tests on the colliders near each ray only (bit-identical outputs);
    // the work counters are defined by the reference's full scans, so counting frames use the brute-force kernels
    bool useGrid = !ctx->gridDisabled && !count && !(prm->flags & ART_FRAME_BRUTE_FORCE);
    // Small batches: the brute-force kernels give a whole warp to every ray (32 colliders per step), the grid kernels one
    // lane; below a few thousand rays the GPU is not full and the brute-force mapping has the lower latency whatever the
    // scene size (tools/path_crossover.py: 4,096 rays vs 1,023 colliders 0.23 ms vs 0.74 ms; 65,536 rays 2.46 vs 0.96 ms).
    if (useGrid && !(prm->flags & ART_FRAME_FORCE_GRID)) {
        const long long nc = (long long)L.ns + L.na + L.no;
        long long minRays = ctx->gridMinRays;
        if (minRays < 0) {
            minRays = nc > 0 ? 32768LL * 1024 / nc : 32768;
            if (minRays > 32768) minRays = 32768;
            if (minRays < 4096) minRays = 4096;
        }
        if (map.nLocal < minRays) useGrid = false;
    }
    if (ctx->rerunNoGrid) useGrid = false;            // second pass of a frame whose grid build overflowed
    ctx->frameGridBuilt = false;
    if (useGrid && !ctx->gridBuilt) {
        // Uniform grid over the current scene. The host works out what is O(colliders) -- bounds, dimensions, margins -- and
        // stages the conservative boxes in pinned memory; the cell lists are filled on the device (k5_grid_build.cu). No
        // cell walk and no stream synchronisation on the host: a scene that changes every frame costs a few tens of us here.
        grid_params(ctx->hostS, ctx->hostA, ctx->hostO, ctx->gridCellScale, ctx->grid);
        if (ctx->grid.ok) {
            const size_t nCells = (size_t)ctx->grid.d.nx * ctx->grid.d.ny * ctx->grid.d.nz;
            const size_t nc = (size_t)L.ns + L.na + L.no;
            // (ART_GRID_ENTRIES_PER_COLLIDER=0, a test knob, leaves room for 16 entries only: the build overflows, the frame is
            // re-run on the brute-force kernels and the next build gets a larger buffer)
            const size_t cap = ctx->gridEntriesPerCollider == 0 ? 16
                             : std::min<size_t>((size_t)1 << 28, ctx->gridEntriesPerCollider * nc + 8 * nCells + 1024);
            const size_t boxBytes = ctx->grid.boxLo.size() * sizeof(float);
            const size_t rangeBytes = ctx->grid.rangeO.size() * sizeof(uint2);
            const size_t rangeOff = (2 * boxBytes + 15) & ~(size_t)15;
            CK(ctx->pinGrid.ensure(rangeOff + rangeBytes + 16));
            CK(ctx->pinGridCtl.ensure(16));
            CK(ctx->fanBoxes.ensure(2 * boxBytes + 32));
            CK(ctx->gridRangeO.ensure(rangeBytes + 16));
            CK(ctx->gridCells.ensure(nCells * sizeof(uint2)));
            CK(ctx->gridEntries.ensure(cap * sizeof(uint16_t) + 16));
            CK(ctx->gridCnt.ensure(nCells * 3 * sizeof(unsigned int)));
            CK(ctx->gridCtl.ensure(16));
            unsigned char* hp = ctx->pinGrid.as<unsigned char>();
            memcpy(hp, ctx->grid.boxLo.data(), boxBytes);
            memcpy(hp + boxBytes, ctx->grid.boxHi.data(), boxBytes);
            memcpy(hp + rangeOff, ctx->grid.rangeO.data(), rangeBytes);
            CK(cudaMemcpyAsync(ctx->fanBoxes.p, hp, 2 * boxBytes, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->gridRangeO.p, hp + rangeOff, rangeBytes, cudaMemcpyHostToDevice, ctx->stream));
            GridBuildArgs ga;
            ga.boxLo = ctx->fanBoxes.as<float4>(); ga.boxHi = ga.boxLo + nc;
            ga.ns = L.ns; ga.na = L.na; ga.no = L.no;
            ga.g0x = ctx->grid.d.g0x; ga.g0y = ctx->grid.d.g0y; ga.g0z = ctx->grid.d.g0z;
            ga.csx = ctx->grid.d.csx; ga.csy = ctx->grid.d.csy; ga.csz = ctx->grid.d.csz;
            ga.nx = ctx->grid.d.nx; ga.ny = ctx->grid.d.ny; ga.nz = ctx->grid.d.nz;
            ga.cnt = ctx->gridCnt.as<unsigned int>(); ga.cells = ctx->gridCells.as<uint2>(); ga.entries = ctx->gridEntries.as<uint16_t>();
            ga.capacity = (unsigned int)cap; ga.ctl = ctx->gridCtl.as<unsigned int>();
            CK(launch_grid_build(ga, ctx->stream));
            ctx->kernelLaunches += 4;
            ctx->grid.d.nEntries = (int)cap;
            ctx->frameGridBuilt = true;
        }
        ctx->gridBuilt = true;
    }
    useGrid = useGrid && ctx->grid.ok;
    if (useGrid) {
        const float dx = prm->rayOrigin[0] - ctx->grid.cx, dy = prm->rayOrigin[1] - ctx->grid.cy, dz = prm->rayOrigin[2] - ctx->grid.cz;
        useGrid = std::sqrt(dx * dx + dy * dy + dz * dz) <= ctx->grid.listenerRange;   // else the error bounds of grid_host.h do not hold
    }
    GridDesc gd = ctx->grid.d;
    gd.cells = ctx->gridCells.as<uint2>(); gd.entries = ctx->gridEntries.as<uint16_t>(); gd.rangeO = ctx->gridRangeO.as<uint2>();

    // Target fans (fan_dev.cuh): every echo / muffle / permeation query ends in the listener or an audio target, so the
    // colliders are binned by direction around those few goals, on the device, every frame (the goals move).
    FanDesc fd{};
    bool fanBeside = false;
    bool useFans = useGrid && !ctx->fansDisabled && !ctx->rerunning && !(prm->flags & ART_FRAME_NO_FANS);
    if (useFans) {
        const size_t nFans = (size_t)Na + 1, nc = (size_t)L.ns + L.na + L.no;
        // a collider can subtend every bin of a fan (a wall seen from nearby), a typical one a few dozen
        const size_t perFan = std::min(nc * (size_t)kFanCells, nc * ctx->fanEntriesPerPair + ((size_t)2048 * ctx->fanEntriesPerPair));
        const size_t cap = nFans * perFan + 4096;
        bool fit = !(cap > ((size_t)1 << 30) || nFans * kFanCells > ((size_t)1 << 28));   // > 2 GiB of lists: walk the grid instead
        if (fit) {
            // ... and so does a frame whose fan buffers cannot be allocated (tens of thousands of targets x thousands of colliders:
            // the build scratch alone is 72 B per (target, collider) pair)
            const bool sorted = nc <= 16384;                 // the per-goal sort runs in one CTA's shared memory
            fit = ctx->fanCells.ensure(nFans * kFanCells * (sizeof(uint4) + sizeof(uint2) + sizeof(uint32_t))) == cudaSuccess &&   // FanDesc::cells4, cells, firstA
                  ctx->fanEntries.ensure(cap * sizeof(uint16_t)) == cudaSuccess &&
                  ctx->fanScratch.ensure(fan_build_scratch_bytes((int)nFans, (int)nc)) == cudaSuccess &&
                  (!sorted || ctx->fanOrder.ensure(nFans * nc * sizeof(uint32_t)) == cudaSuccess);
            if (const char* v = getenv("ART_FAN_ALLOC_FAIL")) { if (atoi(v) != 0) fit = false; }   // (test knob, read per frame)
            if (!fit) {
                cudaGetLastError();
                for (DevBuf* b : { &ctx->fanCells, &ctx->fanEntries, &ctx->fanScratch, &ctx->fanOrder }) b->release();
            }
        }
        if (!fit) useFans = false;
        else {
            CK(ctx->fanCtl.ensure(16));
            CK(ctx->pinFanCtl.ensure(16));
            // (Building the fans on the second stream beside the bounce tracer was measured and dropped: the two kernels'
            // CTAs do not fit one SM's register file together, so the build only ran in the tracer's tail -- C3 / 8 shard
            // 3.26 -> 3.20 ms, but C4's 257 fans 1.55 -> 1.80 ms.)
            cudaStream_t fanStream = ctx->stream;
            CK(cudaMemsetAsync(ctx->fanCtl.p, 0, 16, fanStream));
            FanBuildArgs fa;
            fa.boxLo = ctx->fanBoxes.as<float4>(); fa.boxHi = fa.boxLo + nc;
            fa.ns = L.ns; fa.na = L.na; fa.no = L.no;
            fa.ownS = at.ownS; fa.ownA = at.ownA; fa.ownO = at.ownO;
            fa.targets = ctx->targets.as<float>(); fa.nTargets = Na;
            fa.lx = prm->rayOrigin[0]; fa.ly = prm->rayOrigin[1]; fa.lz = prm->rayOrigin[2];
            fa.nearDist = 1e-3f * ctx->grid.d.errScale;
            {   // covering depths (query_fan_kernel's cull); ART_Q_NO_COVER (experiment knob, read per frame) leaves them out
                const char* nc0 = getenv("ART_Q_NO_COVER");
                const bool cover = !(nc0 && atoi(nc0) != 0);
                const unsigned char* gb = ctx->geom.as<unsigned char>();
                fa.aabbA = cover ? reinterpret_cast<const float4*>(gb + L.offAabbA) : nullptr;
                fa.aabbB = cover ? reinterpret_cast<const float2*>(gb + L.offAabbB) : nullptr;
                fa.coverMinThickness = 1e-4f * ctx->grid.d.errScale;
                // covering-depth codes: c = S * log2(depth / nearDist), 253 code units up to 1.01 * errScale
                fa.coverLogS = 253.0f / log2f(1.01f * ctx->grid.d.errScale / fa.nearDist);
                fa.coverLogK = -log2f(fa.nearDist) * fa.coverLogS;
            }
            fa.cells4 = ctx->fanCells.as<uint4>();
            fa.cells = reinterpret_cast<uint2*>(fa.cells4 + nFans * kFanCells); fa.entries = ctx->fanEntries.as<uint16_t>();
            fa.firstA = reinterpret_cast<uint32_t*>(fa.cells + nFans * kFanCells);
            fa.capacity = (unsigned int)cap; fa.ctl = ctx->fanCtl.as<unsigned int>();
            fa.order = nullptr;
            if (nc <= 16384) {                       // the per-goal sort runs in one CTA's shared memory
                fa.order = ctx->fanOrder.as<uint32_t>();
                ctx->kernelLaunches++;
            }
            fan_build_set_scratch(fa, ctx->fanScratch.p);
            CK(launch_fan_build(fa, fanStream));
            if (fanBeside) CK(cudaEventRecord(ctx->evFan, ctx->stream2));
            ctx->kernelLaunches += 2;                // fan_project_kernel, fan_match_kernel
            fd.nFans = (int)nFans; fd.cells = fa.cells; fd.entries = fa.entries; fd.nEntries = (int)cap; fd.firstA = fa.firstA; fd.cells4 = fa.cells4; fd.coverLogS = fa.coverLogS; fd.coverLogK = fa.coverLogK;
            ctx->frameCoverLogS = fa.coverLogS; ctx->frameCoverLogK = fa.coverLogK;
            ctx->frameGridUsed |= 4u;
        }
    }
    ctx->frameFans = useFans;
    ctx->frameFanBeside = fanBeside;
    if (!fanBeside) CK(cudaEventRecord(ctx->evFan, ctx->stream));
    ctx->frameSplit = false;

    // Small frames (brute-force kernels, GPU far from full): run the permeation job on a second stream beside the trace
    // job, as the reference schedules them (ART:191, 213). Large frames stay serial so that each kernel is timed alone.
    const bool overlapJobs = !useGrid && wantRT && wantPM && map.nLocal <= 16384;
    ctx->frameOverlap = overlapJobs;
    if (overlapJobs) {
        CK(cudaEventRecord(ctx->evReady, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->evReady, 0));
    }
    cudaStream_t pmStream = overlapJobs ? ctx->stream2 : ctx->stream;

    // ---------------- K1 ----------------
    if (wantRT) {
        TraceArgs ta;
        ta.geom = ctx->geom.as<unsigned char>(); ta.L = L; ta.at = at;
        ta.dirs = ctx->dirs.as<uint16_t>(); ta.map = map;
        ta.ox = prm->rayOrigin[0]; ta.oy = prm->rayOrigin[1]; ta.oz = prm->rayOrigin[2];
        ta.targets = ctx->targets.as<float>(); ta.nTargets = Na;
        ta.targetOrder = reinterpret_cast<const int*>(ctx->targets.as<unsigned char>() + 12 * (size_t)Na);
        ta.maxRayLife = prm->maxRayLife; ta.H = H; ta.maxMuffle = prm->maxMuffleHitDistance;
        ta.batchSize = b;
        unsigned char* ob = ctx->outAll.as<unsigned char>();
        ta.echo = reinterpret_cast<uint16_t*>(ob + ctx->offEcho);
        ta.hitPoints = wantHitPts ? reinterpret_cast<uint16_t*>(ob + ctx->offHitPts) : nullptr;
        ta.hitCounts = wantHitCnt ? ob + ctx->offHitCnt : nullptr;
        ta.hitIds = wantHitIds ? reinterpret_cast<uint32_t*>(ob + ctx->offHitIds) : nullptr;
        ta.muffleCounts = reinterpret_cast<uint32_t*>(pb + bl.offMuffle);
        ta.counters = dh->counters;
        ta.nextRay = reinterpret_cast<unsigned int*>(ctx->partials.as<unsigned char>() + queueOff);
        bool geomInSmem = trace_smem_bytes(L, Na, true, false) <= (size_t)ctx->maxSmemOptin;
        bool muffleInSmem = trace_smem_bytes(L, Na, geomInSmem, true) <= (size_t)ctx->maxSmemOptin;
        ta.muffleInSmem = muffleInSmem ? 1 : 0;
        for (int sec = 0; sec < 3; sec++) {
            long long owned = 0;
            for (int t = 0; t < Na; t++) owned += ownedCount[(size_t)sec * Na + t];
            ta.anyOwned[sec] = owned > 0 ? 1 : 0;
        }
        ta.scratch = nullptr;
        ta.migGroups = 0; ta.migSlots = 0; ta.migFlags = nullptr; ta.migState = nullptr;
        trace_grid_plan(map.nLocal, Na, ctx->numSms, &ta.gridWarps, &ta.raysPerWarp);
        ta.recA = nullptr; ta.recB = nullptr; ta.recCount = nullptr;
        if (useGrid && useFans) {
            // Bounce-only tracer + one query kernel over all hit points (k1_query_fan.cu): RT:124-173 only writes
            // EchoRayDistances / MuffleRayHits, nothing there feeds the bounce loop.
            const bool stats = (prm->flags & ART_FRAME_GRID_STATS) != 0;
            const size_t recBytesA = (NH * sizeof(float4) + 255) & ~(size_t)255;
            CK(ctx->hitRecs.ensure(recBytesA + NH * sizeof(float2) + 16));
            CK(ctx->queryScratch.ensure(query_fan_scratch_bytes(ctx->numSms)));
            ta.recA = ctx->hitRecs.as<float4>();
            ta.recB = reinterpret_cast<float2*>(ctx->hitRecs.as<unsigned char>() + recBytesA);
            ta.recCount = ta.nextRay + 4;                        // (zeroed with the queue counters)
            const bool gInSmem = bounce_smem_bytes(L, true) <= (size_t)ctx->maxSmemOptin;
            CK(launch_bounce(ta, gd, ctx->numSms, gInSmem, stats, ctx->stream));
            CK(cudaEventRecord(ctx->evBounce, ctx->stream));
            if (fanBeside) CK(cudaStreamWaitEvent(ctx->stream, ctx->evFan, 0));      // the queries need the fans
            ctx->frameSplit = true;
            QueryArgs qa;
            qa.geom = ta.geom; qa.L = L; qa.recA = ta.recA; qa.recB = ta.recB; qa.recCount = ta.recCount;
            qa.map = map; qa.H = H; qa.batchSize = b;
            qa.ox = ta.ox; qa.oy = ta.oy; qa.oz = ta.oz; qa.targets = ta.targets; qa.nTargets = Na;
            qa.maxMuffle = ta.maxMuffle; qa.errScale = gd.errScale;
            qa.echo = ta.echo; qa.muffleCounts = ta.muffleCounts; qa.muffleRows = T; qa.counters = ta.counters;
            qa.queue = ta.nextRay + 5;
            qa.scratch = ctx->queryScratch.as<float4>();
            qa.tablesInSmem = 0; qa.muffleInSmem = 0;
            const bool qInSmem = query_fan_smem_bytes(L, true) <= (size_t)ctx->maxSmemOptin;
            CK(launch_query_fan(qa, fd, ctx->numSms, qInSmem, stats, ctx->maxSmemOptin, ctx->stream));
            ctx->kernelLaunches++;
            ctx->frameGridUsed |= 1u;
        } else if (useGrid) {
            CK(ctx->gridScratch.ensure(trace_grid_scratch_bytes(ctx->numSms)));
            ta.scratch = ctx->gridScratch.as<uint32_t>();
            // small batches (a shard of a ray-sharded frame): rotate the ray groups through the warps (k1_trace_grid.cu)
            if (!(prm->flags & ART_FRAME_GRID_STATS)) ta.migGroups = trace_grid_rotation(map.nLocal, H, ctx->numSms, ta.gridWarps, ctx->lastHitFill >= 0.9f, &ta.migSlots);
            {   // size the log for the case that a later frame of this batch size turns the rotation on (it depends on how long
                // the previous frame's rays lived): the allocation then happens in the context's first frame, not in the middle
                // of a running session
                unsigned int potSlots = 0;
                if (trace_grid_rotation(map.nLocal, H, ctx->numSms, ta.gridWarps, true, &potSlots) > 0) {
                    const size_t fb = ((size_t)potSlots * sizeof(unsigned int) + 255) & ~(size_t)255;
                    CK(ctx->rotateLog.ensure(fb + (size_t)potSlots * 64 * sizeof(float4)));
                }
            }
            if (ta.migGroups > 0) {
                const size_t flagBytes = ((size_t)ta.migSlots * sizeof(unsigned int) + 255) & ~(size_t)255;
                CK(ctx->rotateLog.ensure(flagBytes + (size_t)ta.migSlots * 64 * sizeof(float4)));
                ta.migFlags = ctx->rotateLog.as<unsigned int>();
                ta.migState = reinterpret_cast<float4*>(ctx->rotateLog.as<unsigned char>() + flagBytes);
                CK(cudaMemsetAsync(ta.migFlags, 0, flagBytes, ctx->stream));
                ctx->frameGridUsed |= 16u;
            }
            const bool gInSmem = trace_grid_smem_bytes(L, true) <= (size_t)ctx->maxSmemOptin;
            CK(launch_trace_grid(ta, gd, ctx->numSms, gInSmem, (prm->flags & ART_FRAME_GRID_STATS) != 0, ctx->stream));
            ctx->frameGridUsed |= 1u;
        } else {
            CK(launch_trace(ta, ctx->numSms, geomInSmem, count, ctx->stream));
        }
        ctx->kernelLaunches++;
    }
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    // per-ray outputs are final once K1 has run: copy them back on a second stream while K2 / K3 execute
    bool copiedEarly = false;
    if (hostOut && wantRT && outputs) {
        CK(cudaEventRecord(ctx->evTraceDone, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copyStream, ctx->evTraceDone, 0));
        // one copy covers every requested array (the echo halves come first and are skipped when not wanted)
        const size_t from = outputs->echoRayDistances ? 0 : ctx->offHitIds;
        if (ctx->outBytes > from) {
            CK(ctx->pinAll.ensure(ctx->outBytes));
            CK(cudaMemcpyAsync(ctx->pinAll.as<unsigned char>() + from, ctx->outAll.as<unsigned char>() + from, ctx->outBytes - from,
                               cudaMemcpyDeviceToHost, ctx->copyStream));
        }
        CK(cudaEventRecord(ctx->evCopyDone, ctx->copyStream));
        copiedEarly = true;
    }
    // ---------------- K2 ----------------
    bool joinPermLast = false;
    if (wantPM) {
        CK(ctx->firstHit.ensure(nLoc * 4 + 16));
        PermArgs pa;
        pa.geom = ctx->geom.as<unsigned char>(); pa.L = L; pa.at = at;
        const float* dn = ctx->perm.as<float>();
        pa.densS = dn; pa.densA = dn + L.nsPad; pa.densO = dn + L.nsPad + L.naPad;
        pa.trueDens = reinterpret_cast<const float*>(ctx->perm.as<unsigned char>() + nPad * 8);
        pa.ownedList = reinterpret_cast<const int*>(ctx->perm.as<unsigned char>() + nPad * 4);
        pa.nOwned = ctx->nOwned;
        pa.dirs = ctx->dirs.as<uint16_t>(); pa.map = map;
        pa.ox = prm->rayOrigin[0]; pa.oy = prm->rayOrigin[1]; pa.oz = prm->rayOrigin[2];
        pa.targets = ctx->targets.as<float>(); pa.nTargets = Na;
        pa.targetOrder = reinterpret_cast<const int*>(ctx->targets.as<unsigned char>() + 12 * (size_t)Na);
        pa.nTimesS = (float)N * prm->permeationStrengthPerRay;                     // PM:260
        pa.batchSize = b;
        pa.firstHitDist = ctx->firstHit.as<float>();
        pa.lastHitRay = reinterpret_cast<int*>(pb + bl.offLastHit);
        pa.permSumInt = reinterpret_cast<long long*>(pb + bl.offSumInt);
        pa.permSumFrac = reinterpret_cast<long long*>(pb + bl.offSumFrac);
        pa.permLast = reinterpret_cast<float*>(pb + bl.offPermLast);
        pa.counters = dh->counters;
        pa.nextRay = reinterpret_cast<unsigned int*>(ctx->partials.as<unsigned char>() + queueOff) + 8;
        pa.raysPerWarp = perm_grid_rays_per_warp(map.nLocal, Na, ctx->numSms);
        pa.hitPts = nullptr;
        // Binned loss lines (k2_permeation_binned.cu): with the target fans and enough (ray, target) lines to fill the GPU, the
        // lines are sorted by (target, direction bin) and evaluated 32 of one bin at a time. ART_K2_BINNED=0/1 forces it off/on.
        bool binned = useGrid && useFans && !(prm->flags & ART_FRAME_GRID_STATS) && (size_t)map.nLocal * Na >= ((size_t)1 << 21);
        if (const char* v = getenv("ART_K2_BINNED")) binned = useGrid && useFans && !(prm->flags & ART_FRAME_GRID_STATS) && atoi(v) != 0;
        if (binned && ((size_t)map.nLocal * Na > ((size_t)1 << 31) || Na > 65535)) binned = false;   // 32-bit offsets inside the pair list; one grid row per target
        if (useGrid && binned) {
            // 4 B per (ray, target) line: when that allocation fails the per-line kernel below runs instead (same results)
            if (ctx->permPairs.ensure((size_t)Na * nLoc * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); binned = false; }
        }
        if (useGrid && binned) {
            PermBinArgs ba;
            ba.slices = perm_binned_slices(map.nLocal, Na, ctx->numSms);
            CK(ctx->permHitPts.ensure((size_t)nLoc * sizeof(float4)));
            const size_t cntBytes = (perm_binned_cnt_bytes(Na, ba.slices) + 255) & ~(size_t)255;
            const size_t startBytes = (perm_binned_start_bytes(Na) + 255) & ~(size_t)255;
            CK(ctx->permBinCnt.ensure(cntBytes + startBytes + perm_binned_block_bytes(map.nLocal, Na)));
            CK(ctx->permPairs.ensure((size_t)Na * nLoc * sizeof(uint32_t)));
            pa.hitPts = ctx->permHitPts.as<float4>();
            pa.raysPerWarp = 32;
            ba.hitPts = pa.hitPts; ba.nLocal = map.nLocal; ba.targets = pa.targets; ba.nTargets = Na;
            ba.cnt = ctx->permBinCnt.as<uint32_t>(); ba.pairRay = ctx->permPairs.as<uint32_t>();
            ba.binStart = reinterpret_cast<uint32_t*>(ctx->permBinCnt.as<unsigned char>() + cntBytes);
            ba.blockBin = reinterpret_cast<uint32_t*>(ctx->permBinCnt.as<unsigned char>() + cntBytes + startBytes);
            ba.jobQueue = pa.nextRay + 1;
            const bool g1 = perm_grid_smem_bytes(L, true, false) <= (size_t)ctx->maxSmemOptin;
            CK(launch_permeation_grid(pa, gd, nullptr, ctx->numSms, g1, false, pmStream));          // first-hit distances + hit points
            // the canonical last-writer slots (PM:85, quirk Q5) need only the first hits: a few hundred warps of serial FP32
            // sums (0.1 ms whatever the batch size) run beside the loss lines on the second stream
            CK(cudaEventRecord(ctx->evReady, pmStream));
            CK(cudaStreamWaitEvent(ctx->stream2, ctx->evReady, 0));
            CK(launch_perm_last(pa, T, ctx->stream2));
            CK(cudaEventRecord(ctx->evP1, ctx->stream2));
            joinPermLast = true;
            const bool g2 = perm_binned_smem_bytes(L, true) <= (size_t)ctx->maxSmemOptin;
            CK(launch_permeation_binned(pa, ba, fd, ctx->numSms, g2, pmStream));
            ctx->frameGridUsed |= 2u | 32u;
            ctx->kernelLaunches += 5;
        } else if (useGrid) {
            const bool gInSmem = perm_grid_smem_bytes(L, true, useFans) <= (size_t)ctx->maxSmemOptin;
            CK(launch_permeation_grid(pa, gd, useFans ? &fd : nullptr, ctx->numSms, gInSmem, (prm->flags & ART_FRAME_GRID_STATS) != 0, pmStream));
            CK(launch_perm_last(pa, T, pmStream));
            ctx->frameGridUsed |= 2u;
        } else {
            const bool geomInSmem = perm_smem_bytes(L, true) <= (size_t)ctx->maxSmemOptin;
            if (overlapJobs) CK(cudaEventRecord(ctx->evP0, pmStream));
            CK(launch_permeation(pa, ctx->numSms, geomInSmem, T, pmStream));
            if (overlapJobs) CK(cudaEventRecord(ctx->evP1, pmStream));
        }
        ctx->kernelLaunches += 2;
    }
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    // ---------------- K3 ----------------
    if (wantRT) {
        CK(launch_echo_stats(reinterpret_cast<const uint16_t*>(ctx->outAll.as<unsigned char>() + ctx->offEcho), NH, &dh->echo, (prm->flags & ART_FRAME_REVERB_SEQ_FP32) != 0, ctx->numSms, ctx->stream));
        ctx->kernelLaunches += (prm->flags & ART_FRAME_REVERB_SEQ_FP32) ? 2 : 1;
    }
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (overlapJobs || joinPermLast) CK(cudaStreamWaitEvent(ctx->stream, ctx->evP1, 0));      // join the permeation job / its last-writer kernel
    CK(cudaEventRecord(ctx->ev[6], ctx->stream));
    // ---------------- multi-process frames: all-gather the ranks' partial blobs on the device ----------------
    ctx->frameComm = ctx->comm != nullptr && ctx->commWorld > 1 && !(prm->flags & ART_FRAME_PARTIALS_ONLY) && !ctx->rerunning;
    if (ctx->frameComm) {
        CK(ctx->gathered.ensure(bl.bytes * (size_t)ctx->commWorld));
        CK(ctx->pinGathered.ensure(bl.bytes * (size_t)ctx->commWorld));
        CK(cudaEventRecord(ctx->evX0, ctx->stream));
        const int nrc = nccl_api().AllGather(ctx->partials.p, ctx->gathered.p, bl.bytes, kNcclChar, ctx->comm, ctx->stream);
        if (nrc != 0) { ctx->poisoned = true; return fail(ctx, ART_E_CUDA, "ncclAllGather: %s", nccl_api().GetErrorString(nrc)); }
        CK(cudaEventRecord(ctx->evX1, ctx->stream));
        CK(cudaMemcpyAsync(ctx->pinGathered.p, ctx->gathered.p, bl.bytes * (size_t)ctx->commWorld, cudaMemcpyDeviceToHost, ctx->stream));
    }
    // ---------------- read back ----------------
    CK(cudaMemcpyAsync(ctx->pinPartials.p, ctx->partials.p, bl.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (useFans) CK(cudaMemcpyAsync(ctx->pinFanCtl.p, ctx->fanCtl.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->frameGridBuilt) CK(cudaMemcpyAsync(ctx->pinGridCtl.p, ctx->gridCtl.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (copiedEarly) CK(cudaStreamWaitEvent(ctx->stream, ctx->evCopyDone, 0));
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));

    ctx->params = *prm;
    ctx->params.audioTargetPositions = ctx->pinTargets.as<float>();   // library-owned copy
    ctx->haveUserOut = outputs != nullptr;
    if (outputs) ctx->userOut = *outputs; else memset(&ctx->userOut, 0, sizeof ctx->userOut);
    ctx->map = map; ctx->frameH = H; ctx->frameNa = Na; ctx->frameT = T;
    ctx->frameFlags = prm->flags; ctx->frameJobs = prm->jobs;
    ctx->frameCopiedEarly = copiedEarly;
    ctx->inFlight = true; ctx->frameDone = false;
    ctx->handle++;
    *outHandle = ctx->handle;
    // perm counters derived on the host (exact arithmetic functions of permHitRays)
    ctx->counters = ArtCounters{};
    ctx->lastBlob.clear();
    ctx->frameOwnedCount = ownedCount;
    return ART_OK;
}

// The fan lists of the frame in flight did not fit their buffer (or a list exceeded its length limit): its results are
// incomplete. Schedule it again on the grid walk (under the same handle) and give the next frame a larger buffer.
// Likewise when the device-built grid overflowed its entry buffer or a cell list outgrew the header format: second pass on
// the brute-force kernels, larger buffer for the next scene upload.
static bool grid_overflowed(const ArtCtx* ctx) { return ctx->frameGridBuilt && ctx->pinGridCtl.as<unsigned int>()[1] != 0; }
static bool needs_rerun(const ArtCtx* ctx)
{
    if (ctx->frameIsRerun) return false;
    return grid_overflowed(ctx) || (ctx->frameFans && ctx->pinFanCtl.as<unsigned int>()[1] != 0);
}
static int32_t start_rerun(ArtCtx* ctx)
{
    if (getenv("ART_DEBUG_LOG"))
        fprintf(stderr, "[audiort] frame re-run: grid ctl {%u, %u} (built %d), fan ctl {%u, %u} (fans %d)\n",
                ctx->pinGridCtl.p ? ctx->pinGridCtl.as<unsigned int>()[0] : 0u, ctx->pinGridCtl.p ? ctx->pinGridCtl.as<unsigned int>()[1] : 0u,
                (int)ctx->frameGridBuilt, ctx->pinFanCtl.p ? ctx->pinFanCtl.as<unsigned int>()[0] : 0u,
                ctx->pinFanCtl.p ? ctx->pinFanCtl.as<unsigned int>()[1] : 0u, (int)ctx->frameFans);
    if (grid_overflowed(ctx)) {
        if (ctx->gridEntriesPerCollider < 16384) ctx->gridEntriesPerCollider = std::max<size_t>(64, ctx->gridEntriesPerCollider * 4);
        ctx->rerunNoGrid = true;
        ctx->gridBuilt = false;                        // the next frame builds the lists again, into the larger buffer
    } else if (ctx->fanEntriesPerPair < 4096) ctx->fanEntriesPerPair *= 4;
    ArtParams prm = ctx->params;
    ArtOutputs uo = ctx->userOut;
    const bool hadOut = ctx->haveUserOut;
    const ArtHandle keep = ctx->handle;
    ctx->inFlight = false;
    ctx->rerunning = true;
    ArtHandle h2 = 0;
    const int32_t rc = art_trace_schedule(ctx, &prm, hadOut ? &uo : nullptr, &h2);
    ctx->rerunning = false;
    ctx->rerunNoGrid = false;
    ctx->handle = keep;
    ctx->frameIsRerun = true;
    return rc;
}

ART_API int32_t art_is_completed(ArtCtx* ctx, ArtHandle h)
{
    if (!ctx) return ART_E_ARG;
    if (h != ctx->handle || h == 0) return fail(ctx, ART_E_STATE, "stale handle");
    if (!ctx->inFlight) return 1;
    if (!ctx->children.empty()) return multi_is_completed(ctx);
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaEventQuery(ctx->ev[5]);
    if (e == cudaErrorNotReady) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) {
        ctx->poisoned = true;
        return fail(ctx, ART_E_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
    }
    if (needs_rerun(ctx)) {
        // keep polling callers (ART:95) from ever blocking in art_complete: the second pass starts here, asynchronously
        const int32_t rc = start_rerun(ctx);
        return rc != ART_OK ? rc : 0;
    }
    return 1;
}

ART_API int32_t art_complete(ArtCtx* ctx, ArtHandle h)
{
    if (!ctx) return ART_E_ARG;
    if (h != ctx->handle || h == 0) return fail(ctx, ART_E_STATE, "stale handle");
    if (!ctx->inFlight) return ctx->frameDone ? ART_OK : fail(ctx, ART_E_STATE, "no frame scheduled");
    if (!ctx->children.empty()) return multi_complete(ctx);
    cudaSetDevice(ctx->device);
    // per-ray outputs reach pinned memory right after the trace job: hand them to the caller's arrays while the
    // permeation job and the reduction are still running
    bool perRayCopied = false;
    auto copy_per_ray = [&]() {
        const ArtOutputs& o = ctx->userOut;
        const size_t nl = (size_t)ctx->map.nLocal, H = (size_t)ctx->frameH, nh = nl * H;
        const unsigned char* pb = ctx->pinAll.as<unsigned char>();
        if (!ctx->scatterGlobal || ctx->map.shardCount <= 1) {
            if (o.echoRayDistances) big_memcpy(o.echoRayDistances, pb + ctx->offEcho, nh * 2);
            if (o.rayHitResults) big_memcpy(o.rayHitResults, pb + ctx->offHitPts, nh * 6);
            if (o.rayHitResultCounts) big_memcpy(o.rayHitResultCounts, pb + ctx->offHitCnt, nl);
            if (o.hitColliderIds) big_memcpy(o.hitColliderIds, pb + ctx->offHitIds, nh * 4);
            return;
        }
        // child of a multi-device context: the caller's arrays are indexed by GLOBAL ray, this context owns every
        // shardCount-th chunk of them
        const size_t chunk = (size_t)ctx->map.chunk;
        for (size_t j0 = 0; j0 < nl; j0 += chunk) {
            const size_t cnt = std::min(chunk, nl - j0), g0 = (size_t)ctx->map.to_global((int)j0);
            if (o.echoRayDistances) memcpy(o.echoRayDistances + g0 * H, pb + ctx->offEcho + j0 * H * 2, cnt * H * 2);
            if (o.rayHitResults) memcpy(o.rayHitResults + g0 * H * 3, pb + ctx->offHitPts + j0 * H * 6, cnt * H * 6);
            if (o.rayHitResultCounts) memcpy(o.rayHitResultCounts + g0, pb + ctx->offHitCnt + j0, cnt);
            if (o.hitColliderIds) memcpy(o.hitColliderIds + g0 * H, pb + ctx->offHitIds + j0 * H * 4, cnt * H * 4);
        }
    };
    if (ctx->frameCopiedEarly && ctx->haveUserOut && cudaEventSynchronize(ctx->evCopyDone) == cudaSuccess) {
        copy_per_ray();
        perRayCopied = true;
    }
    cudaError_t e = cudaEventSynchronize(ctx->ev[5]);
    if (e != cudaSuccess) {
        ctx->poisoned = true; ctx->inFlight = false;
        return fail(ctx, ART_E_CUDA, "frame failed: %s", cudaGetErrorString(e));
    }
    if (needs_rerun(ctx)) {
        const int32_t rc = start_rerun(ctx);
        if (rc != ART_OK) { ctx->inFlight = false; return rc; }
        e = cudaEventSynchronize(ctx->ev[5]);
        if (e != cudaSuccess) {
            ctx->poisoned = true; ctx->inFlight = false;
            return fail(ctx, ART_E_CUDA, "frame failed: %s", cudaGetErrorString(e));
        }
        perRayCopied = false;
    } else if (ctx->frameIsRerun) {
        perRayCopied = false;                      // (the second pass was started by art_is_completed)
    }
    ctx->inFlight = false;
    const bool fanOverflow = ctx->frameIsRerun;
    const int Na = ctx->frameNa, T = ctx->frameT;
    const BlobLayout bl = blob_layout(Na, T);
    BlobHeader* hh = ctx->pinPartials.as<BlobHeader>();
    hh->magic = kBlobMagic; hh->nTargets = Na; hh->batchCount = T; hh->shards = 1;

    // counters + timings
    ArtCounters& c = ctx->counters;
    const unsigned long long* dc = hh->counters;
    c.segments = dc[C_SEGMENTS]; c.segmentHits = dc[C_SEGMENT_HITS];
    for (int k = 0; k < 3; k++) { c.traceTests[k] = dc[C_TRACE_S + k]; c.echoTests[k] = dc[C_ECHO_S + k]; c.muffleTests[k] = dc[C_MUFFLE_S + k]; }
    c.echoQueries = dc[C_ECHO_Q]; c.muffleQueries = dc[C_MUFFLE_Q];
    c.permRays = dc[C_PERM_RAYS]; c.permHitRays = dc[C_PERM_HIT_RAYS];
    for (int k = 0; k < 3; k++) {
        c.gridTraceTests[k] = dc[C_GRID_RT_S + k]; c.gridPermFirstTests[k] = dc[C_GRID_PF_S + k]; c.gridPermLossTests[k] = dc[C_GRID_PL_S + k];
    }
    c.gridTraceCells = dc[C_GRID_RT_CELLS] + dc[C_GRID_Q_LISTS]; c.gridPermCells = dc[C_GRID_PM_CELLS];
    for (int k = 0; k < 3; k++) { c.gridQueryTests[k] = dc[C_GRID_Q_S + k]; c.gridTraceTests[k] += c.gridQueryTests[k]; }
    c.gridQueryLists = dc[C_GRID_Q_LISTS];
    c.debugViolations = dc[C_DEBUG_VIOLATIONS];
    {
        const int* ownedCount = ctx->frameOwnedCount.data();
        const uint64_t nSec[3] = { (uint64_t)ctx->L.ns, (uint64_t)ctx->L.na, (uint64_t)ctx->L.no };
        c.permPairs = c.permHitRays * (uint64_t)Na;
        for (int s = 0; s < 3; s++) {
            c.permFirstTests[s] = c.permRays * nSec[s];
            uint64_t owned = 0;
            for (int a = 0; a < Na; a++) owned += (uint64_t)ownedCount[(size_t)s * Na + a];
            c.permLossTests[s] = c.permHitRays * (nSec[s] * (uint64_t)Na - owned);
        }
        for (int k = 0; k < 3; k++) {   // mirror into the blob so merged blobs carry them
            hh->counters[C_PERM_FIRST_S + k] = c.permFirstTests[k];
            hh->counters[C_PERM_LOSS_S + k] = c.permLossTests[k];
        }
        hh->counters[C_PERM_PAIRS] = c.permPairs;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); c.h2dMs = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); c.traceMs = ms;
    cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); c.permeationMs = ms;
    cudaEventElapsedTime(&ms, ctx->ev[3], ctx->ev[4]); c.reduceMs = ms;
    cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[6]); c.deviceMs = ms;
    if (ctx->frameOverlap) { cudaEventElapsedTime(&ms, ctx->evP0, ctx->evP1); c.permeationMs = ms; }
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); c.d2hMs = ms;
    c.fanBuildMs = c.bounceMs = c.queryMs = c.exchangeMs = 0.0f;
    if (ctx->frameJobs & ART_JOB_RAYTRACE) {
        cudaEventElapsedTime(&ms, ctx->frameFanBeside ? ctx->evF0 : ctx->ev[1], ctx->evFan); c.fanBuildMs = ms;
        if (ctx->frameSplit) {
            // (a fan build beside the bounce tracer: the two intervals overlap; queryMs starts when both have finished)
            cudaEventElapsedTime(&ms, ctx->frameFanBeside ? ctx->ev[1] : ctx->evFan, ctx->evBounce); c.bounceMs = ms;
            cudaEventElapsedTime(&ms, ctx->evBounce, ctx->ev[2]); c.queryMs = ms;
        } else {
            cudaEventElapsedTime(&ms, ctx->evFan, ctx->ev[2]); c.bounceMs = ms;   // bounce rays and queries in one kernel
        }
    }
    c.kernelLaunches = ctx->kernelLaunches;
    c.gridUsed = ctx->frameGridUsed | (fanOverflow ? 8u : 0u);
    if (ctx->frameJobs & ART_JOB_RAYTRACE) {
        const double slotsTotal = (double)ctx->map.nLocal * (double)ctx->frameH;
        ctx->lastHitFill = slotsTotal > 0 ? (float)((double)c.segmentHits / slotsTotal) : 0.0f;
    }
    if ((ctx->frameGridUsed & 16u) && c.debugViolations != 0) {
        ctx->frameDone = false;
        return fail(ctx, ART_E_CUDA, "trace job: a group-rotation wait timed out (%llu); outputs are incomplete", (unsigned long long)c.debugViolations);
    }

    // per-ray outputs: pinned staging -> caller arrays
    const bool hostOut = !(ctx->frameFlags & ART_FRAME_NO_HOST_OUTPUTS);
    if (hostOut && (ctx->frameJobs & ART_JOB_RAYTRACE) && ctx->haveUserOut && !perRayCopied) copy_per_ray();
    ctx->lastBlob.assign(ctx->pinPartials.as<unsigned char>(), ctx->pinPartials.as<unsigned char>() + bl.bytes);
    c.devicesUsed = 1;
    if (ctx->frameComm) {
        // multi-process frame: every rank holds all ranks' blobs (gathered on the device); merge them exactly. The header
        // fields and the derived permeation counters are functions of each blob's own device counters and the shared scene.
        cudaEventElapsedTime(&ms, ctx->evX0, ctx->evX1); c.exchangeMs = ms;
        unsigned char* all = ctx->pinGathered.as<unsigned char>();
        for (int r = 0; r < ctx->commWorld; r++) {
            BlobHeader* hr = reinterpret_cast<BlobHeader*>(all + (size_t)r * bl.bytes);
            hr->magic = kBlobMagic; hr->nTargets = Na; hr->batchCount = T; hr->shards = 1;
            const uint64_t rays = hr->counters[C_PERM_RAYS], hit = hr->counters[C_PERM_HIT_RAYS];
            const uint64_t nSec[3] = { (uint64_t)ctx->L.ns, (uint64_t)ctx->L.na, (uint64_t)ctx->L.no };
            for (int sct = 0; sct < 3; sct++) {
                uint64_t owned = 0;
                for (int a = 0; a < Na; a++) owned += (uint64_t)ctx->frameOwnedCount[(size_t)sct * Na + a];
                hr->counters[C_PERM_FIRST_S + sct] = rays * nSec[sct];
                hr->counters[C_PERM_LOSS_S + sct] = hit * (nSec[sct] * (uint64_t)Na - owned);
            }
            hr->counters[C_PERM_PAIRS] = hit * (uint64_t)Na;
            if (r == 0) ctx->lastBlob.assign(all, all + bl.bytes);
            else if (art_partials_merge(ctx->lastBlob.data(), all + (size_t)r * bl.bytes, (int64_t)bl.bytes) != ART_OK)
                return fail(ctx, ART_E_ARG, "merging the partial results of rank %d failed", r);
        }
        c.devicesUsed = (uint32_t)ctx->commWorld;
    }
    ctx->frameDone = true;
    if (!(ctx->frameFlags & ART_FRAME_PARTIALS_ONLY) && ctx->haveUserOut) {
        std::string err;
        int32_t rc = finalize_blob(ctx->lastBlob.data(), ctx->lastBlob.size(), &ctx->params, ctx->nGlobal, &ctx->userOut, &err);
        if (rc != ART_OK) return fail(ctx, rc, "finalize: %s", err.c_str());
    }
    return ART_OK;
}

ART_API int32_t art_get_counters(ArtCtx* ctx, ArtHandle h, ArtCounters* out)
{
    if (!ctx || !out) return ART_E_ARG;
    if (h != ctx->handle || h == 0) return fail(ctx, ART_E_STATE, "stale handle");
    if (ctx->inFlight || !ctx->frameDone) return fail(ctx, ART_E_PENDING, "frame not complete");
    *out = ctx->counters;
    return ART_OK;
}

ART_API int64_t art_partials_size(int32_t totalAudioTargets, int32_t batchCount)
{
    if (totalAudioTargets <= 0 || batchCount <= 0) return ART_E_ARG;
    return (int64_t)blob_layout(totalAudioTargets, batchCount).bytes;
}

ART_API int32_t art_get_partials(ArtCtx* ctx, ArtHandle h, void* blob, int64_t blobBytes)
{
    if (!ctx || !blob) return ART_E_ARG;
    if (h != ctx->handle || h == 0) return fail(ctx, ART_E_STATE, "stale handle");
    if (ctx->inFlight || !ctx->frameDone) return fail(ctx, ART_E_PENDING, "frame not complete");
    if ((size_t)blobBytes < ctx->lastBlob.size()) return fail(ctx, ART_E_ARG, "blob buffer too small");
    memcpy(blob, ctx->lastBlob.data(), ctx->lastBlob.size());
    return ART_OK;
}

ART_API int32_t art_partials_merge(void* accumBlob, const void* otherBlob, int64_t blobBytes)
{
    if (!accumBlob || !otherBlob || (size_t)blobBytes < sizeof(BlobHeader)) return ART_E_ARG;
    unsigned char* A = static_cast<unsigned char*>(accumBlob);
    const unsigned char* B = static_cast<const unsigned char*>(otherBlob);
    BlobHeader ha, hb; memcpy(&ha, A, sizeof ha); memcpy(&hb, B, sizeof hb);
    if (ha.magic != kBlobMagic || hb.magic != kBlobMagic || ha.nTargets != hb.nTargets || ha.batchCount != hb.batchCount) return ART_E_ARG;
    const int Na = ha.nTargets, T = ha.batchCount;
    const BlobLayout bl = blob_layout(Na, T);
    if ((size_t)blobBytes < bl.bytes) return ART_E_ARG;
    ha.shards += hb.shards;
    ha.echo.fixedLo += hb.echo.fixedLo; ha.echo.fixedHi += hb.echo.fixedHi;
    ha.echo.zeros += hb.echo.zeros; ha.echo.entries += hb.echo.entries;
    ha.echo.posInf += hb.echo.posInf; ha.echo.negInf += hb.echo.negInf; ha.echo.nan += hb.echo.nan;
    ha.echo.seqValid = 0;   // a sequential FP32 sum cannot be merged
    for (int i = 0; i < C_COUNT; i++) ha.counters[i] += hb.counters[i];
    memcpy(A, &ha, sizeof ha);
    int32_t* la = reinterpret_cast<int32_t*>(A + bl.offLastHit);
    const int32_t* lb = reinterpret_cast<const int32_t*>(B + bl.offLastHit);
    float* pa = reinterpret_cast<float*>(A + bl.offPermLast);
    const float* pbv = reinterpret_cast<const float*>(B + bl.offPermLast);
    for (int k = 0; k < T; k++)
        if (lb[k] > la[k]) { la[k] = lb[k]; memcpy(pa + (size_t)k * Na, pbv + (size_t)k * Na, sizeof(float) * (size_t)Na); }
    uint32_t* ma = reinterpret_cast<uint32_t*>(A + bl.offMuffle);
    const uint32_t* mb = reinterpret_cast<const uint32_t*>(B + bl.offMuffle);
    for (size_t i = 0; i < (size_t)T * Na; i++) ma[i] += mb[i];
    long long* sa = reinterpret_cast<long long*>(A + bl.offSumInt);
    const long long* sb = reinterpret_cast<const long long*>(B + bl.offSumInt);
    for (size_t i = 0; i < 2 * (size_t)Na; i++) sa[i] += sb[i];   // int + frac arrays are adjacent
    return ART_OK;
}

ART_API int32_t art_finalize(const void* blob, int64_t blobBytes, const ArtParams* params, int32_t rayCount, const ArtOutputs* outputs)
{
    if (!blob || !params || !outputs || rayCount <= 0) return ART_E_ARG;
    // the same range checks art_trace_schedule applies (PA:34-35 divide by these; RT:115 / ATM:112 index with their products)
    if (params->totalAudioTargets < 1 || params->totalAudioTargets > 32767 || params->batchCount < 1 || params->maxHitsPerRay < 1) return ART_E_ARG;
    if ((long long)rayCount * params->maxHitsPerRay > 0x7FFFFFFFLL || (long long)params->batchCount * params->totalAudioTargets > 0x7FFFFFFFLL) return ART_E_ARG;
    return finalize_blob(static_cast<const unsigned char*>(blob), (size_t)blobBytes, params, rayCount, outputs, nullptr);
}

ART_API int32_t art_comm_unique_id(void* uniqueId128)
{
    if (!uniqueId128) return ART_E_ARG;
    NcclApi& n = nccl_api();
    if (!n.ok()) return fail(nullptr, ART_E_STATE, "art_comm_unique_id: libnccl.so.2 could not be loaded (set ART_NCCL_LIB)");
    NcclUniqueId id;
    const int rc = n.GetUniqueId(&id);
    if (rc != 0) return fail(nullptr, ART_E_CUDA, "ncclGetUniqueId: %s", n.GetErrorString(rc));
    memcpy(uniqueId128, &id, sizeof id);
    return ART_OK;
}

ART_API int32_t art_comm_init(ArtCtx* ctx, const void* uniqueId128, int32_t rank, int32_t world, int32_t chunkRays)
{
    if (!ctx || !uniqueId128) return ART_E_ARG;
    if (!ctx->children.empty() || ctx->scatterGlobal) return fail(ctx, ART_E_STATE, "art_comm_init: not available on a multi-device context");
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "art_comm_init: a frame is in flight");
    if (world < 1 || rank < 0 || rank >= world || chunkRays < 0) return fail(ctx, ART_E_ARG, "art_comm_init: bad rank %d / world %d / chunk %d", rank, world, chunkRays);
    if (ctx->comm) return fail(ctx, ART_E_STATE, "art_comm_init: the context already has a communicator");
    NcclApi& n = nccl_api();
    if (!n.ok()) return fail(ctx, ART_E_STATE, "art_comm_init: libnccl.so.2 could not be loaded (set ART_NCCL_LIB)");
    cudaSetDevice(ctx->device);
    NcclUniqueId id;
    memcpy(&id, uniqueId128, sizeof id);
    void* comm = nullptr;
    const int rc = n.CommInitRank(&comm, world, id, rank);
    if (rc != 0) return fail(ctx, ART_E_CUDA, "ncclCommInitRank: %s", n.GetErrorString(rc));
    ctx->comm = comm; ctx->commRank = rank; ctx->commWorld = world;
    if (ctx->rayHostValid) ctx->raysDirty = true;      // the device holds only the shard's directions
    ctx->shardIndex = rank; ctx->shardCount = world; ctx->chunkRays = chunkRays; ctx->chunkAuto = chunkRays == 0;
    return ART_OK;
}

ART_API int32_t art_grid_build_host(const ArtAABB* aabbs, int32_t nAABB, const ArtOBB* obbs, int32_t nOBB,
                                    const ArtSphere* spheres, int32_t nSphere, float cellScale, ArtGridInfo* info,
                                    uint32_t* cells, int64_t cellsCapacity, uint16_t* entries, int64_t entriesCapacity)
{
    if (!info || nAABB < 0 || nOBB < 0 || nSphere < 0 || (nAABB && !aabbs) || (nOBB && !obbs) || (nSphere && !spheres)) return ART_E_ARG;
    std::vector<uint16_t> hS(reinterpret_cast<const uint16_t*>(spheres), reinterpret_cast<const uint16_t*>(spheres) + 8 * (size_t)nSphere);
    std::vector<uint16_t> hA(reinterpret_cast<const uint16_t*>(aabbs), reinterpret_cast<const uint16_t*>(aabbs) + 10 * (size_t)nAABB);
    std::vector<uint16_t> hO(reinterpret_cast<const uint16_t*>(obbs), reinterpret_cast<const uint16_t*>(obbs) + 13 * (size_t)nOBB);
    HostGrid g;
    build_grid(hS, hA, hO, cellScale > 0.0f ? cellScale : 1.1f, g);
    memset(info, 0, sizeof *info);
    if (!g.ok) return ART_E_STATE;
    info->nx = g.d.nx; info->ny = g.d.ny; info->nz = g.d.nz;
    info->g0[0] = g.d.g0x; info->g0[1] = g.d.g0y; info->g0[2] = g.d.g0z;
    info->g1[0] = g.d.g1x; info->g1[1] = g.d.g1y; info->g1[2] = g.d.g1z;
    info->cellSize[0] = g.d.csx; info->cellSize[1] = g.d.csy; info->cellSize[2] = g.d.csz;
    info->margin = g.margin;
    info->nCells = (int64_t)g.cells.size();
    info->nEntries = (int64_t)g.entries.size();
    if (cells) {
        if (cellsCapacity < 2 * info->nCells) return ART_E_ARG;
        for (size_t i = 0; i < g.cells.size(); i++) { cells[2 * i] = g.cells[i].x; cells[2 * i + 1] = g.cells[i].y; }
    }
    if (entries) {
        if (entriesCapacity < info->nEntries) return ART_E_ARG;
        memcpy(entries, g.entries.data(), g.entries.size() * sizeof(uint16_t));
    }
    return ART_OK;
}

ART_API int32_t art_debug_get_fans(ArtCtx* ctx, ArtFanInfo* info, uint32_t* cells, int64_t cellsCapacity,
                                   uint16_t* entries, int64_t entriesCapacity)
{
    if (!ctx || !info) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "a frame is in flight");
    if (!ctx->children.empty()) return art_debug_get_fans(ctx->children[0], info, cells, cellsCapacity, entries, entriesCapacity);
    if (!ctx->frameDone || !ctx->frameFans) return fail(ctx, ART_E_STATE, "the last frame did not use the target fans");
    cudaSetDevice(ctx->device);
    memset(info, 0, sizeof *info);
    info->nFans = ctx->frameNa + 1; info->binsPerFace = kFanBins; info->cellsPerFan = kFanCells;
    info->nearDist = 1e-3f * ctx->grid.d.errScale;
    info->nCells = (int64_t)info->nFans * kFanCells;
    info->nEntries = (int64_t)ctx->pinFanCtl.as<unsigned int>()[0];
    if (cells) {
        if (cellsCapacity < 2 * info->nCells) return fail(ctx, ART_E_ARG, "cells buffer too small");
        // (the buffer holds FanDesc::cells4, then cells, then firstA)
        CK(cudaMemcpy(cells, ctx->fanCells.as<unsigned char>() + (size_t)info->nCells * sizeof(uint4), (size_t)info->nCells * sizeof(uint2), cudaMemcpyDeviceToHost));
    }
    if (entries) {
        if (entriesCapacity < info->nEntries) return fail(ctx, ART_E_ARG, "entries buffer too small");
        CK(cudaMemcpy(entries, ctx->fanEntries.p, (size_t)info->nEntries * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    }
    return ART_OK;
}

ART_API int32_t art_debug_get_fan_cover(ArtCtx* ctx, uint32_t* codes, int64_t capacity, float* logS, float* logK)
{
    if (!ctx || !codes || !logS || !logK) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "a frame is in flight");
    if (!ctx->children.empty()) return art_debug_get_fan_cover(ctx->children[0], codes, capacity, logS, logK);
    if (!ctx->frameDone || !ctx->frameFans) return fail(ctx, ART_E_STATE, "the last frame did not use the target fans");
    cudaSetDevice(ctx->device);
    const int64_t nCells = (int64_t)(ctx->frameNa + 1) * kFanCells;
    if (capacity < nCells) return fail(ctx, ART_E_ARG, "codes buffer too small");
    // cells4[i].w: one strided copy out of the 16-byte cells
    CK(cudaMemcpy2D(codes, sizeof(uint32_t), ctx->fanCells.as<unsigned char>() + 12, sizeof(uint4), sizeof(uint32_t), (size_t)nCells, cudaMemcpyDeviceToHost));
    *logS = ctx->frameCoverLogS; *logK = ctx->frameCoverLogK;
    return ART_OK;
}

ART_API int32_t art_microbench(ArtCtx* ctx, int32_t kind, double* gops)
{
    if (!ctx || !gops || kind < 0 || kind > 2) return ART_E_ARG;
    if (ctx->inFlight) return fail(ctx, ART_E_PENDING, "a frame is in flight");
    if (!ctx->children.empty()) return art_microbench(ctx->children[0], kind, gops);
    cudaSetDevice(ctx->device);
    CK(ctx->queue.ensure(64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    long long laneOps = 0;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0, ctx->stream));
        CK(launch_microbench(kind, ctx->numSms, ctx->queue.as<float>() + 4, &laneOps, ctx->stream));
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *gops = (double)laneOps / (best * 1e-3) / 1e9;
    return ART_OK;
}

}  // extern "C"
