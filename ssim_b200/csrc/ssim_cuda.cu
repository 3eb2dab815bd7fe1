// ssim_cuda.cu -- runtime behind the C ABI of include/ssim_cuda.h.
//
// Host-side responsibilities (everything the reference's compute_ssim() does around its tile loop,
// src/ssim.cpp:956-1103, re-thought for a GPU):
//   * per-device context: streams, grow-only device scratch (canonical A/B planes, dense map, partial sums),
//     pinned staging, cached kernel geometry
//   * bringing ANY (step, stride, host|device) image into the canonical layout the fused kernel wants
//     (1 byte/pixel, 16-byte aligned base and pitch => TMA-legal): cudaMemcpy2DAsync for the common host case,
//     footprint copy + pack kernel for interleaved / negative-stride / column-major layouts
//   * work decomposition of (frames x rows x 64-column bands) into warp items sized to the 148-SM machine
//   * TMA descriptors (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda needed)
//   * getting the map back out in the caller's layout, and the final float(sum / double(W*H))
// There is deliberately no CPU fallback: without a device every compute entry point returns ENODEV.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "ssim_cuda.h"
#include "ssim_kernels.h"

namespace {

// ------------------------------------------------------------------------------------------------ errors
thread_local char g_err[512] = "";
thread_local int  g_lastLaunches = 0;
std::atomic<int> g_minSlotRows{0};       // tuning knobs (0 = automatic), see ssim_cuda_set_tuning()
std::atomic<int> g_maxPairsPerSm{0};
std::atomic<unsigned long long*> g_dbgTimes{nullptr};   // development aid, see ssim_cuda_debug_slot_times()

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (getenv("SSIM_CUDA_VERBOSE")) fprintf(stderr, "ssim_cuda: %s\n", g_err);
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    if (e == cudaErrorMemoryAllocation) return fail(ENOMEM, "%s: %s", what, cudaGetErrorString(e));
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice)
        return fail(ENODEV, "%s: %s", what, cudaGetErrorString(e));
    return fail(EIO, "%s: %s", what, cudaGetErrorString(e));
}

#define CU_TRY(expr)                                                   \
    do {                                                               \
        cudaError_t e__ = (expr);                                      \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr);          \
    } while (0)

// Every entry point that needs `device` current makes it so through this guard and puts the caller's device back on
// return: a host application (or a PyTorch thread) whose own CUDA work runs on another device is not switched silently.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != device) { err = cudaSetDevice(device); switched = (err == cudaSuccess); }
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_GUARD(dev)                                              \
    DeviceGuard guard__(dev);                                          \
    if (guard__.err != cudaSuccess) return cuda_fail(guard__.err, "cudaSetDevice")

// ------------------------------------------------------------------------------------------------ Gaussian taps
// The window every shipped build of the reference actually applies is its FLOAT 11x11 kernel: the generic blur
// computes it at run time (src/ssim.cpp:272-318, Float = float) and all SIMD blurs carry the same values as a
// literal table, even in the double build (src/ssim_fma.cpp:164-175).  Each of those 121 taps is rounded
// separately, so the window sums to 1 + 1.02e-8 instead of 1; that bias moves every variance by -1e-8 mu^2
// and the global SSIM of the double build by up to 1.7e-6 (tests/golden: ref_f64_auto vs ref_f64_generic).
// To reproduce the reference rather than the ideal Gaussian we
//   1. rebuild that float window exactly as the reference does,
//   2. take its best separable (rank-1, symmetric) approximation g g^T by power iteration in double,
//   3. round g to float and nudge the three outermost taps by a few ulps so that (sum g)^2 equals the window's sum.
// In double arithmetic the resulting separable filter is within 3e-8 of the reference's double build.
void gaussian_taps(float g[6])
{
    const int R = 5, N = 11;
    const float sigma = 1.5f, sigma2 = sigma * sigma;
    float kf[N * N];
    double sum = 0.0;
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const float num = expf(-(float)((x - R) * (x - R) + (y - R) * (y - R)) / (2 * sigma2));
            kf[y * N + x] = num / ((float)(2 * M_PI) * sigma2);
            sum += (double)kf[y * N + x];
        }
    double T[N * N], total = 0.0;
    for (int i = 0; i < N * N; ++i) { T[i] = (double)(kf[i] / (float)sum); total += T[i]; }

    double v[N], t[N];
    for (int i = 0; i < N; ++i) v[i] = 1.0 / std::sqrt((double)N);
    double lambda = 0.0;
    for (int it = 0; it < 200; ++it) {
        double norm = 0.0;
        for (int i = 0; i < N; ++i) { t[i] = 0.0; for (int j = 0; j < N; ++j) t[i] += T[i * N + j] * v[j]; norm += t[i] * t[i]; }
        norm = std::sqrt(norm);
        lambda = 0.0;
        for (int i = 0; i < N; ++i) { lambda += v[i] * t[i]; v[i] = t[i] / norm; }
    }
    for (int d = 0; d <= R; ++d) g[d] = (float)(std::sqrt(lambda) * 0.5 * (v[R - d] + v[R + d]));

    const double target = std::sqrt(total);
    for (int d = 3; d <= R; ++d) {
        double s = (double)g[0];
        for (int k = 1; k <= R; ++k) s += 2.0 * (double)g[k];
        const double ulp = (double)(std::nextafterf(g[d], 1.0f) - g[d]);
        g[d] = (float)((double)g[d] + std::nearbyint((target - s) / (2.0 * ulp)) * ulp);
    }
}

// ------------------------------------------------------------------------------------------------ TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

CUtensorMapL2promotion l2_promotion()
{
    static const CUtensorMapL2promotion promo = [] {
        // 64 bytes: a box row (96 bytes starting 16 bytes left of a 64-byte aligned band) then pulls exactly the 64-byte granules
        // it touches.  Measured on 64 x 4K (profiles/r02_l2_promotion.txt): DRAM reads 1.39 GB with 64 B, 1.71 GB with 128 B
        // (the 16-byte margins at a team's edges each cost a whole 128-byte line) and with NONE, 1.85 GB with 256 B;
        // algorithmic 1.06 GB; same kernel time in all four.  SSIM_CUDA_L2_PROMOTION = 0 / 64 / 128 / 256 overrides it.
        const char* e = getenv("SSIM_CUDA_L2_PROMOTION");
        const int v = e ? atoi(e) : 64;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    return promo;
}

int make_plane_map(CUtensorMap* tm, const uint8_t* base, uint32_t width, uint32_t rows, uint32_t frames, size_t pitch,
                   size_t frameStride, int elemBytes)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(EIO, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3]    = {width, rows, frames};
    const cuuint64_t strides[2] = {pitch, frames > 1 ? frameStride : (cuuint64_t)pitch * rows};
    const cuuint32_t box[3]     = {(cuuint32_t)(elemBytes == 2 ? ssimk::PixGeo<true>::kBoxElems : ssimk::PixGeo<false>::kBoxElems), (cuuint32_t)ssimk::kLoadRows, 1};
    const cuuint32_t estr[3]    = {1, 1, 1};
    CUresult r = enc(tm, elemBytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2_promotion(),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(EIO, "cuTensorMapEncodeTiled failed (%d) for %ux%ux%u pitch %zu", (int)r, width, rows, frames, pitch);
    return 0;
}

// Descriptor cache: a video sweep or a pipelined host call presents the same (pointer, geometry) again and again (frame
// rings, the context's own scratch planes), and encoding a tensor map costs a driver call of a few microseconds -- as much
// as the kernel launch itself.  Per thread (no lock), 512 entries, direct-mapped.  A descriptor only holds the
// address and the geometry, so a hit on a pointer that was freed and re-allocated with the same geometry is still right.
struct MapKey {
    const uint8_t* base; size_t pitch, frameStride; uint32_t width, rows, frames; int elemBytes;
    bool operator==(const MapKey& o) const
    {
        return base == o.base && pitch == o.pitch && frameStride == o.frameStride && width == o.width && rows == o.rows &&
               frames == o.frames && elemBytes == o.elemBytes;
    }
};
struct MapCache {
    static const int kEntries = 512;          // direct-mapped by a hash of the key: a ring of 64 4K frames is 128 descriptors
    MapKey key[kEntries];
    alignas(64) CUtensorMap tm[kEntries];
    MapCache() { memset(key, 0, sizeof(key)); }
};
thread_local std::unique_ptr<MapCache> g_mapCache;

int map_cache_entry(const uint8_t* base, uint32_t width, uint32_t rows, uint32_t frames, size_t pitch, int elemBytes)
{
    uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)width << 32 | rows) * 0xC2B2AE3D27D4EB4Full;
    h ^= (uint64_t)pitch * 0x165667B19E3779F9ull + frames + (uint64_t)elemBytes * 131;
    return (int)((h >> 40) % MapCache::kEntries);
}

// The descriptor is COPIED out (128 bytes): a call looks up two planes, and when both hash to the same entry the second
// look-up replaces the first one's descriptor -- a pointer into the cache would then describe the wrong image.
int get_plane_map(CUtensorMap* out, const uint8_t* base, uint32_t width, uint32_t rows, uint32_t frames, size_t pitch,
                  size_t frameStride, int elemBytes)
{
    if (!g_mapCache) g_mapCache.reset(new MapCache());
    MapCache& mc = *g_mapCache;
    const MapKey k = {base, pitch, frames > 1 ? frameStride : 0, width, rows, frames, elemBytes};
    const int slot = map_cache_entry(base, width, rows, frames, pitch, elemBytes);
    if (mc.key[slot] == k && k.base != nullptr) { *out = mc.tm[slot]; return 0; }
    mc.key[slot].base = nullptr;
    int rc = make_plane_map(&mc.tm[slot], base, width, rows, frames, pitch, frameStride, elemBytes);
    if (rc) return rc;
    mc.key[slot] = k;
    *out = mc.tm[slot];
    return 0;
}

// ------------------------------------------------------------------------------------------------ device context
struct Buffer {
    void*  ptr = nullptr;
    size_t cap = 0;
    bool   pinnedHost = false;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        release();
        const size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = pinnedHost ? cudaMallocHost(&ptr, want) : cudaMalloc(&ptr, want);
        if (e != cudaSuccess) { ptr = nullptr; cap = 0; cudaGetLastError(); return fail(ENOMEM, "allocating %zu bytes of %s memory failed", want, pinnedHost ? "pinned host" : "device"); }
        cap = want;
        return 0;
    }
    void release()
    {
        if (ptr) { if (pinnedHost) cudaFreeHost(ptr); else cudaFree(ptr); }
        ptr = nullptr; cap = 0;
    }
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {             // reduction workspace of one stream, see get_workspace()
    Buffer buf;
};

struct Context {
    int device = -1;
    int numSMs = 0;
    int pairsPerSm = ssimk::kPairsPerCta;     // warp pairs resident per SM
    cudaStream_t stream = nullptr;            // stream of the blocking host-pointer path (compute)
    cudaStream_t streamIn = nullptr;          // pipelined host path: H2D copies
    cudaStream_t streamOut = nullptr;         // pipelined host path: D2H copies
    static const int kMaxChunks = 32;
    cudaEvent_t evIn[kMaxChunks] = {}, evDone[kMaxChunks] = {};
    Buffer chunkSums, chunkSumsHost;          // one double per chunk (device, pinned host)
    std::mutex hostPathMutex;                 // the host path shares the scratch below
    Buffer planeA, planeB, rawA, rawB, map, stage;
    Buffer scalars;                           // double sum + float ssim of the host path
    std::mutex wsMutex;
    std::map<cudaStream_t, Workspace> workspaces;   // per-stream reduction workspace of the fused kernel
    std::vector<void*> retired;               // outgrown workspaces: work queued earlier may still use them, freed at shutdown
    float taps[6];
    float eps2 = 0.f;

    ~Context()
    {
        // RAII: also runs when create_context() fails half-way (nothing leaks on its error paths)
        if (device < 0) return;
        DeviceGuard g(device);
        cudaDeviceSynchronize();
        for (Buffer* b : {&planeA, &planeB, &rawA, &rawB, &map, &stage, &scalars, &chunkSums, &chunkSumsHost}) b->release();
        for (int i = 0; i < kMaxChunks; ++i) { if (evIn[i]) cudaEventDestroy(evIn[i]); if (evDone[i]) cudaEventDestroy(evDone[i]); }
        if (streamIn) cudaStreamDestroy(streamIn);
        if (streamOut) cudaStreamDestroy(streamOut);
        for (auto& p : workspaces) p.second.buf.release();
        for (void* p : retired) cudaFree(p);
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
    }
};

std::mutex g_ctxMutex;
std::map<int, Context*> g_ctx;
const int kFastDevices = 64;
std::atomic<Context*> g_ctxFast[kFastDevices];     // lock-free lookup for the hot device-pointer path

int create_context(int device, Context** out)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) { cudaGetLastError(); return fail(ENODEV, "no usable CUDA device (%s)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)); }
    if (device < 0 || device >= count) return fail(EINVAL, "device %d out of range [0,%d)", device, count);
    DEVICE_GUARD(device);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(ENODEV, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
    std::unique_ptr<Context> c(new Context());
    c->device = device;
    c->numSMs = prop.multiProcessorCount;
    c->stage.pinnedHost = true;
    gaussian_taps(c->taps);
    {
        double s1 = (double)c->taps[0];
        for (int d = 1; d < 6; ++d) s1 += 2.0 * (double)c->taps[d];
        c->eps2 = (float)(2.0 * (s1 * s1 - 1.0));
    }
    CU_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->streamIn, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->streamOut, cudaStreamNonBlocking));
    for (int i = 0; i < Context::kMaxChunks; ++i) {
        CU_TRY(cudaEventCreateWithFlags(&c->evIn[i], cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&c->evDone[i], cudaEventDisableTiming));
    }
    c->chunkSumsHost.pinnedHost = true;
    int regsMap = 0, regsNoMap = 0, pairs = 0;
    CU_TRY(ssimk::fused_kernel_attributes(&regsMap, &regsNoMap, &pairs));
    if (pairs < ssimk::kPairsPerCta) return fail(EIO, "the fused kernel does not fit this device (%d warp pairs per SM)", pairs);
    c->pairsPerSm = pairs;
    *out = c.release();
    return 0;
}

void destroy_context(Context* c) { delete c; }

int get_context(int device, Context** out)
{
    if (device >= 0 && device < kFastDevices) {
        Context* c = g_ctxFast[device].load(std::memory_order_acquire);
        if (c) { *out = c; return 0; }
    }
    std::lock_guard<std::mutex> lock(g_ctxMutex);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) { *out = it->second; return 0; }
    Context* c = nullptr;
    int rc = create_context(device, &c);
    if (rc) return rc;
    g_ctx[device] = c;
    if (device < kFastDevices) g_ctxFast[device].store(c, std::memory_order_release);
    *out = c;
    return 0;
}

// The blocking host-pointer calls use per-context scratch (planes, map, staging, streams), so one context serves one call
// at a time.  Callers that invoke the API from several threads (the reference is re-entrant, SURVEY 8b "Threading") get up
// to kHostContexts calls in flight per device: a busy primary context makes the call take (or lazily create) a sibling,
// whose copies and kernels then overlap the first call's on the GPU -- two threads hide each other's pipeline fill and
// drain and bring the host path close to the PCIe bound.
const size_t kHostContexts = 3;
std::map<int, std::vector<Context*>> g_hostCtx;     // siblings of g_ctx[device], guarded by g_ctxMutex

int acquire_host_context(Context** c, std::unique_lock<std::mutex>* lock)
{
    Context* primary = *c;
    *lock = std::unique_lock<std::mutex>(primary->hostPathMutex, std::try_to_lock);
    if (lock->owns_lock()) return 0;
    {
        std::lock_guard<std::mutex> g(g_ctxMutex);
        std::vector<Context*>& sib = g_hostCtx[primary->device];
        for (Context* s : sib) {
            *lock = std::unique_lock<std::mutex>(s->hostPathMutex, std::try_to_lock);
            if (lock->owns_lock()) { *c = s; return 0; }
        }
        if (sib.size() + 1 < kHostContexts) {
            Context* s = nullptr;
            int rc = create_context(primary->device, &s);
            if (rc) return rc;
            sib.push_back(s);
            *lock = std::unique_lock<std::mutex>(s->hostPathMutex);
            *c = s;
            return 0;
        }
    }
    *lock = std::unique_lock<std::mutex>(primary->hostPathMutex);      // everything busy: queue on the primary
    return 0;
}

// ------------------------------------------------------------------------------------------------ work partition + launch
// See "work partition" in ssim_kernels.h: the persistent grid's warp pairs ("slots") share the rows of all (frame, band)
// columns evenly.  The only choices left to the host are how many CTAs per SM to use and how thin the work may be spread
// (a slot pays 10 start-up rows, so tiny inputs use fewer slots).
const uint32_t kDefaultMinSlotRows = 6;       // measured (tools/dev/minrows_sweep.py): 256x256 9.5 us with 4-6, 11.0 with 12, 13.5 with 16; large inputs do not care

bool plan_for(const Context* c, uint32_t width, uint32_t outRows, uint32_t frames, ssimk::SlotPlan* plan)
{
    const int minRows = g_minSlotRows.load(std::memory_order_relaxed);
    int pairs = g_maxPairsPerSm.load(std::memory_order_relaxed);
    if (pairs <= 0 || pairs > c->pairsPerSm) pairs = c->pairsPerSm;
    return ssimk::plan_slots(std::min<uint32_t>((uint32_t)(c->numSMs * pairs), ssimk::kMaxSlots), width, outRows, frames, minRows > 0 ? (uint32_t)minRows : kDefaultMinSlotRows, plan);
}

// Reduction workspace of a stream: one 64-bit accumulator per frame, zero between launches -- cleared once here, and the
// kernel puts every word it has finished back to zero.  Launches on one stream run in order, so they can share it.  A
// workspace that has become too small is retired, not freed: work queued earlier on the stream may still be using it.
struct WorkspaceView { unsigned long long* frameAcc; };

int get_workspace(Context* c, cudaStream_t stream, size_t frames, WorkspaceView* out)
{
    std::lock_guard<std::mutex> lock(c->wsMutex);
    auto it = c->workspaces.find(stream);
    if (it == c->workspaces.end()) {
        if (c->workspaces.size() >= 64) {
            // stream handles come and go in a long-lived process: start over instead of growing without bound
            CU_TRY(cudaDeviceSynchronize());
            for (auto& p : c->workspaces) p.second.buf.release();
            for (void* p : c->retired) cudaFree(p);
            c->workspaces.clear();
            c->retired.clear();
        }
        it = c->workspaces.emplace(stream, Workspace()).first;
    }
    Workspace& w = it->second;
    if (frames * sizeof(unsigned long long) > w.buf.cap) {
        if (w.buf.ptr) { c->retired.push_back(w.buf.ptr); w.buf.ptr = nullptr; w.buf.cap = 0; }
        const size_t bytes = std::max<size_t>(2 * frames, 1024) * sizeof(unsigned long long);
        int rc = w.buf.ensure(bytes);
        if (rc) return rc;
        CU_TRY(cudaMemsetAsync(w.buf.ptr, 0, w.buf.cap, stream));
    }
    out->frameAcc = (unsigned long long*)w.buf.ptr;
    return 0;
}

int compute_device_impl(Context* c, cudaStream_t stream, uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows,
                        uint32_t frames, const uint8_t* dA, size_t pitchA, size_t frameStrideA, const uint8_t* dB, size_t pitchB,
                        size_t frameStrideB, float* dMap, size_t mapPitch, size_t mapFrameStride, double* dSums, float* dSsim,
                        int elemBytes = 1, const ssimk::ExchangeParams* xchg = nullptr, size_t mapStep = 1)
{
    g_lastLaunches = 0;
    if (width == 0 || srcRows == 0 || outRows == 0 || frames == 0) return fail(EINVAL, "width, rows and frames must be non-zero");
    if (dA == nullptr || dB == nullptr) return fail(EINVAL, "dA or dB is NULL");
    if (dMap == nullptr && dSums == nullptr && dSsim == nullptr && xchg == nullptr) return fail(EINVAL, "no output requested");
    if ((uint64_t)outY0 + outRows > srcRows) return fail(EINVAL, "output rows [%u,%u) exceed the %u source rows", outY0, outY0 + outRows, srcRows);
    if (width > 0x7fffff00u || srcRows > 0x7fffff00u) return fail(EINVAL, "dimensions too large");
    if (((uintptr_t)dA | (uintptr_t)dB | pitchA | pitchB) & 15) return fail(EINVAL, "plane base addresses and pitches must be multiples of 16 bytes");
    if (frames > 1 && ((frameStrideA | frameStrideB) & 15)) return fail(EINVAL, "frame strides must be multiples of 16 bytes");
    if (pitchA < (size_t)width * elemBytes || pitchB < (size_t)width * elemBytes) return fail(EINVAL, "pitch smaller than width");
    if (dMap && (mapStep < 1 || mapPitch < (size_t)width * mapStep)) return fail(EINVAL, "map pitch smaller than width");
    if (dMap && mapStep != 1 && elemBytes != 1) return fail(EINVAL, "maps with a pixel step are supported for 8-bit images only");

    ssimk::SlotPlan plan;
    if (!plan_for(c, width, outRows, frames, &plan)) return fail(EINVAL, "image or batch too large (more than 2^31 row units)");

    WorkspaceView ws;
    int rc = get_workspace(c, stream, frames, &ws);
    if (rc) return rc;
    alignas(64) CUtensorMap tmA, tmB;
    if ((rc = get_plane_map(&tmA, dA, width, srcRows, frames, pitchA, frameStrideA, elemBytes))) return rc;
    if ((rc = get_plane_map(&tmB, dB, width, srcRows, frames, pitchB, frameStrideB, elemBytes))) return rc;

    ssimk::FusedParams p;
    memset(&p, 0, sizeof(p));
    p.u16 = elemBytes == 2;
    p.a = dA; p.b = dB;
    p.pitchA = (long long)pitchA; p.frameStrideA = (long long)frameStrideA;
    p.pitchB = (long long)pitchB; p.frameStrideB = (long long)frameStrideB;
    p.map = dMap; p.mapPitch = (long long)mapPitch; p.mapFrameStride = (long long)mapFrameStride; p.mapStep = (long long)mapStep; p.mapPitchBytes = (long long)(mapPitch * sizeof(float));
    p.width = (int)width; p.srcRows = (int)srcRows; p.outY0 = (int)outY0; p.outRows = (int)outRows; p.frames = (int)frames;
    p.geo = ssimk::make_slot_geo(plan, width);
    p.frameAcc = ws.frameAcc;
    ssimk::acc_format(plan, &p.accBias, &p.accScale, &p.accInvScale);
    p.sums = dSums; p.ssim = dSsim;
    p.invCount = 1.0 / (double)(uint32_t)(width * outRows);    // uint32 product, as src/ssim.cpp:1102
    for (int d = 0; d < 6; ++d) p.g[d] = c->taps[d];
    p.magic = 0x4B000064u;
    {
        static const unsigned backoff = [] { const char* e = getenv("SSIM_CUDA_BACKOFF_NS"); return e ? (unsigned)atoi(e) : ssimk::kBackoffNs; }();
        p.backoffNs = backoff;
    }
    p.eps2 = c->eps2;
    p.dbgTimes = g_dbgTimes.load(std::memory_order_relaxed);
    CU_TRY(ssimk::launch_fused(stream, tmA, tmB, p, xchg));      // the ONLY launch: reduction (and exchange) happen inside
    g_lastLaunches = 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------ general path
enum class Where { Host, Device };

Where classify(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return Where::Host; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? Where::Device : Where::Host;
}

// Brings one strided image into a canonical plane (dense rows, 16-byte aligned pitch) in device memory.
// On return *plane/*pitch describe it; it is either the caller's own memory (already canonical) or ctx scratch.
int canonical_plane(Context* c, cudaStream_t s, const uint8_t* img, ptrdiff_t step, ptrdiff_t stride, uint32_t W, uint32_t H,
                    Buffer& planeBuf, Buffer& rawBuf, const uint8_t** plane, size_t* pitch, bool luma = false, int elemBytes = 1)
{
    // elemBytes == 2: 16-bit pixels; `step`/`stride` are BYTE distances in every case
    // luma: `img` points at the R byte of interleaved RGB(A) pixels, three channels are read per pixel
    auto pack = [&](uint8_t* dst, long long dstPitch, const uint8_t* src, long long st, long long sd) {
        return luma ? ssimk::launch_pack_luma(s, dst, dstPitch, src, st, sd, (int)W, (int)H)
             : elemBytes == 2 ? ssimk::launch_pack_u16(s, dst, dstPitch, src, st, sd, (int)W, (int)H)
                              : ssimk::launch_pack_u8(s, dst, dstPitch, src, st, sd, (int)W, (int)H);
    };
    const ptrdiff_t extra = luma ? 2 : elemBytes - 1;          // bytes read beyond the addressed one
    const size_t rowBytes = (size_t)W * elemBytes;
    const Where where = classify(img);
    const size_t canonPitch = align_up(rowBytes, 16);
    if (!luma && where == Where::Device && step == elemBytes && stride >= (ptrdiff_t)rowBytes && (stride & 15) == 0 && ((uintptr_t)img & 15) == 0) {
        *plane = img; *pitch = (size_t)stride;
        return 0;
    }
    int rc = planeBuf.ensure(canonPitch * H);
    if (rc) return rc;
    uint8_t* dst = (uint8_t*)planeBuf.ptr;
    *plane = dst; *pitch = canonPitch;
    if (!luma && where == Where::Host && step == elemBytes && stride >= (ptrdiff_t)rowBytes) {
        CU_TRY(cudaMemcpy2DAsync(dst, canonPitch, img, (size_t)stride, rowBytes, H, cudaMemcpyHostToDevice, s));
        return 0;
    }
    if (where == Where::Device) {
        CU_TRY(pack(dst, (long long)canonPitch, img, step, stride));
        return 0;
    }
    // host image with a general layout (interleaved channels, bottom-up, column-major ...): copy the byte range that
    // contains every addressed pixel, then gather on the device.
    const ptrdiff_t xExt = (ptrdiff_t)(W - 1) * step, yExt = (ptrdiff_t)(H - 1) * stride;
    const ptrdiff_t lo = std::min<ptrdiff_t>(0, xExt) + std::min<ptrdiff_t>(0, yExt);
    const ptrdiff_t hi = std::max<ptrdiff_t>(0, xExt) + std::max<ptrdiff_t>(0, yExt) + extra;
    const size_t span = (size_t)(std::abs(xExt)) + 1 + (size_t)extra;   // bytes touched in one row
    const size_t absStride = (size_t)std::abs(stride);
    if (absStride >= span && H > 1) {
        // row-major-like: 2-D copy of H rows of `span` bytes
        const size_t rawPitch = align_up(span, 16);
        if ((rc = rawBuf.ensure(rawPitch * H))) return rc;
        const uint8_t* firstRow = img + std::min<ptrdiff_t>(0, xExt) + std::min<ptrdiff_t>(0, yExt);   // lowest row start
        CU_TRY(cudaMemcpy2DAsync(rawBuf.ptr, rawPitch, firstRow, absStride, span, H, cudaMemcpyHostToDevice, s));
        // device address of pixel (0,0): row index in the copy is y (stride>0) or H-1-y (stride<0)
        const uint8_t* d0 = (const uint8_t*)rawBuf.ptr + (stride < 0 ? (size_t)(H - 1) * rawPitch : 0) + (step < 0 ? span - 1 - (size_t)extra : 0);
        CU_TRY(pack(dst, (long long)canonPitch, d0, step, stride < 0 ? -(long long)rawPitch : (long long)rawPitch));
    } else {
        const size_t bytes = (size_t)(hi - lo) + 1;
        if ((rc = rawBuf.ensure(bytes))) return rc;
        CU_TRY(cudaMemcpyAsync(rawBuf.ptr, img + lo, bytes, cudaMemcpyHostToDevice, s));
        CU_TRY(pack(dst, (long long)canonPitch, (const uint8_t*)rawBuf.ptr - lo, step, stride));
    }
    return 0;
}

// One "general job" = one image pair (or one row strip of it) in the caller's layout, enqueued on the context's
// stream: inputs are made canonical, the fused kernel runs, the map starts flowing back.  finish_general() waits
// and completes the host-side part.  `a`/`b` point at source row 0 of the strip, `map` at its first OUTPUT row.
struct GeneralJob {
    Context* c = nullptr;
    uint32_t W = 0, outRows = 0;
    float* map = nullptr;
    ptrdiff_t mapStep = 0, mapStride = 0;
    bool cpuScatter = false;
    // the map's way back to the caller, when it is not written in place: enqueued by return_map()
    float* dMap = nullptr;
    size_t dMapPitch = 0;
    bool mapDirect = false, mapOnDevice = false;
};

int return_map(GeneralJob* job);

int enqueue_general(Context* c, uint32_t W, uint32_t srcRows, uint32_t outY0, uint32_t outRows, const uint8_t* a, ptrdiff_t stepA,
                    ptrdiff_t strideA, const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep,
                    ptrdiff_t mapStride, bool wantSsim, GeneralJob* job, bool luma = false, int elemBytes = 1,
                    const ssimk::ExchangeParams* xchg = nullptr, bool deferMapReturn = false)
{
    DEVICE_GUARD(c->device);
    cudaStream_t s = c->stream;
    const uint8_t *pa, *pb;
    size_t pitchA, pitchB;
    int rc;
    if ((rc = canonical_plane(c, s, a, stepA, strideA, W, srcRows, c->planeA, c->rawA, &pa, &pitchA, luma, elemBytes))) return rc;
    if ((rc = canonical_plane(c, s, b, stepB, strideB, W, srcRows, c->planeB, c->rawB, &pb, &pitchB, luma, elemBytes))) return rc;

    // map destination: write straight into a canonical device map of the caller, else into scratch
    float* dMap = nullptr;
    size_t dMapPitch = 0, dMapStep = 1;
    bool mapDirect = false;
    Where mapWhere = Where::Host;
    if (map) {
        mapWhere = classify(map);
        if (mapWhere == Where::Device && mapStep >= 1 && mapStride >= (ptrdiff_t)W * mapStep && (mapStep == 1 || elemBytes == 1)) {
            // the kernel writes the caller's device map in place, interleaved neighbours (step > 1) stay untouched
            dMap = map; dMapPitch = (size_t)mapStride; dMapStep = (size_t)mapStep; mapDirect = true;
        } else {
            dMapPitch = align_up(W, 4);
            if ((rc = c->map.ensure(dMapPitch * outRows * sizeof(float)))) return rc;
            dMap = (float*)c->map.ptr;
        }
    }
    if ((rc = c->scalars.ensure(16))) return rc;
    double* dSum = (double*)c->scalars.ptr;
    float* dSsim = (float*)((char*)c->scalars.ptr + 8);

    rc = compute_device_impl(c, s, W, srcRows, outY0, outRows, 1, pa, pitchA, 0, pb, pitchB, 0, dMap, dMapPitch, 0, dSum,
                             wantSsim ? dSsim : nullptr, elemBytes, xchg, dMapStep);
    if (rc) return rc;

    job->c = c; job->W = W; job->outRows = outRows; job->map = map; job->mapStep = mapStep; job->mapStride = mapStride;
    job->cpuScatter = false;
    job->dMap = dMap; job->dMapPitch = dMapPitch; job->mapDirect = mapDirect; job->mapOnDevice = mapWhere == Where::Device;
    // A device-to-host copy into pageable memory blocks the calling thread until the kernel before it has finished: callers
    // that still have kernels to launch on OTHER devices which this kernel waits for (strips exchanged over peer memory)
    // defer the map's return until everything is launched.
    return deferMapReturn ? 0 : return_map(job);
}

int return_map(GeneralJob* job)
{
    if (!job->map || job->mapDirect) return 0;
    Context* c = job->c;
    DEVICE_GUARD(c->device);
    cudaStream_t s = c->stream;
    const uint32_t W = job->W, outRows = job->outRows;
    int rc;
    if (job->mapOnDevice) {
        CU_TRY(ssimk::launch_scatter_map(s, job->map, job->mapStep, job->mapStride, job->dMap, (long long)job->dMapPitch, (int)W, (int)outRows));
    } else if (job->mapStep == 1 && job->mapStride >= (ptrdiff_t)W) {
        CU_TRY(cudaMemcpy2DAsync(job->map, (size_t)job->mapStride * sizeof(float), job->dMap, job->dMapPitch * sizeof(float), (size_t)W * sizeof(float),
                                 outRows, cudaMemcpyDeviceToHost, s));
    } else {
        // strided / bottom-up host map: dense copy to pinned staging, scattered by the CPU in finish_general(),
        // touching only the addressed floats (interleaved neighbours stay untouched, as in src/ssim.cpp:661-667)
        if ((rc = c->stage.ensure((size_t)W * outRows * sizeof(float)))) return rc;
        CU_TRY(cudaMemcpy2DAsync(c->stage.ptr, (size_t)W * sizeof(float), job->dMap, job->dMapPitch * sizeof(float), (size_t)W * sizeof(float),
                                 outRows, cudaMemcpyDeviceToHost, s));
        job->cpuScatter = true;
    }
    return 0;
}

int finish_general(const GeneralJob& job)
{
    DEVICE_GUARD(job.c->device);
    CU_TRY(cudaStreamSynchronize(job.c->stream));
    if (job.cpuScatter) {
        const float* src = (const float*)job.c->stage.ptr;
        for (uint32_t y = 0; y < job.outRows; ++y)
            for (uint32_t x = 0; x < job.W; ++x)
                job.map[(ptrdiff_t)x * job.mapStep + (ptrdiff_t)y * job.mapStride] = src[(size_t)y * job.W + x];
    }
    return 0;
}

const uint64_t kPipelineMinBytes = 1u << 21;      // images of at least 2 MB take the chunked copy/compute pipeline

// Large host images in plain row layout: the image is cut into row chunks and three streams overlap the H2D copy of
// chunk k+1, the kernel of chunk k and the map D2H of chunk k-1 (PCIe is full duplex; the kernel is ~10x faster than
// either copy, so a blocking call costs about max(H2D, D2H) instead of H2D + kernel + D2H).  A chunk's kernel reads rows
// up to 5 below its last output row, so its H2D covers 5 extra rows; everything lands in one full-height device plane and
// the kernel addresses it with (srcRows = H, outY0, outRows), i.e. exactly the strip mechanism of the multi-GPU path.
int compute_pipelined_body(Context* c, uint32_t W, uint32_t H, const uint8_t* a, ptrdiff_t strideA, const uint8_t* b, ptrdiff_t strideB,
                           float* map, ptrdiff_t mapStride, float* ssim, int elemBytes)
{
    // a, b: byte pointers; strideA/strideB in BYTES; elemBytes = 1 (8-bit) or 2 (16-bit pixels)
    DEVICE_GUARD(c->device);
    const size_t rowBytes = (size_t)W * elemBytes;
    const size_t pitch = align_up(rowBytes, 16), mapPitch = align_up(W, 4);
    int rc;
    if ((rc = c->planeA.ensure(pitch * H)) || (rc = c->planeB.ensure(pitch * H))) return rc;
    if (map && (rc = c->map.ensure(mapPitch * H * sizeof(float)))) return rc;
    // Uniform chunks of ~2 MB of pixels per image: the per-chunk host API cost (~20 us: three copies, two events, one launch)
    // makes finer chunks lose, and graded chunk sizes (small first/last chunks to shorten pipeline fill and drain) measured
    // no better than 4 uniform chunks for a 4K pair (tools/dev/e2e_sweep.py).
    static const uint32_t chunkBytes = [] { const char* e = getenv("SSIM_CUDA_CHUNK_KB"); return (e ? (uint32_t)atoi(e) : 2048u) << 10; }();
    const uint32_t targetRows = std::max<uint32_t>(64, chunkBytes / std::max<uint32_t>((uint32_t)rowBytes, 1));
    const int nChunks = std::max<int>(2, (int)std::min<uint32_t>(Context::kMaxChunks, (H + targetRows - 1) / targetRows));
    std::vector<uint32_t> bounds(nChunks + 1);                 // chunk k covers rows [bounds[k], bounds[k+1])
    for (int k = 0; k <= nChunks; ++k) bounds[k] = (uint32_t)((uint64_t)H * k / nChunks);
    if ((rc = c->chunkSums.ensure(sizeof(double) * Context::kMaxChunks)) || (rc = c->chunkSumsHost.ensure(sizeof(double) * Context::kMaxChunks))) return rc;
    uint8_t *dA = (uint8_t*)c->planeA.ptr, *dB = (uint8_t*)c->planeB.ptr;
    float* dMap = map ? (float*)c->map.ptr : nullptr;
    double* dSums = (double*)c->chunkSums.ptr;

    // Every chunk is ONE launch: the tensor maps of the whole planes come from the descriptor cache (encoded by the first
    // chunk, the context's scratch planes rarely move), and each launch leaves its chunk's double sum in dSums[k].
    uint32_t copied = 0;                                       // rows [0, copied) are on the device (or in flight on streamIn)
    for (int k = 0; k < nChunks; ++k) {
        const uint32_t y0 = bounds[k], y1 = bounds[k + 1];
        const uint32_t need = std::min<uint64_t>(H, (uint64_t)y1 + ssimk::kHalo);
        if (need > copied) {
            CU_TRY(cudaMemcpy2DAsync(dA + (size_t)copied * pitch, pitch, a + (ptrdiff_t)copied * strideA, (size_t)strideA, rowBytes, need - copied,
                                     cudaMemcpyHostToDevice, c->streamIn));
            CU_TRY(cudaMemcpy2DAsync(dB + (size_t)copied * pitch, pitch, b + (ptrdiff_t)copied * strideB, (size_t)strideB, rowBytes, need - copied,
                                     cudaMemcpyHostToDevice, c->streamIn));
            copied = need;
        }
        if (y1 == y0) continue;
        CU_TRY(cudaEventRecord(c->evIn[k], c->streamIn));
        CU_TRY(cudaStreamWaitEvent(c->stream, c->evIn[k], 0));
        rc = compute_device_impl(c, c->stream, W, H, y0, y1 - y0, 1, dA, pitch, 0, dB, pitch, 0,
                                 dMap ? dMap + (size_t)y0 * mapPitch : nullptr, mapPitch, 0, dSums + k, nullptr, elemBytes);
        if (rc) return rc;
        if (map) {
            CU_TRY(cudaEventRecord(c->evDone[k], c->stream));
            CU_TRY(cudaStreamWaitEvent(c->streamOut, c->evDone[k], 0));
            CU_TRY(cudaMemcpy2DAsync(map + (ptrdiff_t)y0 * mapStride, (size_t)mapStride * sizeof(float), dMap + (size_t)y0 * mapPitch,
                                     mapPitch * sizeof(float), (size_t)W * sizeof(float), y1 - y0, cudaMemcpyDeviceToHost, c->streamOut));
        }
    }
    // the chunk sums in chunk order (deterministic), then the reference's last step (src/ssim.cpp:1102)
    CU_TRY(cudaMemcpyAsync(c->chunkSumsHost.ptr, dSums, sizeof(double) * nChunks, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (map) CU_TRY(cudaStreamSynchronize(c->streamOut));
    if (ssim) {
        double total = 0.0;
        for (int k = 0; k < nChunks; ++k)
            if (bounds[k + 1] > bounds[k]) total += ((const double*)c->chunkSumsHost.ptr)[k];
        *ssim = (float)(total / (double)(uint32_t)(W * H));
    }
    return 0;
}

// A failed call must not leave copies in flight on the caller's memory (include/ssim_cuda.h: "a call never retains caller
// pointers"): error paths of the host-pointer entry points drain the context's streams before returning.
int drained(Context* c, int rc)
{
    if (rc) {
        DeviceGuard g(c->device);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->streamIn);
        cudaStreamSynchronize(c->streamOut);
        cudaGetLastError();
    }
    return rc;
}

int compute_pipelined(Context* c, uint32_t W, uint32_t H, const uint8_t* a, ptrdiff_t strideA, const uint8_t* b, ptrdiff_t strideB,
                      float* map, ptrdiff_t mapStride, float* ssim, int elemBytes = 1)
{
    return drained(c, compute_pipelined_body(c, W, H, a, strideA, b, strideB, map, mapStride, ssim, elemBytes));
}

// Brings the float result of an enqueued general job back (through the context's pinned scalars) and completes the job.
int finish_with_scalar(Context* c, const GeneralJob& job, float* ssim)
{
    int rc = 0;
    {
        DeviceGuard g(c->device);
        if (ssim) {
            rc = c->chunkSumsHost.ensure(sizeof(double) * Context::kMaxChunks);
            if (!rc && cudaMemcpyAsync(c->chunkSumsHost.ptr, (char*)c->scalars.ptr + 8, sizeof(float), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
                rc = cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(ssim)");
        }
    }
    if (!rc) rc = finish_general(job);
    if (rc) return drained(c, rc);
    if (ssim) *ssim = *(const float*)c->chunkSumsHost.ptr;
    return 0;
}

int compute_general(Context* c, uint32_t W, uint32_t H, const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA, const uint8_t* b,
                    ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep, ptrdiff_t mapStride, float* ssim)
{
    std::unique_lock<std::mutex> lock;
    if (int arc = acquire_host_context(&c, &lock)) return arc;      // may switch to a free sibling context
    // pipelined path: plain-row host images (and host map) large enough for the overlap to pay.  Pageable memory works
    // too (the runtime stages it), pinned memory gets the full overlap.
    if ((uint64_t)W * H >= kPipelineMinBytes && stepA == 1 && stepB == 1 && strideA >= (ptrdiff_t)W && strideB >= (ptrdiff_t)W &&
        classify(a) == Where::Host && classify(b) == Where::Host &&
        (map == nullptr || (mapStep == 1 && mapStride >= (ptrdiff_t)W && classify(map) == Where::Host)) &&
        getenv("SSIM_CUDA_NO_PIPELINE") == nullptr)
        return compute_pipelined(c, W, H, a, strideA, b, strideB, map, mapStride, ssim);
    GeneralJob job;
    int rc = enqueue_general(c, W, H, 0, H, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride, ssim != nullptr, &job);
    if (rc) return drained(c, rc);
    return finish_with_scalar(c, job, ssim);
}

// 16-bit pixels: same machinery, element size 2, blocking single-shot path
int compute_general_u16(Context* c, uint32_t W, uint32_t H, const uint16_t* a, ptrdiff_t stepA, ptrdiff_t strideA, const uint16_t* b,
                        ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep, ptrdiff_t mapStride, float* ssim)
{
    std::unique_lock<std::mutex> lock;
    if (int arc = acquire_host_context(&c, &lock)) return arc;      // may switch to a free sibling context
    // same pipelined path as 8-bit images for large plain-row host images (H2D / kernel / D2H overlapped in row chunks);
    // the threshold is in BYTES of one image (2 MB, as for 8-bit pixels), hence half as many pixels
    if ((uint64_t)W * H * 2 >= kPipelineMinBytes && stepA == 1 && stepB == 1 && strideA >= (ptrdiff_t)W && strideB >= (ptrdiff_t)W &&
        classify(a) == Where::Host && classify(b) == Where::Host &&
        (map == nullptr || (mapStep == 1 && mapStride >= (ptrdiff_t)W && classify(map) == Where::Host)) &&
        getenv("SSIM_CUDA_NO_PIPELINE") == nullptr)
        return compute_pipelined(c, W, H, (const uint8_t*)a, 2 * strideA, (const uint8_t*)b, 2 * strideB, map, mapStride, ssim, 2);
    GeneralJob job;
    int rc = enqueue_general(c, W, H, 0, H, (const uint8_t*)a, 2 * stepA, 2 * strideA, (const uint8_t*)b, 2 * stepB, 2 * strideB, map, mapStep,
                             mapStride, ssim != nullptr, &job, false, 2);
    if (rc) return drained(c, rc);
    return finish_with_scalar(c, job, ssim);
}

// ------------------------------------------------------------------------------------------------ NCCL (dlopen'ed)
// Only the scalar partial sums cross GPUs (BASELINE.json north_star), so NCCL is loaded lazily and privately:
// libssim_cuda.so has no link-time dependency on it and coexists with a framework that bundles its own copy.
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::map<std::vector<int>, std::vector<ncclComm_t>> comms;
    std::mutex mutex;
};
Nccl g_nccl;

int nccl_load()
{
    if (g_nccl.handle) return 0;
    const char* names[] = {getenv("SSIM_CUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return fail(EIO, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define NCCL_SYM(field, name)                                                                  \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, name);                                     \
    if (!g_nccl.field) return fail(EIO, "NCCL symbol %s not found", name);
    NCCL_SYM(CommInitAll, "ncclCommInitAll")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    return 0;
}

#define NCCL_TRY(expr)                                                                          \
    do {                                                                                        \
        ncclResult_t r__ = (expr);                                                              \
        if (r__ != ncclSuccess) return fail(EIO, "%s: %s", #expr, g_nccl.GetErrorString(r__));  \
    } while (0)

// ---- exchange buffers of a single process driving several GPUs (ssim_cuda_compute_strips)
std::mutex g_xchgMutex;
std::map<int, void*> g_xchgBuf;                       // device -> its exchange buffer
std::map<std::pair<int, int>, bool> g_peerEnabled;    // (device, peer) pairs with peer access switched on
std::atomic<unsigned long long> g_stripEpoch{0};

int exchange_alloc(int device, void** out)
{
    DEVICE_GUARD(device);
    const size_t bytes = 2 * ssimk::kMaxRanks * sizeof(ssimk::ExchangeSlot);
    CU_TRY(cudaMalloc(out, bytes));
    CU_TRY(cudaMemset(*out, 0, bytes));
    CU_TRY(cudaDeviceSynchronize());
    return 0;
}

int enable_peer(int device, int peerDevice)
{
    if (device == peerDevice) return 0;
    DEVICE_GUARD(device);
    int can = 0;
    CU_TRY(cudaDeviceCanAccessPeer(&can, device, peerDevice));
    if (!can) return fail(ENODEV, "device %d cannot access device %d", device, peerDevice);
    cudaError_t e = cudaDeviceEnablePeerAccess(peerDevice, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    CU_TRY(e);
    return 0;
}

int compute_strips(int n, const int* devices, uint32_t W, uint32_t H, const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                   const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep, ptrdiff_t mapStride, float* ssim)
{
    if ((uint32_t)n > H) n = (int)H;                  // no empty strips: an image of fewer rows than devices uses fewer devices
    std::vector<Context*> ctx(n);
    for (int g = 0; g < n; ++g) {
        for (int k = 0; k < g; ++k)
            if (devices[k] == devices[g]) return fail(EINVAL, "device %d listed twice", devices[g]);
        int rc = get_context(devices[g], &ctx[g]);
        if (rc) return rc;
    }
    // The strip sums are combined inside the kernels over NVLink peer memory (ssimk::ExchangeParams); NCCL stays available
    // as SSIM_CUDA_STRIPS_NCCL=1 (one ncclAllReduce of a double per GPU, as BASELINE.json's north star words it).
    static const bool useNccl = [] { const char* e = getenv("SSIM_CUDA_STRIPS_NCCL"); return e && atoi(e) != 0; }();
    if (n > ssimk::kMaxRanks && !useNccl) return fail(EINVAL, "at most %d devices", ssimk::kMaxRanks);
    std::vector<ncclComm_t>* comms = nullptr;
    std::vector<void*> xbuf(n, nullptr);
    if (n > 1 && useNccl) {
        std::lock_guard<std::mutex> lock(g_nccl.mutex);
        int rc = nccl_load();
        if (rc) return rc;
        std::vector<int> key(devices, devices + n);
        auto it = g_nccl.comms.find(key);
        if (it == g_nccl.comms.end()) {
            std::vector<ncclComm_t> c(n);
            NCCL_TRY(g_nccl.CommInitAll(c.data(), n, devices));
            it = g_nccl.comms.emplace(key, c).first;
        }
        comms = &it->second;
    } else if (n > 1) {
        std::lock_guard<std::mutex> lock(g_xchgMutex);
        for (int g = 0; g < n; ++g) {
            auto it = g_xchgBuf.find(devices[g]);
            if (it == g_xchgBuf.end()) {
                void* buf = nullptr;
                int rc = exchange_alloc(devices[g], &buf);
                if (rc) return rc;
                it = g_xchgBuf.emplace(devices[g], buf).first;
            }
            xbuf[g] = it->second;
            for (int k = 0; k < n; ++k) {
                if (k == g || g_peerEnabled.count({devices[g], devices[k]})) continue;
                int rc = enable_peer(devices[g], devices[k]);
                if (rc) return rc;
                g_peerEnabled[{devices[g], devices[k]}] = true;
            }
        }
    }
    // lock the contexts in device order (no lock-order inversion between concurrent callers)
    std::vector<int> order(n);
    for (int g = 0; g < n; ++g) order[g] = g;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return devices[x] < devices[y]; });
    std::vector<std::unique_lock<std::mutex>> locks;
    for (int g : order) locks.emplace_back(ctx[g]->hostPathMutex);

    auto fail_all = [&](int rc) { for (int g = 0; g < n; ++g) drained(ctx[g], rc); return rc; };

    // strip g produces rows [g*H/n, (g+1)*H/n) and reads 5 more rows on each interior edge (src/ssim.cpp:749-761:
    // every tile of the reference re-reads its halo the same way)
    const unsigned long long epoch = g_stripEpoch.fetch_add(1) + 1;
    static const unsigned long long timeoutMs = [] { const char* e = getenv("SSIM_CUDA_EXCHANGE_TIMEOUT_MS"); return e ? (unsigned long long)atoll(e) : 2000ull; }();
    std::vector<GeneralJob> jobs(n);
    for (int g = 0; g < n; ++g) {
        const uint32_t y0 = (uint32_t)((uint64_t)H * g / n), y1 = (uint32_t)((uint64_t)H * (g + 1) / n);
        DEVICE_GUARD(devices[g]);
        int rc = ctx[g]->scalars.ensure(16);
        if (!rc) rc = ctx[g]->chunkSums.ensure(sizeof(double) * Context::kMaxChunks);
        if (!rc) rc = ctx[g]->chunkSumsHost.ensure(sizeof(double) * Context::kMaxChunks);
        if (rc) return fail_all(rc);
        ssimk::ExchangeParams x;
        memset(&x, 0, sizeof(x));
        if (n > 1 && !useNccl) {
            for (int r = 0; r < n; ++r) x.peers[r] = (ssimk::ExchangeSlot*)xbuf[r];
            x.world = n; x.rank = g; x.epoch = epoch; x.timeoutNs = timeoutMs * 1000000ull;
            x.sumAll = (double*)ctx[g]->chunkSums.ptr;
            x.status = (int*)((char*)ctx[g]->chunkSums.ptr + 8);
            x.invCountAll = 0.0;
        }
        const uint32_t s0 = y0 >= (uint32_t)ssimk::kHalo ? y0 - ssimk::kHalo : 0, s1 = std::min<uint64_t>(H, (uint64_t)y1 + ssimk::kHalo);
        rc = enqueue_general(ctx[g], W, s1 - s0, y0 - s0, y1 - y0, a + (ptrdiff_t)s0 * strideA, stepA, strideA, b + (ptrdiff_t)s0 * strideB,
                             stepB, strideB, map ? map + (ptrdiff_t)y0 * mapStride : nullptr, mapStep, mapStride, false, &jobs[g], false, 1,
                             x.world ? &x : nullptr, /*deferMapReturn*/ true);
        if (rc) return fail_all(rc);
    }
    // every GPU's kernel is launched: now the maps may start flowing back (see enqueue_general)
    for (int g = 0; g < n; ++g) {
        int rc = return_map(&jobs[g]);
        if (rc) return fail_all(rc);
    }
    // where the total ends up on device 0: the exchange wrote it (and a status word) to chunkSums, NCCL reduces scalars in place
    const void* dTotal = ctx[0]->scalars.ptr;
    if (n > 1 && useNccl) {
        if (g_nccl.GroupStart() != ncclSuccess) return fail_all(fail(EIO, "ncclGroupStart failed"));
        for (int g = 0; g < n; ++g)
            if (g_nccl.AllReduce(ctx[g]->scalars.ptr, ctx[g]->scalars.ptr, 1, ncclDouble, ncclSum, (*comms)[g], ctx[g]->stream) != ncclSuccess)
                return fail_all(fail(EIO, "ncclAllReduce failed"));
        if (g_nccl.GroupEnd() != ncclSuccess) return fail_all(fail(EIO, "ncclGroupEnd failed"));
    } else if (n > 1) {
        dTotal = ctx[0]->chunkSums.ptr;
    }
    {
        DEVICE_GUARD(devices[0]);
        if (cudaMemcpyAsync(ctx[0]->chunkSumsHost.ptr, dTotal, 16, cudaMemcpyDeviceToHost, ctx[0]->stream) != cudaSuccess)
            return fail_all(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync(total)"));
    }
    for (int g = 0; g < n; ++g) {
        int rc = finish_general(jobs[g]);
        if (rc) return fail_all(rc);
    }
    const double total = *(const double*)ctx[0]->chunkSumsHost.ptr;
    if (n > 1 && !useNccl && *(const int*)((const char*)ctx[0]->chunkSumsHost.ptr + 8) != 0)
        return fail(EIO, "strip-sum exchange timed out: a peer GPU did not deliver its sum within %llu ms", timeoutMs);
    if (ssim) *ssim = (float)(total / (double)(uint32_t)(W * H));
    return 0;
}

int compute_channels_body(Context* c, uint32_t width, uint32_t height, uint32_t channels, const uint8_t* a, ptrdiff_t strideA,
                          const uint8_t* b, ptrdiff_t strideB, float* map, ptrdiff_t mapStride, float* ssim);

}  // namespace

// ================================================================================================ C ABI
extern "C" {

int ssim_cuda_abi_version(void) { return SSIM_CUDA_ABI_VERSION; }

int ssim_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ssim_cuda_init(int device)
{
    Context* c;
    return get_context(device, &c);
}

void ssim_cuda_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_ctxMutex);
    for (int d = 0; d < kFastDevices; ++d) g_ctxFast[d].store(nullptr, std::memory_order_release);
    for (auto& kv : g_ctx) destroy_context(kv.second);
    for (auto& kv : g_hostCtx)
        for (Context* s : kv.second) destroy_context(s);
    g_hostCtx.clear();
    g_ctx.clear();
    {
        std::lock_guard<std::mutex> xlock(g_xchgMutex);
        for (auto& kv : g_xchgBuf) { DeviceGuard g(kv.first); cudaFree(kv.second); }
        g_xchgBuf.clear();
    }
    {
        std::lock_guard<std::mutex> nlock(g_nccl.mutex);
        for (auto& kv : g_nccl.comms)
            for (ncclComm_t comm : kv.second) g_nccl.CommDestroy(comm);
        g_nccl.comms.clear();
    }
}

const char* ssim_cuda_last_error_string(void) { return g_err; }

void* ssim_cuda_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void ssim_cuda_host_free(void* p) { if (p) cudaFreeHost(p); }

int ssim_cuda_compute(int device, uint32_t width, uint32_t height, const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                      const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep, ptrdiff_t mapStride, float* ssim)
{
    if (ssim == nullptr && map == nullptr) return fail(EINVAL, "both ssim and map are NULL, nothing would be computed");
    if (a == nullptr || b == nullptr) return fail(EINVAL, "image pointer is NULL");
    if (width == 0 || height == 0) return fail(EINVAL, "width and height must be non-zero");
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    return compute_general(c, width, height, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride, ssim);
}

int ssim_cuda_compute_luma(int device, uint32_t width, uint32_t height, const uint8_t* rgbA, ptrdiff_t stepA, ptrdiff_t strideA,
                           const uint8_t* rgbB, ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep, ptrdiff_t mapStride, float* ssim)
{
    if (ssim == nullptr && map == nullptr) return fail(EINVAL, "both ssim and map are NULL, nothing would be computed");
    if (rgbA == nullptr || rgbB == nullptr) return fail(EINVAL, "image pointer is NULL");
    if (width == 0 || height == 0) return fail(EINVAL, "width and height must be non-zero");
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    std::unique_lock<std::mutex> lock;
    if (int arc = acquire_host_context(&c, &lock)) return arc;      // may switch to a free sibling context
    GeneralJob job;
    rc = enqueue_general(c, width, height, 0, height, rgbA, stepA, strideA, rgbB, stepB, strideB, map, mapStep, mapStride, ssim != nullptr, &job, true);
    if (rc) return drained(c, rc);
    return finish_with_scalar(c, job, ssim);
}

int ssim_cuda_compute_channels(int device, uint32_t width, uint32_t height, uint32_t channels, const uint8_t* a, ptrdiff_t strideA,
                               const uint8_t* b, ptrdiff_t strideB, float* map, ptrdiff_t mapStride, float* ssim)
{
    if (ssim == nullptr && map == nullptr) return fail(EINVAL, "both ssim and map are NULL, nothing would be computed");
    if (a == nullptr || b == nullptr) return fail(EINVAL, "image pointer is NULL");
    if (width == 0 || height == 0 || channels == 0 || channels > 16) return fail(EINVAL, "bad dimensions / channel count");
    const size_t rowBytes = (size_t)width * channels;
    if (strideA < (ptrdiff_t)rowBytes || strideB < (ptrdiff_t)rowBytes || (map && mapStride < (ptrdiff_t)rowBytes))
        return fail(EINVAL, "strides must be positive and cover width*channels");
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    std::unique_lock<std::mutex> lock;
    if (int arc = acquire_host_context(&c, &lock)) return arc;      // may switch to a free sibling context
    return drained(c, compute_channels_body(c, width, height, channels, a, strideA, b, strideB, map, mapStride, ssim));
}

}  // extern "C"

namespace {
int compute_channels_body(Context* c, uint32_t width, uint32_t height, uint32_t channels, const uint8_t* a, ptrdiff_t strideA,
                          const uint8_t* b, ptrdiff_t strideB, float* map, ptrdiff_t mapStride, float* ssim)
{
    const size_t rowBytes = (size_t)width * channels;
    int rc;
    DEVICE_GUARD(c->device);
    cudaStream_t s = c->stream;
    const size_t rawPitch = align_up(rowBytes, 16), pitch = align_up(width, 16), plane = pitch * height;
    if ((rc = c->rawA.ensure(rawPitch * height)) || (rc = c->rawB.ensure(rawPitch * height))) return rc;
    if ((rc = c->planeA.ensure(plane * channels)) || (rc = c->planeB.ensure(plane * channels))) return rc;
    if ((rc = c->scalars.ensure((sizeof(double) + sizeof(float)) * 16))) return rc;
    float* dInter = nullptr;
    if (map) {
        if ((rc = c->map.ensure(rowBytes * height * sizeof(float)))) return rc;
        dInter = (float*)c->map.ptr;
    }
    // one upload of the interleaved bytes, one split into planes, ONE fused launch with the channels as frames that writes the
    // interleaved map directly (pixel step = channels, "frame" stride = 1 float)
    CU_TRY(cudaMemcpy2DAsync(c->rawA.ptr, rawPitch, a, (size_t)strideA, rowBytes, height, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpy2DAsync(c->rawB.ptr, rawPitch, b, (size_t)strideB, rowBytes, height, cudaMemcpyHostToDevice, s));
    CU_TRY(ssimk::launch_deinterleave_u8(s, (uint8_t*)c->planeA.ptr, (long long)pitch, (long long)plane, (const uint8_t*)c->rawA.ptr, (long long)rawPitch, (int)channels, (int)width, (int)height));
    CU_TRY(ssimk::launch_deinterleave_u8(s, (uint8_t*)c->planeB.ptr, (long long)pitch, (long long)plane, (const uint8_t*)c->rawB.ptr, (long long)rawPitch, (int)channels, (int)width, (int)height));
    double* dSums = (double*)c->scalars.ptr;
    float* dSsim = (float*)((char*)c->scalars.ptr + sizeof(double) * 16);
    rc = compute_device_impl(c, s, width, height, 0, height, channels, (const uint8_t*)c->planeA.ptr, pitch, plane, (const uint8_t*)c->planeB.ptr, pitch, plane,
                             dInter, rowBytes, 1, dSums, dSsim, 1, nullptr, channels);
    if (rc) return rc;
    float hostSsim[16];
    if (ssim) CU_TRY(cudaMemcpyAsync(hostSsim, dSsim, sizeof(float) * channels, cudaMemcpyDeviceToHost, s));
    if (map) {
        CU_TRY(cudaMemcpy2DAsync(map, (size_t)mapStride * sizeof(float), dInter, rowBytes * sizeof(float), rowBytes * sizeof(float), height, cudaMemcpyDeviceToHost, s));
    }
    CU_TRY(cudaStreamSynchronize(s));
    if (ssim) memcpy(ssim, hostSsim, sizeof(float) * channels);
    return 0;
}
}  // namespace

extern "C" {

int ssim_cuda_compute_device(int device, void* stream, uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows, uint32_t frames,
                             const uint8_t* dA, size_t pitchA, size_t frameStrideA, const uint8_t* dB, size_t pitchB, size_t frameStrideB,
                             float* dMap, size_t mapPitch, size_t mapFrameStride, double* dSums, float* dSsim)
{
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    DEVICE_GUARD(device);
    return compute_device_impl(c, (cudaStream_t)stream, width, srcRows, outY0, outRows, frames, dA, pitchA, frameStrideA, dB, pitchB,
                               frameStrideB, dMap, mapPitch, mapFrameStride, dSums, dSsim);
}

int ssim_cuda_compute_u16(int device, uint32_t width, uint32_t height, const uint16_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                          const uint16_t* b, ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep, ptrdiff_t mapStride, float* ssim)
{
    if (ssim == nullptr && map == nullptr) return fail(EINVAL, "both ssim and map are NULL, nothing would be computed");
    if (a == nullptr || b == nullptr) return fail(EINVAL, "image pointer is NULL");
    if (width == 0 || height == 0) return fail(EINVAL, "width and height must be non-zero");
    if (((uintptr_t)a | (uintptr_t)b) & 1) return fail(EINVAL, "16-bit images must be 2-byte aligned");
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    return compute_general_u16(c, width, height, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride, ssim);
}

int ssim_cuda_compute_device_u16(int device, void* stream, uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows, uint32_t frames,
                                 const uint16_t* dA, size_t pitchA, size_t frameStrideA, const uint16_t* dB, size_t pitchB, size_t frameStrideB,
                                 float* dMap, size_t mapPitch, size_t mapFrameStride, double* dSums, float* dSsim)
{
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    DEVICE_GUARD(device);
    return compute_device_impl(c, (cudaStream_t)stream, width, srcRows, outY0, outRows, frames, (const uint8_t*)dA, pitchA, frameStrideA,
                               (const uint8_t*)dB, pitchB, frameStrideB, dMap, mapPitch, mapFrameStride, dSums, dSsim, 2);
}

// ---- strip sums exchanged over peer memory (NVLink) inside the fused kernel
int ssim_cuda_exchange_create(int device, void** dBuf, void* ipcHandle64)
{
    if (!dBuf) return fail(EINVAL, "dBuf is NULL");
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    if ((rc = exchange_alloc(device, dBuf))) return rc;
    if (ipcHandle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        DEVICE_GUARD(device);
        cudaIpcMemHandle_t h;
        CU_TRY(cudaIpcGetMemHandle(&h, *dBuf));
        memcpy(ipcHandle64, &h, sizeof(h));
    }
    return 0;
}

int ssim_cuda_exchange_open(int device, const void* ipcHandle64, void** dPeerBuf)
{
    if (!ipcHandle64 || !dPeerBuf) return fail(EINVAL, "handle or output pointer is NULL");
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    DEVICE_GUARD(device);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipcHandle64, sizeof(h));
    CU_TRY(cudaIpcOpenMemHandle(dPeerBuf, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int ssim_cuda_exchange_close(int device, void* dPeerBuf)
{
    DEVICE_GUARD(device);
    CU_TRY(cudaIpcCloseMemHandle(dPeerBuf));
    return 0;
}

int ssim_cuda_exchange_destroy(int device, void* dBuf)
{
    DEVICE_GUARD(device);
    CU_TRY(cudaFree(dBuf));
    return 0;
}

int ssim_cuda_exchange_enable_peer(int device, int peerDevice) { return enable_peer(device, peerDevice); }

int ssim_cuda_compute_strip_allreduce(int device, void* stream, uint32_t width, uint32_t srcRows, uint32_t outY0, uint32_t outRows,
                                      uint32_t imageRows, const uint8_t* dA, size_t pitchA, const uint8_t* dB, size_t pitchB,
                                      float* dMap, size_t mapPitch, void* const* peerBufs, int world, int rank, uint64_t epoch,
                                      double* dSumAll, float* dSsimAll, int* dStatus)
{
    if (!peerBufs || !dSumAll) return fail(EINVAL, "peerBufs or dSumAll is NULL");
    if (world < 1 || world > ssimk::kMaxRanks || rank < 0 || rank >= world) return fail(EINVAL, "world %d / rank %d out of range (max %d ranks)", world, rank, ssimk::kMaxRanks);
    if (epoch == 0) return fail(EINVAL, "epoch must be >= 1");
    for (int r = 0; r < world; ++r) if (!peerBufs[r]) return fail(EINVAL, "peerBufs[%d] is NULL", r);
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    DEVICE_GUARD(device);
    ssimk::ExchangeParams x;
    memset(&x, 0, sizeof(x));
    for (int r = 0; r < world; ++r) x.peers[r] = (ssimk::ExchangeSlot*)peerBufs[r];
    x.world = world; x.rank = rank; x.epoch = epoch;
    static const unsigned long long timeoutMs = [] { const char* e = getenv("SSIM_CUDA_EXCHANGE_TIMEOUT_MS"); return e ? (unsigned long long)atoll(e) : 2000ull; }();
    x.timeoutNs = timeoutMs * 1000000ull;
    x.sumAll = dSumAll; x.ssimAll = dSsimAll; x.status = dStatus;
    x.invCountAll = 1.0 / (double)(uint32_t)(width * imageRows);       // uint32 product, as src/ssim.cpp:1102
    return compute_device_impl(c, (cudaStream_t)stream, width, srcRows, outY0, outRows, 1, dA, pitchA, 0, dB, pitchB, 0, dMap, mapPitch, 0,
                               nullptr, nullptr, 1, &x);
}

int ssim_cuda_last_launch_count(void) { return g_lastLaunches; }

int ssim_cuda_compute_strips(int nDevices, const int* devices, uint32_t width, uint32_t height, const uint8_t* a, ptrdiff_t stepA,
                             ptrdiff_t strideA, const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB, float* map, ptrdiff_t mapStep,
                             ptrdiff_t mapStride, float* ssim)
{
    if (nDevices < 1 || nDevices > 64 || devices == nullptr) return fail(EINVAL, "bad device list");
    if (ssim == nullptr && map == nullptr) return fail(EINVAL, "both ssim and map are NULL, nothing would be computed");
    if (a == nullptr || b == nullptr) return fail(EINVAL, "image pointer is NULL");
    if (width == 0 || height == 0) return fail(EINVAL, "width and height must be non-zero");
    return compute_strips(nDevices, devices, width, height, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride, ssim);
}

int ssim_cuda_synth_fill(int device, void* stream, uint8_t* dA, size_t pitchA, uint8_t* dB, size_t pitchB, uint32_t width, uint32_t rows,
                         uint32_t y0, uint32_t frame, uint64_t seed)
{
    Context* c;
    int rc = get_context(device, &c);
    if (rc) return rc;
    if (!dA || !dB || width == 0 || rows == 0) return fail(EINVAL, "bad synth_fill arguments");
    DEVICE_GUARD(device);
    CU_TRY(ssimk::launch_synth_fill((cudaStream_t)stream, dA, (long long)pitchA, dB, (long long)pitchB, (int)width, (int)rows, (int)y0, frame, seed));
    return 0;
}

void ssim_cuda_debug_slot_times(unsigned long long* dTimes) { g_dbgTimes.store(dTimes, std::memory_order_relaxed); }

int ssim_cuda_debug_map_cache_entry(const void* base, uint32_t width, uint32_t rows, uint32_t frames, size_t pitch, int elemBytes)
{
    return map_cache_entry((const uint8_t*)base, width, rows, frames, pitch, elemBytes);
}

void ssim_cuda_set_tuning(int maxPairsPerSm, int minSlotRows)
{
    g_maxPairsPerSm.store(maxPairsPerSm > 0 ? maxPairsPerSm : 0, std::memory_order_relaxed);
    g_minSlotRows.store(minSlotRows > 0 ? minSlotRows : 0, std::memory_order_relaxed);
}

}  // extern "C"
