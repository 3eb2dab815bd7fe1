// rmgr_api.cpp -- the reference's public C / C++ API (include/rmgr/ssim.h, ssim-openmp.h) implemented on
// top of the C ABI of libssim_cuda (include/ssim_cuda.h).  Parameter validation follows the order and the
// errno conventions of the reference's compute_ssim() (reference src/ssim.cpp:933-978) and helpers
// (src/ssim.cpp:156-217,1126-1154); the computation itself is one call into the shim.
#include <rmgr/ssim.h>
#include <rmgr/ssim-openmp.h>
#include <rmgr/ssim-version.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>

#include "ssim_cuda.h"

// Diagnostics go to stderr in debug builds only, through the same user-overridable macro name as the reference
// (src/ssim.cpp:37-43), so a project that already redirects RMGR_SSIM_REPORT_ERROR keeps doing so.
#if !defined(RMGR_SSIM_REPORT_ERROR)
#   if defined(NDEBUG)
#       define RMGR_SSIM_REPORT_ERROR(...) ((void)0)
#   else
#       define RMGR_SSIM_REPORT_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#   endif
#endif

namespace
{

// rejects a call with EINVAL and says why (debug builds)
rmgr_int32_t reject(const char* what) RMGR_NOEXCEPT
{
    RMGR_SSIM_REPORT_ERROR("rmgr-ssim (CUDA): invalid argument: %s\n", what);
    (void)what;
    return EINVAL;
}

// The allocation hooks are part of the reference's Params; the GPU build accepts them and never calls them (scratch is
// device memory), but use_default_allocator() still has to install a working pair (src/ssim.cpp:206-217).
void* aligned_new(size_t bytes, size_t alignment) RMGR_NOEXCEPT
{
    void* p = NULL;
    if (::posix_memalign(&p, alignment, bytes) != 0)
        p = NULL;
    return p;
}

void aligned_delete(void* p) RMGR_NOEXCEPT
{
    ::free(p);
}

// SSIM_CUDA_DEVICE selects the GPU used by the drop-in API (default 0)
int selected_device() RMGR_NOEXCEPT
{
    static const int device = [] {
        const char* env = ::getenv("SSIM_CUDA_DEVICE");
        return env ? ::atoi(env) : 0;
    }();
    return device;
}

// one image description = pointer to pixel (0,0), byte distance between pixels, byte distance between rows
void describe(rmgr_ssim_ImgParams& img, const rmgr_uint8_t* origin, ptrdiff_t step, ptrdiff_t stride) RMGR_NOEXCEPT
{
    img.topLeft = origin;
    img.step    = step;
    img.stride  = stride;
}

} // namespace


// channel `channelNum` of an interleaved image (reference src/ssim.cpp:156-178)
extern "C" rmgr_int32_t rmgr_ssim_init_interleaved(rmgr_ssim_ImgParams* params, const rmgr_uint8_t* data, ptrdiff_t imgStride, rmgr_uint32_t channelCount, rmgr_uint32_t channelNum) RMGR_NOEXCEPT
{
    if (!params || !data)
        return reject("init_interleaved: params and data must not be NULL");
    if (channelNum >= channelCount)
        return reject("init_interleaved: channelNum must be below channelCount");
    describe(*params, data + channelNum, (ptrdiff_t)channelCount, imgStride);
    return 0;
}


// plane `planeNum` of a planar image (reference src/ssim.cpp:181-203)
extern "C" rmgr_int32_t rmgr_ssim_init_planar(rmgr_ssim_ImgParams* params, rmgr_uint8_t const* const planes[], const ptrdiff_t strides[], rmgr_uint32_t planeNum) RMGR_NOEXCEPT
{
    if (!params || !planes || !strides || !planes[planeNum])
        return reject("init_planar: params, planes, strides and the selected plane must not be NULL");
    describe(*params, planes[planeNum], 1, strides[planeNum]);
    return 0;
}


extern "C" rmgr_int32_t rmgr_ssim_use_default_allocator(rmgr_ssim_Params* params) RMGR_NOEXCEPT
{
    if (!params)
        return reject("use_default_allocator: params must not be NULL");
    params->alloc   = aligned_new;
    params->dealloc = aligned_delete;
    return 0;
}


extern "C" rmgr_int32_t rmgr_ssim_get_version(rmgr_ssim_Version* version) RMGR_NOEXCEPT
{
    static const char text[] = RMGR_SSIM_VERSION_STRING;
    if (!version)
        return reject("get_version: version must not be NULL");
    const rmgr_ssim_Version v = {RMGR_SSIM_VERSION_MAJOR, RMGR_SSIM_VERSION_MINOR, RMGR_SSIM_VERSION_PATCH, text};
    *version = v;
    return 0;
}


namespace rmgr { namespace ssim
{

int32_t compute_ssim(float* ssim, const GeneralParams& params, const ThreadPool* threadPool) RMGR_NOEXCEPT
{
    // Same checks in the same order as the reference (src/ssim.cpp:962-978), so a caller sees the same errno for a
    // call that is wrong in several ways at once.
    const bool wantsMap = params.ssimMap != NULL;
    if (!ssim && !wantsMap)
        return reject("compute_ssim: neither a global result nor a map was requested");
    if (!params.imgA.topLeft || !params.imgB.topLeft)
        return reject("compute_ssim: an image pointer is NULL");
    if (threadPool && threadPool->dispatch && threadPool->threadCount == 0)
        return reject("compute_ssim: a thread pool with a dispatch function needs threadCount > 0");
    // Divergence (documented in the header): the reference returns garbage for empty images
    if (params.width == 0 || params.height == 0)
        return reject("compute_ssim: width and height must be non-zero");

    // With no map, ssimStep/ssimStride are ignored (src/ssim.cpp:980-987).  The thread pool and the allocation
    // hooks are not used: tiles are scheduled on the GPU grid and scratch lives in device memory.
    const int32_t rc = ::ssim_cuda_compute(selected_device(), params.width, params.height,
                                           params.imgA.topLeft, params.imgA.step, params.imgA.stride,
                                           params.imgB.topLeft, params.imgB.step, params.imgB.stride,
                                           params.ssimMap, wantsMap ? params.ssimStep : 0, wantsMap ? params.ssimStride : 0,
                                           ssim);
    if (rc != 0)
        RMGR_SSIM_REPORT_ERROR("rmgr-ssim (CUDA): %s\n", ::ssim_cuda_last_error_string());
    return rc;
}


// deprecated overload: threaded parameters in one struct, result or -errno as a float (src/ssim.cpp:1109-1120)
float compute_ssim(const Params& params) RMGR_NOEXCEPT
{
    const ThreadPool pool = {params.threadPool, params.threadPoolContext, params.threadCount};
    float value = 0.0f;
    const int32_t rc = compute_ssim(&value, params, &pool);
    return rc ? -float(rc) : value;
}

}} // namespace rmgr::ssim


extern "C" rmgr_int32_t rmgr_ssim_compute_ssim(float* ssim, const rmgr_ssim_Params* params, const rmgr_ssim_ThreadPool* threadPool) RMGR_NOEXCEPT
{
    if (!params)
        return reject("compute_ssim: params must not be NULL");
    return rmgr::ssim::compute_ssim(ssim, *params, threadPool);
}


extern "C" rmgr_int32_t rmgr_ssim_compute_ssim_openmp(float* ssim, const rmgr_ssim_Params* params) RMGR_NOEXCEPT
{
    // src/ssim-openmp.c:40-47 builds an OpenMP thread pool here; the GPU grid takes that role
    return rmgr_ssim_compute_ssim(ssim, params, NULL);
}
