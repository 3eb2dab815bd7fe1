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

#ifndef RMGR_SSIM_REPORT_ERROR
    #ifndef NDEBUG
        #define RMGR_SSIM_REPORT_ERROR(...)  fprintf(stderr, __VA_ARGS__)
    #else
        #define RMGR_SSIM_REPORT_ERROR(...)
    #endif
#endif

namespace
{

void* default_alloc(size_t size, size_t alignment) RMGR_NOEXCEPT
{
    void* address = NULL;
    return (::posix_memalign(&address, alignment, size) == 0) ? address : NULL;
}

void default_dealloc(void* address) RMGR_NOEXCEPT
{
    ::free(address);
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

} // namespace


extern "C" rmgr_int32_t rmgr_ssim_init_interleaved(rmgr_ssim_ImgParams* params, const rmgr_uint8_t* data, ptrdiff_t imgStride, rmgr_uint32_t channelCount, rmgr_uint32_t channelNum) RMGR_NOEXCEPT
{
    if (params == NULL || data == NULL || channelNum >= channelCount)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: params/data cannot be NULL and channelNum must be < channelCount\n");
        return EINVAL;
    }
    params->topLeft = data + channelNum;
    params->step    = ptrdiff_t(channelCount);
    params->stride  = imgStride;
    return 0;
}


extern "C" rmgr_int32_t rmgr_ssim_init_planar(rmgr_ssim_ImgParams* params, rmgr_uint8_t const* const planes[], const ptrdiff_t strides[], rmgr_uint32_t planeNum) RMGR_NOEXCEPT
{
    if (params == NULL || planes == NULL || planes[planeNum] == NULL || strides == NULL)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: params, planes, planes[planeNum] and strides cannot be NULL\n");
        return EINVAL;
    }
    params->topLeft = planes[planeNum];
    params->step    = 1;
    params->stride  = strides[planeNum];
    return 0;
}


extern "C" rmgr_int32_t rmgr_ssim_use_default_allocator(rmgr_ssim_Params* params) RMGR_NOEXCEPT
{
    if (params == NULL)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: params cannot be NULL\n");
        return EINVAL;
    }
    params->alloc   = default_alloc;
    params->dealloc = default_dealloc;
    return 0;
}


extern "C" rmgr_int32_t rmgr_ssim_get_version(rmgr_ssim_Version* version) RMGR_NOEXCEPT
{
    static const char versionString[] = RMGR_SSIM_VERSION_STRING;
    if (version == NULL)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: version cannot be NULL\n");
        return EINVAL;
    }
    version->major  = RMGR_SSIM_VERSION_MAJOR;
    version->minor  = RMGR_SSIM_VERSION_MINOR;
    version->patch  = RMGR_SSIM_VERSION_PATCH;
    version->string = versionString;
    return 0;
}


namespace rmgr { namespace ssim
{

int32_t compute_ssim(float* ssim, const GeneralParams& params, const ThreadPool* threadPool) RMGR_NOEXCEPT
{
    // Same checks, same order as the reference (src/ssim.cpp:962-978)
    if (ssim == NULL && params.ssimMap == NULL)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameters: both ssim and ssimMap are NULL, nothing will be computed\n");
        return EINVAL;
    }
    if (params.imgA.topLeft == NULL || params.imgB.topLeft == NULL)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: imgA.topLeft or imgB.topLeft is NULL\n");
        return EINVAL;
    }
    if (threadPool != NULL && threadPool->dispatch != NULL && threadPool->threadCount == 0u)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: threadCount cannot be 0 if threadPool is not NULL\n");
        return EINVAL;
    }
    // Divergence (documented in the header): the reference returns garbage for empty images
    if (params.width == 0 || params.height == 0)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: width and height cannot be 0\n");
        return EINVAL;
    }

    // With no map, ssimStep/ssimStride are ignored (src/ssim.cpp:980-987).  The thread pool and the allocation
    // hooks are not used: tiles are scheduled on the GPU grid and scratch lives in device memory.
    const int32_t result = ::ssim_cuda_compute(selected_device(), params.width, params.height,
                                               params.imgA.topLeft, params.imgA.step, params.imgA.stride,
                                               params.imgB.topLeft, params.imgB.step, params.imgB.stride,
                                               params.ssimMap, params.ssimMap ? params.ssimStep : 0, params.ssimMap ? params.ssimStride : 0,
                                               ssim);
    if (result != 0)
        RMGR_SSIM_REPORT_ERROR("ssim_cuda: %s\n", ::ssim_cuda_last_error_string());
    return result;
}


float compute_ssim(const Params& params) RMGR_NOEXCEPT
{
    // src/ssim.cpp:1109-1120
    ThreadPool threadPool;
    threadPool.dispatch    = params.threadPool;
    threadPool.context     = params.threadPoolContext;
    threadPool.threadCount = params.threadCount;

    float ssim = 0.0f;
    const int32_t result = compute_ssim(&ssim, params, &threadPool);
    return (result == 0) ? ssim : float(-result);
}

}} // namespace rmgr::ssim


extern "C" rmgr_int32_t rmgr_ssim_compute_ssim(float* ssim, const rmgr_ssim_Params* params, const rmgr_ssim_ThreadPool* threadPool) RMGR_NOEXCEPT
{
    if (params == NULL)
    {
        RMGR_SSIM_REPORT_ERROR("Invalid parameter: params cannot be NULL\n");
        return EINVAL;
    }
    return rmgr::ssim::compute_ssim(ssim, *params, threadPool);
}


extern "C" rmgr_int32_t rmgr_ssim_compute_ssim_openmp(float* ssim, const rmgr_ssim_Params* params) RMGR_NOEXCEPT
{
    // src/ssim-openmp.c:40-47 builds an OpenMP thread pool here; the GPU grid takes that role
    return rmgr_ssim_compute_ssim(ssim, params, NULL);
}
