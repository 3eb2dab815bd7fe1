/*
 * rmgr/ssim.h -- C and C++ API of rmgr::ssim, re-declared for the B200-native implementation.
 *
 * This header declares, with identical names, argument meaning, struct layouts (LP64) and error
 * conventions, the public interface of the reference (reference include/rmgr/ssim.h:428-731), so
 * that code written against the reference compiles and links unchanged against librmgr-ssim.so
 * from this repository.  Everything below the API -- the reference's ISA dispatch (src/ssim.cpp:798-896),
 * its 256x64 tile driver (src/ssim.cpp:747-791,933-1106) and the OpenMP thread pool
 * (src/ssim-openmp.c) -- is replaced by one fused sm_100a CUDA kernel behind libssim_cuda
 * (include/ssim_cuda.h).  There is no CPU fallback: without a usable CUDA device the compute calls
 * return ENODEV.
 *
 * Units: image `step`/`stride` are in BYTES, map `ssimStep`/`ssimStride` in FLOATS; all may be negative.
 */
#ifndef RMGR_SSIM_H
#define RMGR_SSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
    #if __cplusplus >= 201103L
        #define RMGR_NOEXCEPT  noexcept
    #else
        #define RMGR_NOEXCEPT  throw()
    #endif
    #if __cplusplus >= 201703L
        #define RMGR_NOEXCEPT_TYPEDEF  noexcept   /* exception specifications are part of the type since C++17 */
    #else
        #define RMGR_NOEXCEPT_TYPEDEF
    #endif
#else
    #define RMGR_NOEXCEPT
    #define RMGR_NOEXCEPT_TYPEDEF
#endif

#if defined(__cplusplus) && __cplusplus >= 201402L
    #define RMGR_DEPRECATED_MSG(msg)  [[deprecated(msg)]]
#elif defined(__GNUC__)
    #define RMGR_DEPRECATED_MSG(msg)  __attribute__((deprecated(msg)))
#else
    #define RMGR_DEPRECATED_MSG(msg)
#endif

typedef uint8_t   rmgr_uint8_t;
typedef int32_t   rmgr_int32_t;
typedef uint32_t  rmgr_uint32_t;
typedef uint64_t  rmgr_uint64_t;


/*=================================================================================================
 * C API */

#ifdef __cplusplus
extern "C"
{
#endif

/* Scratch allocation hooks (reference ssim.h:438-439).  The CUDA implementation keeps its scratch in
 * device memory, so these are accepted and validated but never called. */
typedef void* (*rmgr_ssim_AllocFct)(size_t size, size_t alignment) RMGR_NOEXCEPT_TYPEDEF;
typedef void  (*rmgr_ssim_DeallocFct)(void* address) RMGR_NOEXCEPT_TYPEDEF;

/* Thread-pool plug-in types (reference ssim.h:448,466).  Work is scheduled on the GPU grid instead, so a
 * pool is validated (threadCount must not be 0 when dispatch is set) but never dispatched to. */
typedef void         (*rmgr_ssim_ThreadFct)(void* arg, rmgr_uint32_t jobNum) RMGR_NOEXCEPT_TYPEDEF;
typedef rmgr_int32_t (*rmgr_ssim_ThreadPoolFct)(void* context, rmgr_ssim_ThreadFct fct, void* const args[], rmgr_uint32_t threadCount, rmgr_uint32_t jobCount) RMGR_NOEXCEPT_TYPEDEF;

typedef struct rmgr_ssim_Version_
{
    rmgr_uint32_t major;
    rmgr_uint32_t minor;
    rmgr_uint32_t patch;
    const char*   string;
} rmgr_ssim_Version;

/* One channel of one image: address of pixel (x,y) = topLeft + x*step + y*stride  (reference ssim.h:489-499) */
typedef struct rmgr_ssim_ImgParams_
{
    const rmgr_uint8_t* topLeft; /* the considered channel of the top-left pixel */
    ptrdiff_t           step;    /* bytes between a pixel and its right neighbour */
    ptrdiff_t           stride;  /* bytes between a pixel and the one below it    */

#ifdef __cplusplus
    rmgr_int32_t init_interleaved(const rmgr_uint8_t* data, ptrdiff_t imgStride, rmgr_uint32_t channelCount, rmgr_uint32_t channelNum) RMGR_NOEXCEPT;
    rmgr_int32_t init_planar(rmgr_uint8_t const* const planes[], const ptrdiff_t strides[], rmgr_uint32_t planeNum) RMGR_NOEXCEPT;
#endif
} rmgr_ssim_ImgParams;

/* All non-threading parameters (reference ssim.h:505-525); 96 bytes on LP64 */
typedef struct rmgr_ssim_Params_
{
    rmgr_uint32_t        width;      /* pixels */
    rmgr_uint32_t        height;     /* pixels */
    rmgr_ssim_ImgParams  imgA;
    rmgr_ssim_ImgParams  imgB;

    float*               ssimMap;    /* top-left of the per-pixel SSIM map, or NULL */
    ptrdiff_t            ssimStep;   /* floats between horizontally adjacent map pixels */
    ptrdiff_t            ssimStride; /* floats between vertically adjacent map pixels (negative: bottom-up) */

    rmgr_ssim_AllocFct   alloc;      /* accepted, unused (see above) */
    rmgr_ssim_DeallocFct dealloc;

#ifdef __cplusplus
    void use_default_allocator() RMGR_NOEXCEPT;
#endif
} rmgr_ssim_Params;

typedef struct rmgr_ssim_ThreadPool_
{
    rmgr_ssim_ThreadPoolFct dispatch;
    void*                   context;
    rmgr_uint32_t           threadCount;
} rmgr_ssim_ThreadPool;

/* 0, or EINVAL if version is NULL                                           (reference ssim.h:544) */
rmgr_int32_t rmgr_ssim_get_version(rmgr_ssim_Version* version) RMGR_NOEXCEPT;

/* topLeft = data+channelNum, step = channelCount, stride = imgStride; EINVAL on NULL params/data or
 * channelNum >= channelCount                                                (reference ssim.h:560) */
rmgr_int32_t rmgr_ssim_init_interleaved(rmgr_ssim_ImgParams* params, const rmgr_uint8_t* data, ptrdiff_t imgStride, rmgr_uint32_t channelCount, rmgr_uint32_t channelNum) RMGR_NOEXCEPT;

/* topLeft = planes[planeNum], step = 1, stride = strides[planeNum]          (reference ssim.h:575) */
rmgr_int32_t rmgr_ssim_init_planar(rmgr_ssim_ImgParams* params, rmgr_uint8_t const* const planes[], const ptrdiff_t strides[], rmgr_uint32_t planeNum) RMGR_NOEXCEPT;

/* installs malloc/free-style hooks; EINVAL if params is NULL                (reference ssim.h:584) */
rmgr_int32_t rmgr_ssim_use_default_allocator(rmgr_ssim_Params* params) RMGR_NOEXCEPT;

/*
 * Global SSIM of one channel of two images and/or the per-pixel map       (reference ssim.h:605).
 *   ssim       receives the global SSIM, may be NULL
 *   params     must not be NULL
 *   threadPool may be NULL
 * Returns 0, EINVAL (both outputs NULL, a NULL topLeft, dispatch set with threadCount 0, NULL params,
 * and -- divergence from the reference, which returns garbage -- width or height of 0), ENOMEM
 * (host or device allocation failed), ENODEV (no usable CUDA device) or EIO (CUDA runtime failure).
 * Image and map pointers may be host pointers or CUDA device pointers (detected per pointer).
 */
rmgr_int32_t rmgr_ssim_compute_ssim(float* ssim, const rmgr_ssim_Params* params, const rmgr_ssim_ThreadPool* threadPool) RMGR_NOEXCEPT;

#ifdef __cplusplus
} /* extern "C" */
#endif


/*=================================================================================================
 * C++ API */

#ifdef __cplusplus

inline rmgr_int32_t rmgr_ssim_ImgParams::init_interleaved(const rmgr_uint8_t* data, ptrdiff_t imgStride, rmgr_uint32_t channelCount, rmgr_uint32_t channelNum) RMGR_NOEXCEPT
{
    return ::rmgr_ssim_init_interleaved(this, data, imgStride, channelCount, channelNum);
}

inline rmgr_int32_t rmgr_ssim_ImgParams::init_planar(rmgr_uint8_t const* const planes[], const ptrdiff_t strides[], rmgr_uint32_t planeNum) RMGR_NOEXCEPT
{
    return ::rmgr_ssim_init_planar(this, planes, strides, planeNum);
}

inline void rmgr_ssim_Params::use_default_allocator() RMGR_NOEXCEPT
{
    ::rmgr_ssim_use_default_allocator(this);
}

namespace rmgr { namespace ssim
{

typedef ::rmgr_uint8_t             uint8_t;
typedef ::rmgr_int32_t             int32_t;
typedef ::rmgr_uint32_t            uint32_t;
typedef ::rmgr_uint64_t            uint64_t;
typedef ::rmgr_ssim_AllocFct       AllocFct;
typedef ::rmgr_ssim_DeallocFct     DeallocFct;
typedef ::rmgr_ssim_ThreadFct      ThreadFct;
typedef ::rmgr_ssim_ThreadPoolFct  ThreadPoolFct;
typedef ::rmgr_ssim_Version        Version;
typedef ::rmgr_ssim_ImgParams      ImgParams;
typedef ::rmgr_ssim_Params         GeneralParams;
typedef ::rmgr_ssim_ThreadPool     ThreadPool;
typedef GeneralParams              UnthreadedParams;

inline Version get_version() RMGR_NOEXCEPT                                    /* reference ssim.h:660 */
{
    Version version;
    ::rmgr_ssim_get_version(&version);
    return version;
}

/* Same contract as rmgr_ssim_compute_ssim()                                     (reference ssim.h:686) */
int32_t compute_ssim(float* ssim, const GeneralParams& params, const ThreadPool* threadPool=NULL) RMGR_NOEXCEPT;

/* Full parameter set of the deprecated overload; 120 bytes on LP64             (reference ssim.h:692-697) */
struct Params: public rmgr_ssim_Params_
{
    ThreadPoolFct  threadPool;
    void*          threadPoolContext;
    uint32_t       threadCount;
};

/* Returns the SSIM, or float(-errno) on error                                  (reference ssim.h:713) */
RMGR_DEPRECATED_MSG("deprecated overload: call compute_ssim(&ssim, generalParams, threadPool) and test its return code")
float compute_ssim(const Params& params) RMGR_NOEXCEPT;

RMGR_DEPRECATED_MSG("only meaningful with the deprecated float compute_ssim(const Params&): the new overload returns the error code itself")
inline int32_t get_errno(float ssim) RMGR_NOEXCEPT                              /* reference ssim.h:725 */
{
    return (ssim>=0) ? 0 : -int32_t(ssim);
}

}} /* namespace rmgr::ssim */
#endif /* __cplusplus */

#endif /* RMGR_SSIM_H */
