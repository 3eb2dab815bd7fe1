/*
 * rmgr/ssim-openmp.h -- the reference's multi-threaded entry points (reference include/rmgr/ssim-openmp.h:50,77,98;
 * src/ssim-openmp.c:26-47), kept so that callers compile unchanged.  In this implementation the
 * parallelism comes from the CUDA grid, so these are the same GPU path as rmgr_ssim_compute_ssim().
 */
#ifndef RMGR_SSIM_OPENMP_H
#define RMGR_SSIM_OPENMP_H

#include <rmgr/ssim.h>

#ifdef __cplusplus
extern "C"
{
#endif

/* Same contract and return values as rmgr_ssim_compute_ssim(ssim, params, NULL) */
rmgr_int32_t rmgr_ssim_compute_ssim_openmp(float* ssim, const rmgr_ssim_Params* params) RMGR_NOEXCEPT;

#ifdef __cplusplus
} /* extern "C" */

namespace rmgr { namespace ssim
{

inline int32_t compute_ssim_openmp(float* ssim, const GeneralParams& params) RMGR_NOEXCEPT
{
    return ::rmgr_ssim_compute_ssim_openmp(ssim, &params);
}

RMGR_DEPRECATED_MSG("deprecated overload: call compute_ssim_openmp(&ssim, generalParams) and test its return code")
inline float compute_ssim_openmp(const UnthreadedParams& params) RMGR_NOEXCEPT
{
    float ssim;
    const int32_t result = compute_ssim_openmp(&ssim, params);
    return (result == 0) ? ssim : float(-result);
}

}} /* namespace rmgr::ssim */
#endif /* __cplusplus */

#endif /* RMGR_SSIM_OPENMP_H */
