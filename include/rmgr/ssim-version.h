/* rmgr/ssim-version.h -- version of the rmgr::ssim API this library is a drop-in for.
 * Replaces the file the reference generates from src/ssim-version.h.in:25-28 (2.1.0, CMakeLists.txt:42-45). */
#ifndef RMGR_SSIM_VERSION_H
#define RMGR_SSIM_VERSION_H

#define RMGR_SSIM_VERSION_MAJOR   (2)
#define RMGR_SSIM_VERSION_MINOR   (1)
#define RMGR_SSIM_VERSION_PATCH   (0)
#define RMGR_SSIM_VERSION_STRING  "2.1.0"

/* identifies the implementation behind the API (not present in the reference) */
#define RMGR_SSIM_BACKEND_STRING  "ssim_b200 (CUDA sm_100a)"

#endif
