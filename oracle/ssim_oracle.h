/*
 * oracle/ssim_oracle.h -- TEST INFRASTRUCTURE ONLY (see ssim_oracle.c).
 * CPU restatement of rmgr::ssim::compute_ssim() (reference src/ssim.cpp:933-1106) in double.
 */
#ifndef SSIM_ORACLE_H
#define SSIM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSIM_ORACLE_TAPS_RUNTIME 0 /* taps in double at run time: reference double build, IMPL_GENERIC ("O2") */
#define SSIM_ORACLE_TAPS_TABLE   1 /* float-pipeline taps promoted to double: reference double build, default dispatch ("O1") */

/* 121 normalised Gaussian taps, row-major 11x11 (src/ssim.cpp:281-318) */
void ssim_oracle_taps(double taps[121], int mode);

/* returns 0 / EINVAL / ENOMEM like the reference (src/ssim.cpp:962-978,1051-1052) */
int ssim_oracle_compute(uint32_t width, uint32_t height,
                        const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                        const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                        float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                        int tapsMode, float* ssim, double* sumOut);

/* 16-bit pixels (L = 65535); step/stride in uint16 elements.  Parity unpinned against the reference (not implemented there). */
int ssim_oracle_compute_u16(uint32_t width, uint32_t height,
                            const uint16_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                            const uint16_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                            float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                            int tapsMode, float* ssim, double* sumOut);

int ssim_oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
