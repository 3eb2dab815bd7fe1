// oracle/naive_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" entry points around the reference's OWN test oracle, tests/ssim_naive.h (header-only, template on the float
// type F and the pixel type T; dynamic range L = numeric_limits<T>::max(), tests/ssim_naive.h:230-241), compiled in place
// from /root/reference by oracle/Makefile into oracle/_ref/libnaive.so.  Nothing of it is copied into this repository.
// naive::compute_ssim<double, uint16_t> is the reference-held 16-bit (L = 65535) implementation the 16-bit path of this
// repository is pinned to; <double, uint8_t> is exported next to it so that the same shim can be cross-checked against
// the reference's einstein known answers (tests/rmgr-ssim-tests.cpp:354-359 were produced with it).
#include <stdint.h>
#include <stddef.h>
#include "ssim_naive.h"      // found through -I$(REF)/tests at build time

extern "C" double naive_ssim_u8(uint32_t width, uint32_t height, const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                                const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB, double* map, ptrdiff_t mapStep, ptrdiff_t mapStride)
{
    return rmgr::ssim::naive::compute_ssim<double, uint8_t>(width, height, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride);
}

extern "C" double naive_ssim_u16(uint32_t width, uint32_t height, const uint16_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                                 const uint16_t* b, ptrdiff_t stepB, ptrdiff_t strideB, double* map, ptrdiff_t mapStep, ptrdiff_t mapStride)
{
    return rmgr::ssim::naive::compute_ssim<double, uint16_t>(width, height, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride);
}
