// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// A few extern "C" entry points linked into oracle/_ref/libref_f{32,64}.so next to the UNMODIFIED
// reference sources (compiled in place from /root/reference by oracle/Makefile), so that ctypes can
// reach the reference's test-only implementation switch (src/ssim_internal.h:41-53) and learn which
// precision a given build uses.  The reference's own C API (rmgr_ssim_compute_ssim,
// rmgr_ssim_compute_ssim_openmp) is exported by the reference sources themselves.
#include <rmgr/ssim.h>
#include "ssim_internal.h"   // found through -I$(REF)/src at build time; never copied into this repo

extern "C" unsigned ref_select_impl(int impl)
{
    return rmgr::ssim::select_impl(static_cast<rmgr::ssim::Implementation>(impl));
}

extern "C" int ref_uses_double(void)
{
    return RMGR_SSIM_USE_DOUBLE;
}

extern "C" int ref_float_size(void)
{
    return int(sizeof(rmgr::ssim::Float));
}
