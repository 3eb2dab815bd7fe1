/*
 * oracle/ssim_oracle.c  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A plain-C, double-precision CPU restatement of the one hot path this repository
 * accelerates: rmgr::ssim::compute_ssim() (reference: src/ssim.cpp:933-1106).
 * It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
 * check the CUDA path; nothing under ssim_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (1) the six decoder-independent "einstein" goldens of the reference's own test-suite
 *       (tests/rmgr-ssim-tests.cpp:354-359) to <= 1e-13,
 *   (2) the unmodified reference compiled from /root/reference into oracle/_ref/
 *       (libref_f64.so, both IMPL_GENERIC and AUTO dispatch; libref_f32.so as the baseline),
 *   (3) golden vectors generated from that build and committed under tests/golden/.
 *
 * The restatement is written as the direct (non-tiled, non-factored) form of the algorithm:
 * the reference's 256x64 tiling, 4-fold-symmetry factoring and ISA dispatch only change the
 * order of floating-point additions, which in double precision moves results by ~1e-16.
 *
 * Two coefficient modes reproduce the two double-precision behaviours of the reference:
 *   SSIM_ORACLE_TAPS_RUNTIME (0): taps computed at run time in double, as the generic blur
 *       consumes them (src/ssim.cpp:272-318)            -> "O2", IMPL_GENERIC of the double build
 *   SSIM_ORACLE_TAPS_TABLE   (1): taps of the *float* pipeline promoted to double, which is what
 *       the literal coefficient table of every SIMD blur holds (src/ssim_fma.cpp:164-175; we
 *       regenerate the values by running the float computation, they are not copied)
 *                                                       -> "O1", default dispatch of the double build
 */
#include <errno.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "ssim_oracle.h"

#define RADIUS 5
#define TAPS   (2 * RADIUS + 1)

/* src/ssim.cpp:272-278 (Float = double): k(x,y) = exp(-(x^2+y^2) / (2 sigma^2)) / (2 pi sigma^2) */
static double tap_f64(int x, int y, double sigma)
{
    const double sigma2 = sigma * sigma;
    return exp(-(double)(x * x + y * y) / (2 * sigma2)) / ((double)(2 * M_PI) * sigma2);
}

/* same expression evaluated in float, as the float build does (Float = float, std::exp -> expf) */
static float tap_f32(int x, int y, float sigma)
{
    const float sigma2 = sigma * sigma;
    return expf(-(float)(x * x + y * y) / (2 * sigma2)) / ((float)(2 * M_PI) * sigma2);
}

/*
 * src/ssim.cpp:281-318: fill the 11x11 kernel, accumulate the sum of all 121 taps in double,
 * divide every tap by Float(sum).  sigma = 1.5, radius = 5 (src/ssim.cpp:227-228).
 */
void ssim_oracle_taps(double taps[TAPS * TAPS], int mode)
{
    double sum = 0.0;
    if (mode == SSIM_ORACLE_TAPS_TABLE) {
        float k[TAPS * TAPS];
        for (int y = 0; y < TAPS; ++y)
            for (int x = 0; x < TAPS; ++x)
                sum += (double)(k[y * TAPS + x] = tap_f32(x - RADIUS, y - RADIUS, 1.5f));
        for (int i = 0; i < TAPS * TAPS; ++i)
            taps[i] = (double)(k[i] / (float)sum);
    } else {
        for (int y = 0; y < TAPS; ++y)
            for (int x = 0; x < TAPS; ++x)
                sum += (taps[y * TAPS + x] = tap_f64(x - RADIUS, y - RADIUS, 1.5));
        for (int i = 0; i < TAPS * TAPS; ++i)
            taps[i] /= sum;
    }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/*
 * The whole path for one image pair.
 *   - pixel address = topLeft + x*step + y*stride, bytes, signed  (include/rmgr/ssim.h:481-499)
 *   - outside the image the nearest edge pixel is replicated, applied to the SOURCE pixels
 *     before squaring                                             (src/ssim.cpp:515-583)
 *   - five windowed moments E[a], E[b], E[a^2], E[b^2], E[ab]     (src/ssim.cpp:760-779)
 *   - sigma_a^2 = E[a^2]-mu_a^2 ..., ssim = ((2 mu_a mu_b + c1)(2 sigma_ab + c2)) /
 *     ((mu_a^2 + mu_b^2 + c1)(sigma_a^2 + sigma_b^2 + c2))        (src/ssim.cpp:590-704)
 *   - c1 = (0.01*255)^2, c2 = (0.03*255)^2                        (src/ssim.cpp:956-960)
 *   - map[x*mapStep + y*mapStride] = float(ssim), units of float  (src/ssim.cpp:661-667,781)
 *   - mean = sum / double(width*height), width*height a uint32 product; returned as float
 *                                                                 (src/ssim.cpp:1102)
 */
/* pixel (x,y) of an image whose step/stride are in ELEMENTS of 1 or 2 bytes */
static double fetch(const void* img, int elemBytes, ptrdiff_t idx)
{
    return elemBytes == 2 ? (double)((const uint16_t*)img)[idx] : (double)((const uint8_t*)img)[idx];
}

static int oracle_core(uint32_t width, uint32_t height, int elemBytes, double L,
                       const void* a, ptrdiff_t stepA, ptrdiff_t strideA,
                       const void* b, ptrdiff_t stepB, ptrdiff_t strideB,
                       float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                       int tapsMode, float* ssim, double* sumOut)
{
    if (ssim == NULL && map == NULL && sumOut == NULL)
        return EINVAL;                                   /* src/ssim.cpp:962-966 */
    if (a == NULL || b == NULL)
        return EINVAL;                                   /* src/ssim.cpp:968-972 */
    if (width == 0 || height == 0)
        return EINVAL;                                   /* documented divergence: the reference does not validate this */

    double taps[TAPS * TAPS];
    ssim_oracle_taps(taps, tapsMode);

    const double c1 = (0.01 * L) * (0.01 * L);           /* src/ssim.cpp:956-960 with L = 255 */
    const double c2 = (0.03 * L) * (0.03 * L);

    const int W = (int)width, H = (int)height;
    const int PW = W + 2 * RADIUS;

    /* widen both images once into clamp-to-edge padded double planes */
    const size_t padded = (size_t)PW * (size_t)(H + 2 * RADIUS);
    double* pa = (double*)malloc(padded * sizeof(double));
    double* pb = (double*)malloc(padded * sizeof(double));
    double* rowSums = (double*)malloc((size_t)H * sizeof(double));
    if (!pa || !pb || !rowSums) { free(pa); free(pb); free(rowSums); return ENOMEM; }

    for (int y = -RADIUS; y < H + RADIUS; ++y) {
        const int sy = clampi(y, 0, H - 1);
        for (int x = -RADIUS; x < W + RADIUS; ++x) {
            const int sx = clampi(x, 0, W - 1);
            const size_t d = (size_t)(y + RADIUS) * PW + (size_t)(x + RADIUS);
            pa[d] = fetch(a, elemBytes, (ptrdiff_t)sx * stepA + (ptrdiff_t)sy * strideA);
            pb[d] = fetch(b, elemBytes, (ptrdiff_t)sx * stepB + (ptrdiff_t)sy * strideB);
        }
    }

    #pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y) {
        double rowSum = 0.0;
        for (int x = 0; x < W; ++x) {
            double muA = 0, muB = 0, eAA = 0, eBB = 0, eAB = 0;
            for (int ky = 0; ky < TAPS; ++ky) {
                const double* ra = pa + (size_t)(y + ky) * PW + x;
                const double* rb = pb + (size_t)(y + ky) * PW + x;
                const double* k  = taps + ky * TAPS;
                for (int kx = 0; kx < TAPS; ++kx) {
                    const double va = ra[kx], vb = rb[kx], w = k[kx];
                    muA += w * va;
                    muB += w * vb;
                    eAA += w * (va * va);
                    eBB += w * (vb * vb);
                    eAB += w * (va * vb);
                }
            }
            const double muA2 = muA * muA, muB2 = muB * muB, muAB = muA * muB;
            const double sA2 = eAA - muA2, sB2 = eBB - muB2, sAB = eAB - muAB;
            const double num = (2 * muAB + c1) * (2 * sAB + c2);
            const double den = (muA2 + muB2 + c1) * (sA2 + sB2 + c2);
            const double s = num / den;
            rowSum += s;
            if (map != NULL)
                map[(ptrdiff_t)x * mapStep + (ptrdiff_t)y * mapStride] = (float)s;
        }
        rowSums[y] = rowSum;
    }

    double sum = 0.0;
    for (int y = 0; y < H; ++y)
        sum += rowSums[y];

    free(pa); free(pb); free(rowSums);

    if (sumOut != NULL)
        *sumOut = sum;
    if (ssim != NULL)
        *ssim = (float)(sum / (double)(uint32_t)(width * height));
    return 0;
}

int ssim_oracle_compute(uint32_t width, uint32_t height,
                        const uint8_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                        const uint8_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                        float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                        int tapsMode, float* ssim, double* sumOut)
{
    return oracle_core(width, height, 1, 255.0, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride, tapsMode, ssim, sumOut);
}

/* 16-bit pixels, L = 65535: the extension the reference's README names (README.md:107-111); the library does not
 * implement it, but the reference's own test oracle does: tests/ssim_naive.h:230-241 is a template on the pixel type T with
 * L = numeric_limits<T>::max().  PARITY PINNED to naive::compute_ssim<double, uint16_t> (compiled in place into
 * oracle/_ref/libnaive.so): tests/test_oracle.py checks this function against the vectors that shim produced
 * (tests/golden/golden.json "u16_naive", tests/golden/u16_pair.npz) and live against the shim, to 1e-12 on the double
 * mean with TAPS_RUNTIME; the scale invariance SSIM_16(257 a, 257 b) == SSIM_8(a, b) is checked as well. */
int ssim_oracle_compute_u16(uint32_t width, uint32_t height,
                            const uint16_t* a, ptrdiff_t stepA, ptrdiff_t strideA,
                            const uint16_t* b, ptrdiff_t stepB, ptrdiff_t strideB,
                            float* map, ptrdiff_t mapStep, ptrdiff_t mapStride,
                            int tapsMode, float* ssim, double* sumOut)
{
    return oracle_core(width, height, 2, 65535.0, a, stepA, strideA, b, stepB, strideB, map, mapStep, mapStride, tapsMode, ssim, sumOut);
}

int ssim_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
