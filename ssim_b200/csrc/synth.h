/*
 * synth.h -- deterministic synthetic 8-bit grayscale frame pairs (integer-only, so host C, CUDA and
 * numpy produce identical bytes).  Recipe: SURVEY.md section 8(d); it is the input of every
 * BASELINE.json config except the tests/images one.  The pair contains exact-match blocks (SSIM == 1),
 * smooth ramps with tiny noise (worst case for fp32 cancellation) and heavy noise (negative SSIM).
 */
#ifndef SSIM_B200_SYNTH_H
#define SSIM_B200_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SSIM_SYNTH_FN __host__ __device__ static inline
#else
#define SSIM_SYNTH_FN static inline
#endif

#define SSIM_SYNTH_DEFAULT_SEED 0x5517ull

SSIM_SYNTH_FN uint64_t ssim_synth_splitmix(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

SSIM_SYNTH_FN int ssim_synth_tri(int v, int p)
{
    const int m = v % (2 * p);
    return m < p ? p - m : m - p;
}

/* one pixel of frame f at (x, y): *a = reference image, *b = distorted image */
SSIM_SYNTH_FN void ssim_synth_pixel(uint64_t seed, uint32_t f, uint32_t x, uint32_t y, uint8_t* a, uint8_t* b)
{
    const uint64_t h = ssim_synth_splitmix(seed ^ ssim_synth_splitmix(((uint64_t)f << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)x));
    int base = (ssim_synth_tri((int)(2 * x + y), 256) + ssim_synth_tri((int)(x + 3 * y), 512) / 2) / 2;
    if (((x >> 7) + (y >> 7)) & 1)
        base += (int)(h & 15) - 8;
    const int va = base < 0 ? 0 : (base > 255 ? 255 : base);
    const int levels[7] = {0, 1, 2, 4, 8, 16, 32};
    const int k  = levels[((x >> 8) + 3 * (y >> 8) + f) % 7];
    const int vb = va + (int)((h >> 16) % (uint64_t)(2 * k + 1)) - k;
    *a = (uint8_t)va;
    *b = (uint8_t)(vb < 0 ? 0 : (vb > 255 ? 255 : vb));
}

/* order-dependent 64-bit checksum of a pair, pinned in tests/golden/synthetic.json */
SSIM_SYNTH_FN uint64_t ssim_synth_checksum_step(uint64_t cs, uint8_t a, uint8_t b)
{
    return (cs * 1099511628211ull) ^ (uint64_t)a ^ ((uint64_t)b << 8);
}

#endif
