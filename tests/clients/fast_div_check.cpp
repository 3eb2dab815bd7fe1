// Host-only check of ssimk::fast_div() (ssim_b200/csrc/ssim_kernels.h): the multiply-shift constants the fused kernel uses to
// decode work-item indices must reproduce n / d exactly for every divisor the host can produce and every n < 2^31.
#include <cstdint>
#include <cstdio>
#include "ssim_kernels.h"

static uint32_t apply(uint32_t n, uint32_t mul, uint32_t shift) { return mul ? (uint32_t)(((uint64_t)n * mul) >> 32) >> shift : n; }

int main()
{
    uint64_t checked = 0;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    for (uint32_t d = 1; d <= 70000; d = d < 4200 ? d + 1 : d + 977) {
        uint32_t mul, shift;
        ssimk::fast_div(d, &mul, &shift);
        const uint32_t edge[] = {0u, 1u, d - 1, d, d + 1, 2 * d - 1, 2 * d, 0x7fffffffu, 0x7ffffffeu, 0x7fffffffu / d * d, 0x7fffffffu / d * d - 1};
        for (uint32_t n : edge) {
            if (n > 0x7fffffffu) continue;
            if (apply(n, mul, shift) != n / d) { std::printf("FAIL d=%u n=%u\n", d, n); return 1; }
            ++checked;
        }
        for (int i = 0; i < 2000; ++i) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            const uint32_t n = (uint32_t)(rng >> 33);            // < 2^31
            if (apply(n, mul, shift) != n / d) { std::printf("FAIL d=%u n=%u\n", d, n); return 1; }
            ++checked;
        }
    }
    std::printf("fast_div ok (%llu checks)\n", (unsigned long long)checked);
    return 0;
}
