// Host-only check of ssimk::plan_segments(): the partition must cover every output row exactly once with non-empty
// segments for any shape, and reproduce the choices the measurements in DESIGN.md were taken with.
#include <cstdint>
#include <cstdio>
#include "ssim_kernels.h"

static int fails = 0;
static void expect(bool ok, const char* what, uint32_t w, uint32_t h, uint32_t f, int rows, int segs)
{
    if (!ok) { std::printf("FAIL %s: %ux%u x%u -> %d rows x %d segments\n", what, w, h, f, rows, segs); ++fails; }
}

int main()
{
    const long long slots = 148 * 2;          // B200: 148 SMs x 2 CTAs
    const uint32_t widths[] = {1, 63, 64, 65, 640, 1920, 3840, 16384, 100000};
    const uint32_t heights[] = {1, 2, 23, 24, 25, 47, 48, 141, 1080, 2058, 2160, 16384, 1000003};
    const uint32_t frames[] = {1, 3, 64, 512, 4096};
    for (uint32_t w : widths) for (uint32_t h : heights) for (uint32_t f : frames) {
        int rows = 0, segs = 0;
        ssimk::plan_segments(slots, w, h, f, 0, &rows, &segs);
        expect(rows >= 1 && segs >= 1, "positive", w, h, f, rows, segs);
        expect((long long)rows * segs >= h && (long long)rows * (segs - 1) < h, "covers every row exactly once, last segment not empty", w, h, f, rows, segs);
        expect(h < 48 || rows >= 24, "segments of at least 24 rows", w, h, f, rows, segs);
        for (int forced : {1, 7, 100, 5000}) {
            ssimk::plan_segments(slots, w, h, f, forced, &rows, &segs);
            expect((long long)rows * segs >= h && (long long)rows * (segs - 1) < h && rows <= (forced > (int)h ? (int)h : forced), "override", w, h, f, rows, segs);
        }
    }
    int rows, segs;
    ssimk::plan_segments(slots, 3840, 2160, 64, 0, &rows, &segs);   expect(rows == 540 && segs == 4, "64 x 4K: 12.97 waves", 3840, 2160, 64, rows, segs);
    ssimk::plan_segments(slots, 3840, 2160, 1, 0, &rows, &segs);    expect(segs == 19, "one 4K pair: one wave of 285 CTAs", 3840, 2160, 1, rows, segs);
    ssimk::plan_segments(slots, 16384, 2058, 1, 0, &rows, &segs);   expect(segs == 9, "16384 x 2058 strip: 1.95 waves", 16384, 2058, 1, rows, segs);
    ssimk::plan_segments(slots, 1920, 1080, 1, 0, &rows, &segs);    expect(segs == 39, "one 1080p pair: one wave", 1920, 1080, 1, rows, segs);
    std::printf(fails ? "plan_segments: %d failures\n" : "plan_segments ok\n", fails);
    return fails ? 1 : 0;
}
