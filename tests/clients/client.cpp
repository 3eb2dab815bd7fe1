// A caller written against the reference's C++ API (the call forms of sample/rmgr-ssim-sample.cpp:93-95 and
// tests/rmgr-ssim-tests.cpp:273-300 of the reference), compiled against include/rmgr/*.h and linked with librmgr-ssim.so.
// Usage: client <file>   file = "W H C\n" + W*H*C bytes of image A + W*H*C bytes of image B (interleaved channels)
// Prints one line per channel:  <ssim general> <ssim openmp> <ssim deprecated> <map sum> <heap ssim>
// With no argument: only the argument-validation paths (no device needed), prints "validation ok".
#include <rmgr/ssim.h>
#include <rmgr/ssim-openmp.h>
#include <rmgr/ssim-version.h>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <vector>

#if defined(__GNUC__)
#   pragma GCC diagnostic ignored "-Wdeprecated-declarations"
#endif

static int validation()
{
    float s = 0.f;
    rmgr::ssim::GeneralParams p = rmgr::ssim::GeneralParams();
    if (rmgr::ssim::compute_ssim(NULL, p) != EINVAL) return 1;                 // both outputs NULL (src/ssim.cpp:962-966)
    if (rmgr::ssim::compute_ssim(&s, p) != EINVAL) return 2;                   // NULL images (src/ssim.cpp:968-972)
    rmgr::ssim::ImgParams img;
    unsigned char px[12] = {0};
    if (img.init_interleaved(px, 6, 3, 3) != EINVAL) return 3;                 // channelNum >= channelCount (src/ssim.cpp:156-178)
    if (img.init_interleaved(px, 6, 3, 1) != 0 || img.topLeft != px + 1 || img.step != 3 || img.stride != 6) return 4;
    const unsigned char* planes[2] = {px, px + 6};
    const ptrdiff_t strides[2] = {2, 3};
    if (img.init_planar(planes, strides, 1) != 0 || img.topLeft != px + 6 || img.step != 1 || img.stride != 3) return 5;
    const rmgr::ssim::Version v = rmgr::ssim::get_version();
    if (v.major != RMGR_SSIM_VERSION_MAJOR || v.string == NULL) return 6;
    if (rmgr::ssim::get_errno(-22.f) != 22 || rmgr::ssim::get_errno(0.5f) != 0) return 7;
    rmgr::ssim::Params dp = rmgr::ssim::Params();
    if (rmgr::ssim::get_errno(rmgr::ssim::compute_ssim(dp)) != EINVAL) return 8;   // deprecated overload returns -errno (src/ssim.cpp:1119)
    std::puts("validation ok");
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 2) return validation();
    std::FILE* f = std::fopen(argv[1], "rb");
    unsigned w, h, c;
    if (!f || std::fscanf(f, "%u %u %u", &w, &h, &c) != 3 || std::fgetc(f) != '\n') return 10;
    std::vector<unsigned char> a((size_t)w * h * c), b(a.size());
    if (std::fread(&a[0], 1, a.size(), f) != a.size() || std::fread(&b[0], 1, b.size(), f) != b.size()) return 11;
    std::fclose(f);
    for (unsigned ch = 0; ch < c; ++ch) {
        std::vector<float> map((size_t)w * h);
        rmgr::ssim::Params p = rmgr::ssim::Params();
        p.width = w; p.height = h;
        p.imgA.init_interleaved(&a[0], (ptrdiff_t)w * c, c, ch);
        p.imgB.init_interleaved(&b[0], (ptrdiff_t)w * c, c, ch);
        p.ssimMap = &map[0]; p.ssimStep = 1; p.ssimStride = w;
        float s1 = -1.f, s2 = -1.f, s4 = -1.f;
        int rc = rmgr::ssim::compute_ssim(&s1, p);
        if (rc != 0) { std::fprintf(stderr, "compute_ssim failed: %d\n", rc); return 20 + rc; }
        double mapSum = 0;
        for (size_t i = 0; i < map.size(); ++i) mapSum += map[i];
        p.ssimMap = NULL;
        if ((rc = rmgr::ssim::compute_ssim_openmp(&s2, p)) != 0) return 40 + rc;
        const float s3 = rmgr::ssim::compute_ssim(p);                       // deprecated overload
        p.use_default_allocator();                                          // "heap" variant of the reference's test matrix
        if ((rc = rmgr::ssim::compute_ssim(&s4, p)) != 0) return 60 + rc;
        std::printf("%.9g %.9g %.9g %.9g %.9g\n", s1, s2, s3, mapSum, s4);
    }
    return 0;
}
