/* The same through the C API, compiled as C99 (reference include/rmgr/ssim.h:428-560). */
#include <rmgr/ssim.h>
#include <rmgr/ssim-openmp.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv)
{
    rmgr_ssim_Version v;
    rmgr_ssim_Params p;
    float s = 0.f;
    if (rmgr_ssim_get_version(NULL) != EINVAL || rmgr_ssim_get_version(&v) != 0) return 1;
    if (rmgr_ssim_compute_ssim(&s, NULL, NULL) != EINVAL) return 2;
    if (argc < 2) { printf("validation ok %u.%u.%u %s\n", v.major, v.minor, v.patch, v.string); return 0; }
    {
        FILE* f = fopen(argv[1], "rb");
        unsigned w, h, c;
        unsigned char *a, *b;
        const unsigned char* planes[1];
        ptrdiff_t strides[1];
        if (!f || fscanf(f, "%u %u %u", &w, &h, &c) != 3 || fgetc(f) != '\n' || c != 1) return 10;
        a = (unsigned char*)malloc((size_t)w * h); b = (unsigned char*)malloc((size_t)w * h);
        if (fread(a, 1, (size_t)w * h, f) != (size_t)w * h || fread(b, 1, (size_t)w * h, f) != (size_t)w * h) return 11;
        fclose(f);
        p.width = w; p.height = h; p.ssimMap = NULL; p.ssimStep = 0; p.ssimStride = 0; p.alloc = NULL; p.dealloc = NULL;
        planes[0] = a; strides[0] = (ptrdiff_t)w;
        if (rmgr_ssim_init_planar(&p.imgA, planes, strides, 0) != 0) return 12;
        planes[0] = b;
        if (rmgr_ssim_init_planar(&p.imgB, planes, strides, 0) != 0) return 13;
        if (rmgr_ssim_use_default_allocator(&p) != 0) return 14;
        if (rmgr_ssim_compute_ssim(&s, &p, NULL) != 0) return 15;
        printf("%.9g", s);
        if (rmgr_ssim_compute_ssim_openmp(&s, &p) != 0) return 16;
        printf(" %.9g\n", s);
        free(a); free(b);
    }
    return 0;
}
