// How long does __nanosleep(t) really suspend a warp on B200, and how long does one mbarrier.try_wait on an incomplete phase block?
#include <cstdio>
#include <cstdint>
__global__ void k(unsigned ns, long long* out, int mode)
{
    __shared__ uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < 64; ++i) {
        if (mode == 0) __nanosleep(ns);
        else {
            uint32_t ok;
            asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(b), "r"(0u) : "memory");
            if (ok) out[1] = 1;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / 64;
}
int main()
{
    long long* d; cudaMalloc(&d, 16); long long h;
    int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    for (unsigned ns : {0u, 20u, 100u, 200u, 1000u, 3000u}) {
        k<<<1, 32>>>(ns, d, 0); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("__nanosleep(%u): %lld cycles per call (%.0f ns at %d MHz nominal)\n", ns, h, h * 1e6 / mhz, mhz / 1000);
    }
    k<<<1, 32>>>(0, d, 1); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("mbarrier.try_wait on an incomplete phase: %lld cycles per call (%.0f ns)\n", h, h * 1e6 / mhz);
    return 0;
}
