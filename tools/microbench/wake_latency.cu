// Hand-over latency between two warps of a CTA through an mbarrier: time from the arrive of the signalling warp to the
// return of the waiting warp, for (0) a blocking try_wait loop, (1) a test_wait spin, (2) try_wait with a 20 ns suspend hint.
#include <cstdio>
#include <cstdint>
__global__ void k(int mode, unsigned delayNs, long long* out)
{
    __shared__ uint64_t bar;
    __shared__ long long tArrive;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" :: "r"(b));
    __syncthreads();
    long long acc = 0;
    for (int it = 0; it < 32; ++it) {
        const uint32_t parity = it & 1;
        if (threadIdx.x >= 32) {                       // signaller
            __nanosleep(delayNs);
            if (threadIdx.x == 32) tArrive = clock64();
            __syncwarp();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(b) : "memory");
        } else {                                       // waiter
            uint32_t ok = 0;
            while (!ok) {
                if (mode == 0)      asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
                else if (mode == 1) asm volatile("{ .reg .pred P1; mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
                else                asm volatile("{ .reg .pred P1; mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3; selp.u32 %0, 1, 0, P1; }" : "=r"(ok) : "r"(b), "r"(parity), "r"(20u) : "memory");
            }
            const long long t = clock64();
            if (threadIdx.x == 0) acc += t - tArrive;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = acc / 32;
}
int main()
{
    long long* d; cudaMalloc(&d, 8); long long h;
    const char* names[3] = {"try_wait (blocking)", "test_wait spin", "try_wait, 20 ns suspend hint"};
    for (unsigned delay : {500u, 2000u})
        for (int mode = 0; mode < 3; ++mode) {
            k<<<1, 64>>>(mode, delay, d); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            printf("signal after ~%u ns, %-30s: waiter returns %lld cycles after the arrive\n", delay, names[mode], h);
        }
    return 0;
}
