// Does a packed FFMA2 (2 FMA-pipe cycles) leave its second cycle's ISSUE slot free for other pipes?
// Per iteration: NF independent FFMA2 (asm volatile, uniform tap) + NA independent ALU ops (asm volatile lop3 on 8 chains)
// + NS independent scalar FFMA.  Prints cycles/iteration/SMSP-warp-slot against the two models.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
constexpr int ITERS = 4096;
template <int NF, int NA, int NS, int AOP = 0>
__global__ void __launch_bounds__(128) k(float* out, float g, unsigned seed)
{
    u64 q[16]; unsigned a[8]; float s[8];
    #pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("mov.b64 %0, {%1,%2};" : "=l"(q[i]) : "f"(threadIdx.x * 1e-3f + i), "f"(g + i));
    #pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; s[i] = g + i; }
    u64 w, h; asm volatile("mov.b64 %0, {%1,%1};" : "=l"(w) : "f"(g)); asm volatile("mov.b64 %0, {%1,%2};" : "=l"(h) : "f"(g), "f"(1.0f - g));
    #pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int j = 0; j < (NF > NA ? (NF > NS ? NF : NS) : (NA > NS ? NA : NS)); ++j) {
            if (j < NF) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(q[j & 15]) : "l"(h), "l"(w));
            if (j < NA) {
                if (AOP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j & 7]) : "r"(seed), "r"(a[(j + 3) & 7]));
                if (AOP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j & 7]) : "r"(seed));                       // IADD3, independent chains
                if (AOP == 2) asm volatile("prmt.b32 %0, %0, %1, 0x7440;" : "+r"(a[j & 7]) : "r"(seed));             // PRMT
                if (AOP == 3) asm volatile("mov.b32 %0, %1;" : "=r"(a[j & 7]) : "r"(a[(j + 1) & 7]));                  // MOV
                if (AOP == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a[j & 7]) : "r"((unsigned)(threadIdx.x * 4)));   // LDS
            }
            if (j < NS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[j & 7]) : "f"(g), "f"(1.0f - g));
        }
    }
    float r = 0; unsigned x = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) { float lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[i])); r += lo + hi; }
    #pragma unroll
    for (int i = 0; i < 8; ++i) { x ^= a[i]; r += s[i]; }
    out[blockIdx.x * 128 + threadIdx.x] = r + (float)x;
}
template <int NF, int NA, int NS, int AOP = 0> static void run(int ctasPerSm, int sms, float* out)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * ctasPerSm;
    k<NF, NA, NS, AOP><<<blocks, 128>>>(out, 0.5f, 7u); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) { cudaEventRecord(e0); k<NF, NA, NS, AOP><<<blocks, 128>>>(out, 0.5f, 7u); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double cyc = best * 1e-3 * 1965e6 / ((double)ITERS * ctasPerSm);
    int pipe = 2 * NF + NS, issue = NF + NA + NS;
    printf("op%d FFMA2=%2d ALU=%2d FFMA=%2d  warps/SMSP=%d  cycles/iter/warp-slot=%6.1f   model max(pipe,issue)=%3d   model FFMA2-holds-issue-port=%3d\n",
           AOP, NF, NA, NS, ctasPerSm, cyc, pipe > issue ? pipe : issue, 2 * NF + NA + NS);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 128);
    const int w = 4;
    printf("ops: 0=LOP3(3 reg) 1=IADD 2=PRMT 3=MOV 4=LDS.32\n");
    run<0, 32, 0, 0>(w, sms, out); run<0, 32, 0, 1>(w, sms, out); run<0, 32, 0, 2>(w, sms, out); run<0, 32, 0, 3>(w, sms, out); run<0, 32, 0, 4>(w, sms, out);
    printf("\n");
    run<32, 0, 0, 1>(w, sms, out);
    run<32, 8, 0, 1>(w, sms, out); run<32, 16, 0, 1>(w, sms, out); run<32, 32, 0, 1>(w, sms, out);
    run<32, 16, 0, 2>(w, sms, out); run<32, 16, 0, 3>(w, sms, out); run<32, 16, 0, 4>(w, sms, out); run<32, 8, 0, 4>(w, sms, out);
    printf("\n");
    run<0, 16, 32, 1>(w, sms, out); run<0, 32, 32, 1>(w, sms, out); run<0, 32, 64, 1>(w, sms, out);
    return 0;
}
