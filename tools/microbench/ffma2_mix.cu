// Microbenchmark: how well does the B200 SMSP overlap packed FFMA2 (2 FMA-pipe cycles each) with the other instruction
// classes of the SSIM kernel?  Each variant runs 44 in-place FFMA2 (uniform-register tap) per "row" plus an extra mix,
// at 1..4 warps per SMSP, and reports cycles per row against the FMA-pipe minimum.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(u64 v) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

constexpr int ROWS = 2048;
// MODE 0: FFMA2 only; 1: + 18-long dependent scalar FFMA chain; 2: + 18 independent scalar FFMA; 3: + 20 integer ALU ops;
// 4: + 4 LDS.64; 5: chain + ALU + LDS (the consumer's real mix); 6: 44 scalar FFMA pairs instead of FFMA2 (88 FFMA) + chain
template <int MODE>
__global__ void __launch_bounds__(128) k_mix(float* out, float g0, float g1, float g2, float g3, float g4, float g5, int zero)
{
    __shared__ u64 sm[128 * 4];
    for (int i = threadIdx.x; i < 512; i += 128) sm[i] = pack2(1.0f + i * 1e-6f, 0.5f);
    __syncthreads();
    u64 w[6] = {pack2(g0, g0), pack2(g1, g1), pack2(g2, g2), pack2(g3, g3), pack2(g4, g4), pack2(g5, g5)};
    u64 q[44];
    #pragma unroll
    for (int i = 0; i < 44; ++i) q[i] = pack2(threadIdx.x * 1e-3f + i, 1.0f);
    float f[4] = {1.f, 2.f, 3.f, 4.f};
    float c = g0, acc = 0.f;
    int ia = threadIdx.x, ib = zero, ja = 1; float xprev = g2;
    u64 h = pack2(g1, g2);
    const u64* sp = sm + threadIdx.x;
    #pragma unroll 1
    for (int r = 0; r < ROWS; ++r) {
        if (MODE == 4 || MODE == 5) {
            u64 l0 = sp[0], l1 = sp[128 ^ ib], l2 = sp[256], l3 = sp[384 ^ ib];
            h = fma2(l0, w[0], l1); h = fma2(l2, w[1], h); h = fma2(l3, w[2], h);
        }
        if (MODE != 6 && MODE != 7 && MODE != 8) {
            #pragma unroll
            for (int i = 0; i < 44; ++i) q[i] = fma2(h, w[i % 6], q[i]);
        } else {
            #pragma unroll
            for (int i = 0; i < 44; ++i) {
                float lo, hi, hl, hh; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(q[i])); asm("mov.b64 {%0,%1}, %2;" : "=f"(hl), "=f"(hh) : "l"(h));
                lo = fmaf(hl, g0 + 0.f, lo); hi = fmaf(hh, g1, hi); q[i] = pack2(lo, hi);
            }
        }
        if (MODE == 1 || MODE == 5 || MODE == 6) {
            float x = lo2(q[r & 3 ? 0 : 1]) * 0.f + c;   // depends on this row's result
            #pragma unroll
            for (int k = 0; k < 18; ++k) x = fmaf(x, 0.999f, c);
            acc += x;
        }
        if (MODE == 2) {
            #pragma unroll
            for (int k = 0; k < 18; ++k) f[k & 3] = fmaf(f[k & 3], 0.999f, c);
        }
        if (MODE == 3 || MODE == 5) {
            #pragma unroll
            for (int k = 0; k < 20; ++k) { ia = (ia + ib + k) ^ (ia >> 3); }
        }
        if (MODE == 7 || MODE == 8) {
            // software-pipelined: the dependent chain of the PREVIOUS row is interleaved with this row's FFMA2 stream
            float x = xprev;
            #pragma unroll
            for (int i = 0; i < 44; ++i) {
                q[i] = fma2(h, w[i % 6], q[i]);
                if ((i & 1) == 0 && i < 36) x = fmaf(x, 0.999f, c);
                if (MODE == 8 && (i % 4) == 1 && i < 40) { ia = (ia + ib + i); ja = ja ^ (ia >> 3); }
            }
            acc += x;
            xprev = lo2(q[r & 3 ? 0 : 1]) * 0.f + c;
        }
        h = fma2(h, w[5], w[1]);
    }
    float s = acc + f[0] + f[1] + f[2] + f[3] + (float)ia + (float)ja + xprev;
    #pragma unroll
    for (int i = 0; i < 44; ++i) s += lo2(q[i]);
    out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int MODE> static void run(const char* name, int ctasPerSm, int sms, float* out, double fmaCyclesPerRow)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * ctasPerSm;
    for (int i = 0; i < 2; ++i) k_mix<MODE><<<blocks, 128>>>(out, .26f, .21f, .11f, .04f, .008f, .001f, 0);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) { cudaEventRecord(e0); k_mix<MODE><<<blocks, 128>>>(out, .26f, .21f, .11f, .04f, .008f, .001f, 0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double cyc = best * 1e-3 * 1965e6;                  // SM cycles elapsed
    double rowsPerSmsp = (double)ROWS * ctasPerSm;      // each CTA = 4 warps = 1 warp per SMSP
    double cycPerRow = cyc / rowsPerSmsp;
    printf("%-34s warps/SMSP=%d  cycles/row/warp-slot=%7.1f  FMA-pipe min=%5.0f  pipe util=%5.1f%%\n", name, ctasPerSm, cycPerRow, fmaCyclesPerRow, 100.0 * fmaCyclesPerRow / cycPerRow);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 128);
    for (int w : {2, 3, 4}) {
        run<0>("ffma2 x45", w, sms, out, 90);
        run<1>("ffma2 x45 + 18 dependent FFMA", w, sms, out, 90 + 19);
        run<2>("ffma2 x45 + 18 independent FFMA", w, sms, out, 90 + 18);
        run<3>("ffma2 x45 + 40 int ALU", w, sms, out, 90);
        run<4>("ffma2 x48 + 4 LDS.64", w, sms, out, 96);
        run<5>("ffma2 x48 + chain + ALU + LDS", w, sms, out, 96 + 19);
        run<6>("scalar ffma x88 + chain", w, sms, out, 88 + 19 + 2);
        run<7>("ffma2 x45 + chain, sw-pipelined", w, sms, out, 90 + 19);
        run<8>("ffma2 x45 + chain + 20 ALU, sw-pipelined", w, sms, out, 90 + 19);
        printf("\n");
    }
    return 0;
}
