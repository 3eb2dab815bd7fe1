// Microbenchmark: FP32 pipe throughput on sm_100a (B200).
// Measures lane-ops/clk/SM for FFMA (3-reg, const-operand), packed FFMA2/FADD2/FMUL2,
// u8->f32 conversion idioms, MUFU.RCP, and FFMA(2)+LDS co-issue. Output feeds DESIGN.md's
// roofline discussion (which instruction mix the fused SSIM kernel should be built from).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ float lo2(unsigned long long v) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo + hi; }

// ---- 1. scalar FFMA, 16 independent chains, 3 register operands
__global__ void k_ffma(float* out, float w0, float w1) {
    float a[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float x = w0, y = w1;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0; 
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 1b. scalar FFMA in the "shifted accumulator" pattern  q[k] = fma(w[k], h, q[k+1])
__global__ void k_ffma_shift(float* out, float w0, float w1) {
    float q[12], w[11];
    #pragma unroll
    for (int i = 0; i < 12; ++i) q[i] = threadIdx.x * 0.001f + i;
    #pragma unroll
    for (int i = 0; i < 11; ++i) w[i] = w0 + i * w1;
    float h = w1;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int k = 0; k < 11; ++k) q[k] = fmaf(w[k], h, q[k + 1]);
        q[11] = q[0] * 0.5f; h += 1.0f;
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 12; ++i) s += q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 2. packed FFMA2 16 chains
__global__ void k_ffma2(float* out, float w0, float w1) {
    unsigned long long a[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = pack2(threadIdx.x * 0.001f + i, i);
    unsigned long long x = pack2(w0, w0), y = pack2(w1, w1);
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x), "l"(y));
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += lo2(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 2b. packed FFMA2 in the shifted-accumulator pattern, 11 distinct weights
__global__ void k_ffma2_shift(float* out, float w0, float w1) {
    unsigned long long q[12], w[6];
    #pragma unroll
    for (int i = 0; i < 12; ++i) q[i] = pack2(threadIdx.x * 0.001f + i, i);
    #pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = pack2(w0 + i * w1, w0 + i * w1);
    unsigned long long h = pack2(w1, w0);
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int k = 0; k < 11; ++k) {
            int wi = k < 6 ? k : 10 - k;
            asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q[k]) : "l"(w[wi]), "l"(h), "l"(q[k + 1]));
        }
        asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q[11]) : "l"(q[0]), "l"(w[0]));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(h) : "l"(w[1]));
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 12; ++i) s += lo2(q[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 3. packed FADD2 / FMUL2
__global__ void k_fadd2(float* out, float w0, float w1) {
    unsigned long long a[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = pack2(threadIdx.x * 0.001f + i, i);
    unsigned long long x = pack2(w0, w1);
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(x));
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += lo2(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 4. FFMA2 : LDS.128 at 8:1 (co-issue check)
__global__ void k_ffma2_lds(float* out, float w0, float w1) {
    __shared__ float4 sm[256 * 2];
    sm[threadIdx.x] = make_float4(w0, w1, w0, w1); sm[threadIdx.x + 256] = make_float4(w1, w0, w1, w0);
    __syncthreads();
    unsigned long long a[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = pack2(threadIdx.x * 0.001f + i, i);
    unsigned long long y = pack2(w1, w1);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        float4 v0 = sm[idx], v1 = sm[idx ^ 256 ^ (it & 1)];
        unsigned long long x0 = pack2(v0.x, v0.y), x1 = pack2(v1.z, v1.w);
        #pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x0), "l"(y));
        #pragma unroll
        for (int i = 8; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(x1), "l"(y));
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += lo2(a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 4b. FFMA : LDS.128 at 8:1
__global__ void k_ffma_lds(float* out, float w0, float w1) {
    __shared__ float4 sm[256 * 2];
    sm[threadIdx.x] = make_float4(w0, w1, w0, w1); sm[threadIdx.x + 256] = make_float4(w1, w0, w1, w0);
    __syncthreads();
    float a[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        float4 v0 = sm[idx], v1 = sm[idx ^ 256 ^ (it & 1)];
        #pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], v0.x, v0.y);
        #pragma unroll
        for (int i = 8; i < 16; ++i) a[i] = fmaf(a[i], v1.z, v1.w);
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 5. u8 -> f32 idioms: (a) cvt.rn.f32.u8 style via byte extract, (b) PRMT magic + FADD
__global__ void k_cvt_i2f(float* out, const uint32_t* in) {
    uint32_t v = in[threadIdx.x & 31];
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            s0 += (float)(v & 0xff); s1 += (float)((v >> 8) & 0xff); s2 += (float)((v >> 16) & 0xff); s3 += (float)(v >> 24);
            v = v * 1664525u + 1013904223u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
__global__ void k_cvt_prmt(float* out, const uint32_t* in) {
    uint32_t v = in[threadIdx.x & 31];
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const float M = 8388608.0f + 128.0f;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            s0 += __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440)) - M;
            s1 += __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7441)) - M;
            s2 += __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7442)) - M;
            s3 += __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7443)) - M;
            v = v * 1664525u + 1013904223u;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3;
}
// ---- 6. MUFU.RCP mixed 1:8 with FFMA
__global__ void k_rcp(float* out, float w0, float w1) {
    float a[8];
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i + 1.0f;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __frcp_rn(a[i]) ;
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rcp_approx(float* out, float w0, float w1) {
    float a[8];
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i + 1.0f;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// ---- 7. DADD throughput
__global__ void k_dadd(float* out, double w0) {
    double a[8];
    #pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001 + i;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += w0;
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

template <typename F>
static int run(const char* name, F launch, double lane_ops_per_thread, int blocks, int threads, int sms, double* clk_mhz_out = nullptr) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double total = lane_ops_per_thread * (double)blocks * threads;
    double per_s = total / (best * 1e-3);
    printf("%-18s blocks=%5d thr=%4d  %8.3f ms  %9.2f Gops/s  => %7.2f lane-ops/clk/SM @1965MHz\n",
           name, blocks, threads, best, per_s * 1e-9, per_s / (1965e6 * sms));
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device: %s  SMs=%d  smem/SM=%zu  smem/block optin=%zu  regs/SM=%d  clockRate=%d kHz  L2=%d MB\n",
           p.name, sms, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor, p.clockRate, p.l2CacheSize >> 20);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 16 * 256));
    uint32_t* in; CK(cudaMalloc(&in, 128)); CK(cudaMemset(in, 0x5a, 128));
    for (int occ : {1, 2, 4, 8}) {
        int blocks = sms * occ, thr = 256;
        printf("--- %d CTA(s) of 256 threads per SM (%d warps/SM)\n", occ, occ * 8);
        run("ffma", [&] { k_ffma<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 16.0 * ITERS, blocks, thr, sms);
        run("ffma_shift", [&] { k_ffma_shift<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 13.0 * ITERS, blocks, thr, sms);
        run("ffma2 (x2 lanes)", [&] { k_ffma2<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 32.0 * ITERS, blocks, thr, sms);
        run("ffma2_shift (x2)", [&] { k_ffma2_shift<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 26.0 * ITERS, blocks, thr, sms);
        run("fadd2 (x2)", [&] { k_fadd2<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 32.0 * ITERS, blocks, thr, sms);
        run("ffma2+lds 8:1", [&] { k_ffma2_lds<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 32.0 * ITERS, blocks, thr, sms);
        run("ffma+lds 8:1", [&] { k_ffma_lds<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 16.0 * ITERS, blocks, thr, sms);
        run("cvt i2f (4 cvt)", [&] { k_cvt_i2f<<<blocks, thr>>>(out, in); }, 16.0 * ITERS, blocks, thr, sms);
        run("cvt prmt+fadd", [&] { k_cvt_prmt<<<blocks, thr>>>(out, in); }, 16.0 * ITERS, blocks, thr, sms);
        run("rcp_rn", [&] { k_rcp<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 8.0 * ITERS, blocks, thr, sms);
        run("rcp.approx", [&] { k_rcp_approx<<<blocks, thr>>>(out, 1.0001f, 0.5f); }, 8.0 * ITERS, blocks, thr, sms);
        run("dadd", [&] { k_dadd<<<blocks, thr>>>(out, 1.0001); }, 8.0 * ITERS, blocks, thr, sms);
    }
    return 0;
}
