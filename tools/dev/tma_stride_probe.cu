// TMA probe: can a tiled tensor map with elementStrides[0] = C pick ONE channel of an interleaved u8 image (C = 3, 4) on the way
// into shared memory, at which start coordinates, and what lands for out-of-range columns?  (Input-side channel selection
// inside the fused kernel)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int boxBytes, uint8_t* out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = smem_u32(&bar), d = smem_u32(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(boxBytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(d), "l"(&tm), "r"(x), "r"(y), "r"(0), "r"(b) : "memory");
    }
    // bounded wait: a box that never completes must not hang the GPU
    bool done = false;
    for (int it = 0; it < 2000000 && !done; ++it) {
        uint32_t ok;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(ok) : "r"(b), "r"(0) : "memory");
        done = ok != 0;
    }
    for (int i = threadIdx.x; i < boxBytes; i += blockDim.x) out[i] = done ? smem[i] : 0xDD;
}
int main()
{
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int W = 333, H = 141;
    for (int C = 3; C <= 4; ++C) {
        const int pitch = (W * C + 15) / 16 * 16;
        std::vector<uint8_t> h(pitch * H);
        for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) h[y * pitch + x] = (uint8_t)(x * 7 + y * 13 + 3);
        uint8_t *d, *out; cudaMalloc(&d, pitch * H); cudaMalloc(&out, 4096); cudaMemcpy(d, h.data(), pitch * H, cudaMemcpyHostToDevice);
        const int boxPx = 48, boxH = 8;
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)W * C, H, 1}; cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
        cuuint32_t box[3] = {(cuuint32_t)(boxPx * C), (cuuint32_t)boxH, 1}; cuuint32_t es[3] = {(cuuint32_t)C, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("C=%d: encode box %d elements stride %d -> %d\n", C, boxPx * C, C, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int pxStarts[] = {0, 16, -16, 48, 304, 320};
        for (int px0 : pxStarts) for (int ch = 0; ch < C; ++ch) {
            const int x = px0 * C + ch, y = 5, bytes = boxPx * boxH;
            cudaMemset(out, 0xEE, 4096);
            probe<<<1, 32, 4096>>>(tm, x, y, bytes, out);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
            int bad = 0, timeout = 0;
            for (int r2 = 0; r2 < boxH; ++r2) for (int k = 0; k < boxPx; ++k) {
                const int xx = x + k * C, yy = y + r2;
                const uint8_t want = (xx < 0 || xx >= W * C || yy < 0 || yy >= H) ? 0 : h[yy * pitch + xx];
                if (o[r2 * boxPx + k] == 0xDD) ++timeout; else if (o[r2 * boxPx + k] != want) ++bad;
            }
            printf("  C=%d pixel %4d channel %d (coordinate %5d): run=%s mismatches=%d timeouts=%d first bytes %02x %02x %02x want %02x %02x %02x\n", C, px0, ch, x,
                   cudaGetErrorString(e), bad, timeout, o[0], o[1], o[2],
                   (x < 0) ? 0 : h[y * pitch + x], (x + C < 0) ? 0 : h[y * pitch + x + C], (x + 2 * C < 0) ? 0 : h[y * pitch + x + 2 * C]);
            if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
        }
        cudaFree(d); cudaFree(out);
    }
    return 0;
}
