// minimal TMA probe: which (rank, width, box) combinations work for u8 planes
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, int x, int y, int boxBytes, uint8_t* out, int dstOff)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint32_t b = smem_u32(&bar), d = smem_u32(smem) + dstOff;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(boxBytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(d), "l"(&tm), "r"(x), "r"(y), "r"(0), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(d), "l"(&tm), "r"(x), "r"(y), "r"(b) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" :: "r"(b), "r"(0) : "memory");
    for (int i = threadIdx.x; i < boxBytes; i += blockDim.x) out[i] = smem[i + dstOff];
}
int main()
{
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    printf("entry point %p q=%d\n", fp, (int)q);
    const int W = 333, H = 141, pitch = 336;
    std::vector<uint8_t> h(pitch * H);
    for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) h[y * pitch + x] = (uint8_t)(x * 7 + y * 13);
    uint8_t *d, *out; cudaMalloc(&d, pitch * H); cudaMalloc(&out, 4096); cudaMemcpy(d, h.data(), pitch * H, cudaMemcpyHostToDevice);
    struct Case { int rank, boxW, boxH, x, y, off; };
    Case cases[] = {{3, 64, 16, -16, 8, 0}, {3, 64, 1, 16, 5, 64}, {3, 64, 1, 16, 5, 192}, {3, 64, 1, -16, 140, 16}, {3, 64, 1, -16, 140, 32}, {3, 128, 1, 0, 3, 80}};
    for (Case c : cases) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {W, H, 1}; cuuint64_t strides[2] = {pitch, (cuuint64_t)pitch * H};
        cuuint32_t box[3] = {(cuuint32_t)c.boxW, (cuuint32_t)c.boxH, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, c.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        int bytes = c.boxW * c.boxH;
        cudaMemset(out, 0xEE, 4096);
        if (c.rank == 3) probe<3><<<1, 32, 4096>>>(tm, c.x, c.y, bytes, out, c.off); else probe<2><<<1, 32, 4096>>>(tm, c.x, c.y, bytes, out, c.off);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint8_t> o(bytes); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int r2 = 0; r2 < c.boxH; ++r2) for (int k = 0; k < c.boxW; ++k) {
            int xx = c.x + k, yy = c.y + r2; uint8_t want = (xx < 0 || xx >= W || yy < 0 || yy >= H) ? 0 : h[yy * pitch + xx];
            if (o[r2 * c.boxW + k] != want) ++bad;
        }
        printf("rank %d box %dx%d at (%d,%d) smem+%d: encode=%d run=%s mismatches=%d\n", c.rank, c.boxW, c.boxH, c.x, c.y, c.off, (int)r, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
    }
    return 0;
}
